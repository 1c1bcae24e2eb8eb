"""The N>1 path of bench.py on CPU: two processes over gloo shard the windows with no data-path collective and
reduce the timing as the max over ranks (what the nccl path does on the GPU box)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _worker(rank, world, port, n_total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = bench.shard_windows(n_total, rank, world)
    # every rank "processes" its own windows; only the timing is reduced
    t = torch.tensor([10.0 * (rank + 1) + len(mine)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        out.put((float(t.item()), gathered))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [128, 7])
def test_two_rank_sharding_gloo(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n_total
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    t_max, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    flat = [w for part in gathered for w in part]
    assert sorted(flat) == list(range(n_total)) and len(set(flat)) == n_total     # a partition: no overlap, no gap
    assert abs(len(gathered[0]) - len(gathered[1])) <= 1
    assert t_max == 20.0 + len(gathered[1])                                        # max over ranks


def test_shard_windows_partition_property():
    for world in (1, 2, 4, 8):
        for n in (0, 1, 63, 64, 512):
            parts = [bench.shard_windows(n, r, world) for r in range(world)]
            flat = [w for p in parts for w in p]
            assert flat == list(range(n))


def test_library_chunk_partition_is_contiguous_and_balanced():
    """whisper_b200_partition_owner (host logic of the multi-GPU entry point): chunk -> GPU in contiguous blocks whose sizes
    differ by at most one, every GPU used when there are enough chunks."""
    import open_whisper_kit_b200 as pkg
    lib = pkg.load()
    for n_chunks in (1, 2, 5, 8, 15, 64, 120):
        for n_gpus in (1, 2, 3, 4, 8):
            owners = [lib.whisper_b200_partition_owner(i, n_chunks, n_gpus) for i in range(n_chunks)]
            assert owners == sorted(owners) and owners[0] == 0 and max(owners) < n_gpus
            counts = [owners.count(g) for g in range(n_gpus)]
            if n_chunks >= n_gpus:
                assert min(counts) >= 1 and max(counts) - min(counts) <= 1, (n_chunks, n_gpus, counts)
