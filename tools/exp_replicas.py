"""Experiment: does running the 64-window batch as R concurrent half-batches on ONE GPU (R contexts on device 0 through the
group API, each with its own streams and workspaces) beat the single lock-step batch?  Prints ms per pass for R = 1, 2, 3, 4."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import open_whisper_kit_b200 as pkg  # noqa: E402


def main():
    import torch
    lib = pkg.load()
    arch = sys.argv[1] if len(sys.argv) > 1 else "large-v3"
    n_win = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    path = bench.ensure_model(arch, 0, lambda: None)
    wl = bench.Workload(list(range(n_win)))
    p = bench.greedy_params(lib, no_timestamps=True)
    for R in [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "1,2,3,4").split(",")]:
        devs = (C.c_int * R)(*([0] * R))
        cp = lib.whisper_context_default_params()
        g = lib.whisper_b200_group_init_from_file(path.encode(), cp, devs, R)
        assert g
        ts = []
        for it in range(4):
            torch.cuda.synchronize()
            t0 = time.time()
            rc = lib.whisper_b200_group_full_parallel(g, p, wl.host_ptr, wl.n_samples, wl.n_win)
            torch.cuda.synchronize()
            ts.append((time.time() - t0) * 1e3)
            assert rc == 0
        ctx0 = lib.whisper_b200_group_context(g, 0)
        print(f"replicas on one GPU = {R}: ms/pass {['%.1f' % t for t in ts]}  tokens {bench.count_tokens(lib, ctx0)}", flush=True)
        lib.whisper_b200_group_free(g)


if __name__ == "__main__":
    main()
