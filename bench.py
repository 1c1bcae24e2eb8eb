#!/usr/bin/env python
"""bench.py -- batched 30 s-window transcription throughput (audio-seconds per second, RTFx).

Default workload (BASELINE.json configs[2], the configuration the headline metric is quoted on): whisper large-v3 geometry
(128 mel, 32+32 layers, d=1280), random-init weights (seeded, modelgen.py), 64 x 30 s windows of synthetic 16 kHz PCM IN
TOTAL, sharded over the GPUs in contiguous blocks (STRONG scaling: 64 / 32 / 16 / 8 windows per GPU at 1 / 2 / 4 / 8 GPUs),
greedy decoding without temperature fallback (whisper-cli -bs 1 -bo 1 -nf), one 30 s window per chunk.
One "step" = one pass of the whole hot path over the batch: log-mel -> encoder -> cross K/V -> greedy decode loop with the
logit rules on device -> segments on the host.

  value : PCM already resident in HBM (whisper_b200_full_device), CUDA-event timed, max over ranks.
  e2e   : the reference's own API call (whisper_full_parallel, include/whisper.h) with HOST buffers: H2D of the PCM
          from pinned memory and the D2H of the results are inside the timed region.
  weak  : (N > 1 only) the same step with 64 windows PER GPU, reported beside the strong-scaling value.
  parity_check : the first two windows decoded by a second context in the golden numeric mode (flash_attn = false) and
          compared with tokens the UNMODIFIED reference produced for the same model and audio (tests/golden/golden_r2.json).
  --impl reference : the UNMODIFIED reference CPU path (oracle/_ref) through whisper_full on the box's host cores, on a
          bounded sample (two windows, decode loops cut at two lengths and extrapolated linearly to the 220-token window;
          stated in `sample`).

Other BASELINE configurations (run by hand; outputs are kept under profiles/):
  --config turbo-beam5 : configs[3], large-v3-turbo, 1 h = 120 windows in total, "beam search" 5 + timestamps, selection on device.
  --config mel-sweep   : configs[4], log-mel only, 1 h and 10 h of PCM, 80 and 128 bins, GB/s against the HBM peak.

Multi-GPU: one process per GPU (torchrun), windows are independent units -> sharded with no data-path collective.
torch is used for process-group plumbing, pinned/device buffers and events only.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ARCH = "large-v3"
WINDOW = 480000
N_TOKENS_PER_WINDOW = 220            # n_text_ctx/2 - 4: a random-init model never emits EOT
MODEL_DIR = os.environ.get("WHISPER_B200_MODEL_DIR", "/tmp/whisper_b200_models")
METRIC = "audio-sec/sec (RTFx) large-v3 batched"


def shard_windows(n_total, rank, world):
    """Contiguous block partition of window indices (what whisper_full_parallel does with chunks)."""
    per, rem = divmod(n_total, world)
    start = rank * per + min(rank, rem)
    return list(range(start, start + per + (1 if rank < rem else 0)))


def workload_string(arch, n_total, mode="greedy, no fallback, no_timestamps"):
    """The same string in both arms (the driver compares `config` between them)."""
    return f"whisper {arch} random-init, {n_total}x30s windows in total, {mode}"


def read_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except OSError:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def read_ncu_traffic(kernel_class, n_rows):
    """DRAM bytes per launch of a kernel class from the ncu --set full capture recorded in profiles/ncu_traffic.json
    (written by tools/ncu_traffic.py from the raw page of the report), scaled to this launch's row count."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[kernel_class]
        return rec["dram_bytes_per_launch"] * (n_rows / float(rec["rows"]))
    except (OSError, KeyError, ValueError, ZeroDivisionError):
        return None


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region, through NVML in-process (no fork: spawning nvidia-smi
    from a process that holds a CUDA context and GBs of pinned memory stalls the very run it is observing)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.reasons = index, [], False, set()
        self.sm_max = None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            names = {
                getattr(pynvml, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(pynvml, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(pynvml, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(pynvml, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            }
            while not self.stop_flag:
                self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                try:
                    mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for bit, n in names.items():
                        if mask & bit:
                            self.reasons.add(n)
                except Exception:
                    pass
                time.sleep(0.1)
        except Exception as ex:      # the clocks line is evidence, never a reason to fail the bench
            self.reasons.add(f"nvml_unavailable:{type(ex).__name__}")

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons)}


def ensure_model(arch, rank, barrier, with_tensors=True):
    from open_whisper_kit_b200 import modelgen
    os.makedirs(MODEL_DIR, exist_ok=True)
    path = os.path.join(MODEL_DIR, f"{arch}-f16-seed1234{'' if with_tensors else '-header'}.bin")
    if rank == 0 and not os.path.exists(path):
        tmp = path + ".tmp"
        modelgen.write_model(tmp, arch, seed=1234, ftype=1, with_tensors=with_tensors)
        os.replace(tmp, path)
    barrier()
    return path


def greedy_params(lib, no_timestamps=True, n_threads=1):
    from open_whisper_kit_b200 import capi
    p = lib.whisper_full_default_params(capi.GREEDY)
    p.greedy.best_of = 1
    p.temperature_inc = 0.0
    p.no_timestamps = no_timestamps
    p.print_progress = False
    p.n_threads = n_threads
    p.language = b"en"
    return p


def beam_params(lib, beam_size=5, n_threads=1):
    """whisper-cli defaults (-bs 5 -bo 5, timestamps on) with the temperature fallback off (-nf)."""
    from open_whisper_kit_b200 import capi
    p = lib.whisper_full_default_params(capi.BEAM_SEARCH)
    p.beam_search.beam_size = beam_size
    p.greedy.best_of = beam_size
    p.temperature_inc = 0.0
    p.no_timestamps = False
    p.print_progress = False
    p.n_threads = n_threads
    p.language = b"en"
    return p


def count_tokens(lib, ctx):
    return sum(lib.whisper_full_n_tokens(ctx, i) for i in range(lib.whisper_full_n_segments(ctx)))


def segment_tokens(lib, ctx):
    return [[lib.whisper_full_get_token_id(ctx, i, j) for j in range(lib.whisper_full_n_tokens(ctx, i))]
            for i in range(lib.whisper_full_n_segments(ctx))]


# ---------------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path, timed on this box's host cores."""
    if rank != 0:
        return
    from open_whisper_kit_b200 import api, modelgen
    from oracle import reflib
    ref, variant = reflib.load()
    if ref is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref was not built (no /root/reference on this box)"}))
        return
    cores = os.cpu_count() or 1
    n_threads = min(cores, 32)
    arch = "large-v3-turbo" if args.config == "turbo-beam5" else args.arch
    path = ensure_model(arch, 0, lambda: None)
    w = api.Whisper(ref, path, use_gpu=False, flash_attn=True)
    beam = args.config == "turbo-beam5"
    n_total = 120 if beam else args.windows
    pcm = [modelgen.synth_pcm(WINDOW, seed=7, stream=i) for i in range(2)]
    n_tok = (4, 36)

    def one(window, max_tokens):
        p = beam_params(ref, n_threads=n_threads) if beam else greedy_params(ref, n_threads=n_threads)
        p.max_tokens = max_tokens
        t = time.perf_counter()
        rc, _ = w.full(p, pcm[window])
        assert rc == 0
        return time.perf_counter() - t

    one(0, n_tok[0])       # untimed: pages the model in and spins the thread pool up, whatever --warmup says
    steps = []
    for i in range(args.warmup + args.steps):
        # two DIFFERENT windows per step, one cut after 4 and one after 36 decode calls: the fixed part (mel + encoder + prompt)
        # is the same work for every window, the per-token cost comes from the difference
        ta, tb = one(0, n_tok[0]), one(1, n_tok[1])
        # max_tokens = m ends the window after m decode calls beyond the prompt pass (src/whisper.cpp:7402-7404);
        # a full window of a model that never emits EOT runs N_TOKENS_PER_WINDOW - 1 of them (7219, 7436-7460)
        per_tok = max(1e-9, (tb - ta) / (n_tok[1] - n_tok[0]))
        fixed = max(0.0, ta - per_tok * n_tok[0])
        full = fixed + per_tok * (N_TOKENS_PER_WINDOW - 1)
        if i >= args.warmup:
            steps.append((full, fixed, per_tok, ta + tb))
    full = float(np.mean([s[0] for s in steps]))
    rtfx = 30.0 / full
    what = "beam 5 + timestamps (time to the 220-step budget; the random model's windows end earlier)" if beam else "greedy"
    sample = (f"2 of {n_total} windows through whisper_full, {what}: mel+encode+prompt measured on both, decode loop cut at "
              f"{n_tok[0]} (window 0) and {n_tok[1]} (window 1) steps via max_tokens and extrapolated linearly to "
              f"{N_TOKENS_PER_WINDOW} tokens; {n_threads} threads, build {variant}, flash_attn on (the cli default); "
              f"encode+mel {np.mean([s[1] for s in steps]):.2f} s, {np.mean([s[2] for s in steps]) * 1e3:.1f} ms/step")
    mode = "beam 5 + timestamps, no fallback" if beam else "greedy, no fallback, no_timestamps"
    line = {
        "impl": "reference", "metric": METRIC if not beam else "audio-sec/sec (RTFx) large-v3-turbo beam 5", "value": rtfx,
        "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": full * 1e3 * n_total,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"workload": workload_string(arch, n_total, mode), "global_windows": n_total,
                   "flush": "inputs larger than L2"},
        "cpu_baseline": {"value": rtfx, "unit": "audio-s/s", "cores": n_threads, "kind": "reference", "sample": sample},
        "e2e": {"value": rtfx, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def cpu_baseline_subprocess(args, extra=()):
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0",
                              "--arch", args.arch, "--windows", str(args.windows), "--config", args.config, *extra],
                             capture_output=True, text=True, timeout=1200)
        last = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
        if last:
            return json.loads(last[-1]).get("cpu_baseline")
        return {"value": None, "unit": "audio-s/s", "cores": 0, "kind": "reference", "sample": "failed: " + out.stderr[-200:]}
    except Exception as ex:       # the baseline is reported, never required
        return {"value": None, "unit": "audio-s/s", "cores": 0, "kind": "reference", "sample": f"failed: {ex}"}


# ---------------------------------------------------------------------------------------------------------------
def parity_check(lib, model_path, device, cross_kv="f16"):
    """Decode the first two windows in the golden numeric mode and compare with the reference-generated fixture
    (tests/golden/golden_r2.json, case large-v3/synth2/nots48; rule of tests/test_gpu_parity_r2.py)."""
    from open_whisper_kit_b200 import modelgen
    try:
        gold = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_r2.json")))["large-v3/synth2/nots48"]
    except (OSError, KeyError):
        return {"ok": None, "why": "tests/golden/golden_r2.json has no large-v3 case"}
    cp = lib.whisper_context_default_params()
    cp.gpu_device = device
    cp.flash_attn = False
    if cross_kv == "fp8":
        os.environ["WHISPER_B200_CROSS_KV"] = "fp8"
    ctx = lib.whisper_init_from_file_with_params(model_path.encode(), cp)
    os.environ.pop("WHISPER_B200_CROSS_KV", None)
    if not ctx:
        return {"ok": False, "why": "second context failed to load"}
    try:
        p = greedy_params(lib, no_timestamps=True)
        p.max_tokens = gold["max_tokens"]
        pcm = np.concatenate([modelgen.synth_pcm(WINDOW, seed=gold["seed"], stream=i) for i in range(gold["windows"])])
        rc = lib.whisper_full_parallel(ctx, p, pcm.ctypes.data_as(C.POINTER(C.c_float)), len(pcm), gold["windows"])
        if rc != 0:
            return {"ok": False, "why": f"rc {rc}"}
        ours = segment_tokens(lib, ctx)
    finally:
        lib.whisper_free(ctx)
    margin, res, ok = 5e-3, [], len(ours) == len(gold["segments"])
    for wi, ref in enumerate(gold["segments"]):
        if wi >= len(ours):
            break
        gaps, runner = gold["steps"][wi]["gaps"], gold["steps"][wi]["runner_up"]
        strict = next((k for k, g in enumerate(gaps) if g < margin), len(gaps))
        k = next((i for i, (x, y) in enumerate(zip(ours[wi], ref[2])) if x != y), None)
        if k is None and len(ours[wi]) != len(ref[2]):
            k = min(len(ours[wi]), len(ref[2]))
        good = k is None or (k >= strict and k < len(gaps) and gaps[k] < margin and k < len(ours[wi]) and ours[wi][k] == runner[k])
        ok = ok and good
        res.append({"reference_tokens": len(ref[2]), "strict_prefix": strict, "identical_until": len(ref[2]) if k is None else k})
    return {"ok": bool(ok), "against": "tests/golden/golden_r2.json large-v3/synth2/nots48 (unmodified reference, AVX-512 build)",
            "rule": "identical before the reference's first top-2 margin < 5e-3; a later first mismatch must be the reference's runner-up "
                    "on such a step", "windows": res}


# ---------------------------------------------------------------------------------------------------------------
def init_dist():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    def timed(fn, k):
        """k calls of fn bracketed by barrier + synchronize on both sides, CUDA events, max over ranks -> ms."""
        barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    return rank, world, local_rank, barrier, timed


class Workload:
    """This rank's windows of the global synthetic stream: pinned host PCM + a device copy."""

    def __init__(self, windows):
        import torch
        from open_whisper_kit_b200 import modelgen
        self.n_win = len(windows)
        self.n_samples = self.n_win * WINDOW
        self.host = torch.empty(max(1, self.n_samples), dtype=torch.float32).pin_memory()
        for i, wi in enumerate(windows):
            self.host[i * WINDOW:(i + 1) * WINDOW] = torch.from_numpy(modelgen.synth_pcm(WINDOW, seed=7, stream=wi))
        self.dev = self.host.cuda(non_blocking=False)
        self.host_ptr = C.cast(self.host.data_ptr(), C.POINTER(C.c_float))
        self.dev_ptr = C.c_void_p(self.dev.data_ptr())


def main_transcribe(args):
    import torch.distributed as dist
    import open_whisper_kit_b200 as pkg
    rank, world, local_rank, barrier, timed = init_dist()
    beam = args.config == "turbo-beam5"
    arch = "large-v3-turbo" if beam else args.arch
    n_total = 120 if beam else args.windows
    lib = pkg.load()
    path = ensure_model(arch, rank, barrier)
    cp = lib.whisper_context_default_params()
    cp.gpu_device = local_rank
    t0 = time.time()
    if args.cross_kv == "fp8":
        os.environ["WHISPER_B200_CROSS_KV"] = "fp8"        # read when a context is created
    else:
        os.environ.pop("WHISPER_B200_CROSS_KV", None)
    ctx = lib.whisper_init_from_file_with_params(path.encode(), cp)
    os.environ.pop("WHISPER_B200_CROSS_KV", None)
    if not ctx:
        raise SystemExit("model load failed")
    load_s = time.time() - t0

    strong = args.scaling == "strong"
    wl = Workload(shard_windows(n_total, rank, world) if strong else shard_windows(n_total * world, rank, world))
    params = beam_params(lib) if beam else greedy_params(lib, no_timestamps=not args.timestamps)

    def step_device(w=wl, p=params):
        if w.n_win == 0:
            return
        rc = lib.whisper_b200_full_device(ctx, p, w.dev_ptr, w.n_samples, w.n_win)
        assert rc == 0, rc

    def step_e2e(w=wl, p=params):
        if w.n_win == 0:
            return 0
        rc = lib.whisper_full_parallel(ctx, p, w.host_ptr, w.n_samples, w.n_win)
        assert rc == 0, rc
        # read the result back like a caller does: every segment's token ids (D2H already happened inside the call)
        return count_tokens(lib, ctx)

    sampler = ClockSampler(local_rank)
    for _ in range(args.warmup):
        step_device()
    launches0 = lib.whisper_b200_kernel_launches(ctx)
    sampler.start()
    ms_dev = timed(step_device, args.steps)
    sampler.stop_flag = True
    sampler.join(timeout=2.0)
    launches = lib.whisper_b200_kernel_launches(ctx) - launches0
    n_tokens = count_tokens(lib, ctx) if wl.n_win else 0
    step_e2e()
    e2e_steps = max(1, args.steps)
    ms_e2e = timed(step_e2e, e2e_steps)

    n_global = n_total if strong else n_total * world
    audio_s = 30.0 * n_global
    value = audio_s / (ms_dev * 1e-3 / args.steps)
    e2e_value = audio_s / (ms_e2e * 1e-3 / e2e_steps)

    # one extra, untimed, instrumented step: per-kernel-class CUDA-event times for the roofline numbers
    roof, stages = None, {}
    if wl.n_win:
        roof, stages = kernel_profile(lib, pkg, ctx, step_device, ms_dev / args.steps, wl.n_win)

    # second line of the scaling picture: 64 windows PER GPU (weak), same step, same timing rules
    weak = None
    if world > 1 and strong and not args.no_weak:
        wl_w = Workload(shard_windows(n_total * world, rank, world))
        for _ in range(2):
            step_device(wl_w)
        k = max(1, min(args.steps, 5))
        ms_w = timed(lambda: step_device(wl_w), k)
        weak = {"value": 30.0 * n_total * world / (ms_w * 1e-3 / k), "unit": "audio-s/s", "ms_per_step": ms_w / k,
                "windows_per_gpu": n_total, "steps": k}

    greedy_ts = None
    if beam:
        # the same windows with greedy + timestamps: what the sampled ("beam") selection costs on top of the decode itself
        pg = greedy_params(lib, no_timestamps=False)
        step_device(p=pg)
        ms_g = timed(lambda: step_device(p=pg), max(1, args.steps))
        greedy_ts = {"ms_per_step": ms_g / max(1, args.steps), "ratio_beam_over_greedy": (ms_dev / args.steps) / (ms_g / max(1, args.steps))}

    # secondary line of the default single-GPU run: the same step with the opt-in e4m3 cross-K/V pool (reduced storage precision,
    # tests/test_gpu_cross_fp8.py is its parity study) -- what halving the bytes of the roofline kernel buys
    fp8_line = None
    if world == 1 and not beam and args.cross_kv == "f16" and not args.no_fp8_line and wl.n_win:
        try:
            os.environ["WHISPER_B200_CROSS_KV"] = "fp8"
            ctx8 = lib.whisper_init_from_file_with_params(path.encode(), cp)
            os.environ.pop("WHISPER_B200_CROSS_KV", None)
            if ctx8:
                def step8():
                    rc8 = lib.whisper_b200_full_device(ctx8, params, wl.dev_ptr, wl.n_samples, wl.n_win)
                    assert rc8 == 0, rc8
                for _ in range(2):
                    step8()
                ms8 = timed(step8, max(1, args.steps)) / max(1, args.steps)
                t8, t16 = segment_tokens(lib, ctx8), segment_tokens(lib, ctx)
                fp8_line = {"value": audio_s / (ms8 * 1e-3), "unit": "audio-s/s", "ms_per_step": ms8,
                            "segments_identical_to_f16_pool_run": f"{sum(a == b for a, b in zip(t8, t16))} of {len(t16)} "
                                                                  "(220 tokens each; a near-tie flip changes the rest of a window)",
                            "note": "WHISPER_B200_CROSS_KV=fp8: cross K/V stored as e4m3 chunks with per-chunk scales; NOT the reference's "
                                    "F16 cache, off by default; parity study: tests/test_gpu_cross_fp8.py"}
                lib.whisper_free(ctx8)
        except Exception as ex:            # never lets the opt-in line cost the headline
            os.environ.pop("WHISPER_B200_CROSS_KV", None)
            fp8_line = {"value": None, "note": f"failed: {ex}"}

    check = None
    cpu_baseline = None
    if rank == 0 and not beam:
        check = parity_check(lib, path, local_rank, args.cross_kv)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_baseline_subprocess(args)

    if rank == 0:
        mode = "beam 5 + timestamps, no fallback" if beam else \
            f"greedy, no fallback, {'timestamps' if args.timestamps else 'no_timestamps'}"
        line = {
            "metric": METRIC if not beam else "audio-sec/sec (RTFx) large-v3-turbo beam 5", "value": value, "unit": "audio-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f16" if lib.whisper_b200_dtype(ctx) == 0 else "bf16", "data": "synthetic",
            "config": {"workload": workload_string(arch, n_global, mode), "global_windows": n_global,
                       "windows_per_gpu_rank0": wl.n_win, "tokens_decoded_per_step_rank0": n_tokens,
                       "parallelism": f"window-sharded x{world} in contiguous blocks, no collective",
                       "flush": "inputs larger than L2 (weights + cross-K/V stream per decode step >> 126 MB)"},
            "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": wl.n_samples * 4,
                    "d2h_bytes_per_step": n_tokens * 24 + wl.n_win * 4, "ms_per_step": ms_e2e / e2e_steps},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": roof,
            "stages": stages,
            "cpu_baseline": cpu_baseline,
            "model_load_s": round(load_s, 1),
        }
        if args.cross_kv == "fp8":
            line["config"]["cross_kv"] = "e4m3 chunks, per-chunk f32 scale (opt-in; the reference's cache is F16)"
            line["config"]["workload"] += ", cross K/V e4m3 (opt-in)"
        if fp8_line is not None:
            line["opt_in_fp8_cross_kv"] = fp8_line
        if weak is not None:
            line["weak"] = weak
        if greedy_ts is not None:
            line["greedy_timestamps_same_windows"] = greedy_ts
        if check is not None:
            line["parity_check"] = check
        if world > 1:
            line["scaling_limiter"] = ("per decode step a fixed ~1.5 ms chain of dependent small kernels (6 tc_skinny_kernel GEMMs + "
                                       "self-attention per layer, ~5 us each incl. the programmatic hand-over) does not shrink with the "
                                       "batch; only the cross-attention K/V stream does")
        print(json.dumps(line))
    lib.whisper_free(ctx)
    if world > 1:
        dist.destroy_process_group()


def kernel_profile(lib, pkg, ctx, step_device, step_ms, n_win):
    lib.whisper_b200_profile_enable(ctx, 1)
    step_device()
    buf = (C.c_double * (3 * 32))()
    n_cls = lib.whisper_b200_profile_read(ctx, buf, 3 * 32)
    lib.whisper_b200_profile_enable(ctx, 0)
    prof = {}
    for i in range(n_cls):
        name, unit = pkg.PROFILE_CLASSES[i]
        ms, n, work = buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]
        if n > 0:
            prof[name] = {"ms": ms, "launches": int(n), "work": work, "unit": unit}
    peaks, peak_kind = read_peaks()
    # The events bracket every launch, which (a) adds the dependent-launch gap to each kernel and (b) disables the
    # programmatic-dependent-launch overlap of the real run.  The mean gap is what the bracketed times add up to beyond the
    # un-instrumented step, per launch; kernel-time estimates (ms_kernel) subtract it.  They agree with ncu's per-launch
    # durations (profiles/*launches*.summary.txt).
    n_prof_launches = sum(v["launches"] for v in prof.values()) or 1
    gap_ms = max(0.0, (sum(v["ms"] for v in prof.values()) - step_ms) / n_prof_launches)
    for v in prof.values():
        v["ms_kernel"] = max(v["ms"] - gap_ms * v["launches"], 0.25 * v["ms"])
    total_prof_ms = sum(v["ms_kernel"] for v in prof.values()) or 1.0
    dom = max(prof, key=lambda k: prof[k]["ms_kernel"])
    dv = prof[dom]
    if dv["unit"] == "B":
        achieved = dv["work"] / (dv["ms_kernel"] * 1e-3) / 1e9
        roof = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": read_ncu_traffic(dom, n_win),
                "traffic_source": "profiles/ncu_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, scaled by rows)",
                "achieved_incl_launch_gap": dv["work"] / (dv["ms"] * 1e-3) / 1e9}
    else:
        achieved = dv["work"] / (dv["ms_kernel"] * 1e-3) / 1e12
        peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
        roof = {"kernel": dom, "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": read_ncu_traffic(dom, n_win),
                "achieved_incl_launch_gap": dv["work"] / (dv["ms"] * 1e-3) / 1e12}
    if dom == "cross_attention" and n_win >= 32 and os.environ.get("WHISPER_B200_CROSS_PF_CHUNKS", "4") != "0":
        # The K prefix of every (window, head) block (4 x 16 KB) is requested into L2 by the six GEMMs that run before the launch,
        # while HBM is ~85 % idle (csrc/tc_skinny.cu, engine.cu): those bytes are part of the algorithmic stream but reach the SMs
        # from L2, which is how `achieved` can exceed the measured HBM (copy) peak.  `frac_hbm_lower_bound` charges only the
        # remaining bytes to HBM (L2 keeps about half of the requested lines until use, so the truth lies in between).
        pf = float(n_win * lib.whisper_model_n_text_head(ctx) * 4 * 16384) * dv["launches"]
        roof["l2_prefetched_bytes_per_launch"] = pf / dv["launches"]
        roof["frac_hbm_lower_bound"] = (dv["work"] - pf) / (dv["ms_kernel"] * 1e-3) / 1e9 / peaks["hbm_gbs"]
        roof["note"] = ("cross-attention K/V stream; frac > 1 because part of the stream is prefetched into L2 during the preceding "
                        "latency-bound GEMM chain (ncu's serialised, cache-flushed capture shows traffic = algorithmic bytes)")
    roof["peak_source"] = peak_kind
    roof["share_of_step"] = dv["ms_kernel"] / total_prof_ms
    roof["algorithmic_per_launch"] = dv["work"] / dv["launches"]
    roof["us_per_launch"] = 1e3 * dv["ms_kernel"] / dv["launches"]
    roof["launch_gap_us"] = 1e3 * gap_ms
    stages = {}
    for k, v in prof.items():
        rate = v["work"] / (v["ms"] * 1e-3)
        stages[k] = {"ms": round(v["ms"], 3), "ms_kernel": round(v["ms_kernel"], 3), "launches": v["launches"],
                     ("GB/s" if v["unit"] == "B" else "TFLOP/s"): round(rate / (1e9 if v["unit"] == "B" else 1e12), 1)}
    if "mel" in stages:
        stages["mel"]["frac_of_hbm_peak"] = round(stages["mel"]["GB/s"] / peaks["hbm_gbs"], 3)
    enc_flop = sum(prof[k]["work"] for k in ("gemm_conv", "gemm_encoder", "encoder_attention", "gemm_cross_kv") if k in prof)
    enc_ms = sum(prof[k]["ms"] for k in ("gemm_conv", "gemm_encoder", "encoder_attention", "gemm_cross_kv", "im2col", "layernorm")
                 if k in prof)
    if enc_ms > 0:
        tf = enc_flop / (enc_ms * 1e-3) / 1e12
        stages["encoder_total"] = {"TFLOP/s": round(tf, 1),
                                   "tensor_util_of_sustained_peak": round(tf / peaks.get("bf16_tflops_sustained", 1399.0), 3)}
    dec_ms = sum(prof[k]["ms_kernel"] for k in ("gemm_decoder", "layernorm_decoder", "self_attention") if k in prof)
    n_steps = prof.get("gemm_logits", {}).get("launches", 0)
    if n_steps:
        stages["decoder_latency_chain_ms_per_step"] = round(dec_ms / n_steps, 3)
    return roof, stages


# ---------------------------------------------------------------------------------------------------------------
def main_mel_sweep(args):
    """BASELINE configs[4]: log-mel only.  hours x bins; the PCM of one measurement is split over the ranks at window
    boundaries and every rank runs ONE whisper_pcm_to_mel-sized call over its share (the global-max clamp is per call, i.e.
    per rank: state it).  Algorithmic bytes: 64 000 B PCM + 100 * n_mel * 4 B mel per audio-second."""
    import torch
    import torch.distributed as dist
    import open_whisper_kit_b200 as pkg
    from open_whisper_kit_b200 import modelgen
    rank, world, local_rank, barrier, timed = init_dist()
    lib = pkg.load()
    peaks, peak_kind = read_peaks()
    rows = []
    FP = C.POINTER(C.c_float)
    for n_mel, arch in ((80, "tiny.en"), (128, "large-v3")):
        path = ensure_model(arch, rank, barrier, with_tensors=False)       # header only: filters + vocabulary
        cp = lib.whisper_context_default_params()
        cp.gpu_device = local_rank
        ctx = lib.whisper_init_from_file_with_params(path.encode(), cp)
        assert ctx
        filt = modelgen.mel_filters(n_mel)
        for hours in (1, 10):
            n_win_total = hours * 120
            mine = shard_windows(n_win_total, rank, world)
            n_samples = len(mine) * WINDOW
            # kernel only, device-resident PCM, L2 flushed between launches
            k_ms = lib.whisper_b200_kernel_log_mel_bench(1, n_samples, filt.ctypes.data_as(FP), n_mel, max(3, args.steps), 1)
            # through the reference-facing call with pinned host PCM (H2D inside)
            host = torch.empty(n_samples, dtype=torch.float32).pin_memory()
            one = torch.from_numpy(modelgen.synth_pcm(WINDOW, seed=7, stream=rank))
            for i in range(len(mine)):
                host[i * WINDOW:(i + 1) * WINDOW] = one
            hp = C.cast(host.data_ptr(), FP)

            def call():
                assert lib.whisper_pcm_to_mel(ctx, hp, n_samples, 1) == 0

            for _ in range(2):
                call()
            e_ms = timed(call, max(1, args.steps)) / max(1, args.steps)
            if world > 1:
                t = torch.tensor([k_ms], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                k_ms = float(t.item())
            audio_s = 3600.0 * hours
            algo = audio_s * (64000 + 100 * n_mel * 4)
            rows.append({"hours": hours, "n_mel": n_mel, "kernel_ms": round(k_ms, 4), "kernel_GBs": round(algo / (k_ms * 1e-3) / 1e9, 1),
                         # aggregate GB/s over all ranks against the aggregate peak (= the per-GPU fraction of the slowest rank)
                         "kernel_frac_of_hbm_peak": round(algo / (k_ms * 1e-3) / 1e9 / (peaks["hbm_gbs"] * world), 3),
                         "e2e_ms": round(e_ms, 3), "e2e_GBs": round(algo / (e_ms * 1e-3) / 1e9, 1),
                         "h2d_bytes": n_samples * 4 * world, "audio_s_per_s": round(audio_s / (e_ms * 1e-3), 0)})
            del host
        lib.whisper_free(ctx)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = mel_cpu_baseline()
    if rank == 0:
        best = max(rows, key=lambda r: r["kernel_GBs"])
        print(json.dumps({
            "metric": "log-mel GB/s (algorithmic bytes / kernel time)", "value": best["kernel_GBs"], "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": 2, "higher_is_better": True, "scaling": "strong", "dtype": "f32", "data": "synthetic",
            "config": {"workload": "log-mel STFT sweep 1 h and 10 h of 16 kHz PCM, 80 and 128 bins, split over the ranks at window "
                                   "boundaries, one call per rank (global-max clamp per call)", "flush": "L2 flushed between launches"},
            "peak_hbm_GBs": peaks["hbm_gbs"], "peak_source": peak_kind, "rows": rows, "cpu_baseline": cpu}))
    if world > 1:
        dist.destroy_process_group()


def mel_cpu_baseline():
    """whisper_pcm_to_mel of the unmodified reference (src/whisper.cpp:3875-3886) on the host cores, 10 minutes of audio."""
    try:
        from open_whisper_kit_b200 import api, modelgen
        from oracle import reflib
        ref, variant = reflib.load()
        if ref is None:
            return None
        cores = min(os.cpu_count() or 1, 32)
        out = {"kind": "reference", "cores": cores, "unit": "GB/s", "sample": f"10 min of audio per call, {cores} threads, build {variant}"}
        pcm = np.concatenate([modelgen.synth_pcm(WINDOW, seed=7, stream=0)] * 20)
        for n_mel, arch in ((80, "tiny.en"), (128, "large-v3")):
            w = api.Whisper(ref, ensure_model(arch, 0, lambda: None, with_tensors=False), use_gpu=False)
            w.pcm_to_mel(pcm[:WINDOW], cores)
            t = time.perf_counter()
            assert w.pcm_to_mel(pcm, cores) == 0
            dt = time.perf_counter() - t
            out[f"GBs_{n_mel}"] = round(600.0 * (64000 + 100 * n_mel * 4) / dt / 1e9, 4)
            w.close()
        out["value"] = out["GBs_128"]
        return out
    except Exception as ex:
        return {"kind": "reference", "value": None, "sample": f"failed: {ex}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--arch", default=ARCH)
    ap.add_argument("--config", default="large-v3-64", choices=["large-v3-64", "turbo-beam5", "mel-sweep"])
    ap.add_argument("--windows", type=int, default=64, help="30 s windows in total (strong) / per GPU (weak)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--no-weak", action="store_true", help="skip the second (weak-scaling) measurement at N > 1")
    ap.add_argument("--timestamps", action="store_true", help="decode with timestamp tokens (variable work)")
    ap.add_argument("--cross-kv", default="f16", choices=["f16", "fp8"],
                    help="fp8: the whole run with the opt-in e4m3 cross-K/V pool (reduced precision; not the headline configuration)")
    ap.add_argument("--no-fp8-line", action="store_true", help="skip the secondary e4m3 cross-K/V measurement of the default run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))
    elif args.config == "mel-sweep":
        main_mel_sweep(args)
    else:
        main_transcribe(args)


if __name__ == "__main__":
    main()
