#!/usr/bin/env python
"""bench.py -- batched 30 s-window transcription throughput (audio-seconds per second, RTFx).

Workload (BASELINE.json configs[2], the configuration the headline metric is quoted on): whisper large-v3 geometry
(128 mel, 32+32 layers, d=1280), random-init weights (seeded, modelgen.py), 64 x 30 s windows of synthetic 16 kHz PCM
per GPU, greedy decoding without temperature fallback (whisper-cli -bs 1 -bo 1 -nf), one 30 s window per chunk.
One "step" = one pass of the whole hot path over the batch: log-mel -> encoder -> cross K/V -> greedy decode loop
with the logit rules on device -> segments on the host.

  value : PCM already resident in HBM (whisper_b200_full_device), CUDA-event timed, max over ranks.
  e2e   : the reference's own API call (whisper_full_parallel, include/whisper.h) with HOST buffers: H2D of the PCM
          from pinned memory and the D2H of the results are inside the timed region.
  --impl reference : the UNMODIFIED reference CPU path (oracle/_ref) through whisper_full on the box's host cores,
          on a bounded sample (one window, decode loop cut at two lengths and extrapolated linearly to the
          220-token window; stated in `sample`).

Multi-GPU: one process per GPU (torchrun), windows are independent units -> sharded with no data-path collective;
weak scaling (64 windows per GPU).  torch is used for process-group plumbing, pinned/device buffers and events only.
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


def shard_windows(n_total, rank, world):
    """Contiguous block partition of window indices (what whisper_full_parallel does with chunks)."""
    per, rem = divmod(n_total, world)
    start = rank * per + min(rank, rem)
    return list(range(start, start + per + (1 if rank < rem else 0)))


def read_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except OSError:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


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


def ensure_model(arch, rank, barrier):
    from open_whisper_kit_b200 import modelgen
    os.makedirs(MODEL_DIR, exist_ok=True)
    path = os.path.join(MODEL_DIR, f"{arch}-f16-seed1234.bin")
    if rank == 0 and not os.path.exists(path):
        tmp = path + ".tmp"
        modelgen.write_model(tmp, arch, seed=1234, ftype=1)
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


def count_tokens(lib, ctx):
    return sum(lib.whisper_full_n_tokens(ctx, i) for i in range(lib.whisper_full_n_segments(ctx)))


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
    path = ensure_model(args.arch, 0, lambda: None)
    w = api.Whisper(ref, path, use_gpu=False, flash_attn=True)
    pcm = modelgen.synth_pcm(WINDOW, seed=7, stream=0)
    n_tok = (4, 36)

    def one(max_tokens):
        p = greedy_params(ref, n_threads=n_threads)
        p.max_tokens = max_tokens
        t = time.perf_counter()
        rc, _ = w.full(p, pcm)
        assert rc == 0
        return time.perf_counter() - t

    one(n_tok[0])       # untimed: pages the model in and spins the thread pool up, whatever --warmup says
    steps = []
    for i in range(args.warmup + args.steps):
        ta, tb = one(n_tok[0]), one(n_tok[1])
        # max_tokens = m ends the window after m decode calls beyond the prompt pass (src/whisper.cpp:7402-7404);
        # a full window of a model that never emits EOT runs N_TOKENS_PER_WINDOW - 1 of them (7219, 7436-7460)
        per_tok = max(1e-9, (tb - ta) / (n_tok[1] - n_tok[0]))
        fixed = max(0.0, ta - per_tok * n_tok[0])
        full = fixed + per_tok * (N_TOKENS_PER_WINDOW - 1)
        if i >= args.warmup:
            steps.append((full, fixed, per_tok, ta + tb))
    full = float(np.mean([s[0] for s in steps]))
    rtfx = 30.0 / full
    sample = (f"1 of {args.windows} windows through whisper_full (mel+encode+prompt measured, decode loop cut at "
              f"{n_tok[0]} and {n_tok[1]} tokens via max_tokens and extrapolated linearly to {N_TOKENS_PER_WINDOW} tokens); "
              f"{n_threads} threads, build {variant}; encode+mel {np.mean([s[1] for s in steps]):.2f} s, "
              f"{np.mean([s[2] for s in steps]) * 1e3:.1f} ms/token")
    line = {
        "impl": "reference", "metric": "audio-sec/sec (RTFx) large-v3 batched", "value": rtfx, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": full * 1e3 * args.windows,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"workload": f"whisper {args.arch} random-init, {args.windows}x30s windows/GPU, greedy, no fallback",
                   "windows_per_gpu": args.windows, "flush": "inputs larger than L2"},
        "cpu_baseline": {"value": rtfx, "unit": "audio-s/s", "cores": n_threads, "kind": "reference", "sample": sample},
        "e2e": {"value": rtfx, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--arch", default=ARCH)
    ap.add_argument("--windows", type=int, default=64, help="30 s windows per GPU")
    ap.add_argument("--timestamps", action="store_true", help="decode with timestamp tokens (variable work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import open_whisper_kit_b200 as pkg
    from open_whisper_kit_b200 import modelgen

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()

    lib = pkg.load()
    path = ensure_model(args.arch, rank, barrier)
    cp = lib.whisper_context_default_params()
    cp.gpu_device = local_rank
    t0 = time.time()
    ctx = lib.whisper_init_from_file_with_params(path.encode(), cp)
    if not ctx:
        raise SystemExit("model load failed")
    load_s = time.time() - t0

    # this rank's windows of the global synthetic stream (window index = global index -> distinct audio per rank)
    my_windows = shard_windows(args.windows * world, rank, world)
    pcm_host = torch.empty(len(my_windows) * WINDOW, dtype=torch.float32).pin_memory()
    for i, wi in enumerate(my_windows):
        pcm_host[i * WINDOW:(i + 1) * WINDOW] = torch.from_numpy(modelgen.synth_pcm(WINDOW, seed=7, stream=wi))
    pcm_dev = pcm_host.cuda(non_blocking=False)
    n_win = len(my_windows)
    n_samples = n_win * WINDOW
    params = greedy_params(lib, no_timestamps=not args.timestamps)
    FP = C.POINTER(C.c_float)
    host_ptr = C.cast(pcm_host.data_ptr(), FP)

    def step_device():
        rc = lib.whisper_b200_full_device(ctx, params, C.c_void_p(pcm_dev.data_ptr()), n_samples, n_win)
        assert rc == 0, rc

    def step_e2e():
        rc = lib.whisper_full_parallel(ctx, params, host_ptr, n_samples, n_win)
        assert rc == 0, rc
        # read the result back like a caller does: every segment's token ids (D2H already happened inside the call)
        return count_tokens(lib, ctx)

    def timed(fn, k):
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

    sampler = ClockSampler(local_rank)
    for _ in range(args.warmup):
        step_device()
    launches0 = lib.whisper_b200_kernel_launches(ctx)
    sampler.start()
    ms_dev = timed(step_device, args.steps)
    sampler.stop_flag = True
    sampler.join(timeout=2.0)
    launches = lib.whisper_b200_kernel_launches(ctx) - launches0
    n_tokens = count_tokens(lib, ctx)
    step_e2e()
    ms_e2e = timed(step_e2e, max(1, args.steps))
    e2e_steps = max(1, args.steps)

    # one extra, untimed, instrumented step: per-kernel-class CUDA-event times for the roofline numbers
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
    # durations (profiles/r1_launches_*.summary.txt): cross-attention 81.7 us here vs 82.1 us under ncu.
    n_prof_launches = sum(v["launches"] for v in prof.values()) or 1
    step_ms = ms_dev / args.steps
    gap_ms = max(0.0, (sum(v["ms"] for v in prof.values()) - step_ms) / n_prof_launches)
    for v in prof.values():
        v["ms_kernel"] = max(v["ms"] - gap_ms * v["launches"], 0.25 * v["ms"])
    total_prof_ms = sum(v["ms_kernel"] for v in prof.values()) or 1.0
    dom = max(prof, key=lambda k: prof[k]["ms_kernel"])
    dv = prof[dom]
    # DRAM bytes per launch of the dominant kernels from one `ncu --set full` capture each (profiles/r1_ncu_full_summary.txt)
    ncu_traffic = {"cross_attention": 496.32e6 * (n_win / 64.0)}
    if dv["unit"] == "B":
        achieved = dv["work"] / (dv["ms_kernel"] * 1e-3) / 1e9
        roof = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": ncu_traffic.get(dom),
                "achieved_incl_launch_gap": dv["work"] / (dv["ms"] * 1e-3) / 1e9}
    else:
        achieved = dv["work"] / (dv["ms_kernel"] * 1e-3) / 1e12
        peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
        roof = {"kernel": dom, "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": None, "achieved_incl_launch_gap": dv["work"] / (dv["ms"] * 1e-3) / 1e12}
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

    audio_s = 30.0 * n_win * world
    value = audio_s / (ms_dev * 1e-3 / args.steps)
    e2e_value = audio_s / (ms_e2e * 1e-3 / e2e_steps)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0",
                                  "--arch", args.arch, "--windows", str(args.windows)], capture_output=True, text=True, timeout=900)
            last = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
            if last:
                cpu_baseline = json.loads(last[-1]).get("cpu_baseline")
        except Exception as ex:       # the baseline is reported, never required
            cpu_baseline = {"value": None, "unit": "audio-s/s", "cores": 0, "kind": "reference", "sample": f"failed: {ex}"}

    if rank == 0:
        line = {
            "metric": "audio-sec/sec (RTFx) large-v3 batched", "value": value, "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f16" if lib.whisper_b200_dtype(ctx) == 0 else "bf16", "data": "synthetic",
            "config": {"workload": f"whisper {args.arch} random-init, {args.windows}x30s windows/GPU, greedy, no fallback, "
                                   f"{'timestamps' if args.timestamps else 'no_timestamps'}",
                       "windows_per_gpu": args.windows, "global_windows": args.windows * world,
                       "tokens_decoded_per_step_rank0": n_tokens, "parallelism": f"window-sharded x{world}, no collective",
                       "flush": "inputs larger than L2 (3.1 GB weights + 15.7 GB cross-K/V stream per decode step)"},
            "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": n_samples * 4,
                    "d2h_bytes_per_step": n_tokens * 24 + n_win * 4, "ms_per_step": ms_e2e / e2e_steps},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": roof,
            "stages": stages,
            "cpu_baseline": cpu_baseline,
            "model_load_s": round(load_s, 1),
        }
        print(json.dumps(line))
    lib.whisper_free(ctx)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
