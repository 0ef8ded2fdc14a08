#!/usr/bin/env python
"""bench.py -- audio-hours/sec of the B200 spectral front-end on BASELINE.json's configs[1].

Workload (N=1 and per rank at N>1, weak scaling): MFCC 40 coeffs + delta + delta-delta on
4096 synthetic 4 s 16 kHz clips (librosa framing n_fft 2048 / hop 512 / 128 mels), one "step" =
one pass of the hot path over that batch.  Prints ONE JSON line (see the task contract):
  value  : device-resident throughput (inputs already in HBM), CUDA events, max over ranks
  e2e    : same metric through Frontend.extract_host (host buffers in, host features out)
  roofline / cpu_baseline : dominant kernel vs FP32 peak (and HBM), oracle on the host cores
`--impl reference` times the CPU oracle port (the reference's librosa path restated, joblib
over all host cores as ASV_dl_func.py:1036 does) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# loky workers are fresh interpreters: make `import oracle` resolve there too
os.environ["PYTHONPATH"] = ROOT + os.pathsep + os.environ.get("PYTHONPATH", "")

import numpy as np

SR = 16000
CLIP_S = 4.0
CLIPS_PER_GPU = 4096
N_FFT, HOP, N_MELS, N_MFCC, N_DELTA = 2048, 512, 128, 40, 2
WORKLOAD = "configs[1]: MFCC-40 + delta + delta-delta, 4096 x 4 s @16 kHz per GPU (n_fft 2048, hop 512, 128 mels)"
METRIC = "audio-hours/sec (log-mel/MFCC/LFCC front-end)"
UNIT = "audio-hours/s"


# --------------------------------------------------------------------------- helpers
def algorithmic_work(n_fft, hop, n_filt, n_ceps, n_delta, nnz_fb, c_out, in_bytes=4):
    """flops and bytes per frame."""
    k = n_fft // 2 + 1
    stft_fb = n_fft + 2.5 * n_fft * math.log2(n_fft) + 3 * k + 2 * nnz_fb + n_filt
    cep = 2 * n_filt * n_ceps + 2 * 9 * n_ceps * n_delta
    return {"flops_stft_fb": stft_fb, "flops_epilogue": cep, "flops": stft_fb + cep,
            "bytes": in_bytes * hop + 4 * c_out}


def synth_clip(seed: int, n: int) -> np.ndarray:
    """Deterministic noise clip  clip(0.1 N(0,1), -1, 1)  (SURVEY.md 8d synthetic inputs)."""
    rng = np.random.default_rng(seed)
    return np.clip(0.1 * rng.standard_normal(n), -1.0, 1.0).astype(np.float32)


def _cpu_clip_job(seed: int, n: int):
    import oracle
    y = synth_clip(seed, n)
    out = oracle.mfcc_with_deltas_ref(y, SR, n_mfcc=N_MFCC, n_fft=N_FFT, hop_length=HOP, n_mels=N_MELS,
                                      n_delta=N_DELTA)
    return out.shape[1]


def cpu_reference_pass(n_clips: int, n_jobs: int, pool=None) -> float:
    """Seconds for one joblib pass of the oracle over n_clips in-memory clips."""
    from joblib import Parallel, delayed
    n = int(CLIP_S * SR)
    t0 = time.perf_counter()
    runner = pool if pool is not None else Parallel(n_jobs=n_jobs)
    runner(delayed(_cpu_clip_job)(1000 + i, n) for i in range(n_clips))
    return time.perf_counter() - t0


def calibrate_cpu_sample(target_cpu_s: float, cores: int) -> int:
    n = int(CLIP_S * SR)
    _cpu_clip_job(99, n)            # imports + first-call set-up stay out of the calibration
    t0 = time.perf_counter()
    for i in range(4):
        _cpu_clip_job(i, n)
    per_clip = (time.perf_counter() - t0) / 4
    k = int(target_cpu_s / max(per_clip, 1e-4))
    k = max(cores * 4, min(CLIPS_PER_GPU, k))
    return (k + cores - 1) // cores * cores


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.001):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def load_measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    n_sample = calibrate_cpu_sample(args.cpu_seconds, cores)
    from joblib import Parallel
    hours = n_sample * CLIP_S / 3600.0
    with Parallel(n_jobs=cores) as pool:
        for _ in range(max(args.warmup, 1)):
            cpu_reference_pass(min(n_sample, cores * 2), cores, pool)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_reference_pass(n_sample, cores, pool)
        dt = time.perf_counter() - t0
    value = hours * args.steps / dt
    sample = (f"{n_sample} of {CLIPS_PER_GPU} clips per step, in memory (no file decode), joblib loky "
              f"n_jobs={cores}; oracle = numpy/scipy restatement of librosa 0.11 (librosa not installed)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (f64 FFT inside)",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _OUT.emit(json.dumps(line))
    return 0


# --------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import audioanalysisdetector_b200 as aad
    from audioanalysisdetector_b200.frontend import Frontend, FrontendParams
    from audioanalysisdetector_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # one process per GPU: run (and first-touch the pinned host buffers) on the GPU's own NUMA node
    numa_cpus = aad.bind_to_gpu_numa(physical_gpu_index(local_rank)) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    B, Ls = args.clips, int(CLIP_S * SR)
    params = FrontendParams.mfcc(SR, n_mfcc=N_MFCC, n_mels=N_MELS, n_fft=N_FFT, hop_length=HOP, n_delta=N_DELTA)
    fe = Frontend(params, dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    wav = (0.1 * torch.randn((B, Ls), generator=gen, device=dev)).clamp_(-1.0, 1.0)
    lengths = torch.full((B,), Ls, dtype=torch.int32, device=dev)
    t_max, c_out, _ = fe.query(B, Ls)
    out = torch.zeros((B, c_out, t_max), dtype=torch.float32, device=dev)
    frames = B * t_max
    hours_per_step = B * CLIP_S / 3600.0

    # ---- device-resident timed region ------------------------------------------------
    for _ in range(args.warmup):
        fe(wav, lengths, out=out)
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        fe(wav, lengths, out=out)
    e1.record()
    barrier()
    ms_total = max_over_ranks(float(e0.elapsed_time(e1)))
    clocks = sampler.stop()
    ms_per_step = ms_total / args.steps
    value = world * hours_per_step / (ms_per_step * 1e-3)
    _, nf, st = fe(wav, lengths, out=out)
    assert int(st.sum().item()) == 0 and int(nf.min().item()) == t_max

    # ---- per-kernel timing pass (events around each launch, same stream) -------------
    fe.set_profiling(True)
    kt = {"prepare": [], "stft_fb": [], "epilogue": []}
    for _ in range(max(3, min(args.steps, 20))):
        fe(wav, lengths, out=out)
        torch.cuda.synchronize(dev)
        t = fe.kernel_times_ms()
        for k in kt:
            kt[k].append(t[k])
    fe.set_profiling(False)
    kms = {k: float(np.mean(v)) for k, v in kt.items()}

    # ---- end-to-end: pinned host buffers through the library's host entry point -------
    host_wav = torch.empty((B, Ls), dtype=torch.float32, pin_memory=True)
    host_wav.copy_(wav)
    host_out = torch.empty((B, c_out, t_max), dtype=torch.float32, pin_memory=True)
    hw, ho = host_wav.numpy(), host_out.numpy()
    hlen = np.full(B, Ls, dtype=np.int32)
    for _ in range(max(1, min(args.warmup, 2))):
        fe.extract_host(hw, hlen, out=ho)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        _, nf_h, st_h = fe.extract_host(hw, hlen, out=ho)
    torch.cuda.synchronize(dev)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    assert int(st_h.sum()) == 0
    e2e_value = world * hours_per_step * e2e_steps / e2e_s
    e2e_err = float(np.abs(ho[:8] - out[:8].cpu().numpy()).max())

    # ---- same, 16-bit PCM over the bus (the native format of the corpus; pcm/32768 is what
    #      librosa.load decodes): half the H2D bytes, features bit-identical to the float path
    host_pcm = torch.empty((B, Ls), dtype=torch.int16, pin_memory=True)
    host_pcm.copy_((wav * 32767.0).round().clamp_(-32768, 32767).to(torch.int16))
    hp = host_pcm.numpy()
    fe.extract_host(hp, hlen, out=ho)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        fe.extract_host(hp, hlen, out=ho)
    torch.cuda.synchronize(dev)
    pcm_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    ref8, _, _ = fe(host_pcm[:8].to(dev).to(torch.float32).div_(32768.0))
    pcm_err = float(np.abs(ho[:8] - ref8.cpu().numpy()).max())
    pcm_value = world * hours_per_step * e2e_steps / pcm_s

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline ---------------------------------------------------------------------
    fb = fe.table(L.TABLE_FILTERBANK)
    work = algorithmic_work(N_FFT, HOP, N_MELS, N_MFCC, N_DELTA, int((fb != 0).sum()), c_out)
    peaks, peak_src = load_measured_peaks()
    fp32_peak = aad.fp32_peak_tflops(local_rank)
    fp32_nominal = 148 * 128 * 2 * 1.965e9 / 1e12
    k1_tflops = work["flops_stft_fb"] * frames / (kms["stft_fb"] * 1e-3) / 1e12
    step_tflops = work["flops"] * frames / (ms_per_step * 1e-3) / 1e12
    step_gbs = work["bytes"] * frames / (ms_per_step * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "stft_fb_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {
        "bound": "fp32",
        "bound_note": "north_star: the slower of the FP32 compute peak and the HBM roofline; K1 does %.1f flop per byte, "
                      "the ridge is %.1f, so the FP32 (CUDA-core) peak bounds it, not HBM and not the tensor pipe"
                      % (work["flops_stft_fb"] / (4 * HOP + 4 * N_MELS), fp32_peak * 1e3 / peaks.get("hbm_gbs", 6650.0)),
        "kernel": "k_stft_fb<32,f32>", "achieved": k1_tflops, "peak": fp32_peak,
        "unit": "TFLOP/s", "frac": k1_tflops / fp32_peak if fp32_peak else None, "traffic": traffic,
        "peak_source": "FFMA micro-benchmark measured in this run (aad_fp32_peak); nominal 148x128x2x1.965 GHz = %.1f" % fp32_nominal,
        "algorithmic_flops_per_frame": work["flops_stft_fb"], "frames_per_launch": frames,
        "kernel_ms": kms,
        "whole_step": {"achieved_tflops": step_tflops, "frac_fp32": step_tflops / fp32_peak if fp32_peak else None,
                       "flops_per_frame": work["flops"]},
        "hbm": {"achieved": step_gbs, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                "frac": step_gbs / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None,
                "bytes_per_frame": work["bytes"], "peak_source": peak_src},
    }

    # ---- CPU baseline (oracle port on the host cores, bounded sample) -------------------
    cores = os.cpu_count() or 1
    cpu = None
    if world == 1 and not args.no_cpu:
        n_sample = calibrate_cpu_sample(args.cpu_seconds, cores)
        from joblib import Parallel
        with Parallel(n_jobs=cores) as pool:
            cpu_reference_pass(cores * 2, cores, pool)        # warm the loky workers
            dt = cpu_reference_pass(n_sample, cores, pool)
        cpu = {"value": n_sample * CLIP_S / 3600.0 / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n_sample} of {B} clips, in memory, joblib loky n_jobs={cores}, oracle numpy/scipy "
                         f"restatement of librosa 0.11 (librosa not installed here)"}

    # optional second CPU baseline: torchaudio's own MFCC chain on the host, all torch threads,
    # batched (no per-clip process fan-out); same algorithm for the static rows, its own delta edge handling
    cpu_ta = None
    if world == 1 and not args.no_cpu:
        try:
            import torchaudio
            n_ta = min(B, 512)
            yb = torch.clamp(0.1 * torch.randn((n_ta, Ls), generator=torch.Generator().manual_seed(7)), -1, 1)
            mfcc_t = torchaudio.transforms.MFCC(sample_rate=SR, n_mfcc=40, dct_type=2, norm="ortho", log_mels=False,
                                                melkwargs=dict(n_fft=2048, hop_length=512, n_mels=128, center=True,
                                                               pad_mode="constant", power=2.0, norm="slaney",
                                                               mel_scale="slaney", f_min=0.0, f_max=SR / 2))

            def ta_pass():
                with torch.no_grad():
                    m = mfcc_t(yb)
                    d1 = torchaudio.functional.compute_deltas(m, win_length=9)
                    return torch.cat([m, d1, torchaudio.functional.compute_deltas(d1, win_length=9)], dim=1)
            ta_pass()
            t0 = time.perf_counter()
            ta_pass()
            dt_ta = time.perf_counter() - t0
            cpu_ta = {"value": n_ta * CLIP_S / 3600.0 / dt_ta, "unit": UNIT, "threads": torch.get_num_threads(),
                      "kind": "torchaudio " + torchaudio.__version__,
                      "sample": f"{n_ta} of {B} clips as one batch, transforms.MFCC(40) + compute_deltas x2 on the CPU "
                                f"(f32 FFT; delta edges replicate instead of librosa's interp fit): informational"}
        except Exception as e:  # torchaudio missing or a different API: the key stays null
            cpu_ta = {"unavailable": str(e)[:120]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "clips_per_gpu": B, "frames_per_gpu": frames,
                   "l2": "inputs larger than L2 (%.2f GB waveforms per GPU per step, no flush needed)" % (B * Ls * 4 / 1e9),
                   "parallelism": f"utterance-sharded x{world}, no collective in the hot path",
                   "host_affinity": (f"rank bound to {len(numa_cpus)} GPU-local CPUs (NVML)" if numa_cpus else "unbound")},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * Ls * 4 + B * 4,
                "d2h_bytes_per_step": B * c_out * t_max * 4 + 2 * B * 4, "steps": e2e_steps,
                "ms_per_step": e2e_s / e2e_steps * 1e3,
                "h2d_gbs": (B * Ls * 4) / (e2e_s / e2e_steps) / 1e9,
                "api": "Frontend.extract_host -> aad_extract_host (pinned host in/out, chunked H2D/compute/D2H on 3 streams)",
                "note": "PCIe-bound: the box measures 55.6 GB/s H2D (profiles/r1_pcie_probe.log)",
                "max_abs_diff_vs_device_path": e2e_err},
        "e2e_pcm16": {"value": pcm_value, "unit": UNIT, "h2d_bytes_per_step": B * Ls * 2 + B * 4,
                      "d2h_bytes_per_step": B * c_out * t_max * 4 + 2 * B * 4, "steps": e2e_steps,
                      "ms_per_step": pcm_s / e2e_steps * 1e3,
                      "note": "same call with int16 PCM host buffers (pcm * 2^-15 folded into the window); "
                              "informational, the headline e2e above moves float32",
                      "max_abs_diff_vs_float_path_on_pcm_over_32768": pcm_err},
        "gpu_launches": fe.launches_per_call * args.steps,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "cpu_baseline_torchaudio": cpu_ta,
    }
    _OUT.emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


class _StdoutToStderr:
    """Everything written to fd 1 while the bench runs (NCCL prints its version banner there) goes to
    stderr, so that stdout carries exactly one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, line: str):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        print(line, flush=True)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


_OUT = None


def main():
    global _OUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clips", type=int, default=CLIPS_PER_GPU, help="clips per GPU (default: the named workload)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU work per baseline pass (core-seconds)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    with _StdoutToStderr() as _OUT:
        if args.impl == "reference":
            return run_reference(args)
        return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
