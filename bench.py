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
class Ranks:
    """One process per GPU: device, NCCL group, barrier and max-over-ranks timing (the bench contract)."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        import audioanalysisdetector_b200 as aad
        self.torch, self.dist, self.aad = torch, dist, aad
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        # one process per GPU: run (and first-touch the pinned host buffers) on the GPU's own NUMA node
        self.numa_cpus = aad.bind_to_gpu_numa(physical_gpu_index(self.local_rank)) if self.world > 1 else None
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps: int, warmup: int):
        """ms per step of fn() on this rank's current stream: barrier + sync on both sides, CUDA events, max over
        ranks; the SM clock is sampled while the timed region runs."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        sampler = ClockSampler(physical_gpu_index(self.local_rank))
        sampler.start()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        ms = self.max_over_ranks(float(e0.elapsed_time(e1))) / steps
        return ms, sampler.stop()

    def affinity_note(self):
        if not self.numa_cpus:
            return "unbound"
        note = f"rank bound to {len(self.numa_cpus)} GPU-local CPUs (NVML)"
        if self.world > 1 and len(self.numa_cpus) == (os.cpu_count() or 0):
            note += "; every GPU reports the same CPU set on this host (one NUMA node): the binding changes nothing here"
        return note

    def done(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def base_line(R, args, value, ms_per_step, scaling, dtype, workload, extra_config, clocks, launches):
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": R.world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": dtype,
        "data": "synthetic",
        "config": dict({"workload": workload, "parallelism": f"utterance-sharded x{R.world}, no collective in the hot path",
                        "host_affinity": R.affinity_note()}, **extra_config),
        "gpu_launches": launches, "clocks": clocks,
    }


# --------------------------------------------------------------------------- configs[2]: ragged LFCC
def run_c3(args):
    """BASELINE configs[2]: LFCC 20 filters x 3 (delta, delta-delta) on variable-length 1-8 s int16 clips, padded to
    8 s; 4096 clips per GPU, the global batch partitioned over the ranks by frame count (greedy LPT)."""
    R = Ranks()
    torch, aad = R.torch, R.aad
    from audioanalysisdetector_b200.frontend import Frontend, FrontendParams
    from audioanalysisdetector_b200 import _lib as L
    per_gpu, lmax = args.clips, 8 * SR
    G = per_gpu * R.world
    lens_all = np.random.default_rng(3).integers(SR, lmax + 1, size=G).astype(np.int64)
    params = FrontendParams.lfcc(SR, n_ceps=20, nfilts=20, win_len=0.02, n_delta=2, layout=L.LAYOUT_CT)
    frames_all = np.array([params.n_frames(int(n)) for n in lens_all])
    parts = aad.partition_by_frames(frames_all, R.world)
    mine = parts[R.rank]
    lens = torch.from_numpy(lens_all[mine].astype(np.int32)).to(R.dev)
    gen = torch.Generator(device=R.dev)
    gen.manual_seed(300 + R.rank)
    wav = (torch.randn((len(mine), lmax), generator=gen, device=R.dev) * 3000).clamp_(-32767, 32767).to(torch.int16)
    fe = Frontend(params, R.dev)
    t_max, c_out, _ = fe.query(len(mine), lmax)
    out = torch.zeros((len(mine), c_out, t_max), dtype=torch.float32, device=R.dev)
    ms, clocks = R.timed(lambda: fe(wav, lens, out=out), args.steps, args.warmup)
    _, nf, st = fe(wav, lens, out=out)
    assert int(st.sum().item()) == 0 and int(nf.sum().item()) == int(frames_all[mine].sum())
    loads = [int(frames_all[p].sum()) for p in parts]
    hours = float(lens_all.sum()) / SR / 3600.0
    if R.rank == 0:
        line = base_line(R, args, hours / (ms * 1e-3), ms, "weak", "f32 (int16 PCM in)",
                         "configs[2]: LFCC 20 filters x (static, delta, delta-delta), int16 PCM, ragged 1-8 s clips padded "
                         "to 8 s (win 20 ms / hop 10 ms / n_fft 512), 4096 clips per GPU",
                         {"clips_total": G, "frames_total": int(frames_all.sum()),
                          "partition": "greedy LPT on frame counts (sharding.partition_by_frames)",
                          "frames_per_rank_max_over_mean": max(loads) / (sum(loads) / len(loads)),
                          "l2": "inputs larger than L2 (%.2f GB of PCM per GPU per step)" % (len(mine) * lmax * 2 / 1e9)},
                         clocks, fe.launches_per_call * args.steps)
        _OUT.emit(json.dumps(line))
    R.done()
    return 0


# --------------------------------------------------------------------------- configs[3]: corpus -> features -> scores
def run_c4(args):
    """BASELINE configs[3]: ASVspoof-2019-LA-sized synthetic corpus (25 380 two-second chunks @16 kHz) -> MFCC-13 and
    log-mel-64 from ONE STFT -> CNN-BiLSTM scores on the same device, sharded by utterance (strong scaling); the
    all-gather of the scores is timed separately (outside the hot path)."""
    R = Ranks()
    torch, aad = R.torch, R.aad
    from audioanalysisdetector_b200.frontend import Frontend, FrontendParams
    n_chunks, chunk = 25380, 2 * SR
    sl = aad.contiguous_shard(n_chunks, R.rank, R.world)
    n_local = sl.stop - sl.start
    gen = torch.Generator(device=R.dev)
    gen.manual_seed(4242 + R.rank)
    wav = (0.1 * torch.randn((n_local, chunk), generator=gen, device=R.dev)).clamp_(-1, 1)
    g = np.load(os.path.join(ROOT, "tests", "golden", "consumer.npz"))
    weights = {k[3:]: torch.from_numpy(g[k]).to(R.dev) for k in g.files if k.startswith("w::")}
    fe_mfcc = Frontend(FrontendParams.mfcc(SR, n_mfcc=13), R.dev)
    fe_mel = Frontend(FrontendParams.logmel(SR, n_mels=64), R.dev)
    engine = aad.DetectorEngine(weights, feature_dim=13, device=R.dev)
    state = {}

    def step():
        (feats, mel), nf, st = fe_mfcc.extract_pair(fe_mel, wav)
        state["scores"], state["st"] = engine(feats), st

    ms, clocks = R.timed(step, args.steps, args.warmup)
    ms_feat, _ = R.timed(lambda: fe_mfcc.extract_pair(fe_mel, wav), max(3, args.steps // 2), 1)
    idx = torch.arange(sl.start, sl.stop, device=R.dev)
    ms_gather, _ = R.timed(lambda: aad.gather_features(state["scores"], idx, n_chunks), 5, 1)
    all_scores = aad.gather_features(state["scores"], idx, n_chunks)
    assert int(state["st"].ne(0).sum().item()) == 0 and tuple(all_scores.shape)[0] == n_chunks
    hours = n_chunks * chunk / SR / 3600.0
    if R.rank == 0:
        line = base_line(R, args, hours / (ms * 1e-3), ms, "strong", "f32",
                         "configs[3]: 25 380 x 2 s @16 kHz -> MFCC-13 + log-mel-64 (one STFT) -> CNN-BiLSTM scores, "
                         "sharded by utterance",
                         {"chunks_total": n_chunks, "chunks_per_gpu": n_local, "features_ms": ms_feat,
                          "model_ms": ms - ms_feat, "scores_all_gather_ms_outside_the_step": ms_gather,
                          "scores_mean": float(all_scores.mean().item())},
                         clocks, (fe_mfcc.launches_per_call + 1 + 3) * args.steps)
        _OUT.emit(json.dumps(line))
    R.done()
    return 0


# --------------------------------------------------------------------------- the notebook's whole extractor map
def run_map(args):
    """The reference notebook's feature_extractors_map (ASV_deep_learning.ipynb:152-160: cqcc, gtcc, mel-spect, mfcc,
    lfcc) over the configs[3] corpus: 25 380 two-second chunks @16 kHz resident in HBM as int16 PCM, sharded by
    utterance (strong scaling).  mel-spect and mfcc share one STFT (aad_extract_pair)."""
    R = Ranks()
    torch, aad = R.torch, R.aad
    from audioanalysisdetector_b200.frontend import Frontend, FrontendParams
    n_chunks, chunk = 25380, 2 * SR
    sl = aad.contiguous_shard(n_chunks, R.rank, R.world)
    n_local = sl.stop - sl.start
    gen = torch.Generator(device=R.dev)
    gen.manual_seed(777 + R.rank)
    wav = (3276.8 * torch.randn((n_local, chunk), generator=gen, device=R.dev)).clamp_(-32768, 32767).to(torch.int16)
    cq = aad.CqccFrontend(SR, device=R.dev)
    gt = Frontend(FrontendParams.gtcc(SR), R.dev)
    mf = Frontend(FrontendParams.mfcc(SR, n_mfcc=13), R.dev)
    ml = Frontend(FrontendParams.logmel(SR, n_mels=64), R.dev)
    lf = Frontend(FrontendParams.lfcc(SR), R.dev)
    parts = {"cqcc": lambda: cq(wav), "gtcc": lambda: gt(wav), "mfcc + mel-spect (one STFT)": lambda: mf.extract_pair(ml, wav),
             "lfcc": lambda: lf(wav)}
    state = {}

    def step():
        state["out"] = [f() for f in parts.values()]

    ms, clocks = R.timed(step, args.steps, args.warmup)
    split = {k: R.timed(f, max(3, args.steps // 2), 1)[0] for k, f in parts.items()}
    bad = sum(int(o[-1].ne(0).sum().item()) if not isinstance(o[-1], tuple) else 0 for o in
              [(state["out"][0][2],), (state["out"][1][2],), (state["out"][2][2],), (state["out"][3][2],)])
    assert bad == 0
    hours = n_chunks * chunk / SR / 3600.0
    if R.rank == 0:
        launches = 14 + 1 + gt.launches_per_call + mf.launches_per_call + 1 + lf.launches_per_call   # cqcc: 7 octaves + 6 resamplers + epilogue (+ memset)
        line = base_line(R, args, hours / (ms * 1e-3), ms, "strong", "f32 (int16 PCM in; CQT octaves 3xTF32 on the tensor pipe)",
                         "the reference notebook's whole extractor map (cqcc-19, gtcc-13, mel-spect-64, mfcc-13, lfcc-13) over "
                         "25 380 x 2 s @16 kHz, sharded by utterance",
                         {"chunks_total": n_chunks, "chunks_per_gpu": n_local, "ms_per_feature": split,
                          "reference_minutes_for_28408_chunks_8_workers": 42.2}, clocks, launches * args.steps)
        _OUT.emit(json.dumps(line))
    R.done()
    return 0


# --------------------------------------------------------------------------- configs[4]: long-form batch sweep
def run_c5(args):
    """BASELINE configs[4]: 10 min @48 kHz, n_fft 2048, hop 480, 128 mels, batch sweep B = 1 .. 64.  B >= N GPUs:
    utterances are sharded; B < N: every utterance is split along time over N / B ranks (one MAX all-reduce of a
    float per utterance group: power_to_db's ref=np.max)."""
    R = Ranks()
    torch, dist, aad = R.torch, R.dist, R.aad
    from audioanalysisdetector_b200.frontend import Frontend, FrontendParams
    sr, L_ = 48000, 48000 * 600
    params = FrontendParams.logmel(sr, n_mels=128, n_fft=2048, hop_length=480)
    fe = Frontend(params, R.dev)
    groups = {}
    if R.world > 1:  # sub-groups for the time split (every rank creates every group, in the same order)
        for per in (2, 4, 8):
            if per <= R.world:
                for g0 in range(0, R.world, per):
                    grp = dist.new_group(list(range(g0, g0 + per)))
                    if g0 <= R.rank < g0 + per:
                        groups[per] = grp
    sweep = []
    gen = torch.Generator(device=R.dev)
    gen.manual_seed(55 + R.rank)
    steps = max(3, min(args.steps, 10))
    for B in (1, 2, 4, 8, 16, 32, 64):
        if B >= R.world:
            nb = B // R.world
            wav = (0.1 * torch.randn((nb, L_), generator=gen, device=R.dev)).clamp_(-1, 1)
            t_max, c_out, _ = fe.query(nb, L_)
            out = torch.empty((nb, c_out, t_max), dtype=torch.float32, device=R.dev)
            ms, clocks = R.timed(lambda: fe(wav, out=out), steps, 3)
            mode = "utterance-sharded"
            del out
        else:
            per = R.world // B                       # ranks per utterance
            wav = (0.1 * torch.randn(L_, generator=gen, device=R.dev)).clamp_(-1, 1)
            sub = R.rank % per
            ms, clocks = R.timed(lambda: aad.long_form_logmel(params, wav, sub, per, group=groups.get(per),
                                                              check_status=False), steps, 3)
            mode = f"time-split over {per} ranks per utterance"
        hours = B * 600 / 3600.0
        sweep.append({"B": B, "ms_per_step": ms, "audio_hours_per_s": hours / (ms * 1e-3), "mode": mode,
                      "sm_mhz": clocks.get("sm_mhz"), "reasons": clocks.get("reasons")})
        del wav
        torch.cuda.empty_cache()
    if R.rank == 0:
        best = max(sweep, key=lambda r: r["audio_hours_per_s"])
        line = base_line(R, args, best["audio_hours_per_s"], best["ms_per_step"], "strong", "f32",
                         "configs[4]: log-mel 128 of 10 min @48 kHz (n_fft 2048, hop 480), batch sweep B = 1..64",
                         {"sweep": sweep, "value_is": f"the best point of the sweep (B = {best['B']})",
                          "frames_per_utterance": params.n_frames(L_)},
                         {"sm_mhz": best["sm_mhz"], "reasons": best["reasons"]}, fe.launches_per_call * steps)
        line["steps"] = steps
        _OUT.emit(json.dumps(line))
    R.done()
    return 0


# --------------------------------------------------------------------------- configs[1]: the bench workload
def run_b200(args):
    R = Ranks()
    torch, dist, aad = R.torch, R.dist, R.aad
    from audioanalysisdetector_b200.frontend import Frontend, FrontendParams
    from audioanalysisdetector_b200 import _lib as L
    world, rank, local_rank, dev, numa_cpus = R.world, R.rank, R.local_rank, R.dev, R.numa_cpus
    barrier, max_over_ranks = R.barrier, R.max_over_ranks

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
    # Headline `e2e`: 16-bit PCM in, float32 features out.  PCM16 is the corpus's native format: the reference
    # decodes 16-bit FLAC for every call (librosa.load, ASV_dl_func.py:406,425,524) and pcm / 32768 is exactly the
    # float librosa hands on, so the features are bit-identical to the float path (checked below) for half the
    # bytes over PCIe.  `e2e_f32` is the same call with the decoded float32 waveform as host input.
    host_out = torch.empty((B, c_out, t_max), dtype=torch.float32, pin_memory=True)
    ho = host_out.numpy()
    hlen = np.full(B, Ls, dtype=np.int32)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))

    def time_host_path(hin):
        for _ in range(max(1, min(args.warmup, 2))):
            fe.extract_host(hin, hlen, out=ho)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            _, nf_h, st_h = fe.extract_host(hin, hlen, out=ho)
        torch.cuda.synchronize(dev)
        dt = max_over_ranks(time.perf_counter() - t0)
        barrier()
        assert int(st_h.sum()) == 0
        return dt

    pcm_dev = (wav * 32767.0).round().clamp_(-32768, 32767).to(torch.int16)
    host_pcm = torch.empty((B, Ls), dtype=torch.int16, pin_memory=True)
    host_pcm.copy_(pcm_dev)
    fe.reserve_host(B, Ls, np.int16)
    pcm_s = time_host_path(host_pcm.numpy())
    ref8, _, _ = fe(pcm_dev[:8].to(torch.float32).div_(32768.0))
    pcm_err = float(np.abs(ho[:8] - ref8.cpu().numpy()).max())
    e2e_value = world * hours_per_step * e2e_steps / pcm_s
    # the same with the PCM staged in write-combined pinned memory (aad.pinned_empty(write_combined=True)): the
    # corpus loader writes it once, the GPUs read it without cache snooping -- matters when several GPUs pull
    # from one host memory system
    wc_s = None
    try:
        wc = aad.pinned_empty((B, Ls), np.int16, write_combined=True)
        wc[...] = host_pcm.numpy()
        wc_s = time_host_path(wc)
        del wc
    except Exception as e:  # allocation refused: the key stays null
        sys.stderr.write(f"write-combined staging unavailable: {e}\n")
    host_wav = torch.empty((B, Ls), dtype=torch.float32, pin_memory=True)
    host_wav.copy_(wav)
    f32_s = time_host_path(host_wav.numpy())
    f32_err = float(np.abs(ho[:8] - out[:8].cpu().numpy()).max())
    del host_wav

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
                   "host_affinity": R.affinity_note()},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * Ls * 2 + B * 4,
                "d2h_bytes_per_step": B * c_out * t_max * 4 + 2 * B * 4, "steps": e2e_steps,
                "ms_per_step": pcm_s / e2e_steps * 1e3,
                "h2d_gbs": (B * Ls * 2) / (pcm_s / e2e_steps) / 1e9,
                "d2h_gbs": (B * c_out * t_max * 4) / (pcm_s / e2e_steps) / 1e9,
                "input": "int16 PCM host buffers (pinned), the corpus's native format: the reference decodes 16-bit FLAC "
                         "per call (librosa.load, ASV_dl_func.py:406); pcm * 2^-15 is folded into the window",
                "api": "Frontend.extract_host -> aad_extract_host (pinned host in/out, chunked H2D/compute/D2H on 3 streams)",
                "note": "PCIe-bound: the box measures 55.6 GB/s H2D and 57.2 GB/s D2H per GPU (profiles/r1_pcie_probe.log)",
                "max_abs_diff_vs_float_path_on_pcm_over_32768": pcm_err},
        "e2e_wc": None if wc_s is None else {
            "value": world * hours_per_step * e2e_steps / wc_s, "unit": UNIT, "ms_per_step": wc_s / e2e_steps * 1e3,
            "note": "same call, PCM staged in write-combined pinned memory (aad.pinned_empty(write_combined=True))"},
        "e2e_f32": {"value": world * hours_per_step * e2e_steps / f32_s, "unit": UNIT,
                    "h2d_bytes_per_step": B * Ls * 4 + B * 4, "d2h_bytes_per_step": B * c_out * t_max * 4 + 2 * B * 4,
                    "steps": e2e_steps, "ms_per_step": f32_s / e2e_steps * 1e3,
                    "h2d_gbs": (B * Ls * 4) / (f32_s / e2e_steps) / 1e9,
                    "note": "same call with the decoded float32 waveform as host input (twice the H2D bytes)",
                    "max_abs_diff_vs_device_path": f32_err},
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
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5", "map"],
                    help="c2 = BASELINE configs[1] (the default the driver runs); c3 / c4 / c5 = configs[2] / [3] / [4]; "
                         "map = the reference notebook's five-extractor map over the configs[3] corpus")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    with _StdoutToStderr() as _OUT:
        if args.impl == "reference":
            return run_reference(args)
        return {"c2": run_b200, "c3": run_c3, "c4": run_c4, "c5": run_c5, "map": run_map}[args.workload](args)


if __name__ == "__main__":
    sys.exit(main())
