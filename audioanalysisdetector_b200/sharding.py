"""Multi-GPU sharding of utterance batches: one process per GPU, no data-path collective.

The reference's only parallelism is one joblib task per clip (ASV_dl_func.py:1036-1045):
utterances are independent, so ranks get disjoint sets of utterances balanced by frame
count (greedy longest-processing-time) and run the same kernels on their own HBM shard.
The optional gather of features / n_frames (outside the hot path) goes through
torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import heapq
from typing import List, Optional, Sequence

import numpy as np
import torch


def partition_by_frames(n_frames: Sequence[int], world_size: int) -> List[np.ndarray]:
    """Greedy LPT: indices per rank such that sum(n_frames) per rank is balanced.

    Deterministic (ties broken by index), every rank's indices are sorted ascending."""
    n_frames = np.asarray(n_frames, dtype=np.int64)
    order = np.lexsort((np.arange(len(n_frames)), -n_frames))
    heap = [(0, r) for r in range(world_size)]
    heapq.heapify(heap)
    parts: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        load, r = heapq.heappop(heap)
        parts[r].append(int(i))
        heapq.heappush(heap, (load + int(n_frames[i]), r))
    return [np.array(sorted(p), dtype=np.int64) for p in parts]


def contiguous_shard(n_items: int, rank: int, world_size: int) -> slice:
    """Equal contiguous split for fixed-length batches (BASELINE configs 2 and 4)."""
    per = (n_items + world_size - 1) // world_size
    return slice(min(rank * per, n_items), min((rank + 1) * per, n_items))


def gather_features(local: torch.Tensor, local_index: torch.Tensor, n_total: int,
                    group=None) -> Optional[torch.Tensor]:
    """All-gather per-rank feature shards back into utterance order (rank-agnostic result).

    local [n_local, ...], local_index [n_local] global utterance ids.  Shards are padded to
    the largest shard so a single all_gather suffices."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        out = torch.zeros((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        out[local_index.long()] = local
        return out
    ws = dist.get_world_size(group)
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(n_local) for _ in range(ws)]
    dist.all_gather(counts, n_local, group=group)
    n_max = int(max(int(c.item()) for c in counts))
    pad = torch.zeros((n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    idx = torch.full((n_max,), -1, dtype=torch.int64, device=local.device)
    idx[: local.shape[0]] = local_index.long()
    feats = [torch.empty_like(pad) for _ in range(ws)]
    idxs = [torch.empty_like(idx) for _ in range(ws)]
    dist.all_gather(feats, pad, group=group)
    dist.all_gather(idxs, idx, group=group)
    out = torch.zeros((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for f, i in zip(feats, idxs):
        m = i >= 0
        out[i[m]] = f[m]
    return out


# --------------------------------------------------------------------------------------------------
# Long-form audio with fewer utterances than GPUs (BASELINE config 5): split ONE utterance
# along time.  Frames are independent except for power_to_db's per-utterance maximum (ref=np.max and
# the top_db floor, ASV_dl_func.py:534), so the only exchange step is one MAX all-reduce of a float.
# --------------------------------------------------------------------------------------------------
def time_split(length: int, n_fft: int, hop: int, rank: int, world_size: int) -> dict:
    """Rank's share of the frames of one centred STFT (T = 1 + length // hop frames).

    Returns the frame range [t0, t1) it owns, the sample range [s0, s1) it has to read so that each of
    those frames sees exactly the samples it sees in the unsplit signal, and `skip`: local frame
    `skip + i` of the slice (extracted with the usual centre padding) IS global frame `t0 + i`.
    The slice starts `ceil((n_fft/2) / hop)` hops early (s0 is a multiple of hop, so aligned loads stay
    aligned) and ends n_fft/2 samples after the centre of the last owned frame; local frames before
    `skip` or after `skip + (t1 - t0)` touch the slice's artificial zero padding and are dropped."""
    T = 1 + length // hop
    fr = contiguous_shard(T, rank, world_size)
    t0, t1 = fr.start, fr.stop
    if t1 <= t0:
        return {"t0": t0, "t1": t0, "s0": 0, "s1": 0, "skip": 0}
    k = -(-(n_fft // 2) // hop)
    skip = min(k, t0)
    s0 = (t0 - skip) * hop
    s1 = min(length, (t1 - 1) * hop + n_fft // 2)
    return {"t0": t0, "t1": t1, "s0": s0, "s1": max(s1, s0), "skip": skip}


_split_tables: dict = {}


def long_form_logmel(params, wav: torch.Tensor, rank: int, world_size: int, group=None,
                     local_max_hook=None, check_status: bool = True):
    """Log-mel (dB) of ONE long utterance `wav` [L] (on this rank's GPU) split along time over the ranks.

    Every rank extracts its frames with the reference disabled (raw 10 log10(max(amin, S))), the
    utterance maximum is MAX-all-reduced (the one exchange step), then `aad_db_reference` applies
    `ref=np.max` / `top_db` exactly as power_to_db does.  Returns (features [n_mels, t1 - t0], (t0, t1)):
    the concatenation over ranks equals the unsplit extraction bit for bit.
    The features are a view into the piece's output (row stride = frames of the piece).  Nothing in the
    call waits for the GPU unless `check_status` (a device->host read of the piece's status word).
    `local_max_hook(local_max) -> global_max` replaces the all-reduce (tests on one GPU)."""
    import ctypes as C
    import torch.distributed as dist
    from . import _lib as L
    from .extractors import get_frontend
    from .frontend import _ptr
    if params.kind != L.KIND_LOGMEL or not params.center or params.time_mean or params.znorm or params.layout != L.LAYOUT_CT:
        raise L.AadError("long_form_logmel needs a centred log-mel plan in CT layout")
    part = time_split(int(wav.numel()), params.n_fft, params.hop_length, rank, world_size)
    n = part["t1"] - part["t0"]
    dev = wav.device
    raw = params.replace(ref_type=L.REF_ONE, top_db=-1.0)
    fe = get_frontend(raw, dev)          # cached plan: repeated calls build no tables
    key = (int(wav.numel()), params.n_fft, params.hop_length, rank, world_size, str(dev))
    tabs = _split_tables.get(key)
    if tabs is None:                     # the one-row chunk table of this piece, built once (no per-call H2D)
        tabs = _split_tables[key] = (torch.tensor([part["s0"]], dtype=torch.int64, device=dev),
                                     torch.tensor([max(part["s1"] - part["s0"], 0)], dtype=torch.int32, device=dev),
                                     torch.tensor([n], dtype=torch.int32, device=dev))
    off, ln, nfr = tabs
    status = None
    if n > 0:
        feats, _, status = fe.extract_indexed(wav, off, ln, max_len=part["s1"] - part["s0"], validate=False)
        mine = feats[0, :, part["skip"]:part["skip"] + n]          # a view: rows keep the stride of the piece
        local_max = mine.amax().reshape(1)
    else:
        mine = torch.zeros((params.n_filt, 0), dtype=torch.float32, device=dev)
        local_max = torch.full((1,), float("-inf"), dtype=torch.float32, device=dev)
    if local_max_hook is not None:
        gmax = local_max_hook(local_max)
    else:
        gmax = local_max
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(gmax, op=dist.ReduceOp.MAX, group=group)
    if n > 0:
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            L.check(fe.lib.aad_db_reference(_ptr(mine), 0, mine.stride(0), _ptr(nfr), _ptr(gmax.contiguous()), 1,
                                            params.n_filt, n, int(params.ref_type), float(params.top_db),
                                            C.c_void_p(stream)), "aad_db_reference")
    if check_status and status is not None and int(status[0]) != 0:
        raise L.AadError(f"piece {rank}: {L.ITEM_STATUS_NAMES.get(int(status[0]), 'item failed')}")
    return mine, (part["t0"], part["t1"])


def bind_to_gpu_numa(physical_gpu_index: int) -> Optional[List[int]]:
    """Pin the calling process to the CPUs NVML reports as local to the GPU (same NUMA node / PCIe
    root), so that pinned host buffers allocated afterwards are first-touched next to it.  With one
    process per GPU this keeps every rank's H2D/D2H traffic off the inter-socket link.  Returns the
    CPU list, or None when NVML / sched_setaffinity are unavailable (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(physical_gpu_index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None
