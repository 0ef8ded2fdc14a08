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
