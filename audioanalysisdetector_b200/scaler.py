"""Feature standardisation on the device: the step right after the extractors in the reference.

`prepare_train_test_data` / `prepare_train_test_data_multi` (ASV_dl_func.py:1090-1129) and
`train_all_features` (:963-973) fit `sklearn.preprocessing.StandardScaler` on
`np.vstack(df[col].values)` -- every utterance's 2-D feature array stacked along its first axis -- and
then `transform` each utterance.  Here the features never leave the GPU: column sums and sums of squares
come from one pass of a CUDA kernel in double precision; with one process per GPU the partial sums are
all-reduced (the only exchange step near the path: 2*W doubles) so that every rank holds the statistics
of the whole corpus; `transform` is one in-place pass.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib as L


def merge_stats(count: int, sums: np.ndarray, sumsq: np.ndarray):
    """(count, column sums, column sums of squares) -> (mean_, var_, scale_) exactly as sklearn defines
    them: population variance, scale 1 where the variance is 0."""
    mean = sums / count
    var = np.maximum(sumsq / count - mean * mean, 0.0)
    scale = np.sqrt(var)
    scale[scale == 0.0] = 1.0
    return mean, var, scale


class DeviceStandardScaler:
    """fit / transform with sklearn StandardScaler semantics on [B, R, W] CUDA tensors.

    W (the standardised columns) must be the LAST axis, i.e. the extractor's time-major TC layout
    `[B, T, C]` (or any `[rows, W]` stacking); a CT tensor has to be transposed by the caller.  For ragged
    batches pass the extractor's `n_frames` (and `status`) to `fit`: the reference fits on
    `np.vstack(per-utterance arrays)`, valid frames only, so padding rows must not enter the statistics."""

    def __init__(self):
        self.mean_: Optional[np.ndarray] = None
        self.var_: Optional[np.ndarray] = None
        self.scale_: Optional[np.ndarray] = None
        self.n_samples_seen_ = 0
        self._mean_d = self._inv_d = None

    @staticmethod
    def _rows(x: torch.Tensor):
        if not x.is_cuda or x.dtype != torch.float32 or x.dim() < 2 or not x.is_contiguous():
            raise L.AadError("scaler needs a contiguous float32 CUDA tensor [..., rows, W]")
        return x.numel() // x.shape[-1], x.shape[-1]

    def fit(self, x: torch.Tensor, group=None, n_frames: Optional[torch.Tensor] = None,
            status: Optional[torch.Tensor] = None) -> "DeviceStandardScaler":
        """x: this rank's features; with torch.distributed initialised the statistics are those of the
        union over all ranks of `group` (one all-reduce of 2*W + 1 doubles).  n_frames / status (int32 CUDA,
        [B]): x is a ragged [B, T, W] batch, only the valid rows of utterances with status 0 are counted."""
        import torch.distributed as dist
        lib = L.load()
        n_rows, W = self._rows(x)
        stats = torch.zeros(2 * W + 1, dtype=torch.float64, device=x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(x.device):
            if n_frames is not None:
                if x.dim() != 3 or n_frames.numel() != x.shape[0]:
                    raise L.AadError("ragged fit needs x [B, T, W] and n_frames [B]")
                nf = n_frames.to(device=x.device, dtype=torch.int32).contiguous()
                st = None if status is None else status.to(device=x.device, dtype=torch.int32).contiguous()
                L.check(lib.aad_scaler_accumulate_ragged(C.c_void_p(x.data_ptr()), x.shape[0], x.shape[1], W, W,
                                                         C.c_void_p(nf.data_ptr()),
                                                         C.c_void_p(st.data_ptr() if st is not None else 0),
                                                         C.c_void_p(stats.data_ptr()), C.c_void_p(stream)),
                        "aad_scaler_accumulate_ragged")
            else:
                L.check(lib.aad_scaler_accumulate(C.c_void_p(x.data_ptr()), n_rows, W, W, C.c_void_p(stats.data_ptr()),
                                                  C.c_void_p(stream)), "aad_scaler_accumulate")
                stats[2 * W] = float(n_rows)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
        h = stats.cpu().numpy()
        self.n_samples_seen_ = int(round(h[2 * W]))
        self.mean_, self.var_, self.scale_ = merge_stats(self.n_samples_seen_, h[:W].copy(), h[W:2 * W].copy())
        self._mean_d = torch.from_numpy(self.mean_.astype(np.float32)).to(x.device)
        self._inv_d = torch.from_numpy((1.0 / self.scale_).astype(np.float32)).to(x.device)
        return self

    def transform(self, x: torch.Tensor, inplace: bool = False) -> torch.Tensor:
        if self._mean_d is None:
            raise L.AadError("scaler is not fitted")
        lib = L.load()
        n_rows, W = self._rows(x)
        if W != self._mean_d.numel():
            raise L.AadError(f"scaler was fitted on {self._mean_d.numel()} columns, got {W}")
        out = x if inplace else x.clone()
        stream = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(x.device):
            L.check(lib.aad_scaler_apply(C.c_void_p(out.data_ptr()), n_rows, W, W, C.c_void_p(self._mean_d.data_ptr()),
                                         C.c_void_p(self._inv_d.data_ptr()), C.c_void_p(stream)), "aad_scaler_apply")
        return out

    def fit_transform(self, x: torch.Tensor, group=None, inplace: bool = False) -> torch.Tensor:
        return self.fit(x, group).transform(x, inplace)
