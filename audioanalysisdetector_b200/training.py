"""Training-side hand-off: the front-end's features as the batches the reference's training loops iterate over.

The reference wraps the DataFrame's object column in `CQCCDataset` (cnn_bilstm_hybrid.py:4-15: one
`torch.tensor(features[idx])` per item, label as a `(1,)` float tensor) and a `DataLoader(batch_size=200)`, and
`train_loop` (ASV_dl_func.py:751-829) moves every batch to the device (`X_batch.to(device)`, :762).  Here the
features already are one `(N, F, 63)` tensor in device memory (Frontend / DeviceCorpus output, optionally
standardised by DeviceStandardScaler), so a loader is an index permutation: `DeviceFeatureLoader` yields
`(X_batch, y_batch)` views / gathers on the device with the shapes and dtypes `train_loop` expects -- it can be
passed to the reference's `train_loop` unchanged as `train_loader` / `test_loader`.
"""
from __future__ import annotations

from typing import Iterator, Optional, Tuple

import torch


class DeviceFeatureLoader:
    """Iterable of `(X [n, F, T] float32, y [n, 1] float32)` batches over device-resident features.

    `status` (the extractor's per-item status): items with a non-zero status are dropped, as the reference drops the
    rows whose extractor returned None (`filtr_nan`, ASV_dl_func.py:1065-1071).  `label_shape="column"` gives the
    `(n, 1)` float labels of `CQCCDataset` (BCE with a sigmoid output); `"flat"` gives `(n,)` int64 labels for the
    cross-entropy models (`FeatureColumnDataset`, ASV_dl_func.py:691-706)."""

    def __init__(self, features: torch.Tensor, labels, batch_size: int = 200, shuffle: bool = False,
                 status: Optional[torch.Tensor] = None, seed: Optional[int] = None, drop_last: bool = False,
                 label_shape: str = "column"):
        if features.dim() < 2:
            raise ValueError("features must be [N, ...]")
        dev = features.device
        labels = torch.as_tensor(labels, device=dev)
        if labels.shape[0] != features.shape[0]:
            raise ValueError("one label per item")
        keep = torch.arange(features.shape[0], device=dev)
        if status is not None:
            keep = keep[status.to(dev) == 0]
        self.features, self.keep = features, keep
        if label_shape == "column":
            self.labels = labels.to(torch.float32).reshape(-1, 1)
        elif label_shape == "flat":
            self.labels = labels.to(torch.int64).reshape(-1)
        else:
            raise ValueError("label_shape is 'column' or 'flat'")
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), bool(shuffle), bool(drop_last)
        self.gen = None
        if seed is not None:
            self.gen = torch.Generator(device=dev)
            self.gen.manual_seed(int(seed))

    def __len__(self) -> int:
        n = int(self.keep.numel())
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    @property
    def dataset_size(self) -> int:
        return int(self.keep.numel())

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        idx = self.keep
        if self.shuffle:
            idx = idx[torch.randperm(idx.numel(), device=idx.device, generator=self.gen)]
        for k in range(len(self)):
            sel = idx[k * self.batch_size:(k + 1) * self.batch_size]
            yield self.features.index_select(0, sel), self.labels.index_select(0, sel)
