"""The consumer of the features on the device: inference of the reference's CNN-BiLSTM.

`DetectorEngine(state_dict)` takes the state dict of the reference's `AudioDeepfakeDetector`
(cnn_bilstm_hybrid.py:20-52; any mapping name -> tensor / ndarray with the reference's parameter names) and
runs its eval-mode forward pass (cnn_bilstm_hybrid.py:54-68) with the hand-written kernels of
`csrc/aad_detector.cu` on features that are already in device memory -- the (B, F, 63) tensor the front-end
emits goes in as it is: no DataFrame, no per-item `torch.tensor` (CQCCDataset, :4-15), no permutes.
Training is out of scope; torch is used for device memory and the stream only.
"""
from __future__ import annotations

import ctypes as C
from typing import Mapping

import numpy as np
import torch

from . import _lib as L

_NAMES = {
    "conv_w": "feature_extractor.0.weight", "conv_b": "feature_extractor.0.bias",
    "bn_w": "feature_extractor.1.weight", "bn_b": "feature_extractor.1.bias",
    "bn_mean": "feature_extractor.1.running_mean", "bn_var": "feature_extractor.1.running_var",
    "w_ih": "bilstm.weight_ih_l0", "w_hh": "bilstm.weight_hh_l0", "b_ih": "bilstm.bias_ih_l0", "b_hh": "bilstm.bias_hh_l0",
    "w_ih_r": "bilstm.weight_ih_l0_reverse", "w_hh_r": "bilstm.weight_hh_l0_reverse",
    "b_ih_r": "bilstm.bias_ih_l0_reverse", "b_hh_r": "bilstm.bias_hh_l0_reverse",
    "attn_w": "attention.weight", "attn_b": "attention.bias", "ln_w": "layer_norm.weight", "ln_b": "layer_norm.bias",
    "fc1_w": "classifier.0.weight", "fc1_b": "classifier.0.bias", "fc2_w": "classifier.3.weight", "fc2_b": "classifier.3.bias",
}
_SHAPES = {"conv_w": (64, 63, 3), "conv_b": (64,), "bn_w": (64,), "bn_b": (64,), "bn_mean": (64,), "bn_var": (64,),
           "w_ih": (128, 64), "w_hh": (128, 32), "b_ih": (128,), "b_hh": (128,),
           "w_ih_r": (128, 64), "w_hh_r": (128, 32), "b_ih_r": (128,), "b_hh_r": (128,),
           "attn_w": (1, 64), "attn_b": (1,), "ln_w": (1,), "ln_b": (1,),
           "fc1_w": (64, 64), "fc1_b": (64,), "fc2_w": (1, 64), "fc2_b": (1,)}


def pack_weights(state: Mapping[str, object], feature_dim: int):
    """state dict -> (ctypes struct, list of the float32 host arrays it points to).  Shapes are those of the
    reference's defaults (lstm_units 32, dense_units 64); anything else is rejected."""
    w = L.AadDetectorWeights()
    w.struct_size = C.sizeof(L.AadDetectorWeights)
    w.feature_dim = int(feature_dim)
    w.bn_eps = 1e-5
    keep = []
    for field, name in _NAMES.items():
        if name not in state:
            raise L.AadError(f"state dict has no '{name}'")
        v = state[name]
        a = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
        a = np.ascontiguousarray(a, dtype=np.float32)
        if tuple(a.shape) != _SHAPES[field]:
            raise L.AadError(f"'{name}' has shape {tuple(a.shape)}, the kernels are built for {_SHAPES[field]}")
        keep.append(a)
        setattr(w, field, a.ctypes.data_as(C.POINTER(C.c_float)))
    return w, keep


class DetectorEngine:
    def __init__(self, state: Mapping[str, object], feature_dim: int, device=None):
        self.lib = L.load()
        if not torch.cuda.is_available():
            raise L.AadError("DetectorEngine needs a CUDA device (there is no CPU path)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.feature_dim = int(feature_dim)
        w, keep = pack_weights(state, feature_dim)
        h = C.c_void_p()
        L.check(self.lib.aad_detector_create(C.byref(w), self.device.index or 0, C.byref(h)), "aad_detector_create")
        self._h = h
        self._ws = None
        self.launches_per_call = 3

    def close(self):
        if getattr(self, "_h", None):
            self.lib.aad_detector_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __call__(self, feats: torch.Tensor) -> torch.Tensor:
        """feats (B, F, T >= 63) float32 on this device, as the front-end returns it (only frames 0..62 of every
        row are read, as `x[:, :, :63]` would select) -> scores (B, 1) float32, the reference's output shape."""
        if (feats.dim() != 3 or not feats.is_cuda or feats.device != self.device or feats.dtype != torch.float32
                or feats.shape[1] != self.feature_dim or feats.shape[2] < 63 or feats.stride(2) != 1
                or feats.stride(1) < 63):
            raise L.AadError(f"feats must be a float32 (B, {self.feature_dim}, >= 63) tensor on {self.device} with unit frame stride")
        B = int(feats.shape[0])
        scores = torch.empty((B, 1), dtype=torch.float32, device=self.device)
        if B == 0:
            return scores
        ws = C.c_size_t()
        L.check(self.lib.aad_detector_query(self._h, B, C.byref(ws)), "aad_detector_query")
        if self._ws is None or self._ws.numel() < ws.value:
            self._ws = torch.empty(ws.value, dtype=torch.uint8, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = self.lib.aad_detector_forward(self._h, C.c_void_p(feats.data_ptr()), feats.stride(0), feats.stride(1), B,
                                               C.c_void_p(scores.data_ptr()), C.c_void_p(self._ws.data_ptr()),
                                               self._ws.numel(), C.c_void_p(stream))
        L.check(rc, "aad_detector_forward")
        return scores
