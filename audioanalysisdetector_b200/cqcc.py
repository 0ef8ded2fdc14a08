"""CQCC on the device (`aad_cqcc`): the extractor the reference's CNN-BiLSTM is trained on.

`extract_cqcc` (ASV_dl_func.py:442-481): librosa.cqt (hop 512, C1 upwards) -> |.| -> amplitude_to_db(ref=np.max)
-> per-frame linear interpolation onto a uniform frequency grid -> log(x^2 + 1e-12) -> DCT-II ortho -> (n_ceps, T).
The octave recursion of librosa.cqt is followed on the GPU; its soxr resampler is replaced by a documented
half-band FIR (csrc/aad_cqcc.cu), so parity with librosa is unpinned (oracle/cqcc_ref.py says what is and is not
established).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L
from .frontend import _ptr


class CqccFrontend:
    """One CQCC plan on one GPU: `(features [B, n_ceps, Tmax], n_frames, status) = fe(wav, lengths)`."""

    def __init__(self, sample_rate: int, bins_per_octave: int = 12, n_ceps: int = 19, device=None):
        if not torch.cuda.is_available():
            raise L.AadError("CqccFrontend needs a CUDA device; there is no CPU fallback")
        self.lib = L.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.sample_rate, self.bins_per_octave, self.n_ceps = int(sample_rate), int(bins_per_octave), int(n_ceps)
        h = C.c_void_p()
        L.check(self.lib.aad_cqcc_plan_create(self.sample_rate, self.bins_per_octave, self.n_ceps, self.device.index,
                                              C.byref(h)), "aad_cqcc_plan_create")
        self._h, self._ws = h, None
        self.launches_per_call = None

    def close(self):
        if getattr(self, "_h", None):
            self.lib.aad_cqcc_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def query(self, B: int, max_len: int):
        t, c, nb, ws = C.c_int32(), C.c_int32(), C.c_int32(), C.c_size_t()
        L.check(self.lib.aad_cqcc_query(self._h, B, max_len, C.byref(t), C.byref(c), C.byref(nb), C.byref(ws)),
                "aad_cqcc_query")
        return t.value, c.value, nb.value, ws.value

    def __call__(self, wav: torch.Tensor, lengths: Optional[torch.Tensor] = None, return_cqt: bool = False):
        """wav [B, Lmax] float32|int16 on this device -> (features [B, n_ceps, Tmax], n_frames, status[, |CQT|])."""
        if wav.dim() != 2 or not wav.is_cuda or wav.device != self.device:
            raise L.AadError(f"wav must be a 2-D tensor on {self.device}")
        if wav.stride(1) != 1:
            wav = wav.contiguous()
        B, Lmax = wav.shape
        if lengths is None:
            lengths = torch.full((B,), Lmax, dtype=torch.int32, device=self.device)
        return self._run(wav, wav.stride(0), None, lengths, B, Lmax, return_cqt)

    def extract_indexed(self, pcm: torch.Tensor, offsets: torch.Tensor, lengths: torch.Tensor, max_len: Optional[int] = None,
                        return_cqt: bool = False):
        """pcm: 1-D float32|int16 on this device (decoded files back to back); utterance b is
        pcm[offsets[b] : offsets[b] + lengths[b]] -- the chunk table of DeviceCorpus, as Frontend.extract_indexed."""
        if pcm.dim() != 1 or not pcm.is_cuda or pcm.device != self.device or not pcm.is_contiguous():
            raise L.AadError(f"pcm must be a contiguous 1-D tensor on {self.device}")
        offsets = offsets.to(device=self.device, dtype=torch.int64).contiguous()
        B = int(offsets.numel())
        Lmax = int(max_len) if max_len is not None else max(int(lengths.max()), 1)
        return self._run(pcm, 0, offsets, lengths, B, Lmax, return_cqt)

    def _run(self, wav, stride, offsets, lengths, B, Lmax, return_cqt):
        dt = L.F32 if wav.dtype == torch.float32 else (L.I16 if wav.dtype == torch.int16 else None)
        if dt is None:
            raise L.AadError("wav must be float32 or int16")
        lengths = lengths.to(device=self.device, dtype=torch.int32).contiguous()
        t_max, n_ceps, n_bins, ws_bytes = self.query(B, Lmax)
        t_alloc = max(t_max, 1)
        out = torch.zeros((B, n_ceps, t_alloc), dtype=torch.float32, device=self.device)
        mag = torch.zeros((B, n_bins, t_alloc), dtype=torch.float32, device=self.device) if return_cqt else None
        n_frames = torch.empty(B, dtype=torch.int32, device=self.device)
        status = torch.empty(B, dtype=torch.int32, device=self.device)
        if self._ws is None or self._ws.numel() < ws_bytes:
            self._ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = self.lib.aad_cqcc(self._h, _ptr(wav), dt, int(stride), _ptr(offsets), _ptr(lengths), B, Lmax, _ptr(out),
                                   out.stride(0), t_alloc, _ptr(n_frames), _ptr(status), _ptr(mag), _ptr(self._ws),
                                   self._ws.numel(), C.c_void_p(stream))
        L.check(rc, "aad_cqcc")
        return (out, n_frames, status, mag) if return_cqt else (out, n_frames, status)
