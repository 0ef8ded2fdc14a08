"""Batched spectral front-end: PyTorch tensors in/out over the C ABI (include/aad.h).

`FrontendParams` mirrors `aad_params`; the three constructors reproduce the
parameters the reference's extractors pass to librosa / spafe:

  FrontendParams.logmel(sr, n_mels=64)  extract_mel_spectrogram  ASV_dl_func.py:522-538
  FrontendParams.mfcc(sr, n_mfcc=13)    extract_mfcc             ASV_dl_func.py:404-420
  FrontendParams.lfcc(sr, n_ceps=13)    extract_lfcc             ASV_dl_func.py:423-439

`Frontend` owns one plan (device tables) and runs batches on the caller's current
CUDA stream.  torch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib as L


@dataclass
class FrontendParams:
    kind: int = L.KIND_LOGMEL
    sample_rate: int = 16000
    n_fft: int = 2048
    win_length: int = 2048
    hop_length: int = 512
    window: int = L.WIN_HANN_PERIODIC
    center: bool = True
    quantize_i16: bool = False
    pre_emph: float = 0.0
    n_filt: int = 64
    fb_type: int = L.FB_MEL_SLANEY
    fmin: float = 0.0
    fmax: float = 0.0            # <= 0: sr / 2
    power_scale: float = 1.0
    log_type: int = L.LOG_DB10
    ref_type: int = L.REF_UTT_MAX
    amin: float = 1e-10
    top_db: float = 80.0         # < 0 disables
    n_ceps: int = 0
    n_delta: int = 0
    delta_width: int = 9
    layout: int = L.LAYOUT_CT
    time_mean: bool = False
    znorm: bool = False          # (x - mean) / std over the utterance's feature matrix (ASV_dataset.ipynb compute_melspec)
    i16_scale: float = 0.0       # int16 input: sample = int16 * i16_scale; 0 -> 1/32768 (log-mel / MFCC), 1 (LFCC)
    custom_fb: Optional[np.ndarray] = None   # (n_filt, n_fft//2+1) float32, FB_CUSTOM / FB_CUSTOM_DENSE only
    spectrum: int = L.SPEC_POWER             # SPEC_MAGNITUDE: filter bank on |X| (dense filter banks only)

    # ---- reference presets ---------------------------------------------------
    @classmethod
    def logmel(cls, sample_rate, n_mels=64, fmax=None, n_fft=2048, hop_length=512, **kw):
        """librosa.feature.melspectrogram(n_mels, fmax=fmax or sr/2) + power_to_db(ref=np.max)."""
        return cls(kind=L.KIND_LOGMEL, sample_rate=int(sample_rate), n_fft=n_fft, win_length=n_fft,
                   hop_length=hop_length, n_filt=n_mels, fmax=float(fmax or 0.0),
                   ref_type=L.REF_UTT_MAX, n_ceps=0, **kw)

    @classmethod
    def mfcc(cls, sample_rate, n_mfcc=13, n_mels=128, n_fft=2048, hop_length=512, **kw):
        """librosa.feature.mfcc(n_mfcc): 128 mels, power_to_db(ref=1.0, top_db=80), DCT-II ortho."""
        return cls(kind=L.KIND_MFCC, sample_rate=int(sample_rate), n_fft=n_fft, win_length=n_fft,
                   hop_length=hop_length, n_filt=n_mels, ref_type=L.REF_ONE, n_ceps=n_mfcc, **kw)

    @classmethod
    def lfcc(cls, sample_rate, n_ceps=13, nfilts=24, nfft=512, win_len=0.025, win_hop=0.01,
             pre_emph=0.97, quantize_i16=True, layout=L.LAYOUT_TC, fb_type=L.FB_LINEAR_CONT, **kw):
        """(y*32767).astype(int16) + spafe lfcc(num_ceps, nfilts=24, nfft=512, 25/10 ms hamming)."""
        sr = int(sample_rate)
        return cls(kind=L.KIND_LFCC, sample_rate=sr, n_fft=nfft, win_length=int(win_len * sr),
                   hop_length=int(win_hop * sr), window=L.WIN_HAMMING_SYMMETRIC, center=False,
                   quantize_i16=quantize_i16, pre_emph=pre_emph, n_filt=nfilts, fb_type=fb_type,
                   power_scale=1.0 / nfft, log_type=L.LOG_LN, ref_type=L.REF_ONE, top_db=-1.0,
                   n_ceps=n_ceps, layout=layout, **kw)

    @classmethod
    def gtcc(cls, sample_rate, n_ceps=13, nfilts=40, nfft=512, win_len=0.025, win_hop=0.01, pre_emph=0.97,
             layout=L.LAYOUT_TC, fb_type=L.FB_GAMMATONE, spectrum=L.SPEC_POWER, **kw):
        """spafe gfcc(sig=y, fs, num_ceps, nfilts) as extract_gtcc calls it (ASV_dl_func.py:484-499): float waveform,
        25/10 ms hamming, gammatone bank on |X|^2 / nfft, cube root, DCT-II ortho."""
        sr = int(sample_rate)
        return cls(kind=L.KIND_GTCC, sample_rate=sr, n_fft=nfft, win_length=int(win_len * sr),
                   hop_length=int(win_hop * sr), window=L.WIN_HAMMING_SYMMETRIC, center=False,
                   quantize_i16=False, pre_emph=pre_emph, n_filt=nfilts, fb_type=fb_type,
                   power_scale=(1.0 / nfft if spectrum == L.SPEC_POWER else 1.0), spectrum=spectrum,
                   log_type=L.LOG_CBRT, ref_type=L.REF_ONE, top_db=-1.0, n_ceps=n_ceps, layout=layout, **kw)

    def replace(self, **kw) -> "FrontendParams":
        return dataclasses.replace(self, **kw)

    def key(self):
        d = dataclasses.asdict(self)
        fb = d.pop("custom_fb")
        return tuple(sorted(d.items())) + ((None if fb is None else np.asarray(fb).tobytes()),)

    def to_c(self):
        p = L.AadParams()
        p.struct_size = C.sizeof(L.AadParams)
        for f in ("kind", "sample_rate", "n_fft", "win_length", "hop_length", "window", "n_filt",
                  "fb_type", "log_type", "ref_type", "n_ceps", "n_delta", "delta_width", "layout", "spectrum"):
            setattr(p, f, int(getattr(self, f)))
        p.center = int(bool(self.center))
        p.quantize_i16 = int(bool(self.quantize_i16))
        p.time_mean = int(bool(self.time_mean))
        p.znorm = int(bool(self.znorm))
        for f in ("pre_emph", "fmin", "fmax", "power_scale", "amin", "top_db", "i16_scale"):
            setattr(p, f, float(getattr(self, f)))
        keep = None
        if self.custom_fb is not None:
            keep = np.ascontiguousarray(self.custom_fb, dtype=np.float32)
            p.custom_fb = keep.ctypes.data_as(C.POINTER(C.c_float))
        return p, keep

    # ---- geometry (host mirror of the kernel's frame arithmetic) -------------
    def n_frames(self, length: int) -> int:
        if length <= 0:
            return 0
        if self.center:
            return 1 + length // self.hop_length
        return (length - self.win_length) // self.hop_length + 1 if length >= self.win_length else 0

    @property
    def c_out(self) -> int:
        return (self.n_ceps if self.n_ceps > 0 else self.n_filt) * (1 + self.n_delta)


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Frontend:
    """One plan on one GPU.  Thread-compatible: use one instance per host thread / stream."""

    def __init__(self, params: FrontendParams, device=None):
        if not torch.cuda.is_available():
            raise L.AadError("Frontend needs a CUDA device; there is no CPU fallback")
        self.lib = L.load()
        self.params = params
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise L.AadError("Frontend runs on CUDA devices only")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        cp, keep = params.to_c()
        h = C.c_void_p()
        L.check(self.lib.aad_plan_create(C.byref(cp), self.device.index, C.byref(h)), "aad_plan_create")
        self._h = h
        self._ws = None
        self.launches_per_call = int(self.lib.aad_plan_launches(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self.lib.aad_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- per-kernel device timing (bench roofline) --------------------------------
    def set_profiling(self, enable: bool):
        L.check(self.lib.aad_plan_set_profiling(self._h, int(enable)), "aad_plan_set_profiling")

    def kernel_times_ms(self):
        """{prepare, stft_fb, epilogue, time_mean} of the last call; synchronise first."""
        buf = (C.c_float * 4)()
        L.check(self.lib.aad_plan_kernel_times(self._h, buf), "aad_plan_kernel_times")
        return {"prepare": buf[0], "stft_fb": buf[1], "epilogue": buf[2], "time_mean": buf[3]}

    # ---- introspection --------------------------------------------------------
    def table(self, which: int) -> np.ndarray:
        n = int(self.lib.aad_plan_table(self._h, which, None, 0))
        if n < 0:
            L.check(n, "aad_plan_table")
        out = np.empty(n, dtype=np.float32)
        self.lib.aad_plan_table(self._h, which, out.ctypes.data_as(C.c_void_p), n)
        p = self.params
        if which == L.TABLE_FILTERBANK:
            return out.reshape(p.n_filt, p.n_fft // 2 + 1)
        if which == L.TABLE_DCT:
            return out.reshape(p.n_ceps, p.n_filt)
        if which == L.TABLE_DELTA_TAPS:
            return out.reshape(2, -1)
        return out

    def query(self, B: int, max_len: int) -> Tuple[int, int, int]:
        t, c, ws = C.c_int32(), C.c_int32(), C.c_size_t()
        L.check(self.lib.aad_query(self._h, B, max_len, C.byref(t), C.byref(c), C.byref(ws)), "aad_query")
        return t.value, c.value, ws.value

    def _workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=self.device)
        return self._ws

    # ---- the batched op ----------------------------------------------------------
    def __call__(self, wav: torch.Tensor, lengths: Optional[torch.Tensor] = None,
                 out: Optional[torch.Tensor] = None):
        """wav [B, Lmax] float32|int16 on this device, lengths [B] int32 (default: all Lmax).

        Returns (features, n_frames, status):
          features  CT: [B, C, Tmax]   TC: [B, Tmax, C]   time_mean: [B, C]   float32
          n_frames  [B] int32, status [B] int32 (0 = ok; rows with status != 0 are zeros)
        """
        if wav.dim() != 2 or not wav.is_cuda or wav.device != self.device:
            raise L.AadError(f"wav must be a 2-D tensor on {self.device}")
        if wav.dtype == torch.float32:
            dt = L.F32
        elif wav.dtype == torch.int16:
            dt = L.I16
        else:
            raise L.AadError("wav must be float32 or int16")
        if wav.stride(1) != 1:
            wav = wav.contiguous()
        B, Lmax = wav.shape
        if lengths is None:
            lengths = torch.full((B,), Lmax, dtype=torch.int32, device=self.device)
        elif lengths.dtype != torch.int32 or lengths.device != self.device or not lengths.is_contiguous():
            lengths = lengths.to(device=self.device, dtype=torch.int32).contiguous()
        t_max, c_out, ws_bytes = self.query(B, Lmax)
        t_alloc = max(t_max, 1)
        p = self.params
        if out is None:
            if p.time_mean:
                shape = (B, c_out)
            elif p.layout == L.LAYOUT_CT:
                shape = (B, c_out, t_alloc)
            else:
                shape = (B, t_alloc, c_out)
            out = torch.zeros(shape, dtype=torch.float32, device=self.device)
        else:
            # the C ABI takes a raw pointer and one batch stride: everything else must be dense
            if p.time_mean:
                shape = (B, c_out)
            elif p.layout == L.LAYOUT_CT:
                shape = (B, c_out, t_alloc)
            else:
                shape = (B, t_alloc, c_out)
            if (out.dtype != torch.float32 or out.device != self.device or tuple(out.shape) != shape
                    or not out[0].is_contiguous() or (B > 1 and out.stride(0) < out[0].numel())):
                raise L.AadError(f"out must be a float32 tensor of shape {shape} on {self.device} with dense rows")
        n_frames = torch.empty(B, dtype=torch.int32, device=self.device)
        status = torch.empty(B, dtype=torch.int32, device=self.device)
        ws = self._workspace(ws_bytes)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = self.lib.aad_extract(self._h, _ptr(wav), dt, wav.stride(0), _ptr(lengths), B, Lmax,
                                      _ptr(out), out.stride(0), t_alloc, _ptr(n_frames), _ptr(status),
                                      _ptr(ws), ws.numel(), C.c_void_p(stream))
        L.check(rc, "aad_extract")
        return out, n_frames, status

    # ---- repeated shapes: the whole call as one CUDA graph ------------------------------------------
    def graph(self, wav: torch.Tensor, lengths: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None):
        """Capture this call (K0 -> K1 -> K2, programmatic-dependent-launch edges included) into a CUDA graph for
        repeated batches of one shape, e.g. a serving loop on small batches where the three launches and the
        per-call tensor allocations are a large part of the step.  `wav` / `lengths` / `out` are the STATIC buffers
        of the graph: copy each new batch into them, then `replay()`.

        Returns an object with `.replay()`, `.out`, `.n_frames`, `.status`, `.wav`, `.lengths`."""
        B, Lmax = wav.shape
        if lengths is None:
            lengths = torch.full((B,), Lmax, dtype=torch.int32, device=self.device)
        if out is None:
            out = self(wav, lengths)[0]           # allocates the output (and grows the workspace) outside the capture
        else:
            self(wav, lengths, out=out)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.graph(g, stream=side):
            res = self(wav, lengths, out=out)

        class Graphed:
            pass
        r = Graphed()
        r.graph, r.replay = g, g.replay
        r.wav, r.lengths = wav, lengths
        r.out, r.n_frames, r.status = res
        return r

    # ---- one STFT, two features -------------------------------------------------------
    def extract_pair(self, other: "Frontend", wav: torch.Tensor, lengths: Optional[torch.Tensor] = None,
                     offsets: Optional[torch.Tensor] = None, max_len: Optional[int] = None):
        """This plan's features AND `other`'s (a plain log filter-bank plan over the same STFT, e.g. the 64-mel
        log-mel next to this MFCC plan) from ONE pass over the waveforms: `other`'s filter bank runs on the power
        spectra this plan computes, in the same kernel launch (`aad_extract_pair`).
        `wav` is a padded batch [B, Lmax] (+ `lengths`), or -- with `offsets` -- a 1-D buffer of decoded files and a
        chunk table (offsets int64, lengths int32, as in `extract_indexed`; the caller vouches for the table).
        Returns ((features, other_features), n_frames, status); other_features is [B, n_filt, Tmax]."""
        if not wav.is_cuda or wav.device != self.device or other.device != self.device:
            raise L.AadError(f"wav and both plans must be on {self.device}")
        dt = L.F32 if wav.dtype == torch.float32 else (L.I16 if wav.dtype == torch.int16 else None)
        if dt is None:
            raise L.AadError("wav must be float32 or int16")
        if offsets is None:
            if wav.dim() != 2:
                raise L.AadError("wav must be [B, Lmax] (or 1-D with a chunk table)")
            if wav.stride(1) != 1:
                wav = wav.contiguous()
            B, Lmax = wav.shape
            stride = wav.stride(0)
            if lengths is None:
                lengths = torch.full((B,), Lmax, dtype=torch.int32, device=self.device)
            off_ptr = None
        else:
            if wav.dim() != 1 or not wav.is_contiguous() or lengths is None:
                raise L.AadError("a chunk table needs a contiguous 1-D buffer, offsets and lengths")
            offsets = offsets.to(device=self.device, dtype=torch.int64).contiguous()
            B, stride = int(offsets.numel()), 0
            Lmax = int(max_len) if max_len is not None else max(int(lengths.max()), 1)
            off_ptr = _ptr(offsets)
        lengths = lengths.to(device=self.device, dtype=torch.int32).contiguous()
        t_max, c_out, ws_bytes = self.query(B, Lmax)
        _, c2, ws2_bytes = other.query(B, Lmax)
        t_alloc = max(t_max, 1)
        p = self.params
        shape = (B, c_out) if p.time_mean else ((B, c_out, t_alloc) if p.layout == L.LAYOUT_CT else (B, t_alloc, c_out))
        out = torch.zeros(shape, dtype=torch.float32, device=self.device)
        out2 = torch.zeros((B, c2, t_alloc), dtype=torch.float32, device=self.device)
        n_frames = torch.empty(B, dtype=torch.int32, device=self.device)
        status = torch.empty(B, dtype=torch.int32, device=self.device)
        ws, ws2 = self._workspace(ws_bytes), other._workspace(ws2_bytes)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = self.lib.aad_extract_pair(self._h, other._h, _ptr(wav), dt, stride, off_ptr, _ptr(lengths), B, Lmax,
                                           _ptr(out), out.stride(0), _ptr(out2), out2.stride(0), t_alloc, _ptr(n_frames),
                                           _ptr(status), _ptr(ws), ws.numel(), _ptr(ws2), ws2.numel(), C.c_void_p(stream))
        L.check(rc, "aad_extract_pair")
        return (out, out2), n_frames, status

    # ---- chunks of decoded files that already sit in device memory ------------------
    def extract_indexed(self, pcm: torch.Tensor, offsets: torch.Tensor, lengths: torch.Tensor,
                        max_len: Optional[int] = None, validate: bool = True):
        """pcm: 1-D float32|int16 on this device (decoded files back to back); utterance b is
        pcm[offsets[b] : offsets[b] + lengths[b]] (offsets int64, lengths int32; chunks may overlap).
        Same returns as __call__.  The reference slices `y[start_sample:end_sample]` after decoding the
        whole file once per chunk (ASV_dl_func.py:406-411); here the slice is an offset in a table.
        `max_len` (default lengths.max(), which costs a device->host read) bounds the output width;
        `validate=False` skips the bounds check of the table (another device->host read)."""
        if pcm.dim() != 1 or not pcm.is_cuda or pcm.device != self.device or not pcm.is_contiguous():
            raise L.AadError(f"pcm must be a contiguous 1-D tensor on {self.device}")
        if pcm.dtype == torch.float32:
            dt = L.F32
        elif pcm.dtype == torch.int16:
            dt = L.I16
        else:
            raise L.AadError("pcm must be float32 or int16")
        offsets = offsets.to(device=self.device, dtype=torch.int64).contiguous()
        lengths = lengths.to(device=self.device, dtype=torch.int32).contiguous()
        B = int(offsets.numel())
        if B == 0 or lengths.numel() != B:
            raise L.AadError("offsets and lengths must be non-empty and of equal size")
        if validate:   # a device->host read: every chunk inside the buffer (skip it for tables built on the host)
            lo, hi = int(offsets.min()), int((offsets + lengths.clamp(min=0).to(torch.int64)).max())
            if lo < 0 or hi > pcm.numel():
                raise L.AadError(f"chunk table reaches outside the pcm buffer ([{lo}, {hi}) vs {pcm.numel()} samples)")
        Lmax = int(max_len) if max_len is not None else max(int(lengths.max()), 1)
        t_max, c_out, ws_bytes = self.query(B, Lmax)
        t_alloc = max(t_max, 1)
        p = self.params
        if p.time_mean:
            shape = (B, c_out)
        elif p.layout == L.LAYOUT_CT:
            shape = (B, c_out, t_alloc)
        else:
            shape = (B, t_alloc, c_out)
        out = torch.zeros(shape, dtype=torch.float32, device=self.device)
        n_frames = torch.empty(B, dtype=torch.int32, device=self.device)
        status = torch.empty(B, dtype=torch.int32, device=self.device)
        ws = self._workspace(ws_bytes)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = self.lib.aad_extract_indexed(self._h, _ptr(pcm), dt, _ptr(offsets), _ptr(lengths), B, Lmax,
                                              _ptr(out), out.stride(0), t_alloc, _ptr(n_frames), _ptr(status),
                                              _ptr(ws), ws.numel(), C.c_void_p(stream))
        L.check(rc, "aad_extract_indexed")
        return out, n_frames, status

    # ---- host buffers in / out (pipelined H2D -> kernels -> D2H inside the library) ----
    def reserve_host(self, B: int, max_len: int, dtype=np.float32, chunk_utts: int = 0):
        """Allocate the host path's internal buffers ahead of the first `extract_host` of that shape
        (aad_host_reserve): the timed / latency-critical calls then allocate nothing."""
        t_max, _, _ = self.query(B, max_len)
        dt = L.I16 if np.dtype(dtype) == np.int16 else L.F32
        L.check(self.lib.aad_host_reserve(self._h, dt, int(B), int(max_len), max(t_max, 1), int(chunk_utts)),
                "aad_host_reserve")

    def extract_host(self, wav: np.ndarray, lengths: Optional[np.ndarray] = None,
                     out: Optional[np.ndarray] = None, chunk_utts: int = 0):
        """wav [B, Lmax] float32|int16 HOST array (numpy, or a pinned torch CPU tensor's .numpy())."""
        if isinstance(wav, torch.Tensor):
            wav = wav.numpy()
        if wav.ndim != 2 or wav.strides[1] != wav.itemsize:
            wav = np.ascontiguousarray(wav)
        if wav.dtype == np.float32:
            dt = L.F32
        elif wav.dtype == np.int16:
            dt = L.I16
        else:
            raise L.AadError("wav must be float32 or int16")
        B, Lmax = wav.shape
        if lengths is None:
            lengths = np.full(B, Lmax, dtype=np.int32)
        lengths = np.ascontiguousarray(lengths, dtype=np.int32)
        t_max, c_out, _ = self.query(B, Lmax)
        t_alloc = max(t_max, 1)
        p = self.params
        if p.time_mean:
            shape = (B, c_out)
        elif p.layout == L.LAYOUT_CT:
            shape = (B, c_out, t_alloc)
        else:
            shape = (B, t_alloc, c_out)
        if out is None:
            out = np.zeros(shape, dtype=np.float32)
        else:
            if isinstance(out, torch.Tensor):
                out = out.numpy()
            # the library writes whole rows through this pointer: shape, dtype and row contiguity must be right
            if out.dtype != np.float32 or tuple(out.shape) != shape or out[0].strides != np.empty(shape[1:], np.float32).strides:
                raise L.AadError(f"out must be float32 of shape {shape} with C-contiguous rows, got {out.dtype} {tuple(out.shape)}")
        n_frames = np.empty(B, dtype=np.int32)
        status = np.empty(B, dtype=np.int32)
        rc = self.lib.aad_extract_host(
            self._h, C.c_void_p(wav.ctypes.data), dt, wav.strides[0] // wav.itemsize,
            C.c_void_p(lengths.ctypes.data), B, Lmax, C.c_void_p(out.ctypes.data),
            out.strides[0] // 4, t_alloc, C.c_void_p(n_frames.ctypes.data),
            C.c_void_p(status.ctypes.data), int(chunk_utts))
        L.check(rc, "aad_extract_host")
        return out, n_frames, status


class _PinnedBlock:
    """Owner of one cudaHostAlloc block (freed when the last numpy view of it is gone)."""

    def __init__(self, nbytes: int, write_combined: bool):
        self.lib = L.load()
        p = C.c_void_p()
        L.check(self.lib.aad_host_alloc(C.byref(p), max(int(nbytes), 1), int(bool(write_combined))), "aad_host_alloc")
        self.ptr, self.nbytes = p.value, max(int(nbytes), 1)
        self.buf = (C.c_char * self.nbytes).from_address(self.ptr)

    def __del__(self):
        try:
            if getattr(self, "ptr", None):
                self.lib.aad_host_free(C.c_void_p(self.ptr))
                self.ptr = None
        except Exception:
            pass


class PinnedArray(np.ndarray):
    """numpy array over pinned host memory from aad_host_alloc (the block lives as long as any view of it)."""
    _block = None

    def __array_finalize__(self, obj):
        if obj is not None:
            self._block = getattr(obj, "_block", None)


def pinned_empty(shape, dtype=np.float32, write_combined: bool = False) -> np.ndarray:
    """Pinned host array for `Frontend.extract_host` (aad_host_alloc).  `write_combined=True` suits the INPUT
    staging buffer of a corpus (CPU writes it once, the GPUs read it: no cache snooping on the way out); never
    use it for buffers the CPU reads back."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    blk = _PinnedBlock(n, write_combined)
    arr = np.frombuffer(blk.buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape).view(PinnedArray)
    arr._block = blk
    return arr


def delta(x: torch.Tensor, n_frames: Optional[torch.Tensor] = None, width: int = 9, order: int = 1):
    """librosa.feature.delta(x, width, order, axis=-1, mode='interp') on [B, C, T] (CUDA, float32)."""
    lib = L.load()
    if x.dim() == 2:
        return delta(x.unsqueeze(0), n_frames, width, order)[0]
    if not x.is_cuda or x.dtype != torch.float32:
        raise L.AadError("delta needs a CUDA float32 tensor")
    x = x.contiguous()
    B, Cc, T = x.shape
    if n_frames is None:
        n_frames = torch.full((B,), T, dtype=torch.int32, device=x.device)
    n_frames = n_frames.to(device=x.device, dtype=torch.int32).contiguous()
    out = torch.zeros_like(x)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    with torch.cuda.device(x.device):
        rc = lib.aad_delta(_ptr(x), _ptr(n_frames), B, Cc, T, width, order, _ptr(out), C.c_void_p(stream))
    L.check(rc, "aad_delta")
    return out


def fp32_peak_tflops(device: int = 0, iters: int = 4096) -> float:
    v = C.c_double()
    L.check(L.load().aad_fp32_peak(device, iters, C.byref(v)), "aad_fp32_peak")
    return v.value
