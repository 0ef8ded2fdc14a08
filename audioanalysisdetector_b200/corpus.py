"""Decode once, upload once, slice on the device: the input side of the extractors.

In the reference every 2-second chunk row (`prepare_dataframe`, ASV_dl_func.py:287-293) calls
`librosa.load(filepath)` on its whole file and then slices `y[start_sample:end_sample]`
(ASV_dl_func.py:406-411, 425-429, 524-528), once per feature -- file decode dominates its wall time
(SURVEY.md section 6).  `DeviceCorpus` decodes each distinct file once (16-bit PCM, WAV or FLAC, stays int16: half the
PCIe bytes), lays all files back to back in ONE device buffer -- streamed up through two pinned staging
buffers, the decode of one segment overlapping the H2D copy of the previous one, so the host never holds the
whole corpus -- and hands the kernels a chunk table (element offset + length per row) instead of padded
copies: `aad_extract_indexed` (include/aad.h).  Noise augmentation (`augment_audio(mode="noise")`, ASV_dl_func.py:78-93) is applied on
the device to the rows that ask for it.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import _lib as L
from . import audio_io
from .frontend import Frontend

FILE_ALIGN = 8          # elements: every file starts on a 16-byte (int16) / 32-byte (float32) boundary
STAGE_BYTES = 64 << 20  # each of the two pinned staging buffers of the streamed upload
NOISE_FACTOR = 1.022    # augment_audio's default factor for mode="noise" (ASV_dl_func.py:85-86)


def _isnan(v) -> bool:
    return isinstance(v, float) and math.isnan(v)


def chunk_bounds(n_samples: int, sr: int, chunk_start, chunk_end) -> Tuple[int, int]:
    """[start, end) of `y[start_sample:end_sample]` exactly as the extractors compute it
    (ASV_dl_func.py:407-410): start = int(chunk_start * sr), end = min(int(chunk_end * sr), len(y)),
    then Python slice semantics (a start past the end gives an empty clip).  No chunk -> whole file."""
    if chunk_start is None or chunk_end is None or _isnan(chunk_start) or _isnan(chunk_end):
        return 0, n_samples
    start = int(chunk_start * sr)
    end = min(int(chunk_end * sr), n_samples)
    start, end, _ = slice(start, end).indices(n_samples)
    return start, max(end, start)


def two_second_chunks(n_samples: int, sr: int, chunk_s: float = 2.0) -> List[Tuple[float, float]]:
    """The chunk rows `prepare_dataframe` makes for one file (ASV_dl_func.py:281-293): files shorter
    than one chunk are dropped, the tail past the last full chunk is ignored."""
    duration = n_samples / sr
    if duration < chunk_s:
        return []
    return [(i * chunk_s, (i + 1) * chunk_s) for i in range(int(duration // chunk_s))]


class DeviceCorpus:
    """Decoded files in device memory + chunk tables over them."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise L.AadError("DeviceCorpus needs a CUDA device (there is no CPU path)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._index: Dict[object, int] = {}
        self._host: List[np.ndarray] = []
        self._keep: list = []
        self.sample_rates: List[int] = []
        self.n_samples: List[int] = []
        self._is_i16: List[bool] = []
        self.base: List[int] = []            # element offset of every file inside the device buffer
        self.pcm: Optional[torch.Tensor] = None
        self._pcm_f32: Optional[torch.Tensor] = None
        self.h2d_bytes = 0

    # ---- building -----------------------------------------------------------------
    def add(self, source) -> int:
        """Register `source` (path, or an in-memory (waveform, sr) pair) unless it is already here.  Files are only
        measured now (`audio_io.info`: header read, as soundfile.info in ASV_dl_func.py:280) and decoded while the
        corpus is streamed to the device."""
        if isinstance(source, tuple) and len(source) == 2:
            key = (id(source[0]), int(source[1]))     # the waveform object identifies an in-memory clip
        else:
            key = str(source)
        idx = self._index.get(key)
        if idx is None:
            self._keep.append(source)                 # ids stay unique while the corpus lives
            if isinstance(source, tuple):
                y, sr = audio_io.load_pcm(source)
                y = y if y.dtype == np.int16 else np.ascontiguousarray(y, dtype=np.float32)
                n, i16 = len(y), y.dtype == np.int16
            else:
                y = None                              # decoded in upload()
                n, sr, i16 = audio_io.info_ex(source)
            idx = self._index[key] = len(self._host)
            self._host.append(y)
            self.sample_rates.append(int(sr))
            self.n_samples.append(int(n))
            self._is_i16.append(bool(i16))
            self.pcm = None
        return idx

    def _decoded(self, i: int) -> np.ndarray:
        y = self._host[i]
        if y is None:
            y, sr = audio_io.load_pcm(self._keep[i])
            if len(y) != self.n_samples[i] or int(sr) != self.sample_rates[i]:
                raise L.AadError(f"{self._keep[i]}: header says {self.n_samples[i]} samples @ {self.sample_rates[i]} Hz, "
                                 f"decoded {len(y)} @ {sr}")
        return y

    def upload(self, dtype: Optional[torch.dtype] = None, stage_bytes: int = STAGE_BYTES,
               decode_threads: Optional[int] = None) -> torch.Tensor:
        """Stream the corpus to the device: files are decoded into one of two pinned staging buffers while the other
        one's H2D copy is in flight (copy stream + events), segment after segment.  int16 when every file is 16-bit
        PCM (or `dtype=torch.int16` is forced), else float32 (int16 / 32768).
        Decoding runs `decode_threads` files ahead on a thread pool (default: the host's cores, at most 16): the FLAC
        decoder and the MD5 check release the GIL, and decoding -- not the copy, not the kernels -- is what a corpus
        of files costs (1.4 ms per 4-second clip and core against 24 ms of GPU time for 25 380 chunks)."""
        if self.pcm is not None and (dtype is None or self.pcm.dtype == dtype):
            return self.pcm
        if not self._host:
            raise L.AadError("empty corpus")
        self.base, total = layout_files(self.n_samples)
        if dtype is None:
            dtype = torch.int16 if all(self._is_i16) else torch.float32
        esz = 2 if dtype == torch.int16 else 4
        seg = max(int(stage_bytes) // esz // FILE_ALIGN * FILE_ALIGN, FILE_ALIGN)
        dev = torch.empty(total, dtype=dtype, device=self.device)
        stages = [torch.empty(seg, dtype=dtype, pin_memory=True) for _ in range(2 if total > seg else 1)]
        views = [st.numpy() for st in stages]
        events = [None] * len(stages)
        copy_stream = torch.cuda.Stream(self.device)
        n_thr = min(16, os.cpu_count() or 1) if decode_threads is None else max(0, int(decode_threads))
        if audio_io.custom_loader_installed():
            n_thr = 0                                  # a caller's loader is not known to be thread-safe
        pool = ThreadPoolExecutor(n_thr) if n_thr > 1 and any(y is None for y in self._host) else None
        pending: Dict[int, object] = {}                # file index -> future of its decoded samples
        ahead = 0                                      # files [0, ahead) have been handed to the pool

        def decoded(i: int) -> np.ndarray:
            nonlocal ahead
            if pool is None:
                return self._decoded(i)
            while ahead < len(self._host) and ahead < i + 4 * n_thr:
                if self._host[ahead] is None:
                    pending[ahead] = pool.submit(self._decoded, ahead)
                ahead += 1
            fut = pending.pop(i, None)
            return self._decoded(i) if fut is None else fut.result()

        try:
            dev = self._stream_segments(total, seg, dtype, dev, stages, views, events, copy_stream, decoded)
        finally:
            if pool is not None:
                pool.shutdown(wait=True, cancel_futures=True)
        torch.cuda.current_stream(self.device).wait_stream(copy_stream)
        copy_stream.synchronize()                      # the staging buffers are released below
        self.pcm = dev
        self.h2d_bytes = total * esz
        self._pcm_f32 = None
        return self.pcm

    def _stream_segments(self, total, seg, dtype, dev, stages, views, events, copy_stream, decoded):
        fi, fpos, cur = 0, 0, None
        for k, s0 in enumerate(range(0, total, seg)):
            n = min(seg, total - s0)
            b = k % len(stages)
            if events[b] is not None:
                events[b].synchronize()                # the copy that last read this staging buffer is done
            sv = views[b]
            sv[:n] = 0
            while fi < len(self.base) and self.base[fi] < s0 + n:
                if cur is None:
                    cur = decoded(fi)
                    if dtype == torch.int16 and cur.dtype != np.int16:
                        raise L.AadError("int16 upload of a corpus with non-16-bit files")
                    if dtype == torch.float32 and cur.dtype == np.int16:
                        cur = cur.astype(np.float32) / np.float32(32768.0)
                d0 = self.base[fi] + fpos - s0          # destination inside this segment
                take = min(len(cur) - fpos, n - d0)
                sv[d0:d0 + take] = cur[fpos:fpos + take]
                fpos += take
                if fpos < len(cur):
                    break                              # the file continues in the next segment
                fi, fpos, cur = fi + 1, 0, None
            with torch.cuda.stream(copy_stream):
                dev[s0:s0 + n].copy_(stages[b][:n], non_blocking=True)
                events[b] = torch.cuda.Event()
                events[b].record(copy_stream)
        return dev

    def as_float32(self) -> torch.Tensor:
        """The corpus as librosa.load would return it (int16 / 32768 is exact in float32)."""
        pcm = self.upload()
        if pcm.dtype == torch.float32:
            return pcm
        if self._pcm_f32 is None:
            self._pcm_f32 = pcm.to(torch.float32) * (1.0 / 32768.0)
        return self._pcm_f32

    # ---- chunk tables ---------------------------------------------------------------
    def table(self, rows: Sequence[Tuple[int, object, object]]) -> Tuple[np.ndarray, np.ndarray]:
        """rows of (file index, chunk_start s, chunk_end s) -> (element offsets int64, lengths int32)."""
        if not self.base:
            self.base, _ = layout_files(self.n_samples)
        off = np.empty(len(rows), dtype=np.int64)
        ln = np.empty(len(rows), dtype=np.int32)
        for i, (f, cs, ce) in enumerate(rows):
            s, e = chunk_bounds(self.n_samples[f], self.sample_rates[f], cs, ce)
            off[i], ln[i] = self.base[f] + s, e - s
        return off, ln

    # ---- extraction -----------------------------------------------------------------
    def extract(self, fe: Frontend, offsets: np.ndarray, lengths: np.ndarray,
                noise_rows: Optional[Sequence[int]] = None, generator: Optional[torch.Generator] = None):
        """Features of every chunk of the table, on the device: (features, n_frames, status).

        `noise_rows`: indices of rows with augmentationType == "noise": they become
        `y + 1.022 * randn(len(y))` in float32 (ASV_dl_func.py:84-89; the reference draws the noise from
        numpy's unseeded global generator, so only its distribution can be reproduced)."""
        pcm = self.upload()
        if fe.params.quantize_i16 or (noise_rows is not None and len(noise_rows)):
            pcm = self.as_float32()   # LFCC re-quantises the decoded float (y*32767 -> int16) itself
        off = torch.from_numpy(np.ascontiguousarray(offsets, dtype=np.int64)).to(self.device)
        ln = torch.from_numpy(np.ascontiguousarray(lengths, dtype=np.int32)).to(self.device)
        if noise_rows is not None and len(noise_rows):
            pcm, off = self._with_noise(pcm, off, ln, np.asarray(noise_rows, dtype=np.int64), generator)
        return fe.extract_indexed(pcm, off, ln, max_len=max(int(np.max(lengths)), 1))

    def extract_pair(self, fe: Frontend, fe2: Frontend, offsets: np.ndarray, lengths: np.ndarray):
        """Two features of every chunk from ONE STFT (`Frontend.extract_pair` over the chunk table):
        ((features, features2), n_frames, status).  Both plans must read the corpus the same way (no
        re-quantising LFCC plan here) and no row may ask for augmentation."""
        if fe.params.quantize_i16 or fe2.params.quantize_i16:
            raise L.AadError("paired extraction is for plans over the same STFT (mel-type plans)")
        pcm = self.upload()
        off = torch.from_numpy(np.ascontiguousarray(offsets, dtype=np.int64)).to(self.device)
        ln = torch.from_numpy(np.ascontiguousarray(lengths, dtype=np.int32)).to(self.device)
        return fe.extract_pair(fe2, pcm, ln, offsets=off, max_len=max(int(np.max(lengths)), 1))

    def _with_noise(self, pcm, off, ln, rows, generator):
        """Append noisy copies of the chosen chunks behind the corpus and point their rows at them."""
        r = torch.from_numpy(rows).to(self.device)
        rl = ln[r].to(torch.int64)
        padded = (rl + FILE_ALIGN - 1) // FILE_ALIGN * FILE_ALIGN
        dst0 = torch.cumsum(padded, 0) - padded                       # start of every copy in the appendix
        total = int(padded.sum())
        ext = torch.zeros(pcm.numel() + total, dtype=torch.float32, device=self.device)
        ext[:pcm.numel()] = pcm
        if int(rl.sum()) > 0:
            which = torch.repeat_interleave(torch.arange(len(rows), device=self.device), rl)
            within = torch.arange(int(rl.sum()), device=self.device) - torch.repeat_interleave(torch.cumsum(rl, 0) - rl, rl)
            src = off[r][which] + within
            noise = torch.randn(src.numel(), dtype=torch.float32, device=self.device, generator=generator)
            ext[pcm.numel() + dst0[which] + within] = pcm[src] + NOISE_FACTOR * noise
        off = off.clone()
        off[r] = pcm.numel() + dst0
        return ext, off


def layout_files(n_samples: Sequence[int]) -> Tuple[List[int], int]:
    """Element offset of every file (aligned to FILE_ALIGN) and the total buffer size."""
    base, pos = [], 0
    for n in n_samples:
        base.append(pos)
        pos += (int(n) + FILE_ALIGN - 1) // FILE_ALIGN * FILE_ALIGN
    return base, max(pos, FILE_ALIGN)
