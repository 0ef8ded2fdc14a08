"""Minimal audio decode for the drop-in extractors (host side, outside the hot path).

The reference decodes with `librosa.load(filepath, sr=sr)` (ASV_dl_func.py:406,425,524):
float32 mono at the file's native rate when sr is None.  soundfile/librosa are not part
of this image: PCM WAV is decoded with the standard library, FLAC (the container of the
ASVspoof corpora) by the library's own decoder (`aad_flac_decode`, csrc/aad_flac.cpp; every
frame CRC-checked, the stream's MD5 verified here); other containers go through `soundfile`
when it is importable.  A custom loader can be installed with `set_loader(fn)` where
fn(path) -> (float32 mono ndarray, sample_rate).
"""
from __future__ import annotations

import ctypes as C
import hashlib
import wave
from typing import Callable, Optional, Tuple

import numpy as np

from . import _lib as L

_loader: Optional[Callable[[str], Tuple[np.ndarray, int]]] = None
_verify_flac_md5 = True


def set_flac_md5(verify: bool):
    """FLAC frames are always CRC-checked; the MD5 of the whole decoded stream (STREAMINFO) is checked on top unless
    switched off here (a quarter of the decode time of a 16-bit file; libsndfile, the reference's decoder, does not
    check it either)."""
    global _verify_flac_md5
    _verify_flac_md5 = bool(verify)


def set_loader(fn: Optional[Callable[[str], Tuple[np.ndarray, int]]]):
    global _loader
    _loader = fn


def custom_loader_installed() -> bool:
    return _loader is not None


def _load_wav(path: str, keep_pcm16: bool = False) -> Tuple[np.ndarray, int]:
    with wave.open(path, "rb") as w:
        sr, nch, sw, n = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
        raw = w.readframes(n)
    if sw == 2 and nch == 1 and keep_pcm16:
        return np.frombuffer(raw, dtype="<i2").copy(), sr    # sample value = int16 / 32768 (as decoded below)
    if sw == 2:
        y = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif sw == 4:
        y = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif sw == 1:
        y = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise ValueError(f"unsupported PCM sample width {sw}")
    if nch > 1:
        y = y.reshape(-1, nch).mean(axis=1).astype(np.float32)   # librosa.to_mono
    return y, sr


def decode_flac(data: bytes, verify_md5: bool = True) -> Tuple[np.ndarray, int, int]:
    """FLAC stream in memory -> (interleaved int32 samples [n, channels], sample_rate, bits_per_sample).
    (16-bit streams come back as an int32 VIEW-compatible array too; `_load_flac` narrows them.)"""
    lib = L.load()
    buf = C.cast(C.c_char_p(data), C.c_void_p)     # no copy: the bytes object outlives the calls below
    info = L.AadFlacInfo()
    L.check(lib.aad_flac_info(buf, len(data), C.byref(info)), "aad_flac_info")
    if info.total_samples <= 0:
        raise ValueError("FLAC stream without a sample count in STREAMINFO")
    out = np.empty((int(info.total_samples), int(info.channels)), dtype=np.int32)
    n = C.c_int64(0)
    L.check(lib.aad_flac_decode(buf, len(data), C.c_void_p(out.ctypes.data), int(info.total_samples), C.byref(n)),
            "aad_flac_decode")
    if n.value != info.total_samples:
        raise ValueError(f"FLAC stream ends after {n.value} of {info.total_samples} samples")
    md5 = bytes(info.md5)
    if verify_md5 and any(md5):
        width = (int(info.bits_per_sample) + 7) // 8
        if width == 2:
            raw = out.astype("<i2")                  # hashed through the buffer protocol (no second copy; GIL released)
        elif width == 4:
            raw = out.astype("<i4")
        else:
            raw = out.astype("<i4").view(np.uint8).reshape(-1, 4)[:, :width].tobytes()
        if hashlib.md5(raw).digest() != md5:
            raise ValueError("FLAC MD5 mismatch: decoded audio differs from what the encoder saw")
    return out, int(info.sample_rate), int(info.bits_per_sample)


def decode_flac_pcm16(data: bytes) -> Optional[Tuple[np.ndarray, int]]:
    """Mono 16-bit FLAC stream -> (int16 samples, sample_rate) in ONE library call (decode, narrowing and the MD5 check
    run without the interpreter lock: this is the call the corpus upload fans out over threads); None when the stream
    is not mono 16-bit."""
    lib = L.load()
    buf = C.cast(C.c_char_p(data), C.c_void_p)
    info = L.AadFlacInfo()
    L.check(lib.aad_flac_info(buf, len(data), C.byref(info)), "aad_flac_info")
    if info.channels != 1 or info.bits_per_sample != 16 or info.total_samples <= 0:
        return None
    out = np.empty(int(info.total_samples), dtype=np.int16)
    n, state = C.c_int64(0), C.c_int32(0)
    L.check(lib.aad_flac_decode_pcm16(buf, len(data), C.c_void_p(out.ctypes.data), out.size, C.byref(n),
                                      C.byref(state) if _verify_flac_md5 else None), "aad_flac_decode_pcm16")
    if n.value != info.total_samples:
        raise ValueError(f"FLAC stream ends after {n.value} of {info.total_samples} samples")
    if state.value < 0:
        raise ValueError("FLAC MD5 mismatch: decoded audio differs from what the encoder saw")
    return out, int(info.sample_rate)


def _load_flac(path: str, keep_pcm16: bool = False) -> Tuple[np.ndarray, int]:
    with open(path, "rb") as f:
        data = f.read()
    if keep_pcm16:
        got = decode_flac_pcm16(data)
        if got is not None:
            return got
    pcm, sr, bps = decode_flac(data, verify_md5=_verify_flac_md5)
    if keep_pcm16 and bps == 16 and pcm.shape[1] == 1:
        return pcm[:, 0].astype(np.int16), sr                 # sample value = int16 / 32768
    y = pcm.astype(np.float32) / np.float32(1 << (bps - 1))   # what libsndfile hands to librosa.load
    if pcm.shape[1] > 1:
        y = y.mean(axis=1).astype(np.float32)                 # librosa.to_mono
    else:
        y = y[:, 0]
    return np.ascontiguousarray(y), sr


def info_ex(source) -> Tuple[int, int, bool]:
    """(n_samples per channel, sample_rate, is mono 16-bit PCM) without decoding -- soundfile.info in
    prepare_dataframe (ASV_dl_func.py:280).  In-memory clips and custom loaders are measured by loading them."""
    if isinstance(source, tuple) and len(source) == 2:
        return len(source[0]), int(source[1]), np.asarray(source[0]).dtype == np.int16
    path = str(source)
    if _loader is None and path.lower().endswith(".wav"):
        with wave.open(path, "rb") as w:
            return w.getnframes(), w.getframerate(), (w.getsampwidth() == 2 and w.getnchannels() == 1)
    if _loader is None and path.lower().endswith(".flac"):
        with open(path, "rb") as f:
            head = f.read(1 << 16)
        fi = L.AadFlacInfo()
        lib = L.load()
        rc = lib.aad_flac_info(C.cast(C.c_char_p(head), C.c_void_p), len(head), C.byref(fi))
        if rc == 0 and fi.total_samples > 0:
            return int(fi.total_samples), int(fi.sample_rate), (fi.bits_per_sample == 16 and fi.channels == 1)
    y, sr = load_pcm(source) if _loader is None else load(source)
    return len(y), sr, y.dtype == np.int16


def info(source) -> Tuple[int, int]:
    """(n_samples per channel, sample_rate) without decoding."""
    return info_ex(source)[:2]


def load_pcm(source) -> Tuple[np.ndarray, int]:
    """Like `load` at the native rate, but mono 16-bit PCM (WAV or FLAC) comes back as the raw int16 samples
    (value = int16 / 32768, exactly what `load` would return as float32): half the bytes to upload."""
    if _loader is None and not isinstance(source, tuple):
        low = str(source).lower()
        if low.endswith(".wav"):
            return _load_wav(str(source), keep_pcm16=True)
        if low.endswith(".flac"):
            return _load_flac(str(source), keep_pcm16=True)
    return load(source)


def load(source, sr: Optional[int] = None) -> Tuple[np.ndarray, int]:
    """-> (float32 mono waveform, sample_rate).

    `source` is a file path, or an in-memory clip given as (waveform, sample_rate)."""
    if isinstance(source, tuple) and len(source) == 2:
        y, native = np.asarray(source[0], dtype=np.float32), int(source[1])
    elif _loader is not None:
        y, native = _loader(source)
        y = np.asarray(y, dtype=np.float32)
    else:
        path = str(source)
        if path.lower().endswith(".wav"):
            y, native = _load_wav(path)
        elif path.lower().endswith(".flac"):
            y, native = _load_flac(path)
        else:
            try:
                import soundfile as sf  # noqa: WPS433 (optional dependency)
            except ImportError as e:  # pragma: no cover
                raise RuntimeError(f"cannot decode {path}: soundfile is not installed") from e
            y, native = sf.read(path, dtype="float32", always_2d=False)
            if y.ndim > 1:
                y = y.mean(axis=1).astype(np.float32)
    if sr is not None and int(sr) != native:
        raise ValueError(f"resampling {native} -> {sr} Hz is outside the front-end (load at native rate)")
    return y, native
