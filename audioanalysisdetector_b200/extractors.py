"""Drop-in replacements for the reference's feature extractors and dispatcher.

Same names, arguments, return shapes/dtypes and error convention as
  extract_mel_spectrogram  ASV_dl_func.py:522-538  -> (n_mels, T) float32
  extract_mfcc             ASV_dl_func.py:404-420  -> (n_mfcc, T) float32
  extract_lfcc             ASV_dl_func.py:423-439  -> (T, n_ceps) float64
  extract_features         ASV_dl_func.py:1031-1049
but the arithmetic runs in the sm_100a kernels of libaad_b200.so.  The per-file
functions are thin B=1 wrappers over the batched op; `extract_features` recognises
them in the extractor map and runs ONE batched GPU call per feature instead of a
joblib process pool (a CUDA context per loky worker is exactly what the GPU path must
avoid).  Any other callable in the map is called per row, as the reference does.

On any per-item failure the reference prints "[BŁĄD ...]" and returns None; so do we.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from . import audio_io
from .frontend import Frontend, FrontendParams

_plans: Dict[tuple, Frontend] = {}


def get_frontend(params: FrontendParams, device=None) -> Frontend:
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    key = (params.key(), str(dev))
    fe = _plans.get(key)
    if fe is None:
        fe = _plans[key] = Frontend(params, dev)
    return fe


def augment_audio(data, sr, mode="change pitch", factor=None):
    """ASV_dl_func.py:78-93.  'noise' is reproduced; 'change pitch' needs
    librosa.effects.pitch_shift (resampling + phase vocoder), which is outside the
    front-end: it raises, and the caller's error convention turns that into None."""
    if mode == "change pitch":
        raise NotImplementedError("pitch-shift augmentation is outside the spectral front-end")
    if mode == "noise":
        if factor is None:
            factor = 1.022
        noise = np.random.randn(len(data))
        return (data + factor * noise).astype(data.dtype), sr
    return data, sr


def _prepare_clip(filepath, chunk_start, chunk_end, sr, augment):
    y, sr = audio_io.load(filepath, sr=sr)
    if chunk_start is not None and chunk_end is not None and not (
            isinstance(chunk_start, float) and math.isnan(chunk_start)):
        start_sample = int(chunk_start * sr)
        end_sample = min(int(chunk_end * sr), len(y))
        y = y[start_sample:end_sample]
    if augment is not None and not (isinstance(augment, float) and math.isnan(augment)):
        y, sr = augment_audio(y, sr, mode=augment)
    return np.ascontiguousarray(y, dtype=np.float32), sr


_STAGE: dict = {}   # per host thread: one pinned staging buffer, grown geometrically (a cudaHostAlloc per call
#                      would dominate the per-file `func(path)` form of the reference's extractors)


def _pinned_stage(n_elems: int) -> torch.Tensor:
    import threading
    key = threading.get_ident()
    buf = _STAGE.get(key)
    if buf is None or buf.numel() < n_elems:
        buf = torch.empty(max(n_elems, 2 * (buf.numel() if buf is not None else 0), 1 << 16), dtype=torch.float32,
                          pin_memory=True)
        _STAGE[key] = buf
    return buf


def _run_batch(params: FrontendParams, clips: Sequence[np.ndarray]):
    """Pad clips to one [B, Lmax] batch, run the plan, return per-clip arrays (None on item error)."""
    fe = get_frontend(params)
    B = len(clips)
    lens = np.array([len(c) for c in clips], dtype=np.int32)
    Lmax = max(int(lens.max()), 1)
    Lmax = (Lmax + 3) // 4 * 4          # keeps rows 16-byte aligned for the vector loads
    host = _pinned_stage(B * Lmax)[:B * Lmax].view(B, Lmax)
    hv = host.numpy()
    for i, c in enumerate(clips):
        hv[i, :len(c)] = c
        hv[i, len(c):] = 0.0
    wav = host.to(fe.device, non_blocking=True)
    feats, n_frames, status = fe(wav, torch.from_numpy(lens).to(fe.device))
    feats, n_frames, status = feats.cpu().numpy(), n_frames.cpu().numpy(), status.cpu().numpy()   # syncs: the stage is free again
    out: List[Optional[np.ndarray]] = []
    for i in range(B):
        if status[i] != 0:
            out.append(None)
        elif params.time_mean:
            out.append(feats[i].copy())
        elif params.layout == L.LAYOUT_CT:
            out.append(feats[i, :, :n_frames[i]].copy())
        else:
            out.append(feats[i, :n_frames[i], :].copy())
    return out, status


# --------------------------------------------------------------------------- params per call
def _mel_params(sr, n_mels, fmax, mean):
    return FrontendParams.logmel(sr, n_mels=n_mels, fmax=fmax or sr / 2, time_mean=bool(mean))


def _mfcc_params(sr, n_mfcc, mean):
    return FrontendParams.mfcc(sr, n_mfcc=n_mfcc, time_mean=bool(mean))


def _lfcc_params(sr, n_ceps):
    return FrontendParams.lfcc(sr, n_ceps=n_ceps)


def _lfcc_post(x, mean):
    x = x.astype(np.float64)                       # spafe returns float64 (T, n_ceps)
    return np.mean(x, axis=1) if mean else x       # ASV_dl_func.py:436 averages axis=1


# --------------------------------------------------------------------------- per-file drop-ins
def extract_mel_spectrogram(filepath, chunk_start=None, chunk_end=None, sr=None, n_mels=64, fmax=None,
                            mean=False, augment=None):
    try:
        y, sr = _prepare_clip(filepath, chunk_start, chunk_end, sr, augment)
        out, status = _run_batch(_mel_params(sr, n_mels, fmax, mean), [y])
        if out[0] is None:
            raise ValueError(L.ITEM_STATUS_NAMES.get(int(status[0]), "item failed"))
        return out[0]
    except Exception as e:
        print(f"[BŁĄD MEL] {filepath if isinstance(filepath, str) else '<array>'}: {e}")
        return None


def compute_melspec(row, n_mels=128, hop_length=512, n_fft=2048):
    """ASV_dataset.ipynb:1151 (cell [27]): z-normalised log-mel of a whole file -> (n_mels, T) float32.
    Unlike the extractors above the notebook function has no try/except: errors propagate."""
    y, sr = _prepare_clip(row, None, None, None, None)
    params = FrontendParams.logmel(sr, n_mels=n_mels, n_fft=n_fft, hop_length=hop_length, znorm=True)
    out, status = _run_batch(params, [y])
    if out[0] is None:
        raise ValueError(L.ITEM_STATUS_NAMES.get(int(status[0]), "item failed"))
    return out[0]


def extract_mfcc(filepath, chunk_start=None, chunk_end=None, sr=None, n_mfcc=13, mean=False, augment=None):
    try:
        y, sr = _prepare_clip(filepath, chunk_start, chunk_end, sr, augment)
        out, status = _run_batch(_mfcc_params(sr, n_mfcc, mean), [y])
        if out[0] is None:
            raise ValueError(L.ITEM_STATUS_NAMES.get(int(status[0]), "item failed"))
        return out[0]
    except Exception as e:
        print(f"[BŁĄD MFCC] {filepath if isinstance(filepath, str) else '<array>'}: {e}")
        return None


def extract_lfcc(filepath, chunk_start=None, chunk_end=None, n_ceps=13, mean=False, augment=None):
    try:
        y, sr = _prepare_clip(filepath, chunk_start, chunk_end, None, augment)
        out, status = _run_batch(_lfcc_params(sr, n_ceps), [y])
        if out[0] is None:
            raise ValueError(L.ITEM_STATUS_NAMES.get(int(status[0]), "item failed"))
        return _lfcc_post(out[0], mean)
    except Exception as e:
        print(f"[BŁĄD LFCC] {filepath if isinstance(filepath, str) else '<array>'}: {e}")
        return None


def _gtcc_params(sr, n_filters, n_ceps):
    return FrontendParams.gtcc(sr, n_ceps=n_ceps, nfilts=n_filters)


def extract_gtcc(filepath, chunk_start=None, chunk_end=None, sr=None, n_filters=40, n_ceps=13, mean=False, augment=None):
    """spafe gfcc of the float waveform (ASV_dl_func.py:484-499) -> (T, n_ceps) float64"""
    try:
        y, sr = _prepare_clip(filepath, chunk_start, chunk_end, sr, augment)
        out, status = _run_batch(_gtcc_params(sr, n_filters, n_ceps), [y])
        if out[0] is None:
            raise ValueError(L.ITEM_STATUS_NAMES.get(int(status[0]), "item failed"))
        return _lfcc_post(out[0], mean)            # same orientation and mean axis as extract_lfcc (:496)
    except Exception as e:
        print(f"[BŁĄD GTCC] {filepath if isinstance(filepath, str) else '<array>'}: {e}")
        return None


_CQCC_PLANS: dict = {}


def get_cqcc_frontend(sr: int, bins_per_octave: int = 12, n_ceps: int = 19, device=None):
    """Cached CQCC plan per (sample rate, bins per octave, n_ceps, device)."""
    from .cqcc import CqccFrontend
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    key = (int(sr), int(bins_per_octave), int(n_ceps), str(dev))
    fe = _CQCC_PLANS.get(key)
    if fe is None:
        fe = _CQCC_PLANS[key] = CqccFrontend(sr, bins_per_octave, n_ceps, dev)
    return fe


def extract_cqcc_batch(clips: Sequence[np.ndarray], sr: int, bins_per_octave: int = 12, n_ceps: int = 19):
    """CQCCs of a list of float32 clips in one GPU call -> list of (n_ceps, T) arrays (None on item error)."""
    fe = get_cqcc_frontend(sr, bins_per_octave, n_ceps)
    B = len(clips)
    lens = np.array([len(c) for c in clips], dtype=np.int32)
    Lmax = (max(int(lens.max()), 1) + 3) // 4 * 4
    host = _pinned_stage(B * Lmax)[:B * Lmax].view(B, Lmax)
    hv = host.numpy()
    for i, c in enumerate(clips):
        hv[i, :len(c)] = c
        hv[i, len(c):] = 0.0
    feats, n_frames, status = fe(host.to(fe.device, non_blocking=True), torch.from_numpy(lens).to(fe.device))
    feats, n_frames, status = feats.cpu().numpy(), n_frames.cpu().numpy(), status.cpu().numpy()
    return [feats[i, :, :n_frames[i]].copy() if status[i] == 0 else None for i in range(B)], status


def extract_cqcc(filepath, chunk_start=None, chunk_end=None, sr=None, bins_per_octave=12, n_ceps=19, mean=False,
                 augment=None):
    """ASV_dl_func.py:442-481: (n_ceps, T) float32, T = 1 + len(y) // 512 (63 for a 2-second 16 kHz chunk: the
    `(19, 63)` input of the reference's CNN-BiLSTM), or its time mean; None (and a printed line) on any error."""
    try:
        y, sr = _prepare_clip(filepath, chunk_start, chunk_end, sr, augment)
        out, status = extract_cqcc_batch([y], sr, bins_per_octave, n_ceps)
        if out[0] is None:
            raise ValueError(L.ITEM_STATUS_NAMES.get(int(status[0]), "item failed"))
        return np.mean(out[0], axis=1) if mean else out[0]
    except Exception as e:
        print(f"[BŁĄD CQCC] {filepath if isinstance(filepath, str) else '<array>'}: {e}")
        return None


# how extract_features batches each recognised extractor: (params builder, post-processing, tag)
_BATCHED = {
    extract_mel_spectrogram: ("MEL", lambda sr, mean: _mel_params(sr, 64, None, mean), lambda x, mean: x),
    extract_mfcc: ("MFCC", lambda sr, mean: _mfcc_params(sr, 13, mean), lambda x, mean: x),
    extract_lfcc: ("LFCC", lambda sr, mean: _lfcc_params(sr, 13), _lfcc_post),
    extract_gtcc: ("GTCC", lambda sr, mean: _gtcc_params(sr, 40, 13), _lfcc_post),
}


def _row_get(row, key, default=None):
    try:
        v = row.get(key, default)
    except AttributeError:
        v = row[key] if key in row else default
    if isinstance(v, float) and math.isnan(v):
        return default
    return v


def _split_rows(feats, n_frames, status, params, post, mean, rows_idx, results, tag):
    feats, n_frames, status = feats.cpu().numpy(), n_frames.cpu().numpy(), status.cpu().numpy()
    for k, i in enumerate(rows_idx):
        if status[k] != 0:
            print(f"[BŁĄD {tag}] row {i}: {L.ITEM_STATUS_NAMES.get(int(status[k]), 'item failed')}")
        elif params.time_mean:
            results[i] = post(feats[k].copy(), mean)
        elif params.layout == L.LAYOUT_CT:
            results[i] = post(feats[k, :, :n_frames[k]].copy(), mean)
        else:
            results[i] = post(feats[k, :n_frames[k], :].copy(), mean)


def extract_features(final_df, feature_extractors_map: Dict[str, Callable], col_name="filepath",
                     mean=False, aug_col="augmentationType", max_batch_samples: int = 1 << 28):
    """Reference dispatcher (ASV_dl_func.py:1031-1049): adds one object column per map key.

    Recognised extractors (log-mel, MFCC, LFCC and CQCC) run as batched GPU calls over a `DeviceCorpus`: every
    distinct file is decoded ONCE for all features and uploaded once (16-bit PCM as int16); the chunk
    rows become a table of offsets into that buffer (`aad_extract_indexed`), so no chunk is copied or
    padded on the host; "noise" rows are augmented on the device.  Any other callable in the map is
    called per row, as the reference does."""
    from .corpus import DeviceCorpus
    n_rows = len(final_df)

    def column(name):   # one pass per column (a pandas look-up per row and field costs more than the GPU work)
        if name not in final_df.columns:
            return [None] * n_rows
        return [None if (isinstance(v, float) and math.isnan(v)) else v for v in final_df[name].tolist()]

    srcs, c_start, c_end, augs = column(col_name), column("chunk_start"), column("chunk_end"), column(aug_col)
    rows = range(n_rows)
    corpus: Optional[DeviceCorpus] = None
    file_of: List[Optional[int]] = [None] * n_rows
    funcs = list(feature_extractors_map.values())
    # MFCC and log-mel over the same rows share one STFT: the second one rides on the first one's kernel launch
    pair_ok = (extract_mfcc in funcs and extract_mel_spectrogram in funcs and
               funcs.index(extract_mfcc) < funcs.index(extract_mel_spectrogram) and
               not any(a == "noise" for a in augs))
    paired_mel: Dict[tuple, tuple] = {}      # (sr, start, end) -> (features, n_frames, status) of the mel plan
    for name, func in feature_extractors_map.items():
        print(f"   - Ekstrahuję: {name}")
        if func is extract_cqcc:
            # CQCC over the same decoded corpus: one batched call per sample rate through the chunk table
            results = [None] * n_rows
            if corpus is None:
                corpus = DeviceCorpus()
                for i in rows:
                    src = srcs[i]
                    try:
                        file_of[i] = corpus.add(src)
                    except Exception as e:
                        print(f"[BŁĄD CQCC] {src if isinstance(src, str) else '<array>'}: {e}")
            by_sr_c: Dict[int, List[int]] = {}
            for i in rows:
                if file_of[i] is None:
                    continue
                if augs[i] is not None:                       # augmented rows: per row, as the reference does
                    results[i] = func(srcs[i], chunk_start=c_start[i], chunk_end=c_end[i], mean=mean, augment=augs[i])
                    continue
                by_sr_c.setdefault(corpus.sample_rates[file_of[i]], []).append(i)
            for sr, idxs in by_sr_c.items():
                cq = get_cqcc_frontend(sr)
                off, ln = corpus.table([(file_of[i], c_start[i], c_end[i]) for i in idxs])
                step = max(1, max_batch_samples // max(int(ln.max()), 1))
                for a in range(0, len(idxs), step):
                    try:
                        o, l = torch.from_numpy(off[a:a + step].copy()), torch.from_numpy(ln[a:a + step].copy())
                        feats, nf, st = cq.extract_indexed(corpus.upload(), o, l, max_len=max(int(l.max()), 1))
                        feats, nf, st = feats.cpu().numpy(), nf.cpu().numpy(), st.cpu().numpy()
                        for k, i in enumerate(idxs[a:a + step]):
                            if st[k] != 0:
                                print(f"[BŁĄD CQCC] row {i}: {L.ITEM_STATUS_NAMES.get(int(st[k]), 'item failed')}")
                            else:
                                x = feats[k, :, :nf[k]].copy()
                                results[i] = np.mean(x, axis=1) if mean else x
                    except Exception as e:
                        print(f"[BŁĄD CQCC] batch {a}:{a + step}: {e}")
            final_df[name] = results
            continue
        if func not in _BATCHED:
            final_df[name] = [func(srcs[i], chunk_start=c_start[i], chunk_end=c_end[i], mean=mean, augment=augs[i])
                              for i in rows]
            continue
        tag, mk_params, post = _BATCHED[func]
        results: List[Optional[np.ndarray]] = [None] * n_rows
        if corpus is None:                      # decode each distinct file once, for every feature
            corpus = DeviceCorpus()
            for i in rows:
                src = srcs[i]
                try:
                    file_of[i] = corpus.add(src)
                except Exception as e:
                    print(f"[BŁĄD {tag}] {src if isinstance(src, str) else '<array>'}: {e}")
        by_sr: Dict[int, List[int]] = {}
        for i in rows:
            if file_of[i] is None:
                continue
            aug = augs[i]
            if aug == "change pitch":
                print(f"[BŁĄD {tag}] row {i}: pitch-shift augmentation is outside the spectral front-end")
                continue
            by_sr.setdefault(corpus.sample_rates[file_of[i]], []).append(i)
        for sr, idxs in by_sr.items():
            params = mk_params(sr, mean)
            fe = get_frontend(params)
            off, ln = corpus.table([(file_of[i], c_start[i], c_end[i]) for i in idxs])
            start = 0
            while start < len(idxs):            # bound the output (rows x longest chunk) per GPU call
                end, lmax = start, 0
                while end < len(idxs):
                    lm = max(lmax, int(ln[end]))
                    if end > start and lm * (end - start + 1) > max_batch_samples:
                        break
                    lmax, end = lm, end + 1
                part = idxs[start:end]
                try:
                    key = (sr, start, end)
                    if pair_ok and func is extract_mfcc:
                        mel_params = _BATCHED[extract_mel_spectrogram][1](sr, mean)
                        if not mel_params.time_mean:        # the rider must be a plain filter-bank plan
                            (feats, mfeats), nf, st = corpus.extract_pair(fe, get_frontend(mel_params), off[start:end], ln[start:end])
                            paired_mel[key] = (mfeats, nf, st)
                            _split_rows(feats, nf, st, params, post, mean, part, results, tag)
                            start = end
                            continue
                    if func is extract_mel_spectrogram and key in paired_mel:
                        feats, nf, st = paired_mel.pop(key)
                        _split_rows(feats, nf, st, params, post, mean, part, results, tag)
                        start = end
                        continue
                    noise = [k for k, i in enumerate(part) if augs[i] == "noise"]
                    feats, nf, st = corpus.extract(fe, off[start:end], ln[start:end], noise_rows=noise)
                    _split_rows(feats, nf, st, params, post, mean, part, results, tag)
                except Exception as e:
                    print(f"[BŁĄD {tag}] batch {start}:{end}: {e}")
                start = end
        final_df[name] = results
    return final_df
