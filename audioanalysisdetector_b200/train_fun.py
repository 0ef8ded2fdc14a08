"""Drop-ins for the classical pipeline's extractors (reference `train_fun.py`).

`train_fun.py` has its own, simpler plug-in point: `func(path)` callables that return ONE averaged
vector per file, run as `Parallel(n_jobs=-1)(delayed(func)(path) for path in final_df['file_path'])`
and stored as a DataFrame column (train_fun.py:334-344).  Same names and results here:

  extract_mfcc(filepath)  train_fun.py:69-77  librosa.feature.mfcc(n_mfcc=13).mean(axis=1)  -> (13,) float32
  extract_lfcc(filepath)  train_fun.py:80-88  np.mean(spafe lfcc(num_ceps=13), axis=0)      -> (13,) float64
  run_feature_extractors(final_df, feature_extractors)  the loop at train_fun.py:339-344

(The averaging axis differs from ASV_dl_func.extract_lfcc, which averages the time-major LFCC matrix
over axis 1; ASV_func.py:70 and train_fun.py:86 average over time.)  Both means are computed on the
device (`time_mean` in the plan); the dispatcher decodes every file once and runs one batched call per
feature instead of a process pool.  Errors return None, as the reference's blanket `except` does.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import numpy as np

from . import _lib as L
from .extractors import _prepare_clip, _run_batch, get_frontend
from .frontend import FrontendParams


def _mfcc_params(sr):
    return FrontendParams.mfcc(sr, n_mfcc=13, time_mean=True)


def _lfcc_params(sr):
    return FrontendParams.lfcc(sr, n_ceps=13, time_mean=True)


def extract_mfcc(filepath):
    try:
        y, sr = _prepare_clip(filepath, None, None, None, None)
        out, _ = _run_batch(_mfcc_params(sr), [y])
        if out[0] is None:
            raise ValueError("item failed")
        return out[0]
    except Exception:
        print(Exception)          # train_fun.py:76 prints the class object
        return None


def extract_lfcc(filepath):
    try:
        y, sr = _prepare_clip(filepath, None, None, None, None)
        out, _ = _run_batch(_lfcc_params(sr), [y])
        if out[0] is None:
            raise ValueError("item failed")
        return out[0].astype(np.float64)      # spafe computes in float64
    except Exception:
        return None


_BATCHED = {extract_mfcc: (_mfcc_params, np.float32), extract_lfcc: (_lfcc_params, np.float64)}


def run_feature_extractors(final_df, feature_extractors: Dict[str, Callable], path_col: str = "file_path"):
    """train_fun.py:339-344: one column of per-file vectors per extractor (None where it failed; the
    reference drops those rows right after, :347-348).  The two extractors above run batched over a
    DeviceCorpus (every file decoded and uploaded once for both); any other callable is called per path."""
    from .corpus import DeviceCorpus
    paths = list(final_df[path_col])
    corpus: Optional[DeviceCorpus] = None
    file_of: List[Optional[int]] = [None] * len(paths)
    for name, func in feature_extractors.items():
        print(f"   - Ekstrahuję: {name}")
        if func not in _BATCHED:
            final_df[name] = [func(p) for p in paths]
            continue
        mk_params, dtype = _BATCHED[func]
        if corpus is None:
            corpus = DeviceCorpus()
            for i, p in enumerate(paths):
                try:
                    file_of[i] = corpus.add(p)
                except Exception:
                    file_of[i] = None
        results: List[Optional[np.ndarray]] = [None] * len(paths)
        by_sr: Dict[int, List[int]] = {}
        for i, f in enumerate(file_of):
            if f is not None:
                by_sr.setdefault(corpus.sample_rates[f], []).append(i)
        for sr, idxs in by_sr.items():
            fe = get_frontend(mk_params(sr))
            off, ln = corpus.table([(file_of[i], None, None) for i in idxs])
            try:
                feats, _, status = corpus.extract(fe, off, ln)
                feats, status = feats.cpu().numpy(), status.cpu().numpy()
                for k, i in enumerate(idxs):
                    if status[k] == 0:
                        results[i] = feats[k].astype(dtype)
            except L.AadError:
                pass
        final_df[name] = results
    return final_df
