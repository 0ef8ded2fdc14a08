"""Files -> chunk scores without leaving the device: the reference's inference flow in one call.

The reference builds 2-second chunk rows (`prepare_dataframe`, ASV_dl_func.py:281-293), extracts one feature per
row through `extract_features` (:1031-1049), optionally standardises it (`prepare_train_test_data`, :1113-1129),
wraps the object column in a Dataset that converts every item with `torch.tensor` (cnn_bilstm_hybrid.py:4-15)
and calls the model (:54-68).  `score_files` does the same with the pieces of this package: DeviceCorpus (decode
once, one upload, chunk table) -> Frontend (one batched extraction) -> [DeviceStandardScaler] -> DetectorEngine.
"""
from __future__ import annotations

from typing import List, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from .corpus import DeviceCorpus, two_second_chunks
from .detector import DetectorEngine
from .extractors import get_frontend
from .frontend import FrontendParams
from .scaler import DeviceStandardScaler


def score_files(sources: Sequence[object], state_dict: Mapping[str, object], feature: str = "mfcc", n_features: int = 13,
                scaler: Optional[DeviceStandardScaler] = None, device=None, chunk_s: float = 2.0,
                max_rows_per_call: int = 1 << 16) -> Tuple[np.ndarray, List[Tuple[int, float, float]]]:
    """sources: file paths or in-memory (waveform, sr) pairs -> (scores [n_rows], rows [(source index, chunk_start,
    chunk_end)]) for every full `chunk_s`-second chunk, in source order (files shorter than one chunk give no row,
    as in the reference).  `feature`: "mfcc" (librosa.feature.mfcc, n_features coefficients), "mel" (log-mel dB,
    n_features bands) or "cqcc" (extract_cqcc, n_features coefficients: what the reference trains this model on,
    cnn_bilstm_hybrid.py:6,21 with n_features = 19) -- the model needs exactly 63 frames per chunk, i.e. 2-second chunks at 16 kHz with the
    librosa framing.  `scaler`: a fitted DeviceStandardScaler applied to every chunk's (63, n_features) matrix the
    way the reference standardises time-major rows; None: raw features, as cnn_bilstm_hybrid's training loop uses."""
    if feature not in ("mfcc", "mel", "cqcc"):
        raise L.AadError("feature must be 'mfcc', 'mel' or 'cqcc'")
    corpus = DeviceCorpus(device)
    rows: List[Tuple[int, float, float]] = []
    file_rows = []
    for i, src in enumerate(sources):
        f = corpus.add(src)
        for cs, ce in two_second_chunks(corpus.n_samples[f], corpus.sample_rates[f], chunk_s):
            rows.append((i, cs, ce))
            file_rows.append((f, cs, ce))
    if not rows:
        return np.zeros((0,), dtype=np.float32), rows
    srs = {corpus.sample_rates[f] for f, _, _ in file_rows}
    if len(srs) != 1:
        raise L.AadError(f"all sources must share one sample rate, got {sorted(srs)}")
    sr = srs.pop()
    if 1 + int(chunk_s * sr) // 512 != 63:   # librosa's default hop for all three features
        raise L.AadError("the model's first layer takes 63 frames per chunk (2-second chunks at 16 kHz)")
    if feature == "cqcc":
        from .extractors import get_cqcc_frontend
        cq = get_cqcc_frontend(sr, 12, n_features, corpus.device)
    else:
        params = (FrontendParams.mfcc(sr, n_mfcc=n_features) if feature == "mfcc" else FrontendParams.logmel(sr, n_mels=n_features))
        fe = get_frontend(params, corpus.device)
    engine = DetectorEngine(state_dict, feature_dim=n_features, device=corpus.device)
    off, ln = corpus.table(file_rows)
    out = []
    for a in range(0, len(rows), max_rows_per_call):
        if feature == "cqcc":
            o = torch.from_numpy(np.ascontiguousarray(off[a:a + max_rows_per_call]))
            l = torch.from_numpy(np.ascontiguousarray(ln[a:a + max_rows_per_call]))
            feats, nf, st = cq.extract_indexed(corpus.upload(), o, l, max_len=max(int(l.max()), 1))
        else:
            feats, nf, st = corpus.extract(fe, off[a:a + max_rows_per_call], ln[a:a + max_rows_per_call])
        if int(st.ne(0).sum()) != 0:
            raise L.AadError("a chunk failed in the front-end (non-finite audio?)")
        if scaler is not None:     # time-major rows, columns = coefficients (BiLSTM collate, ASV_dl_func.py:1206-1227)
            feats = scaler.transform(feats.transpose(1, 2).contiguous(), inplace=True).transpose(1, 2).contiguous()
        out.append(engine(feats)[:, 0])
    return torch.cat(out).cpu().numpy(), rows
