"""CPU oracle for the spectral front-end hot path (TEST INFRASTRUCTURE ONLY).

This package is a numpy/scipy restatement of the arithmetic the reference
(IzaP1k/AudioAnalysisDetector) runs on its hot path through third-party
libraries that are NOT vendored in the reference tree and NOT installed in
this image:

* librosa ~= 0.11.0  (requirements.txt:3)  -> oracle/librosa_ref.py
* spafe   ~= 0.3.3   (requirements.txt:5)  -> oracle/spafe_ref.py
* scipy   ~= 1.13.1  (requirements.txt:7)  -> DCT-II ortho / savgol, used directly

PARITY UNPINNED: the reference holds no golden vectors, known-answer tests or
fixtures for this path and neither librosa nor spafe
can be imported here, so the restatement is anchored on (a) the reference's
call sites (ASV_dl_func.py:404-439,522-538; ASV_func.py:43-73,142-156;
train_fun.py:69-88), (b) the published algorithms of the pinned library
versions, (c) independent in-container implementations (torchaudio, scipy) and
(d) analytic known-answer tests -- see tests/test_oracle_*.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product package
(audioanalysisdetector_b200) never does.
"""

from . import librosa_ref, spafe_ref, delta_ref  # noqa: F401
from .extractors_ref import (  # noqa: F401
    extract_mel_spectrogram_ref,
    compute_melspec_ref,
    extract_mfcc_ref,
    extract_lfcc_ref,
    extract_gtcc_ref,
    mfcc_with_deltas_ref,
    lfcc_with_deltas_ref,
)
