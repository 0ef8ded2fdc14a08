"""audioanalysisdetector_b200 -- B200-native spectral front-end (log-mel / MFCC / LFCC + deltas).

Drop-in for the feature-extraction hot path of IzaP1k/AudioAnalysisDetector
(ASV_dl_func.py:404-439,522-538,1031-1049): same extractor names and call surface,
arithmetic in hand-written sm_100a CUDA kernels behind the C ABI of include/aad.h.
"""
from . import _lib
from ._lib import AadError
from .frontend import Frontend, FrontendParams, delta, fp32_peak_tflops, pinned_empty
from .extractors import (compute_melspec, extract_cqcc, extract_gtcc, extract_cqcc_batch, extract_features, extract_lfcc,
                         extract_mel_spectrogram, extract_mfcc, get_frontend)
from .cqcc import CqccFrontend
from .detector import DetectorEngine
from .corpus import DeviceCorpus, chunk_bounds, layout_files, two_second_chunks
from . import asv_func, train_fun  # noqa: F401  (drop-ins for the older ASV_func.py / train_fun.py signatures)
from .pipeline import score_files
from .scaler import DeviceStandardScaler, merge_stats
from .training import DeviceFeatureLoader
from .sharding import (bind_to_gpu_numa, contiguous_shard, gather_features, long_form_logmel, partition_by_frames,
                       time_split)

__all__ = [
    "AadError", "Frontend", "FrontendParams", "delta", "fp32_peak_tflops", "pinned_empty",
    "compute_melspec", "extract_cqcc", "extract_gtcc", "extract_cqcc_batch", "CqccFrontend", "extract_features", "extract_lfcc", "extract_mel_spectrogram", "extract_mfcc", "get_frontend",
    "DetectorEngine", "score_files", "DeviceCorpus", "chunk_bounds", "layout_files", "two_second_chunks", "DeviceStandardScaler", "merge_stats", "DeviceFeatureLoader", "bind_to_gpu_numa", "contiguous_shard", "gather_features", "long_form_logmel", "partition_by_frames", "time_split",
]
