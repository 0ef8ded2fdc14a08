"""Generates tests/golden/*.npz (run once in the authoring container; committed).

The reference cannot be imported here (librosa/spafe are absent), so the
fixtures pin (a) the oracle's outputs on small seeded inputs -- a regression anchor for the
restatement itself -- and (b) INDEPENDENT implementations of the same steps available in
this image (torchaudio 2.11 MelSpectrogram / melscale_fbanks, scipy savgol / dct), so the
GPU box can check the oracle and the CUDA path without torchaudio's version mattering.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402
from oracle import librosa_ref as LR, spafe_ref as SR, delta_ref as DR  # noqa: E402


def noise(seed, n):
    rng = np.random.default_rng(seed)
    return np.clip(0.1 * rng.standard_normal(n), -1, 1).astype(np.float32)


def speech(seed, n, sr=16000):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / sr
    return (0.3 * np.sin(2 * np.pi * 140 * t) * (1 + 0.5 * np.sin(2 * np.pi * 3 * t))
            + 0.05 * np.sin(2 * np.pi * 2300 * t) + 0.003 * rng.standard_normal(n)).astype(np.float32)


def main():
    sr = 16000
    clips = {"noise": noise(11, 24000), "speech": speech(12, 20000)}
    out = {}
    for name, y in clips.items():
        out[f"{name}_wave"] = y
        out[f"{name}_logmel64"] = oracle.extract_mel_spectrogram_ref(y, sr)
        out[f"{name}_mfcc13"] = oracle.extract_mfcc_ref(y, sr)
        out[f"{name}_lfcc13"] = oracle.extract_lfcc_ref(y, sr)
        out[f"{name}_mfcc40_d2"] = oracle.mfcc_with_deltas_ref(y, sr, n_mfcc=40)
        out[f"{name}_lfcc20_d2"] = oracle.lfcc_with_deltas_ref(SR.quantize_int16(y), sr)
        out[f"{name}_logmel80_c1"] = LR.logmel_db(y, sr, n_mels=80, n_fft=512, hop_length=160)
    np.savez_compressed(os.path.join(HERE, "oracle_outputs.npz"), **out)

    # independent implementations (torchaudio / scipy), evaluated here and frozen
    import torch
    import torchaudio
    import torchaudio.functional as F
    import scipy.signal
    import scipy.fftpack
    ind = {}
    for n_fft, n_mels, srr in [(2048, 64, 16000), (2048, 128, 16000), (512, 80, 16000), (2048, 128, 48000)]:
        fb = F.melscale_fbanks(n_fft // 2 + 1, 0.0, srr / 2, n_mels, srr, norm="slaney", mel_scale="slaney")
        ind[f"ta_melfb_{n_fft}_{n_mels}_{srr}"] = fb.T.numpy()
    y = clips["noise"]
    ms = torchaudio.transforms.MelSpectrogram(sr, n_fft=2048, hop_length=512, n_mels=128, center=True,
                                              pad_mode="constant", norm="slaney", mel_scale="slaney", power=2.0)
    ind["ta_melspec_noise_2048_128"] = ms(torch.from_numpy(y)).numpy()
    ms = torchaudio.transforms.MelSpectrogram(sr, n_fft=512, hop_length=160, n_mels=80, center=True,
                                              pad_mode="constant", norm="slaney", mel_scale="slaney", power=2.0)
    ind["ta_melspec_noise_512_80"] = ms(torch.from_numpy(y)).numpy()
    x = np.random.default_rng(5).standard_normal((6, 40)).astype(np.float32)
    ind["delta_in"] = x
    ind["scipy_savgol_d1"] = scipy.signal.savgol_filter(x, 9, deriv=1, polyorder=1, axis=-1, mode="interp")
    ind["scipy_savgol_d2"] = scipy.signal.savgol_filter(x, 9, deriv=2, polyorder=2, axis=-1, mode="interp")
    ind["scipy_fftpack_dct"] = scipy.fftpack.dct(x, type=2, norm="ortho", axis=0)
    ind["ta_create_dct_13_128"] = F.create_dct(13, 128, norm="ortho").T.numpy()
    np.savez_compressed(os.path.join(HERE, "independent.npz"), **ind)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
