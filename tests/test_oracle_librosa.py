"""CPU: the librosa restatement against independent implementations and analytic answers."""
import numpy as np
import pytest
import scipy.signal

from oracle import librosa_ref as LR
from helpers import golden, noise, speech


def test_mel_filterbank_matches_torchaudio_frozen():
    ind = golden("independent.npz")
    for n_fft, n_mels, sr in [(2048, 64, 16000), (2048, 128, 16000), (512, 80, 16000), (2048, 128, 48000)]:
        fb = LR.mel_filterbank(sr, n_fft, n_mels)
        assert fb.dtype == np.float32 and fb.shape == (n_mels, n_fft // 2 + 1)
        np.testing.assert_allclose(fb, ind[f"ta_melfb_{n_fft}_{n_mels}_{sr}"], atol=5e-7, rtol=0)


def test_mel_filterbank_structure():
    fb = LR.mel_filterbank(16000, 2048, 128)
    assert int((fb > 0).sum()) == 2020
    nz = (fb > 0).sum(axis=0)
    assert nz.max() <= 2                        # banded: at most two (adjacent) filters per bin
    for k in np.nonzero(nz == 2)[0]:
        j = np.nonzero(fb[:, k])[0]
        assert j[1] - j[0] == 1


def test_mel_filterbank_live_torchaudio():
    ta = pytest.importorskip("torchaudio")
    fb = LR.mel_filterbank(22050, 1024, 40)
    ref = ta.functional.melscale_fbanks(513, 0.0, 11025.0, 40, 22050, norm="slaney", mel_scale="slaney").T.numpy()
    np.testing.assert_allclose(fb, ref, atol=5e-7, rtol=0)


def test_melspectrogram_matches_torchaudio_frozen():
    ind = golden("independent.npz")
    y = golden("oracle_outputs.npz")["noise_wave"]
    for key, kw in [("ta_melspec_noise_2048_128", dict(n_fft=2048, hop_length=512, n_mels=128)),
                    ("ta_melspec_noise_512_80", dict(n_fft=512, hop_length=160, n_mels=80))]:
        S = LR.melspectrogram(y, 16000, **kw)
        ref = ind[key]
        assert S.shape == ref.shape and S.dtype == np.float32
        assert np.abs(S - ref).max() / ref.max() <= 1e-4          # peak-normalised, linear spectra
        assert np.abs(S / ref - 1).max() <= 1e-3                   # f32 FFT in torchaudio vs f64 here


def test_stft_shape_and_padding():
    y = noise(0, 5000)
    D = LR.stft(y, n_fft=512, hop_length=160)
    assert D.shape == (257, 1 + 5000 // 160) and D.dtype == np.complex64
    # centre padding: frame 0 sees n_fft/2 zeros then the first n_fft/2 samples
    w = scipy.signal.get_window("hann", 512, fftbins=True)
    fr = np.concatenate([np.zeros(256), y[:256]]) * w
    np.testing.assert_allclose(D[:, 0], np.fft.rfft(fr), atol=1e-5)


def test_bin_centred_sinusoid_periodic_hann():
    n_fft, k0, A = 512, 37, 0.5
    n = np.arange(4096)
    y = (A * np.cos(2 * np.pi * k0 * n / n_fft)).astype(np.float32)
    D = np.abs(LR.stft(y, n_fft=n_fft, hop_length=128, dtype="f64"))
    mid = D[:, 10]
    assert abs(mid[k0] - A * n_fft / 4) < 1e-4          # A*N/2 * 1/2
    assert abs(mid[k0 - 1] - A * n_fft / 8) < 1e-4 and abs(mid[k0 + 1] - A * n_fft / 8) < 1e-4
    rest = np.delete(mid, [k0 - 1, k0, k0 + 1])
    assert rest.max() < 1e-4


def test_parseval_per_frame():
    y = noise(3, 8192)
    n_fft = 1024
    D = LR.stft(y, n_fft=n_fft, hop_length=256, dtype="f64")
    w = scipy.signal.get_window("hann", n_fft, fftbins=True)
    yp = np.pad(y.astype(np.float64), n_fft // 2)
    for t in (0, 5, 17):
        fr = yp[t * 256:t * 256 + n_fft] * w
        full = np.abs(D[:, t]) ** 2
        energy = (full[0] + full[-1] + 2 * full[1:-1].sum()) / n_fft
        assert abs(energy - (fr ** 2).sum()) <= 1e-9 * max(1.0, (fr ** 2).sum())


def test_power_to_db_known_answers():
    z = np.zeros((4, 7), dtype=np.float32)
    assert np.all(LR.power_to_db(z, ref=np.max) == 0.0)       # 10log10(amin) - 10log10(amin)
    np.testing.assert_allclose(LR.power_to_db(z), -100.0, atol=1e-4)   # float32(1e-10) is not exact
    S = np.array([[1.0, 1e-12, 1e3]], dtype=np.float32)
    db = LR.power_to_db(S, ref=np.max)
    np.testing.assert_allclose(db, [[-30.0, -80.0, 0.0]], atol=1e-5)   # top_db floor at -80
    assert db.dtype == np.float32


def test_zero_input_known_answers():
    z = np.zeros(32000, dtype=np.float32)
    lm = LR.logmel_db(z, 16000)
    assert lm.shape == (64, 63) and np.all(lm == 0.0)
    mf = LR.mfcc(z, 16000, n_mfcc=13)
    np.testing.assert_allclose(mf[0], -100.0 * np.sqrt(128), rtol=1e-6)
    assert np.abs(mf[1:]).max() < 1e-3


def test_dct_matches_scipy_fftpack_and_torchaudio():
    ind = golden("independent.npz")
    x = ind["delta_in"]
    np.testing.assert_allclose(LR.dct_ortho(x, axis=0), ind["scipy_fftpack_dct"], atol=1e-5)
    M = 128
    k = np.arange(13)[:, None]
    m = np.arange(M)[None, :]
    D = 2 * np.where(k == 0, np.sqrt(1 / (4 * M)), np.sqrt(1 / (2 * M))) * np.cos(np.pi * k * (2 * m + 1) / (2 * M))
    np.testing.assert_allclose(D, ind["ta_create_dct_13_128"], atol=1e-6)
    v = np.random.default_rng(0).standard_normal((M, 3))
    np.testing.assert_allclose(D @ v, LR.dct_ortho(v, axis=0, n_out=13), atol=1e-10)
    c = LR.dct_ortho(np.full((M, 1), 3.0), axis=0)
    assert abs(c[0, 0] - 3.0 * np.sqrt(M)) < 1e-9 and np.abs(c[1:]).max() < 1e-9


def test_reference_shapes_for_two_second_chunk():
    import oracle
    y = speech(1, 32000)
    assert oracle.extract_mel_spectrogram_ref(y, 16000).shape == (64, 63)      # cnn_bilstm_hybrid.py:26 needs T=63
    assert oracle.extract_mfcc_ref(y, 16000).shape == (13, 63)
    assert oracle.extract_mel_spectrogram_ref(y, 16000, mean=True).shape == (64,)
    assert oracle.extract_mfcc_ref(y, 16000, chunk_start=0.5, chunk_end=1.0).shape == (13, 1 + 8000 // 512)
    assert oracle.extract_mfcc_ref(np.zeros(0, np.float32), 16000) is None
    assert oracle.extract_mfcc_ref(np.array([np.nan, 0.0], np.float32), 16000) is None


def test_f32_chain_vs_f64_truth_within_tolerance():
    y = speech(2, 32000)
    a, b = LR.logmel_db(y, 16000, dtype="ref"), LR.logmel_db(y, 16000, dtype="f64")
    assert np.abs(a - b).max() <= 1e-3
    a, b = LR.mfcc(y, 16000, dtype="ref"), LR.mfcc(y, 16000, dtype="f64")
    assert np.abs(a - b).max() <= 1e-3


def test_compute_melspec_ref_is_standardised():
    """ASV_dataset.ipynb:1151: the z-normalised variant has zero mean / unit population std."""
    import oracle
    from helpers import noise
    z = oracle.compute_melspec_ref(noise(3, 20000), 16000)
    assert z.shape == (128, 1 + 20000 // 512) and z.dtype == np.float32
    assert abs(float(z.mean())) < 1e-5 and abs(float(z.std()) - 1) < 1e-5


def test_full_chains_match_torchaudio_live():
    """Independent implementation of the whole chain (the oracle is otherwise pinned stage by stage): torchaudio's
    MFCC transform (MelSpectrogram with Slaney scale / norm -> AmplitudeToDB(power, top_db 80, ref 1) -> ortho DCT-II)
    is the same algorithm as librosa.feature.mfcc (ASV_dl_func.py:416), and amplitude_to_DB with the utterance
    maximum as reference is power_to_db(ref=np.max) (:534).  torchaudio runs its FFT in float32: 1e-3 on dB values."""
    ta = pytest.importorskip("torchaudio")
    import torch
    import oracle
    from helpers import noise, speech
    sr = 16000
    for y in (speech(1, 32000), noise(2, 40000), speech(3, 2 * sr) * np.float32(1e-3)):
        mel_kw = dict(n_fft=2048, hop_length=512, center=True, pad_mode="constant", power=2.0, norm="slaney",
                      mel_scale="slaney", f_min=0.0, f_max=sr / 2)
        t = ta.transforms.MFCC(sample_rate=sr, n_mfcc=13, dct_type=2, norm="ortho", log_mels=False,
                               melkwargs=dict(n_mels=128, **mel_kw))
        got = t(torch.from_numpy(y)).numpy()
        want = oracle.extract_mfcc_ref(y, sr)
        assert got.shape == want.shape and np.abs(got - want).max() <= 1e-3
        S = ta.transforms.MelSpectrogram(sample_rate=sr, n_mels=64, **mel_kw)(torch.from_numpy(y))
        db = ta.functional.amplitude_to_DB(S[None], multiplier=10.0, amin=1e-10,
                                           db_multiplier=float(torch.log10(torch.clamp(S.max(), min=1e-10))), top_db=80.0)[0]
        want = oracle.extract_mel_spectrogram_ref(y, sr)
        assert db.shape == want.shape and np.abs(db.numpy() - want).max() <= 1e-3
