"""CPU: the spafe restatement and the delta semantics against independent checks."""
import numpy as np
import pytest
import scipy.signal

from oracle import spafe_ref as SR, delta_ref as DR
from helpers import golden, noise


def test_quantize_int16_truncates_toward_zero():
    y = np.array([0.99999, -0.99999, 1.0, -1.0, 0.5, -0.5, 1e-5, -1e-5, 0.0], dtype=np.float32)
    q = SR.quantize_int16(y)
    assert q.dtype == np.int16
    assert list(q) == [32766, -32766, 32767, -32767, 16383, -16383, 0, 0, 0]
    assert SR.quantize_int16(np.array([1.5], np.float32))[0] == np.int16((int(1.5 * 32767) + 2**15) % 2**16 - 2**15)


def test_pre_emphasis_and_framing():
    s = np.arange(1, 11, dtype=np.int16)
    e = SR.pre_emphasis(s)
    assert e.dtype == np.float64 and e[0] == 1.0
    np.testing.assert_allclose(e[1:], s[1:] - 0.97 * s[:-1])
    fr, flen = SR.framing(np.arange(1000.0), fs=16000)
    assert flen == 400 and fr.shape == ((1000 - 400) // 160 + 1, 400)
    assert fr[2, 0] == 320.0
    with pytest.raises(ValueError):
        SR.framing(np.arange(399.0), fs=16000)


def test_linear_filterbank_structure():
    for fbk in (SR.linear_filter_banks(24, 512, 16000)[0], SR.linear_filter_banks(20, 512, 16000)[0],
                SR.linear_filter_banks_intbin(24, 512, 16000), SR.linear_filter_banks_intbin(20, 512, 16000)):
        assert fbk.shape[1] == 257 and fbk.min() >= 0 and fbk.max() <= 1.0
        nz = (fbk > 0).sum(axis=0)
        assert nz.max() <= 2
        for k in np.nonzero(nz == 2)[0]:
            j = np.nonzero(fbk[:, k])[0]
            assert j[1] - j[0] == 1
    fbk = SR.linear_filter_banks_intbin(24, 512, 16000)
    # unit peaks at the centre bins, adjacent falling/rising slopes sum to one
    bins = np.floor(513 * np.linspace(0, 8000, 26) / 16000).astype(int)
    for j in range(24):
        assert fbk[j, bins[j + 1]] == 1.0
    # spafe 0.3.x construction (the default): tuple return, centres 320 Hz apart for 24 filters at 16 kHz (10.24
    # bins: no centre falls on a bin, so no peak reaches 1); inside the band rising + falling slopes sum to 1
    fbk, centers = SR.linear_filter_banks(24, 512, 16000)
    np.testing.assert_allclose(centers, 320.0 * np.arange(1, 25))
    freqs = np.arange(257) * 31.25
    assert 0.95 < fbk.max(axis=1).min() and fbk.max() < 1.0
    inside = (freqs >= centers[0]) & (freqs <= centers[-1])
    np.testing.assert_allclose(fbk.sum(axis=0)[inside], 1.0, atol=1e-12)
    # triangle values against the closed form on continuous frequencies
    j = 7
    want = np.clip(np.minimum((freqs - (centers[j] - 320)) / 320, ((centers[j] + 320) - freqs) / 320), 0, None)
    np.testing.assert_allclose(fbk[j], want, atol=1e-12)


def test_lfcc_known_answers_and_shapes():
    z = np.zeros(32000, dtype=np.int16)
    lf = SR.lfcc(z, fs=16000, num_ceps=13)
    assert lf.shape == (198, 13) and lf.dtype == np.float64
    np.testing.assert_allclose(lf[:, 0], np.log(SR.EPS) * np.sqrt(24), rtol=1e-12)
    assert np.abs(lf[:, 1:]).max() < 1e-9
    with pytest.raises(ValueError):
        SR.lfcc(z, num_ceps=30)


def test_lfcc_matches_direct_dft_formulation():
    y = SR.quantize_int16(noise(4, 4000))
    lf = SR.lfcc(y, fs=16000, num_ceps=13)
    # independent re-derivation of frame 3 with an explicit DFT matrix
    e = np.concatenate([[float(y[0])], y[1:].astype(np.float64) - 0.97 * y[:-1].astype(np.float64)])
    fr = e[3 * 160:3 * 160 + 400] * (0.54 - 0.46 * np.cos(2 * np.pi * np.arange(400) / 399))
    n, k = np.arange(400)[None, :], np.arange(257)[:, None]
    X = (fr[None, :] * np.exp(-2j * np.pi * k * n / 512)).sum(axis=1)
    en = (np.abs(X) ** 2 / 512) @ SR.linear_filter_banks(24, 512, 16000)[0].T
    m = np.arange(24)
    c = [np.sqrt((1 if q == 0 else 2) / 24) * (np.log(en) * np.cos(np.pi * q * (2 * m + 1) / 48)).sum() for q in range(13)]
    np.testing.assert_allclose(lf[3], c, rtol=1e-9, atol=1e-9)


def test_spafe_front_half_matches_scipy_signal():
    """Independent anchor for the spafe half of the oracle that shares no code with oracle/spafe_ref.py: int16
    truncation with np.trunc, pre-emphasis as the FIR filter [1, -0.97] (scipy.signal.lfilter, zero initial state
    => first sample unchanged), framing + symmetric Hamming + FFT by scipy.signal.spectrogram (no padding, tail
    dropped), |X|^2 / nfft; then spafe's linear bank + ln + DCT-II ortho through scipy.fftpack."""
    import scipy.fftpack
    y = noise(21, 9000)
    q = np.trunc(y.astype(np.float32) * np.float32(32767)).astype(np.int16)
    np.testing.assert_array_equal(q, SR.quantize_int16(y))
    e = scipy.signal.lfilter([1.0, -0.97], [1.0], q.astype(np.float64))
    np.testing.assert_allclose(e, SR.pre_emphasis(q), rtol=0, atol=1e-9)
    win = scipy.signal.get_window("hamming", 400, fftbins=False)
    _, _, mag = scipy.signal.spectrogram(e, fs=16000, window=win, nperseg=400, noverlap=240, nfft=512,
                                         detrend=False, return_onesided=True, scaling="spectrum", mode="magnitude")
    power = (mag.T * win.sum()) ** 2 / 512                                   # (T, 257)
    fr, _ = SR.framing(SR.pre_emphasis(q), 16000)
    own = np.abs(np.fft.fft(np.hamming(400) * fr, 512))[:, :257] ** 2 / 512
    assert power.shape == own.shape == ((9000 - 400) // 160 + 1, 257)
    np.testing.assert_allclose(power, own, rtol=1e-9, atol=1e-9 * own.max())
    fbk = SR.linear_filter_banks(24, 512, 16000)[0]
    want = scipy.fftpack.dct(np.log(power @ fbk.T), type=2, axis=1, norm="ortho")[:, :13]
    np.testing.assert_allclose(SR.lfcc(q, fs=16000, num_ceps=13), want, rtol=1e-9, atol=1e-9)


def test_gammatone_bank_structure():
    """spafe gammatone_filter_banks: ERB-spaced centres from low_freq upwards (Slaney's ERBSpace: equal steps of the
    ERB-rate EarQ * ln(1 + f / (EarQ * minBW))), every filter peaks (value 1) at the bin next to its centre."""
    for fs, nf in ((16000, 40), (44100, 24)):
        fb, fc = SR.gammatone_filter_banks(nf, 512, fs)
        assert fb.shape == (nf, 257) and abs(fc[0]) < 1e-9 and np.all(np.diff(fc) > 0) and fc[-1] < fs / 2
        rate = SR.EAR_Q * np.log(1 + fc / (SR.EAR_Q * SR.MIN_BW))
        np.testing.assert_allclose(np.diff(rate), np.diff(rate)[0], rtol=1e-9)
        np.testing.assert_allclose(fb.max(axis=1), 1.0)
        assert np.all(np.abs(np.argmax(fb, axis=1) - fc / (fs / 512)) <= 0.75)
        assert np.all(fb > 0)                                   # dense: no bin is exactly outside a filter


def test_gammatone_rows_follow_the_analytic_gammatone_response():
    """Independent anchor of the filter shape: the magnitude response of the 4th-order gammatone impulse response
    t^3 exp(-B t) cos(2 pi fc t), evaluated by a long FFT, normalised like the bank.  Slaney's digital cascade is the
    impulse-invariant design of exactly that filter, so the rows must follow it where aliasing is negligible."""
    fs, nf = 16000, 40
    fb, fc = SR.gammatone_filter_banks(nf, 512, fs)
    t = np.arange(1 << 15) / fs
    for i in (12, 20, 28, 34):
        erb = ((fc[i] / SR.EAR_Q) ** 4 + SR.MIN_BW ** 4) ** 0.25
        g = t ** 3 * np.exp(-1.019 * 2 * np.pi * erb * t) * np.cos(2 * np.pi * fc[i] * t)
        H = np.abs(np.fft.rfft(g))[:: (1 << 15) // 512][:257]
        H /= H.max()
        sel = fb[i] > 0.03                                       # down to -30 dB around the peak
        assert sel.sum() >= 5
        np.testing.assert_allclose(fb[i][sel] / fb[i][sel].max(), H[sel] / H[sel].max(), rtol=0.08)


def test_gfcc_chain_matches_an_independent_formulation():
    """pre-emphasis as an FIR filter, frames + symmetric Hamming + FFT through scipy.signal.spectrogram, the bank as a
    matrix product, cube root, scipy.fftpack DCT: shares no code with oracle/spafe_ref.py except the bank itself."""
    import scipy.fftpack
    fs = 16000
    y = (noise(7, 20000) * 0.3).astype(np.float32)
    fb, _ = SR.gammatone_filter_banks(40, 512, fs)
    x = scipy.signal.lfilter([1.0, -0.97], [1.0], y.astype(np.float64))
    x[0] = y[0]
    _, _, Z = scipy.signal.spectrogram(x, fs=fs, window=scipy.signal.get_window("hamming", 400, fftbins=False), nperseg=400,
                                       noverlap=240, nfft=512, detrend=False, scaling="spectrum", mode="complex",
                                       return_onesided=True)
    spec = (np.abs(Z.T) * np.hamming(400).sum()) ** 2 / 512       # undo scipy's window normalisation
    want = scipy.fftpack.dct(np.cbrt(spec @ fb.T), type=2, axis=1, norm="ortho")[:, :13]
    got = SR.gfcc(y, fs, 13, nfilts=40)
    assert got.shape == want.shape == ((20000 - 400) // 160 + 1, 13)
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9)
    assert SR.gfcc(y, fs, 13, nfilts=40, spectrum="magnitude").shape == want.shape


def test_delta_taps_and_known_answers():
    np.testing.assert_allclose(DR.savgol_taps(9, 1) * 60, np.arange(-4, 5), atol=1e-12)
    np.testing.assert_allclose(DR.savgol_taps(9, 2) * 462, [28, 7, -8, -17, -20, -17, -8, 7, 28], atol=1e-10)
    t = np.arange(30, dtype=np.float64)
    np.testing.assert_allclose(DR.delta(3.5 * t[None, :] + 1, order=1), 3.5, atol=1e-10)      # ramp incl. edges
    np.testing.assert_allclose(DR.delta((t ** 2)[None, :], order=2), 2.0, atol=1e-9)           # t^2 -> 2
    x = golden("independent.npz")["delta_in"]
    ind = golden("independent.npz")
    np.testing.assert_allclose(DR.delta(x, order=1), ind["scipy_savgol_d1"], atol=1e-6)
    np.testing.assert_allclose(DR.delta(x, order=2), ind["scipy_savgol_d2"], atol=1e-6)
    d = DR.delta(x, order=1)
    assert np.allclose(d[:, :4], d[:, 4:5], atol=1e-6) and np.allclose(d[:, -4:], d[:, -5:-4], atol=1e-6)
    with pytest.raises(ValueError):
        DR.delta(x[:, :8])


def test_delta_interior_matches_torchaudio():
    ta = pytest.importorskip("torchaudio")
    import torch
    x = golden("independent.npz")["delta_in"]
    ref = ta.functional.compute_deltas(torch.from_numpy(x), win_length=9).numpy()
    np.testing.assert_allclose(DR.delta(x, order=1)[:, 4:-4], ref[:, 4:-4], atol=1e-6)


def test_oracle_regression_against_committed_outputs():
    # oracle_outputs.npz = the oracle's own outputs (tests/golden/make_golden.py): regression anchor, not parity evidence
    import oracle
    g = golden("oracle_outputs.npz")
    for name in ("noise", "speech"):
        y = g[f"{name}_wave"]
        np.testing.assert_allclose(oracle.extract_mel_spectrogram_ref(y, 16000), g[f"{name}_logmel64"], atol=2e-4)
        np.testing.assert_allclose(oracle.extract_mfcc_ref(y, 16000), g[f"{name}_mfcc13"], atol=2e-4)
        np.testing.assert_allclose(oracle.extract_lfcc_ref(y, 16000), g[f"{name}_lfcc13"], atol=1e-8)
        np.testing.assert_allclose(oracle.mfcc_with_deltas_ref(y, 16000, n_mfcc=40), g[f"{name}_mfcc40_d2"], atol=2e-4)
