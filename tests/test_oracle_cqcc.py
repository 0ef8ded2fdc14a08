"""CPU: the CQCC restatement (oracle/cqcc_ref.py) against analytic properties of a constant-Q transform.

The reference's `librosa.cqt` cannot be run here and its soxr resampler cannot be restated (see the oracle's header):
these tests pin what can be pinned without it -- the ortho scaling of every octave, the frequency resolution, the
resampler's response, the dB / interpolation / DCT chain -- not parity with librosa."""
import numpy as np
import pytest
import scipy.fft
import scipy.interpolate

from oracle import cqcc_ref as C

SR = 16000


def test_geometry_matches_the_reference_call():
    assert C.n_bins_for(16000) == 84 and C.n_bins_for(48000) == 108 and C.n_bins_for(8000) == 72
    f = C.cqt_frequencies(84, C.FMIN_C1)
    assert abs(f[0] - 32.7032) < 1e-3 and abs(f[12] / f[0] - 2.0) < 1e-12
    y = np.zeros(32000, np.float32)
    y[100] = 1.0
    out = C.cqcc(y, SR)
    assert out.shape == (19, 63) and out.dtype == np.float32       # the (19, 63) the CNN-BiLSTM consumes


@pytest.mark.parametrize("k", [3, 10, 30, 47, 59, 70, 83])
def test_pure_tone_has_ortho_scaled_magnitude_in_every_octave(k):
    """A cosine of amplitude A at the centre of bin k must give |CQT[k]| = (A / 2) sqrt(length_k) (scale=True is the
    analogue of norm='ortho'), whichever octave -- i.e. however many resampling stages -- the bin lives in."""
    f = C.cqt_frequencies(84, C.FMIN_C1)
    lengths, _ = C.wavelet_lengths(f, SR)
    t = np.arange(4 * SR) / SR
    mag = np.abs(C.cqt((0.5 * np.cos(2 * np.pi * f[k] * t)).astype(np.float32), SR))
    mid = mag[:, mag.shape[1] // 2]
    assert mid.argmax() == k
    assert abs(mid[k] / (0.25 * np.sqrt(lengths[k])) - 1.0) < 2e-3
    assert mid[(k + 6) % 84] < 0.02 * mid[k]                        # half an octave away: below -34 dB


def test_resampler_is_a_flat_half_band():
    h = C.resample_taps()
    H = np.abs(np.fft.rfft(h, 1 << 14))
    fr = np.arange(len(H)) / (1 << 14)                              # cycles per input sample; new Nyquist = 0.25
    assert np.abs(H[fr <= 0.2] - 1).max() < 1e-6
    assert 20 * np.log10(H[fr >= 0.3].max()) < -120
    assert np.abs(h[1::2][np.arange(127) != 63]).max() < 1e-12      # half band: even offsets from the centre vanish
    y = np.random.default_rng(1).standard_normal(1001).astype(np.float32)
    assert len(C.resample2(y)) == 501                               # ceil(n / 2), librosa.resample's length


def test_db_interpolation_and_dct_chain_against_scipy():
    rng = np.random.default_rng(2)
    mag = np.abs(rng.standard_normal((84, 20))).astype(np.float32) + 1e-3
    db = C.amplitude_to_db(mag)
    assert db.max() == 0.0 and db.min() >= -80.0
    np.testing.assert_allclose(db, np.maximum(20 * np.log10(mag / mag.max()), -80.0), atol=2e-5)
    f = C.cqt_frequencies(84, C.FMIN_C1)
    got = C.interp_to_linear_freqs(db, f)
    lin = np.linspace(f.min(), f.max(), num=84)
    want = np.stack([scipy.interpolate.interp1d(f, db[:, t], kind="linear", fill_value="extrapolate")(lin)
                     for t in range(db.shape[1])], axis=1)           # the reference's loop, ASV_dl_func.py:465-468
    np.testing.assert_allclose(got, want.astype(np.float32), atol=1e-5)


def test_error_convention():
    assert C.extract_cqcc_ref(np.zeros(0, np.float32), SR) is None
    y = np.random.default_rng(3).standard_normal(3 * SR).astype(np.float32) * 0.1
    a = C.extract_cqcc_ref(y, SR, chunk_start=0.5, chunk_end=2.5)
    np.testing.assert_array_equal(a, C.cqcc(y[8000:40000], SR))
    assert C.extract_cqcc_ref(y, SR, mean=True).shape == (19,)


def _direct_constant_q(y, sr, n_bins=84, hop=512):
    """The constant-Q transform by its definition, at the ORIGINAL sample rate for every bin (no octave recursion, no
    resampler, no FFT basis, no sparsification): X[k, t] = sqrt(len_k) * sum_n y[t hop + n] conj(w_k[n]) with the
    L1-normalised Hann-windowed exponential of length len_k = Q sr / f_k centred on the frame."""
    import scipy.signal
    f = C.cqt_frequencies(n_bins, C.FMIN_C1)
    lengths, _ = C.wavelet_lengths(f, sr)
    T = 1 + len(y) // hop
    out = np.zeros((n_bins, T), complex)
    for k in range(n_bins):
        n = np.arange(-lengths[k] // 2, lengths[k] // 2, dtype=float)
        w = scipy.signal.get_window("hann", len(n), fftbins=True) * np.exp(2j * np.pi * f[k] * n / sr)
        w = np.conj(w / np.sum(np.abs(w))) * np.sqrt(lengths[k])
        yp = np.concatenate([np.zeros(len(n)), y.astype(np.float64), np.zeros(len(n))])
        idx = (np.arange(T) * hop + int(n[0]) + len(n))[:, None] + np.arange(len(n))[None, :]
        out[k] = yp[idx] @ w
    return out


@pytest.mark.parametrize("kind", ["noise", "speech", "chirp"])
def test_octave_recursion_with_the_stand_in_resampler_follows_the_direct_transform(kind):
    """Independent anchor of the part of the oracle that cannot be pinned on librosa: the recursive algorithm (FFT
    bases sparsified at 1 %, six stages of the stand-in half-band resampler, fractional window lengths at the decimated
    rates) must reproduce the transform computed straight from its definition.  Measured: 1.4 - 4.4 % of the peak
    (1.2 - 2.0 % rms), the same in the octave that is never resampled as in the one that is resampled six times."""
    import scipy.signal
    from helpers import noise, speech
    y = {"noise": lambda: noise(3, 32000), "speech": lambda: speech(4, 32000),
         "chirp": lambda: (0.5 * scipy.signal.chirp(np.arange(32000) / SR, 50, 2.0, 7000, method="logarithmic")).astype(np.float32)}[kind]()
    a, b = np.abs(C.cqt(y, SR)), np.abs(_direct_constant_q(y, SR))
    err = np.abs(a - b[:, :a.shape[1]])
    assert err.max() <= 0.06 * a.max() and np.sqrt((err ** 2).mean()) <= 0.03 * np.sqrt((b ** 2).mean())
    low, top = err[:12].max(), err[-12:].max()          # six resampling stages vs none
    assert low <= max(2.5 * top, 0.03 * a.max())
