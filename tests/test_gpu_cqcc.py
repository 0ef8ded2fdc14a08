"""GPU: CQCC on the device (aad_cqcc) against the oracle restatement (same stand-in resampler on both sides; parity
with librosa + soxr is unpinned, see oracle/cqcc_ref.py).

|CQT| is the well-conditioned quantity: it is compared element-wise relative to the utterance maximum.  The
cepstra pass through log(dB^2 + 1e-12), whose derivative 2/|dB| is unbounded near the utterance maximum, so they are
compared with the tolerance that the |CQT| tolerance implies for every frame (propagated through the oracle's own
chain), plus an absolute floor."""
import numpy as np
import pytest
import torch

from oracle import cqcc_ref as C
from helpers import noise, pad_batch, speech

pytestmark = pytest.mark.gpu
SR = 16000


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _run(clips, dev, sr=SR, dtype=np.float32, **kw):
    from audioanalysisdetector_b200 import CqccFrontend
    fe = CqccFrontend(sr, device=dev, **kw)
    w, lens = pad_batch(clips, dtype)
    out, nf, st, mag = fe(torch.from_numpy(w).to(dev), torch.from_numpy(lens).to(dev), return_cqt=True)
    torch.cuda.synchronize()
    return out.cpu().numpy(), nf.cpu().numpy(), st.cpu().numpy(), mag.cpu().numpy()


def _cqcc_from_mag(mag, n_ceps=19, sr=SR):
    import scipy.fft
    n_bins = mag.shape[0]
    db = C.amplitude_to_db(mag)
    interp = C.interp_to_linear_freqs(db, C.cqt_frequencies(n_bins, C.FMIN_C1))
    lp = np.log(np.square(interp) + np.float32(1e-12)).astype(np.float32)
    return scipy.fft.dct(lp, type=2, axis=0, norm="ortho")[:n_ceps].astype(np.float32)


def test_cqt_magnitudes_and_cepstra_match_the_oracle(dev):
    clips = [noise(1, 32000), speech(2, 32000), noise(3, 20001), speech(4, 47000), noise(5, 700)]
    out, nf, st, mag = _run(clips, dev)
    for i, y in enumerate(clips):
        want_c = np.abs(C.cqt(y, SR))
        T = want_c.shape[1]
        assert st[i] == 0 and nf[i] == T == 1 + len(y) // 512
        got_c = mag[i, :, :T]
        assert np.abs(got_c - want_c).max() <= 2e-5 * want_c.max()          # peak-normalised, like the linear spectra
        # the cepstra of the DEVICE magnitudes through the oracle's own dB / interpolation / log / DCT chain: isolates the
        # epilogue kernel from the conditioning of log(dB^2)
        np.testing.assert_allclose(out[i, :, :T], _cqcc_from_mag(got_c), atol=5e-3, rtol=0)
        want = C.cqcc(y, SR)
        err = np.abs(out[i, :, :T] - want)
        assert np.median(err) <= 2e-3 and err.max() <= 0.5                 # ill-conditioned cells near 0 dB allowed for


def test_pure_tone_ortho_scaling_on_the_device(dev):
    f = C.cqt_frequencies(84, C.FMIN_C1)
    lengths, _ = C.wavelet_lengths(f, SR)
    t = np.arange(3 * SR) / SR
    ks = (5, 28, 52, 83)
    clips = [(0.5 * np.cos(2 * np.pi * f[k] * t)).astype(np.float32) for k in ks]
    _, nf, st, mag = _run(clips, dev)
    for i, k in enumerate(ks):
        mid = mag[i, :, nf[i] // 2]
        assert mid.argmax() == k and abs(mid[k] / (0.25 * np.sqrt(lengths[k])) - 1.0) < 2e-3


def test_int16_input_status_and_other_rates(dev):
    y = speech(7, 32000)
    pcm = np.round(y * 32767).astype(np.int16)
    a, nfa, sta, _ = _run([pcm], dev, dtype=np.int16)
    b, nfb, stb, _ = _run([pcm.astype(np.float32) / 32768.0], dev)
    assert np.array_equal(a, b) and nfa[0] == nfb[0] == 63
    out, nf, st, _ = _run([noise(8, 1000), np.zeros(0, np.float32), np.full(4000, np.nan, np.float32)], dev)
    assert list(st) == [0, 1, 5] and nf[0] == 2 and nf[1] == 0
    for sr in (8000, 22050, 48000):
        y = noise(9, int(1.5 * sr))
        out, nf, st, mag = _run([y], dev, sr=sr)
        want = np.abs(C.cqt(y, sr))
        assert st[0] == 0 and mag.shape[1] == C.n_bins_for(sr) == want.shape[0]
        assert np.abs(mag[0, :, :nf[0]] - want).max() <= 2e-5 * want.max()


def test_drop_in_extract_cqcc(dev):
    import audioanalysisdetector_b200 as aad
    y = speech(11, 3 * SR)
    got = aad.extract_cqcc((y, SR), chunk_start=0.5, chunk_end=2.5)
    want = C.extract_cqcc_ref(y, SR, chunk_start=0.5, chunk_end=2.5)
    assert got.shape == want.shape == (19, 63) and got.dtype == np.float32
    assert np.median(np.abs(got - want)) <= 2e-3
    m = aad.extract_cqcc((y, SR), mean=True)
    assert m.shape == (19,)
    assert aad.extract_cqcc((np.zeros(0, np.float32), SR)) is None
