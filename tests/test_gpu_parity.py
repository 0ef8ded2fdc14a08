"""GPU: parity of the CUDA path (through the C ABI) with the oracle.

Tolerances are BASELINE.json's: max abs error <= 1e-3 on log features (dB, ln, cepstra,
deltas); <= 1e-4 on linear spectra, peak-normalised per utterance, and element-wise on the
flat-spectrum noise input.
"""
import numpy as np
import pytest
import torch

import oracle
from oracle import librosa_ref as LR, spafe_ref as SR, delta_ref as DR
from helpers import golden, noise, pad_batch, speech

pytestmark = pytest.mark.gpu

TOL_LOG = 1e-3
TOL_LIN = 1e-4


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def run(params, clips, dev, dtype=np.float32):
    from audioanalysisdetector_b200.frontend import Frontend
    fe = Frontend(params, dev)
    w, lens = pad_batch(clips, dtype)
    out, nf, st = fe(torch.from_numpy(w).to(dev), torch.from_numpy(lens).to(dev))
    torch.cuda.synchronize()
    return out.cpu().numpy(), nf.cpu().numpy(), st.cpu().numpy(), fe


def FP():
    from audioanalysisdetector_b200.frontend import FrontendParams
    return FrontendParams


def LIB():
    from audioanalysisdetector_b200 import _lib
    return _lib


# ----------------------------------------------------------------------------- log-mel
@pytest.mark.parametrize("n_fft,hop,n_mels,sr", [
    (2048, 512, 64, 16000),      # reference default (ASV_dl_func.py:533)
    (512, 160, 80, 16000),       # BASELINE configs[0]
    (1024, 256, 40, 16000),
    (256, 64, 32, 8000),
    (2048, 480, 128, 48000),     # BASELINE configs[4] framing
])
def test_logmel_matches_oracle(dev, n_fft, hop, n_mels, sr):
    clips = [noise(1, 32000), speech(2, 47999, sr), noise(3, 5000), noise(4, 300)]
    out, nf, st, fe = run(FP().logmel(sr, n_mels=n_mels, n_fft=n_fft, hop_length=hop), clips, dev)
    for i, c in enumerate(clips):
        want = LR.logmel_db(c, sr, n_mels=n_mels, n_fft=n_fft, hop_length=hop)
        assert st[i] == 0 and nf[i] == want.shape[1]
        assert np.abs(out[i, :, :nf[i]] - want).max() <= TOL_LOG
        assert np.all(out[i, :, nf[i]:] == 0)
    np.testing.assert_allclose(fe.table(LIB().TABLE_FILTERBANK), LR.mel_filterbank(sr, n_fft, n_mels), atol=1e-7)


def test_linear_mel_energies_tolerance(dev):
    L = LIB()
    p = FP().logmel(16000, n_mels=128).replace(log_type=L.LOG_LN, top_db=-1.0)
    clips = [noise(5, 64000), speech(6, 64000)]
    out, nf, st, _ = run(p, clips, dev)
    for i, c in enumerate(clips):
        want = LR.melspectrogram(c, 16000, n_mels=128).astype(np.float64)
        got = np.exp(out[i, :, :nf[i]].astype(np.float64))
        assert np.abs(got - want).max() / want.max() <= TOL_LIN          # peak-normalised
        if i == 0:                                                        # flat spectrum: element-wise
            assert np.abs(got / want - 1).max() <= TOL_LIN


# ----------------------------------------------------------------------------- MFCC (+ deltas)
@pytest.mark.parametrize("n_mfcc,n_delta", [(13, 0), (40, 2), (20, 1), (64, 2)])
def test_mfcc_matches_oracle(dev, n_mfcc, n_delta):
    clips = [noise(7, 64000), speech(8, 64000), noise(9, 32000), speech(10, 5000)]
    out, nf, st, fe = run(FP().mfcc(16000, n_mfcc=n_mfcc, n_delta=n_delta), clips, dev)
    assert out.shape[1] == n_mfcc * (1 + n_delta)
    for i, c in enumerate(clips):
        want = oracle.mfcc_with_deltas_ref(c, 16000, n_mfcc=n_mfcc, n_delta=n_delta)
        assert st[i] == 0 and nf[i] == want.shape[1]
        assert np.abs(out[i, :, :nf[i]] - want).max() <= TOL_LOG
    k = np.arange(n_mfcc)[:, None]
    m = np.arange(128)[None, :]
    D = 2 * np.where(k == 0, np.sqrt(1 / 512), np.sqrt(1 / 256)) * np.cos(np.pi * k * (2 * m + 1) / 256)
    np.testing.assert_allclose(fe.table(LIB().TABLE_DCT), D, atol=1e-7)


def test_mfcc_small_fft_variant(dev):
    clips = [noise(11, 48000), speech(12, 30000)]
    p = FP().mfcc(16000, n_mfcc=40, n_mels=80, n_fft=512, hop_length=160, n_delta=2)
    out, nf, st, _ = run(p, clips, dev)
    for i, c in enumerate(clips):
        want = oracle.mfcc_with_deltas_ref(c, 16000, n_mfcc=40, n_fft=512, hop_length=160, n_mels=80)
        assert st[i] == 0 and np.abs(out[i, :, :nf[i]] - want).max() <= TOL_LOG


# ----------------------------------------------------------------------------- LFCC
def test_lfcc_reference_defaults(dev):
    clips = [noise(13, 32000), speech(14, 40001), noise(15, 16000), noise(16, 500)]
    out, nf, st, fe = run(FP().lfcc(16000, n_ceps=13), clips, dev)
    assert out.shape[2] == 13                                     # time-major like spafe
    for i, c in enumerate(clips):
        want = oracle.extract_lfcc_ref(c, 16000)
        assert st[i] == 0 and nf[i] == want.shape[0]
        assert np.abs(out[i, :nf[i], :] - want).max() <= TOL_LOG
    np.testing.assert_allclose(fe.table(LIB().TABLE_FILTERBANK), SR.linear_filter_banks(24, 512, 16000)[0] / 512,
                               rtol=1e-6, atol=1e-12)


@pytest.mark.parametrize("fb", ["intbin", "cont", "custom"])
def test_lfcc_config3_int16_with_deltas(dev, fb):
    L = LIB()
    clips = [SR.quantize_int16(noise(20 + i, n)) for i, n in enumerate((16000, 77777, 128000, 31999))]
    kw = dict(n_ceps=20, nfilts=20, win_len=0.02, n_delta=2, layout=L.LAYOUT_CT)
    if fb == "intbin":   # spafe 0.1.x construction, selectable
        p, fbm = FP().lfcc(16000, fb_type=L.FB_LINEAR_INTBIN, **kw), SR.linear_filter_banks_intbin(20, 512, 16000)
    elif fb == "cont":   # spafe 0.3.x construction, the default
        p = FP().lfcc(16000, **kw)
        fbm = SR.linear_filter_banks(20, 512, 16000)[0]
    else:                # the deployment recipe of INTEGRATION.md: spafe's own matrix as a custom bank
        fbm = SR.linear_filter_banks(20, 512, 16000)[0]
        p = FP().lfcc(16000, fb_type=L.FB_CUSTOM, custom_fb=fbm.astype(np.float32), **kw)
    out, nf, st, _ = run(p, clips, dev, dtype=np.int16)
    for i, c in enumerate(clips):
        want = oracle.lfcc_with_deltas_ref(c, 16000, fbanks=fbm)
        assert st[i] == 0 and nf[i] == want.shape[1]
        assert np.abs(out[i, :, :nf[i]] - want).max() <= TOL_LOG


@pytest.mark.parametrize("sr", [22050, 44100, 48000])
def test_lfcc_above_20khz_window_longer_than_nfft(dev, sr):
    """spafe frames with the full 25 ms window (551 / 1102 / 1200 samples) and np.fft.fft(frames, 512) keeps the
    first 512 windowed samples: the drop-in must return the same LFCCs instead of refusing the plan."""
    clips = [noise(70, sr), speech(71, int(1.3 * sr) + 7, sr), noise(72, int(0.025 * sr)), noise(73, int(0.025 * sr) - 1)]
    out, nf, st, fe = run(FP().lfcc(sr, n_ceps=13), clips, dev)
    assert fe.params.win_length == int(0.025 * sr) > 512
    for i, c in enumerate(clips[:3]):
        want = oracle.extract_lfcc_ref(c, sr)
        assert st[i] == 0 and nf[i] == want.shape[0] == (len(c) - int(0.025 * sr)) // int(0.01 * sr) + 1
        assert np.abs(out[i, :nf[i], :] - want).max() <= TOL_LOG
    assert st[3] == 2 and oracle.extract_lfcc_ref(clips[3], sr) is None      # shorter than one frame


@pytest.mark.parametrize("width", [3, 5, 7])
@pytest.mark.parametrize("poison", [False, True])
def test_fused_deltas_of_narrow_widths(dev, width, poison):
    """The fused stencil of k_cepstra walks a fixed 9-wide window; the positions outside a narrower delta width lie
    in the slack in front of a row / in its pad columns and must be neither read into the sum nor multiplied by
    a zero tap (0 * NaN).  `poison` first runs a batch whose status-5 rows leave NaN patterns behind in shared
    memory-sized workspaces, then the clean batch."""
    from audioanalysisdetector_b200.frontend import Frontend
    p = FP().mfcc(16000, n_mfcc=13, n_delta=2, delta_width=width)
    clips = [noise(80, 32000), speech(81, 24000), noise(82, 512 * (width - 1) + 1), noise(83, 70000)]
    if poison:
        bad = [c.copy() for c in clips]
        for c in bad:
            c[::97] = np.nan
        run(p, bad, dev)
    out, nf, st, _ = run(p, clips, dev)
    h = width // 2
    for i, c in enumerate(clips):
        want = oracle.mfcc_with_deltas_ref(c, 16000, n_mfcc=13, n_delta=2, width=width)
        assert st[i] == 0 and nf[i] == want.shape[1]
        got = out[i, :, :nf[i]]
        assert np.isfinite(got).all()
        assert np.abs(got - want).max() <= TOL_LOG
        for sl in (slice(0, h + 1), slice(nf[i] - 1 - h, nf[i])):                # the edge frames the advisory names
            assert np.abs(got[:, sl] - want[:, sl]).max() <= TOL_LOG


# ----------------------------------------------------------------------------- GTCC (spafe gfcc, dense gammatone bank)
@pytest.mark.parametrize("sr", [16000, 22050, 48000])
@pytest.mark.parametrize("spectrum", ["power", "magnitude"])
def test_gtcc_matches_oracle(dev, sr, spectrum):
    """extract_gtcc (ASV_dl_func.py:484-499): gfcc(sig=y, fs=sr, num_ceps=13, nfilts=40) of the float waveform.  The
    cube root keeps the features linear-ish in the spectrum: tolerance 1e-4 of the utterance's largest coefficient
    (and the 1e-3 absolute bound of the log features on top)."""
    L = LIB()
    clips = [speech(90, 2 * sr, sr), noise(91, int(1.3 * sr) + 5), noise(92, int(0.025 * sr)), speech(93, 3 * sr, sr),
             noise(94, int(0.025 * sr) - 1)]
    p = FP().gtcc(sr, spectrum=L.SPEC_POWER if spectrum == "power" else L.SPEC_MAGNITUDE)
    out, nf, st, fe = run(p, clips, dev)
    np.testing.assert_allclose(fe.table(L.TABLE_FILTERBANK) * (512 if spectrum == "power" else 1),
                               SR.gammatone_filter_banks(40, 512, sr)[0], rtol=2e-6, atol=1e-9)
    for i, c in enumerate(clips[:4]):
        want = SR.gfcc(c, sr, 13, nfilts=40, spectrum=spectrum)
        assert st[i] == 0 and nf[i] == want.shape[0]
        err = np.abs(out[i, :nf[i], :] - want).max()
        assert err <= TOL_LIN * max(1.0, np.abs(want).max()) and err <= TOL_LOG
    assert st[4] == 2 and oracle.extract_gtcc_ref(clips[4], sr) is None


def test_gtcc_custom_dense_bank_int16_and_ct_layout(dev):
    """The deployment recipe: spafe's own (dense) matrix as AAD_FB_CUSTOM_DENSE; 16-bit PCM input (value / 32768, as
    librosa decodes it), feature-major layout, an odd number of filters (the last entry holds one filter)."""
    L = LIB()
    rng = np.random.default_rng(5)
    fbm = np.abs(rng.standard_normal((23, 257))) * np.exp(-np.abs(np.arange(257)[None, :] - np.linspace(5, 250, 23)[:, None]) / 30)
    pcm = [(c * 32768).astype(np.int16) for c in (speech(95, 24000), noise(96, 16001) * 0.5)]
    p = FP().gtcc(16000, n_ceps=11, nfilts=23, fb_type=L.FB_CUSTOM_DENSE, custom_fb=fbm.astype(np.float32), layout=L.LAYOUT_CT)
    out, nf, st, _ = run(p, pcm, dev, dtype=np.int16)
    for i, c in enumerate(pcm):
        want = SR.gfcc(c.astype(np.float32) / 32768, 16000, 11, nfilts=23, fbanks=fbm.astype(np.float32).astype(np.float64)).T
        assert st[i] == 0 and nf[i] == want.shape[1]
        assert np.abs(out[i, :, :nf[i]] - want).max() <= TOL_LIN * max(1.0, np.abs(want).max())


def test_dense_bank_is_limited_to_nfft_512_and_cube_root_to_dense_banks(dev):
    from audioanalysisdetector_b200.frontend import Frontend
    from audioanalysisdetector_b200._lib import AadError
    L = LIB()
    with pytest.raises(AadError):
        Frontend(FP().gtcc(16000, nfft=1024), dev)
    with pytest.raises(AadError):
        Frontend(FP().lfcc(16000).replace(log_type=L.LOG_CBRT), dev)
    with pytest.raises(AadError):
        Frontend(FP().gtcc(16000, nfilts=65, n_ceps=13), dev)


@pytest.mark.parametrize("width", [9, 5])
def test_short_utterance_batches_take_the_64_frame_cepstra_tile(dev, width):
    """No clip longer than 64 frames (the reference's 2-second chunks have 63): k_cepstra runs with 64-frame tiles and
    four warps.  MFCC + deltas in both layouts, and LFCC (ln, no reference) -- against the oracle like the long tiles."""
    L = LIB()
    clips = [noise(120, 32000), speech(121, 31999), noise(122, 512 * 9 + 1), speech(123, 20000), noise(124, 32767)]
    for layout in (L.LAYOUT_CT, L.LAYOUT_TC):
        out, nf, st, _ = run(FP().mfcc(16000, n_mfcc=20, n_delta=2, delta_width=width, layout=layout), clips, dev)
        assert nf.max() <= 64
        for i, c in enumerate(clips):
            want = oracle.mfcc_with_deltas_ref(c, 16000, n_mfcc=20, n_delta=2, width=width)
            got = out[i, :, :nf[i]] if layout == L.LAYOUT_CT else out[i, :nf[i], :].T
            assert st[i] == 0 and nf[i] == want.shape[1] and np.abs(got - want).max() <= TOL_LOG
    short = [SR.quantize_int16(noise(125 + i, n)) for i, n in enumerate((10400, 4000, 401, 9999))]   # <= 64 LFCC frames
    out, nf, st, _ = run(FP().lfcc(16000, n_ceps=13), [c.astype(np.float32) / 32767 for c in short], dev)
    assert nf.max() <= 64
    for i, c in enumerate(short):
        want = SR.lfcc(c, 16000, 13)
        assert st[i] == 0 and nf[i] == want.shape[0] and np.abs(out[i, :nf[i], :] - want).max() <= TOL_LOG


def test_int16_input_equals_float_quantised_input(dev):
    clips = [noise(30, 20000), speech(31, 33333)]
    p = FP().lfcc(16000)
    a, nfa, _, _ = run(p, clips, dev)
    b, nfb, _, _ = run(p, [SR.quantize_int16(c) for c in clips], dev, dtype=np.int16)
    assert np.array_equal(nfa, nfb) and np.array_equal(a, b)      # bit-exact: same arithmetic after the cast


def test_pcm16_input_is_bit_identical_to_decoded_float(dev):
    """16-bit PCM over the bus (half the H2D bytes): int16 * 2^-15 is folded into the window, so the
    features equal those of the float32 waveform librosa.load / soundfile would decode (pcm / 32768)."""
    from audioanalysisdetector_b200.frontend import Frontend
    rng = np.random.default_rng(77)
    pcm = np.clip(np.round(3000 * rng.standard_normal((3, 40000))), -32768, 32767).astype(np.int16)
    pcm[1, 30000:] = 0
    lens = torch.tensor([40000, 30000, 40000], dtype=torch.int32, device=dev)
    for params in (FP().mfcc(16000, n_mfcc=20, n_delta=2), FP().logmel(16000, n_mels=80, n_fft=512, hop_length=160)):
        fe = Frontend(params, dev)
        a, nfa, sta = fe(torch.from_numpy(pcm).to(dev), lens)
        b, nfb, stb = fe(torch.from_numpy(pcm.astype(np.float32) / 32768.0).to(dev), lens)
        torch.cuda.synchronize()
        assert torch.equal(nfa, nfb) and int(sta.sum()) == 0 and int(stb.sum()) == 0
        assert torch.equal(a, b)
        want = (oracle.mfcc_with_deltas_ref(pcm[0].astype(np.float32) / 32768.0, 16000, n_mfcc=20, n_delta=2)
                if params.n_ceps else LR.logmel_db(pcm[0].astype(np.float32) / 32768.0, 16000, n_mels=80, n_fft=512, hop_length=160))
        assert np.abs(a[0, :, :int(nfa[0])].cpu().numpy() - want).max() <= TOL_LOG


def test_non_banded_custom_filterbank_is_rejected(dev):
    from audioanalysisdetector_b200 import AadError, Frontend
    L = LIB()
    fbm = np.ones((20, 257), dtype=np.float32)
    with pytest.raises(AadError, match="banded"):
        Frontend(FP().lfcc(16000, n_ceps=20, nfilts=20, fb_type=L.FB_CUSTOM, custom_fb=fbm), dev)
    with pytest.raises(AadError, match="unsupported"):
        Frontend(FP().mfcc(16000, n_fft=4096), dev)


# ----------------------------------------------------------------------------- golden fixtures
def test_regression_anchor_oracle_outputs(dev):
    """oracle_outputs.npz holds the ORACLE's own outputs (tests/golden/make_golden.py): a regression anchor for the
    restatement and the CUDA path together, not evidence of parity with librosa / spafe."""
    L = LIB()
    g = golden("oracle_outputs.npz")
    for name in ("noise", "speech"):
        y = g[f"{name}_wave"]
        for params, key, tc in [
            (FP().logmel(16000), "logmel64", False),
            (FP().mfcc(16000), "mfcc13", False),
            (FP().lfcc(16000), "lfcc13", True),
            (FP().mfcc(16000, n_mfcc=40, n_delta=2), "mfcc40_d2", False),
            (FP().lfcc(16000, n_ceps=20, nfilts=20, win_len=0.02, n_delta=2, layout=L.LAYOUT_CT), "lfcc20_d2", False),
            (FP().logmel(16000, n_mels=80, n_fft=512, hop_length=160), "logmel80_c1", False),
        ]:
            out, nf, st, _ = run(params, [y], dev)
            want = g[f"{name}_{key}"]
            got = out[0, :nf[0], :] if tc else out[0, :, :nf[0]]
            assert st[0] == 0 and got.shape == want.shape
            assert np.abs(got - want).max() <= TOL_LOG, (name, key)


# ----------------------------------------------------------------------------- known answers
def test_zero_input_known_answers(dev):
    z = [np.zeros(32000, np.float32)]
    out, nf, st, _ = run(FP().logmel(16000), z, dev)
    assert nf[0] == 63 and np.all(out[0] == 0.0)                  # 10log10(amin) - 10log10(amin)
    out, nf, st, _ = run(FP().mfcc(16000), z, dev)
    np.testing.assert_allclose(out[0, 0, :63], -100.0 * np.sqrt(128), rtol=1e-6)
    assert np.abs(out[0, 1:, :63]).max() < 1e-3
    out, nf, st, _ = run(FP().lfcc(16000), z, dev)
    np.testing.assert_allclose(out[0, :198, 0], np.log(np.finfo(float).eps) * np.sqrt(24), rtol=1e-6)
    assert np.abs(out[0, :198, 1:]).max() < 1e-4


def test_bin_centred_sinusoid(dev):
    L = LIB()
    n_fft, k0, A = 512, 37, 0.5
    y = (A * np.cos(2 * np.pi * k0 * np.arange(8192) / n_fft)).astype(np.float32)
    # identity "filterbank": one filter per bin pair is not banded, so probe through a 3-filter custom bank
    fbm = np.zeros((3, 257), np.float32)
    fbm[0, k0 - 1], fbm[1, k0], fbm[2, k0 + 1] = 1, 1, 1
    p = FP().logmel(16000, n_mels=3, n_fft=n_fft, hop_length=128).replace(
        fb_type=L.FB_CUSTOM, custom_fb=fbm, log_type=L.LOG_LN, top_db=-1.0)
    out, nf, st, _ = run(p, [y], dev)
    pw = np.exp(out[0, :, 10].astype(np.float64))
    np.testing.assert_allclose(np.sqrt(pw), [A * n_fft / 8, A * n_fft / 4, A * n_fft / 8], rtol=1e-4)


# ----------------------------------------------------------------------------- edges / status
def test_length_edge_cases_and_status(dev):
    n_fft, hop = 512, 160
    lens = [1, hop - 1, hop, n_fft - 1, n_fft, n_fft + 1, 2 * n_fft + 3]
    clips = [noise(40 + i, n) for i, n in enumerate(lens)]
    out, nf, st, _ = run(FP().logmel(16000, n_mels=40, n_fft=n_fft, hop_length=hop), clips, dev)
    for i, c in enumerate(clips):
        want = LR.logmel_db(c, 16000, n_mels=40, n_fft=n_fft, hop_length=hop)
        assert st[i] == 0 and nf[i] == want.shape[1] == 1 + len(c) // hop
        assert np.abs(out[i, :, :nf[i]] - want).max() <= TOL_LOG
    # delta needs T >= 9 (librosa raises -> the reference returns None)
    out, nf, st, _ = run(FP().mfcc(16000, n_delta=2), [noise(50, 512 * 8), np.zeros(0, np.float32), noise(51, 512 * 7)], dev)
    assert list(st) == [0, 1, 3] and list(nf) == [9, 0, 8]
    assert np.all(out[1] == 0) and np.all(out[2] == 0)
    # spafe framing needs one full window
    out, nf, st, _ = run(FP().lfcc(16000), [noise(52, 399), noise(53, 400), noise(54, 559), noise(55, 560)], dev)
    assert list(st) == [2, 0, 0, 0] and list(nf) == [0, 1, 1, 2]


@pytest.mark.parametrize("kind", ["mfcc", "logmel512", "lfcc_noquant"])
@pytest.mark.parametrize("poison", [np.nan, np.inf])
def test_non_finite_audio_is_flagged_and_contained(dev, kind, poison):
    """librosa.util.valid_audio raises on NaN/Inf (the reference then returns None for that row):
    status NONFINITE for that utterance only; utterances sharing its tiles are bit-identical to a clean run.
    (The reference's LFCC path casts to int16 first, which maps NaN to an integer without raising, so
    only the un-quantised variant of that plan can see a non-finite sample.)"""
    params = {"mfcc": FP().mfcc(16000, n_mfcc=13, n_delta=2),
              "logmel512": FP().logmel(16000, n_mels=80, n_fft=512, hop_length=160),
              "lfcc_noquant": FP().lfcc(16000, n_ceps=13, quantize_i16=False)}[kind]
    clips = [noise(40, 6000), noise(41, 5000), noise(42, 7000)]
    clean, nf0, st0, _ = run(params, clips, dev)
    bad = [c.copy() for c in clips]
    bad[1][2500] = poison
    out, nf, st, _ = run(params, bad, dev)
    assert list(st0) == [0, 0, 0] and list(st) == [0, 5, 0]
    assert np.array_equal(out[0], clean[0]) and np.array_equal(out[2], clean[2])


def test_padding_content_is_ignored_and_batch_invariant(dev):
    from audioanalysisdetector_b200.frontend import Frontend
    p = FP().mfcc(16000, n_mfcc=20, n_delta=2)
    fe = Frontend(p, dev)
    clips = [noise(60, 16000), speech(61, 50001), noise(62, 128000), noise(63, 7000), speech(64, 99999)]
    w, lens = pad_batch(clips)
    garbage = w.copy()
    for i, c in enumerate(clips):
        garbage[i, len(c):] = 1e6 * (i + 1)                      # junk beyond lengths must not matter
    a, nfa, _ = fe(torch.from_numpy(w).to(dev), torch.from_numpy(lens).to(dev))
    b, nfb, _ = fe(torch.from_numpy(garbage).to(dev), torch.from_numpy(lens).to(dev))
    assert torch.equal(a, b)
    perm = [3, 0, 4, 2, 1]
    c_, _, _ = fe(torch.from_numpy(w[perm]).to(dev), torch.from_numpy(lens[perm]).to(dev))
    assert torch.equal(c_, a[perm])                              # order/batch independent, bit-exact
    for i, c in enumerate(clips):
        single, nf1, _ = fe(torch.from_numpy(pad_batch([c])[0]).to(dev),
                            torch.tensor([len(c)], dtype=torch.int32, device=dev))
        t = int(nf1[0])
        assert t == int(nfa[i]) and torch.equal(single[0, :, :t], a[i, :, :t])


def test_layouts_and_time_mean(dev):
    L = LIB()
    clips = [noise(70, 30000), speech(71, 64000)]
    ct, nf, _, _ = run(FP().mfcc(16000, n_mfcc=13, n_delta=1), clips, dev)
    tc, nf2, _, _ = run(FP().mfcc(16000, n_mfcc=13, n_delta=1, layout=L.LAYOUT_TC), clips, dev)
    assert np.array_equal(ct.transpose(0, 2, 1), tc)
    mean, _, _, _ = run(FP().mfcc(16000, n_mfcc=13, time_mean=True), clips, dev)
    full, nf3, _, _ = run(FP().mfcc(16000, n_mfcc=13), clips, dev)
    for i in range(2):
        np.testing.assert_allclose(mean[i], full[i, :, :nf3[i]].mean(axis=1), atol=2e-4)
    lm_mean, _, _, _ = run(FP().logmel(16000, time_mean=True), clips, dev)
    for i, c in enumerate(clips):
        np.testing.assert_allclose(lm_mean[i], oracle.extract_mel_spectrogram_ref(c, 16000, mean=True), atol=TOL_LOG)


@pytest.mark.parametrize("layout", ["CT", "TC"])
def test_znorm_compute_melspec_variant(dev, layout):
    """ASV_dataset.ipynb compute_melspec: log-mel (128 mels, ref=np.max) then (S - S.mean()) / S.std()."""
    L = LIB()
    clips = [noise(50, 32000), speech(51, 47999), noise(52, 70000)]       # the last one spans several K2/znorm chunks
    p = FP().logmel(16000, n_mels=128, znorm=True, layout=L.LAYOUT_CT if layout == "CT" else L.LAYOUT_TC)
    out, nf, st, fe = run(p, clips, dev)
    assert fe.launches_per_call == 5
    for i, c in enumerate(clips):
        want = oracle.compute_melspec_ref(c, 16000)
        got = out[i, :, :nf[i]] if layout == "CT" else out[i, :nf[i], :].T
        assert st[i] == 0 and np.abs(got - want).max() <= TOL_LOG
        assert abs(float(got.mean())) <= 1e-4 and abs(float(got.std()) - 1.0) <= 1e-4
    # also on cepstra (not a reference path, same epilogue)
    out2, nf2, _, _ = run(FP().mfcc(16000, n_mfcc=13, n_delta=2, znorm=True), clips[:1], dev)
    w = oracle.mfcc_with_deltas_ref(clips[0], 16000, n_mfcc=13, n_delta=2)
    assert np.abs(out2[0, :, :nf2[0]] - (w - w.mean()) / w.std()).max() <= TOL_LOG


def test_seeded_parameter_sweep_matches_oracle(dev):
    """Random (seeded) plans across every kernel variant: n_fft, hop, filters, cepstra, deltas, layout,
    log-mel / MFCC / LFCC, float / int16 input, ragged lengths including very short clips."""
    L = LIB()
    rng = np.random.default_rng(2026)
    for trial in range(14):
        n_fft = int(rng.choice([256, 512, 1024, 2048]))
        sr = int(rng.choice([8000, 16000, 22050, 48000]))
        lens = [int(rng.integers(n_fft // 2, 6 * n_fft)) for _ in range(3)] + [int(rng.integers(1, n_fft // 2))]
        clips = [noise(1000 + 10 * trial + i, n) if i % 2 == 0 else speech(1000 + 10 * trial + i, n, sr)
                 for i, n in enumerate(lens)]
        kind = trial % 3
        if kind == 0:      # log-mel
            hop = int(rng.integers(n_fft // 8, n_fft // 2 + 1))
            n_mels = int(rng.integers(8, min(129, n_fft // 4)))
            out, nf, st, _ = run(FP().logmel(sr, n_mels=n_mels, n_fft=n_fft, hop_length=hop), clips, dev)
            for i, c in enumerate(clips):
                want = LR.logmel_db(c, sr, n_mels=n_mels, n_fft=n_fft, hop_length=hop)
                assert st[i] == 0 and nf[i] == want.shape[1], (trial, i)
                assert np.abs(out[i, :, :nf[i]] - want).max() <= TOL_LOG, (trial, i, n_fft, hop, n_mels, sr)
        elif kind == 1:    # MFCC + deltas
            hop = int(rng.integers(n_fft // 8, n_fft // 2 + 1))
            n_mels = int(rng.integers(16, min(129, n_fft // 4)))
            n_mfcc = int(rng.integers(1, n_mels + 1))
            n_delta = int(rng.integers(0, 3))
            layout = L.LAYOUT_TC if trial % 2 else L.LAYOUT_CT
            out, nf, st, _ = run(FP().mfcc(sr, n_mfcc=n_mfcc, n_mels=n_mels, n_fft=n_fft, hop_length=hop,
                                           n_delta=n_delta, layout=layout), clips, dev)
            for i, c in enumerate(clips):
                T = 1 + len(c) // hop
                if n_delta and T < 9:
                    assert st[i] == 3, (trial, i)
                    continue
                want = oracle.mfcc_with_deltas_ref(c, sr, n_mfcc=n_mfcc, n_fft=n_fft, hop_length=hop, n_mels=n_mels,
                                                   n_delta=n_delta)
                got = out[i, :nf[i], :].T if layout == L.LAYOUT_TC else out[i, :, :nf[i]]
                assert st[i] == 0 and nf[i] == want.shape[1], (trial, i)
                assert np.abs(got - want).max() <= TOL_LOG, (trial, i, n_fft, hop, n_mels, n_mfcc, n_delta, sr)
        else:              # LFCC on int16 PCM
            win = int(rng.integers(n_fft // 2, n_fft + 1))
            hop = int(rng.integers(win // 4, win // 2 + 1))
            nfilts = int(rng.integers(8, 41))
            n_ceps = int(rng.integers(1, nfilts + 1))
            pcm = [SR.quantize_int16(c) for c in clips]
            p = FP().lfcc(sr, n_ceps=n_ceps, nfilts=nfilts, nfft=n_fft, win_len=win / sr + 1e-9, win_hop=hop / sr + 1e-9)
            assert p.win_length == win and p.hop_length == hop
            out, nf, st, _ = run(p, pcm, dev, dtype=np.int16)
            for i, c in enumerate(pcm):
                if len(c) < win:
                    assert st[i] == 2, (trial, i)
                    continue
                want = SR.lfcc(sig=c, fs=sr, num_ceps=n_ceps, nfilts=nfilts, nfft=n_fft, win_len=win / sr + 1e-9,
                               win_hop=hop / sr + 1e-9)
                assert st[i] == 0 and nf[i] == want.shape[0], (trial, i)
                assert np.abs(out[i, :nf[i], :] - want).max() <= TOL_LOG, (trial, i, n_fft, win, hop, nfilts, n_ceps, sr)


def test_out_buffer_is_validated(dev):
    from audioanalysisdetector_b200.frontend import Frontend
    from audioanalysisdetector_b200 import AadError
    fe = Frontend(FP().mfcc(16000, n_mfcc=13), dev)
    wav = torch.zeros((2, 8000), device=dev)
    t, c, _ = fe.query(2, 8000)
    good = torch.zeros((2, c, t), device=dev)
    fe(wav, out=good)
    padded = torch.zeros((2, c + 3, t), device=dev)[:, :c, :]       # dense rows, larger batch stride: fine
    fe(wav, out=padded)
    for bad in (torch.zeros((2, c, t + 1), device=dev), torch.zeros((2, c, t), device=dev, dtype=torch.float64),
                torch.zeros((2, c, t)), torch.zeros((2, t, c), device=dev).transpose(1, 2)):
        with pytest.raises(AadError):
            fe(wav, out=bad)
    with pytest.raises(AadError):
        fe(wav.double())
    with pytest.raises(AadError):
        fe(wav.cpu())


def test_standalone_delta_matches_oracle(dev):
    import audioanalysisdetector_b200 as aad
    x = np.random.default_rng(5).standard_normal((3, 7, 50)).astype(np.float32)
    nfr = np.array([50, 9, 8], dtype=np.int32)
    for order in (1, 2):
        got = aad.delta(torch.from_numpy(x).to(dev), torch.from_numpy(nfr).to(dev), order=order).cpu().numpy()
        for b in (0, 1):
            want = DR.delta(x[b, :, :nfr[b]], order=order)
            assert np.abs(got[b, :, :nfr[b]] - want).max() <= 1e-5
        assert np.all(got[2] == 0)                               # T < width: left untouched
    t = np.arange(40, dtype=np.float32)
    got = aad.delta(torch.from_numpy((t ** 2)[None, None, :]).to(dev), order=2).cpu().numpy()
    np.testing.assert_allclose(got, 2.0, atol=1e-3)
    got = aad.delta(torch.from_numpy((3 * t + 1)[None, None, :]).to(dev), order=1).cpu().numpy()
    np.testing.assert_allclose(got, 3.0, atol=1e-4)


def test_long_form_48k_multi_tile(dev):
    sr = 48000
    y = speech(80, 30 * sr, sr)
    out, nf, st, _ = run(FP().logmel(sr, n_mels=128, n_fft=2048, hop_length=480), [y], dev)
    want = LR.logmel_db(y, sr, n_mels=128, n_fft=2048, hop_length=480)
    assert nf[0] == want.shape[1] == 3001 and np.abs(out[0, :, :nf[0]] - want).max() <= TOL_LOG
    out, nf, st, _ = run(FP().mfcc(sr, n_mfcc=20, n_fft=2048, hop_length=480, n_delta=2), [y, y[:100000]], dev)
    for i, c in enumerate((y, y[:100000])):
        want = oracle.mfcc_with_deltas_ref(c, sr, n_mfcc=20, hop_length=480)
        assert np.abs(out[i, :, :nf[i]] - want).max() <= TOL_LOG


def test_config5_ten_minutes_at_48k_full_length(dev):
    """BASELINE configs[4] at its stated length: ONE 10-minute 48 kHz utterance, 60 001 frames = 1 876 tiles of a
    single utterance (the long-utterance path of k_prepare / k_stft_fb), against the oracle over the whole matrix,
    with a quiet stretch so that the top_db floor is active."""
    sr, n = 48000, 48000 * 600
    rng = np.random.default_rng(51)
    y = np.clip(0.1 * rng.standard_normal(n), -1, 1).astype(np.float32)
    y[: 20 * sr] *= 1e-5
    t = np.arange(5 * sr) / sr
    y[100 * sr:105 * sr] += (0.5 * np.sin(2 * np.pi * 1000 * t)).astype(np.float32)
    out, nf, st, _ = run(FP().logmel(sr, n_mels=128, n_fft=2048, hop_length=480), [y], dev)
    want = LR.logmel_db(y, sr, n_mels=128, n_fft=2048, hop_length=480)
    assert st[0] == 0 and nf[0] == want.shape[1] == 60001
    err = np.abs(out[0, :, :nf[0]] - want)
    assert err.max() <= TOL_LOG, float(err.max())
    assert (want == want.max() - 80.0).mean() > 0.01 and (out[0] == out[0].max() - 80.0).mean() > 0.01   # floor active


def test_config3_ragged_4096_clips_full_size(dev):
    """BASELINE configs[2] at its stated size: 4096 int16 clips of 1-8 s padded to 8 s, LFCC 20 x 3.  Every clip must
    equal its own extraction as a batch of one (batch invariance over ~1.8 M ragged frames: tiles span utterances),
    and a random sample of them the oracle."""
    from audioanalysisdetector_b200.frontend import Frontend
    L = LIB()
    B, lmax = 4096, 128000
    lens = np.random.default_rng(3).integers(16000, lmax + 1, size=B).astype(np.int32)
    gen = torch.Generator(device=dev).manual_seed(33)
    wav = (torch.randn((B, lmax), generator=gen, device=dev) * 3000).clamp_(-32767, 32767).to(torch.int16)
    p = FP().lfcc(16000, n_ceps=20, nfilts=20, win_len=0.02, n_delta=2, layout=L.LAYOUT_CT)
    fe = Frontend(p, dev)
    out, nf, st = fe(wav, torch.from_numpy(lens).to(dev))
    torch.cuda.synchronize()
    assert int(st.sum()) == 0
    nfh = nf.cpu().numpy()
    assert np.array_equal(nfh, (lens - 320) // 160 + 1)
    pick = np.random.default_rng(4).choice(B, size=48, replace=False)
    for i in pick:
        one, nf1, st1 = fe(wav[i:i + 1, :int(lens[i] + 3) // 4 * 4].contiguous(), torch.from_numpy(lens[i:i + 1]).to(dev))
        assert int(nf1[0]) == nfh[i] and torch.equal(one[0, :, :nfh[i]], out[i, :, :nfh[i]])
    host = wav[torch.from_numpy(pick[:12]).to(dev)].cpu().numpy()
    for k, i in enumerate(pick[:12]):
        want = oracle.lfcc_with_deltas_ref(host[k, :lens[i]], 16000)
        assert np.abs(out[i, :, :nfh[i]].cpu().numpy() - want).max() <= TOL_LOG


def test_host_path_equals_device_path(dev):
    from audioanalysisdetector_b200.frontend import Frontend
    p = FP().mfcc(16000, n_mfcc=40, n_delta=2)
    fe = Frontend(p, dev)
    clips = [noise(90 + i, n) for i, n in enumerate((64000, 12345, 0, 64000, 33000, 4000, 64000))]
    w, lens = pad_batch(clips)
    a, nfa, sta = fe(torch.from_numpy(w).to(dev), torch.from_numpy(lens).to(dev))
    for chunk in (0, 2, 3):
        b, nfb, stb = fe.extract_host(w, lens, chunk_utts=chunk)
        assert np.array_equal(a.cpu().numpy(), b) and np.array_equal(nfa.cpu().numpy(), nfb)
        assert np.array_equal(sta.cpu().numpy(), stb) and stb[2] == 1


# ----------------------------------------------------------------------------- full size (configs[1])
def test_full_size_config2_properties(dev):
    """4096 x 4 s clips, MFCC-40 + d + dd: replicas of 8 distinct clips must be bit-identical
    to their first occurrence (no cross-utterance leakage at scale), those 8 match the oracle,
    and the delta rows equal the standalone stencil applied to the static rows."""
    import audioanalysisdetector_b200 as aad
    from audioanalysisdetector_b200.frontend import Frontend
    B, Ls = 4096, 64000
    base = np.stack([noise(100 + i, Ls) if i % 2 == 0 else speech(100 + i, Ls) for i in range(8)])
    wav = torch.from_numpy(base).to(dev).repeat(B // 8, 1)
    fe = Frontend(FP().mfcc(16000, n_mfcc=40, n_delta=2), dev)
    out, nf, st = fe(wav)
    torch.cuda.synchronize()
    assert int(st.sum()) == 0 and int(nf.min()) == int(nf.max()) == 126
    ref8 = out[:8]
    assert torch.equal(out.view(B // 8, 8, 120, 126), ref8.unsqueeze(0).expand(B // 8, -1, -1, -1))
    got = ref8.cpu().numpy()
    for i in range(8):
        want = oracle.mfcc_with_deltas_ref(base[i], 16000, n_mfcc=40)
        assert np.abs(got[i] - want).max() <= TOL_LOG
    static = out[:64, :40].contiguous()
    d1 = aad.delta(static, order=1)
    d2 = aad.delta(static, order=2)
    assert (out[:64, 40:80] - d1).abs().max().item() <= 1e-5
    assert (out[:64, 80:120] - d2).abs().max().item() <= 1e-5


def test_paired_extraction_equals_the_two_separate_calls(dev):
    """One STFT, two features (aad_extract_pair): the second plan's filter bank runs on the first plan's power
    spectra in the same launch; both outputs equal the separate calls bit for bit, ragged batch included."""
    from audioanalysisdetector_b200 import AadError, Frontend
    clips = [speech(70, 32000), noise(71, 20000), speech(72, 47001), noise(73, 3000), np.zeros(0, np.float32)]
    w, lens = pad_batch(clips)
    wav, ln = torch.from_numpy(w).to(dev), torch.from_numpy(lens).to(dev)
    for pa, pb in [(FP().mfcc(16000, n_mfcc=13), FP().logmel(16000, n_mels=64)),
                   (FP().mfcc(16000, n_mfcc=40, n_delta=2), FP().logmel(16000, n_mels=64)),
                   (FP().logmel(16000, n_mels=128), FP().logmel(16000, n_mels=64, fmax=4000.0)),
                   (FP().mfcc(16000, n_mfcc=20, n_mels=80, n_fft=512, hop_length=160),
                    FP().logmel(16000, n_mels=40, n_fft=512, hop_length=160))]:
        fa, fb = Frontend(pa, dev), Frontend(pb, dev)
        a_alone, nf_a, st_a = fa(wav, ln)
        b_alone, nf_b, st_b = fb(wav, ln)
        (a_pair, b_pair), nf, st = fa.extract_pair(fb, wav, ln)
        assert torch.equal(nf, nf_a) and torch.equal(st, st_a)
        assert torch.equal(a_pair, a_alone)
        ok = st == 0                                                # n_frames and status are the first plan's
        assert torch.equal(b_pair[ok], b_alone[ok]) and float(b_pair[~ok].abs().max()) == 0.0
    with pytest.raises(AadError):                                   # different STFT
        Frontend(FP().mfcc(16000), dev).extract_pair(Frontend(FP().logmel(16000, n_fft=512, hop_length=160), dev), wav, ln)
    with pytest.raises(AadError):                                   # the second plan must be a plain filter bank
        Frontend(FP().logmel(16000), dev).extract_pair(Frontend(FP().mfcc(16000), dev), wav, ln)


def test_cuda_path_matches_torchaudio_directly(dev):
    """Not through the oracle: the CUDA path against torchaudio's own MFCC / log-mel chain (an independent
    implementation of the algorithm of librosa.feature.mfcc / power_to_db(ref=np.max)), 1e-3 on dB values."""
    ta = pytest.importorskip("torchaudio")
    sr = 16000
    clips = [speech(81, 32000), noise(82, 40000), speech(83, 52345)]
    mel_kw = dict(n_fft=2048, hop_length=512, center=True, pad_mode="constant", power=2.0, norm="slaney",
                  mel_scale="slaney", f_min=0.0, f_max=sr / 2)
    mfcc_t = ta.transforms.MFCC(sample_rate=sr, n_mfcc=20, dct_type=2, norm="ortho", log_mels=False,
                                melkwargs=dict(n_mels=128, **mel_kw))
    mel_t = ta.transforms.MelSpectrogram(sample_rate=sr, n_mels=64, **mel_kw)
    out, nf, st, _ = run(FP().mfcc(sr, n_mfcc=20), clips, dev)
    mel, _, _, _ = run(FP().logmel(sr, n_mels=64), clips, dev)
    for i, y in enumerate(clips):
        want = mfcc_t(torch.from_numpy(y)).numpy()
        assert nf[i] == want.shape[1] and np.abs(out[i, :, :nf[i]] - want).max() <= 1e-3
        S = mel_t(torch.from_numpy(y))
        db = ta.functional.amplitude_to_DB(S[None], multiplier=10.0, amin=1e-10,
                                           db_multiplier=float(torch.log10(torch.clamp(S.max(), min=1e-10))), top_db=80.0)[0]
        assert np.abs(mel[i, :, :nf[i]] - db.numpy()).max() <= 1e-3


def test_linear_filterbank_cepstra_match_torchaudio_lfcc(dev):
    """The linear (continuous-frequency) filter bank + dB + DCT chain against torchaudio.transforms.LFCC, an
    independent implementation (centred Hann STFT, unit-peak linear triangles, AmplitudeToDB(top_db 80), ortho
    DCT-II): pins AAD_FB_LINEAR_CONT and the n_fft 512 kernels without going through the oracle."""
    ta = pytest.importorskip("torchaudio")
    from audioanalysisdetector_b200.frontend import FrontendParams
    L = LIB()
    sr = 16000
    clips = [speech(91, 32000), noise(92, 24000)]
    p = FrontendParams(kind=L.KIND_MFCC, sample_rate=sr, n_fft=512, win_length=512, hop_length=160,
                       window=L.WIN_HANN_PERIODIC, center=True, n_filt=64, fb_type=L.FB_LINEAR_CONT, fmin=0.0, fmax=sr / 2,
                       log_type=L.LOG_DB10, ref_type=L.REF_ONE, amin=1e-10, top_db=80.0, n_ceps=20, layout=L.LAYOUT_CT)
    out, nf, st, _ = run(p, clips, dev)
    t = ta.transforms.LFCC(sample_rate=sr, n_filter=64, f_min=0.0, f_max=sr / 2, n_lfcc=20, dct_type=2, norm="ortho",
                           log_lf=False, speckwargs=dict(n_fft=512, hop_length=160, center=True, pad_mode="constant", power=2.0))
    for i, y in enumerate(clips):
        want = t(torch.from_numpy(y)).numpy()
        assert st[i] == 0 and nf[i] == want.shape[1]
        assert np.abs(out[i, :, :nf[i]] - want).max() <= 1e-3
