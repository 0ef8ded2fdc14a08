"""GPU: the reference-facing drop-ins (same names / arguments / shapes / dtypes / None-on-error
as ASV_dl_func.py:404-439,522-538,1031-1049), checked against the oracle."""
import wave

import numpy as np
import pytest
import torch

import oracle
from helpers import noise, speech

pytestmark = pytest.mark.gpu
SR = 16000


def _write_wav(path, y, sr=SR):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(sr)
        w.writeframes(np.round(y * 32767).astype("<i2").tobytes())


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("wavs")
    paths = []
    for i, n in enumerate((4 * SR + 123, 2 * SR, 6 * SR)):
        p = d / f"clip{i}.wav"
        _write_wav(p, speech(i, n) if i % 2 else noise(i, n))
        paths.append(str(p))
    return paths


def _decoded(path):
    from audioanalysisdetector_b200 import audio_io
    return audio_io.load(path)


def test_per_file_extractors_shapes_dtypes_and_values(files):
    import audioanalysisdetector_b200 as aad
    y, sr = _decoded(files[1])                                       # 2-s chunk -> T = 63 / 198
    mel = aad.extract_mel_spectrogram(files[1])
    assert mel.shape == (64, 63) and mel.dtype == np.float32
    assert np.abs(mel - oracle.extract_mel_spectrogram_ref(y, sr)).max() <= 1e-3
    mf = aad.extract_mfcc(files[1])
    assert mf.shape == (13, 63) and mf.dtype == np.float32
    assert np.abs(mf - oracle.extract_mfcc_ref(y, sr)).max() <= 1e-3
    lf = aad.extract_lfcc(files[1])
    assert lf.shape == (198, 13) and lf.dtype == np.float64
    assert np.abs(lf - oracle.extract_lfcc_ref(y, sr)).max() <= 1e-3
    # keyword variants of the reference signatures
    y0, _ = _decoded(files[0])
    got = aad.extract_mfcc(files[0], chunk_start=1.0, chunk_end=3.0, n_mfcc=20)
    want = oracle.extract_mfcc_ref(y0, sr, chunk_start=1.0, chunk_end=3.0, n_mfcc=20)
    assert got.shape == want.shape == (20, 63) and np.abs(got - want).max() <= 1e-3
    got = aad.extract_mel_spectrogram(files[0], n_mels=80, fmax=4000, mean=True)
    want = oracle.extract_mel_spectrogram_ref(y0, sr, n_mels=80, fmax=4000, mean=True)
    assert got.shape == (80,) and np.abs(got - want).max() <= 1e-3
    got = aad.extract_lfcc(files[0], n_ceps=20, mean=True)
    want = oracle.extract_lfcc_ref(y0, sr, n_ceps=20, mean=True)
    assert got.shape == want.shape and np.abs(got - want).max() <= 1e-3


def test_error_convention_returns_none(files, tmp_path, capsys):
    import audioanalysisdetector_b200 as aad
    assert aad.extract_mfcc(str(tmp_path / "missing.wav")) is None
    assert "[BŁĄD MFCC]" in capsys.readouterr().out
    assert aad.extract_mel_spectrogram(files[0], chunk_start=100.0, chunk_end=102.0) is None   # empty slice
    assert aad.extract_lfcc((np.zeros(100, np.float32), SR)) is None                            # < one window
    assert aad.extract_mfcc(files[0], augment="change pitch") is None                          # out of scope -> None
    out = aad.extract_mfcc(files[0], augment="noise")
    assert out is not None and out.shape[0] == 13


def test_extract_features_dispatcher(files, tmp_path):
    pd = pytest.importorskip("pandas")
    import audioanalysisdetector_b200 as aad
    rows = []
    for p in files:
        y, sr = _decoded(p)
        for k in range(int(len(y) / sr // 2.0)):                    # the reference's 2-s chunk index
            rows.append({"filepath": p, "chunk_index": k, "chunk_start": 2.0 * k, "chunk_end": 2.0 * (k + 1),
                         "augmentationType": None})
    rows.append({"filepath": str(tmp_path / "nope.wav"), "chunk_index": 0, "chunk_start": 0.0, "chunk_end": 2.0,
                 "augmentationType": None})
    df = pd.DataFrame(rows)

    def custom(path, chunk_start=None, chunk_end=None, mean=False, augment=None):
        return np.array([chunk_start, chunk_end])

    fmap = {"mel-spect": aad.extract_mel_spectrogram, "mfcc": aad.extract_mfcc, "lfcc": aad.extract_lfcc,
            "custom": custom}
    df = aad.extract_features(df, fmap)
    assert list(df.columns[-4:]) == list(fmap)
    assert df["mfcc"].iloc[-1] is None and df["mel-spect"].iloc[-1] is None and df["lfcc"].iloc[-1] is None
    cache = {p: _decoded(p) for p in files}
    for _, r in df.iloc[:-1].iterrows():
        y, sr = cache[r["filepath"]]
        assert r["mel-spect"].shape == (64, 63) and r["mfcc"].shape == (13, 63) and r["lfcc"].shape == (198, 13)
        kw = dict(chunk_start=r["chunk_start"], chunk_end=r["chunk_end"])
        assert np.abs(r["mel-spect"] - oracle.extract_mel_spectrogram_ref(y, sr, **kw)).max() <= 1e-3
        assert np.abs(r["mfcc"] - oracle.extract_mfcc_ref(y, sr, **kw)).max() <= 1e-3
        assert np.abs(r["lfcc"] - oracle.extract_lfcc_ref(y, sr, **kw)).max() <= 1e-3
        assert list(r["custom"]) == [r["chunk_start"], r["chunk_end"]]
    # 'mfcc' before 'mel-spect' in the map: both come from ONE STFT (aad_extract_pair); same values as above
    df3 = aad.extract_features(pd.DataFrame(rows), {"mfcc": aad.extract_mfcc, "mel-spect": aad.extract_mel_spectrogram})
    for i in range(len(rows) - 1):
        assert np.array_equal(df3["mfcc"].iloc[i], df["mfcc"].iloc[i])
        assert np.array_equal(df3["mel-spect"].iloc[i], df["mel-spect"].iloc[i])
    assert df3["mfcc"].iloc[-1] is None and df3["mel-spect"].iloc[-1] is None
    # the notebook's full map (ASV_deep_learning.ipynb:152-160) has 'cqcc' first: batched over the same corpus
    from oracle import cqcc_ref
    df4 = aad.extract_features(pd.DataFrame(rows), {"cqcc": aad.extract_cqcc, "gtcc": aad.extract_gtcc, "mfcc": aad.extract_mfcc})
    assert df4["cqcc"].iloc[-1] is None and df4["gtcc"].iloc[-1] is None
    for i in range(len(rows) - 1):
        y, sr = cache[rows[i]["filepath"]]
        want = cqcc_ref.extract_cqcc_ref(y, sr, chunk_start=rows[i]["chunk_start"], chunk_end=rows[i]["chunk_end"])
        got = df4["cqcc"].iloc[i]
        assert got.shape == want.shape == (19, 63) and np.median(np.abs(got - want)) <= 2e-3
        assert np.array_equal(df4["mfcc"].iloc[i], df["mfcc"].iloc[i])
        wg = oracle.extract_gtcc_ref(y, sr, chunk_start=rows[i]["chunk_start"], chunk_end=rows[i]["chunk_end"])
        gg = df4["gtcc"].iloc[i]
        assert gg.dtype == np.float64 and gg.shape == wg.shape == (198, 13) and np.abs(gg - wg).max() <= 1e-4 * max(1, np.abs(wg).max())
    # mean=True variant of the dispatcher (ASV_func.py defaults)
    df2 = aad.extract_features(pd.DataFrame(rows[:3]), {"mfcc": aad.extract_mfcc}, mean=True)
    y, sr = cache[rows[0]["filepath"]]
    want = oracle.extract_mfcc_ref(y, sr, chunk_start=0.0, chunk_end=2.0, mean=True)
    assert df2["mfcc"].iloc[0].shape == (13,) and np.abs(df2["mfcc"].iloc[0] - want).max() <= 1e-3


def test_features_drive_the_reference_consumer_to_the_same_scores():
    """the path drops in ahead of cnn_bilstm_hybrid.AudioDeepfakeDetector.  Features
    stay on the device in the (B, F, 63) layout the model consumes; scores must equal those the REAL
    reference model gave on the oracle's features (tests/golden/consumer.npz)."""
    from audioanalysisdetector_b200 import Frontend, FrontendParams
    from oracle import consumer_ref
    from test_oracle_consumer import consumer_clip, load_fixture
    g, weights = load_fixture()
    dev = torch.device("cuda:0")
    wav = torch.from_numpy(np.stack([consumer_clip(int(s)) for s in g["seeds"]])).to(dev)
    fe = Frontend(FrontendParams.mfcc(16000, n_mfcc=13), dev)
    feats, nf, st = fe(wav)
    assert feats.shape == (8, 13, 63) and int(st.sum()) == 0 and int(nf.min()) == 63
    assert float((feats.cpu() - torch.from_numpy(g["features"])).abs().max()) <= 1e-3
    with torch.no_grad():
        scores = consumer_ref.forward(weights, feats)          # no host round trip, no transpose
    assert scores.is_cuda
    np.testing.assert_allclose(scores.cpu().numpy(), g["scores"], rtol=0, atol=2e-5)
    # the comparison has teeth: shuffled features move the scores by far more than the tolerance
    with torch.no_grad():
        wrong = consumer_ref.forward(weights, feats.flip(0)).cpu().numpy()
    assert np.abs(wrong - g["scores"]).max() > 2e-4


def test_device_standard_scaler_matches_sklearn():
    """prepare_train_test_data (ASV_dl_func.py:1113-1129): StandardScaler.fit(np.vstack(features)) then
    transform per utterance -- on the device, on the features the CUDA path just produced."""
    from sklearn.preprocessing import StandardScaler
    from audioanalysisdetector_b200 import DeviceStandardScaler, Frontend, FrontendParams
    dev = torch.device("cuda:0")
    clips = np.stack([noise(60 + i, 32000) if i % 2 else speech(60 + i, 32000) for i in range(24)])
    fe = Frontend(FrontendParams.logmel(16000, n_mels=64), dev)
    feats, nf, st = fe(torch.from_numpy(clips).to(dev))                    # (24, 64, 63)
    assert int(st.sum()) == 0 and int(nf.min()) == 63
    host = feats.cpu().numpy()
    ref = StandardScaler().fit(np.vstack(list(host)))
    sc = DeviceStandardScaler().fit(feats)
    assert sc.n_samples_seen_ == 24 * 64
    np.testing.assert_allclose(sc.mean_, ref.mean_, rtol=0, atol=1e-5)
    np.testing.assert_allclose(sc.scale_, ref.scale_, rtol=1e-6, atol=1e-6)
    got = sc.transform(feats).cpu().numpy()
    want = np.stack([ref.transform(x) for x in host])
    assert np.abs(got - want).max() <= 1e-4
    assert torch.equal(feats.cpu(), torch.from_numpy(host))                # transform(inplace=False) left the input alone
    # time-major features (T, C), e.g. LFCC for the BiLSTM: columns = coefficients
    fl = Frontend(FrontendParams.lfcc(16000, n_ceps=13), dev)
    lf, _, _ = fl(torch.from_numpy(clips).to(dev))                         # (24, 198, 13)
    ref2 = StandardScaler().fit(np.vstack(list(lf.cpu().numpy())))
    got2 = DeviceStandardScaler().fit_transform(lf).cpu().numpy()
    assert np.abs(got2 - np.stack([ref2.transform(x) for x in lf.cpu().numpy()])).max() <= 1e-4


def test_device_standard_scaler_ragged_counts_valid_frames_only():
    """np.vstack of the per-utterance arrays stacks valid frames only: padding rows of a ragged batch (and rows of
    utterances with a non-zero status) must not enter mean_ / var_ / n_samples_seen_."""
    from sklearn.preprocessing import StandardScaler
    from audioanalysisdetector_b200 import DeviceStandardScaler, Frontend, FrontendParams
    from helpers import pad_batch
    dev = torch.device("cuda:0")
    clips = [noise(90 + i, n) for i, n in enumerate((32000, 17000, 48000, 300, 9000, 24001))]   # one too short
    w, lens = pad_batch(clips)
    fl = Frontend(FrontendParams.lfcc(16000, n_ceps=13), dev)
    lf, nf, st = fl(torch.from_numpy(w).to(dev), torch.from_numpy(lens).to(dev))   # (6, Tmax, 13), time-major
    nfh, sth, host = nf.cpu().numpy(), st.cpu().numpy(), lf.cpu().numpy()
    assert sth[3] != 0 and (sth[[0, 1, 2, 4, 5]] == 0).all()
    valid = [host[i, :nfh[i]] for i in range(len(clips)) if sth[i] == 0]
    ref = StandardScaler().fit(np.vstack(valid))
    sc = DeviceStandardScaler().fit(lf, n_frames=nf, status=st)
    assert sc.n_samples_seen_ == sum(len(v) for v in valid)
    np.testing.assert_allclose(sc.mean_, ref.mean_, rtol=0, atol=1e-5)
    np.testing.assert_allclose(sc.scale_, ref.scale_, rtol=1e-6, atol=1e-6)
    dense = DeviceStandardScaler().fit(lf)                                          # counts the padding: different
    assert dense.n_samples_seen_ == lf.shape[0] * lf.shape[1] > sc.n_samples_seen_


def test_train_fun_extractors_and_dispatcher(files, tmp_path):
    """train_fun.py:69-88 (one averaged vector per file; LFCC averaged over TIME, unlike ASV_dl_func) and the
    `func(path)` loop at :339-344."""
    pd = pytest.importorskip("pandas")
    from audioanalysisdetector_b200 import train_fun as tf
    for p in files:
        y, sr = _decoded(p)
        mf = tf.extract_mfcc(p)
        want = oracle.extract_mfcc_ref(y, sr, mean=True)
        assert mf.shape == (13,) and mf.dtype == np.float32 and np.abs(mf - want).max() <= 1e-3
        lf = tf.extract_lfcc(p)
        want = oracle.extract_lfcc_ref(y, sr, mean=True, mean_axis=0)
        assert lf.shape == (13,) and lf.dtype == np.float64 and np.abs(lf - want).max() <= 1e-3
    assert tf.extract_mfcc(str(tmp_path / "missing.wav")) is None and tf.extract_lfcc(str(tmp_path / "missing.wav")) is None
    df = pd.DataFrame({"file_path": files + [str(tmp_path / "missing.wav")], "label": ["a", "b", "a", "b"]})
    df = tf.run_feature_extractors(df, {"MFCC": tf.extract_mfcc, "LFCC": tf.extract_lfcc, "len": lambda p: len(p)})
    assert df["MFCC"].iloc[-1] is None and df["LFCC"].iloc[-1] is None and df["len"].iloc[0] == len(files[0])
    for i, p in enumerate(files):
        y, sr = _decoded(p)
        assert np.abs(df["MFCC"].iloc[i] - oracle.extract_mfcc_ref(y, sr, mean=True)).max() <= 1e-3
        assert np.abs(df["LFCC"].iloc[i] - oracle.extract_lfcc_ref(y, sr, mean=True, mean_axis=0)).max() <= 1e-3
    assert len(df.dropna(subset=["MFCC", "LFCC"])) == len(files)               # train_fun.py:347


def test_asv_func_signatures_default_to_the_time_mean(files):
    """ASV_func.py:43-73,142-156: mean=True by default, LFCC averaged over time (axis 0)."""
    from audioanalysisdetector_b200 import asv_func as af
    y, sr = _decoded(files[0])
    kw = dict(chunk_start=1.0, chunk_end=3.0)
    got = af.extract_mfcc(files[0], **kw)
    assert got.shape == (13,) and np.abs(got - oracle.extract_mfcc_ref(y, sr, mean=True, **kw)).max() <= 1e-3
    got = af.extract_mel_spectrogram(files[0], **kw)
    assert got.shape == (64,) and np.abs(got - oracle.extract_mel_spectrogram_ref(y, sr, mean=True, **kw)).max() <= 1e-3
    got = af.extract_lfcc(files[0], **kw)
    want = oracle.extract_lfcc_ref(y, sr, mean=True, mean_axis=0, **kw)
    assert got.shape == want.shape == (13,) and got.dtype == np.float64 and np.abs(got - want).max() <= 1e-3
    got = af.extract_lfcc(files[0], mean=False, **kw)
    want = oracle.extract_lfcc_ref(y, sr, **kw)
    assert got.shape == want.shape == (198, 13) and np.abs(got - want).max() <= 1e-3
    assert af.extract_lfcc(files[0], chunk_start=100.0, chunk_end=102.0) is None


def test_cuda_graph_replay_equals_the_direct_call():
    """Frontend.graph: K0 -> K1 -> K2 (with their programmatic-dependent-launch edges) captured once, replayed on new
    contents of the static buffers."""
    from audioanalysisdetector_b200 import Frontend, FrontendParams
    dev = torch.device("cuda:0")
    fe = Frontend(FrontendParams.mfcc(16000, n_mfcc=13, n_delta=2), dev)
    a = torch.from_numpy(np.stack([noise(200 + i, 24000) for i in range(6)])).to(dev)
    b = torch.from_numpy(np.stack([speech(300 + i, 24000) for i in range(6)])).to(dev)
    lens = torch.tensor([24000, 20000, 5000, 24000, 100, 12345], dtype=torch.int32, device=dev)
    g = fe.graph(a.clone(), lens.clone())
    for src in (a, b, a):
        g.wav.copy_(src)
        g.replay()
        want, nf, st = fe(src, lens)
        torch.cuda.synchronize()
        assert torch.equal(g.n_frames, nf) and torch.equal(g.status, st)
        for i in range(6):
            if int(st[i]) == 0:
                assert torch.equal(g.out[i, :, :int(nf[i])], want[i, :, :int(nf[i])])
