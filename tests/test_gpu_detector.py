"""GPU: the CNN-BiLSTM consumer as hand-written kernels (DetectorEngine, csrc/aad_detector.cu) against the scores
of the REAL reference class (tests/golden/consumer.npz, made by tests/golden/make_consumer_golden.py from
/root/reference/cnn_bilstm_hybrid.py) and against its functional restatement (oracle/consumer_ref.py) on
seeded random weights and inputs.  Tolerance on the sigmoid scores: 2e-5."""
import numpy as np
import pytest
import torch

from oracle import consumer_ref
from test_oracle_consumer import consumer_clip, load_fixture

pytestmark = pytest.mark.gpu
TOL = 2e-5


def random_state(seed, ln_bias=0.7):
    from audioanalysisdetector_b200.detector import _NAMES, _SHAPES
    rng = np.random.default_rng(seed)
    st = {}
    for field, name in _NAMES.items():
        shape = _SHAPES[field]
        fan = max(int(np.prod(shape[1:])) if len(shape) > 1 else 8, 1)
        st[name] = torch.from_numpy((rng.standard_normal(shape) / np.sqrt(fan)).astype(np.float32))
    st["feature_extractor.1.running_var"] = torch.from_numpy(rng.uniform(0.3, 2.0, 64).astype(np.float32))
    st["feature_extractor.1.weight"] = torch.from_numpy(rng.uniform(0.5, 1.5, 64).astype(np.float32))
    st["layer_norm.bias"] = torch.tensor([ln_bias], dtype=torch.float32)
    return st


def test_scores_equal_the_real_reference_model_on_the_golden_fixture():
    from audioanalysisdetector_b200 import DetectorEngine, Frontend, FrontendParams
    g, weights = load_fixture()
    dev = torch.device("cuda:0")
    eng = DetectorEngine(weights, feature_dim=13, device=dev)
    got = eng(torch.from_numpy(g["features"]).to(dev))
    assert got.shape == (8, 1) and got.dtype == torch.float32
    np.testing.assert_allclose(got.cpu().numpy(), g["scores"], rtol=0, atol=TOL)
    # the whole hand-off on the device: waveform -> CUDA front-end -> CUDA detector, no host round trip
    wav = torch.from_numpy(np.stack([consumer_clip(int(s)) for s in g["seeds"]])).to(dev)
    feats, nf, st = Frontend(FrontendParams.mfcc(16000, n_mfcc=13), dev)(wav)
    np.testing.assert_allclose(eng(feats).cpu().numpy(), g["scores"], rtol=0, atol=TOL)
    assert np.abs(eng(feats.flip(0)).cpu().numpy() - g["scores"]).max() > 2e-4      # the comparison has teeth


@pytest.mark.parametrize("F,B,ln_bias", [(13, 1, 1.0), (13, 1000, 0.7), (19, 333, -0.5), (64, 257, 1.3), (12, 65, 1.0)])
def test_matches_the_functional_restatement_on_random_weights(F, B, ln_bias):
    from audioanalysisdetector_b200 import DetectorEngine
    dev = torch.device("cuda:0")
    state = random_state(100 + F, ln_bias)
    g = torch.Generator(device=dev).manual_seed(F * 1000 + B)
    x = 5.0 * torch.randn((B, F, 63), generator=g, device=dev) - 2.0
    with torch.no_grad():
        want = consumer_ref.forward(state, x.cpu())     # on the CPU: cuDNN's convolution would run in TF32
    got = DetectorEngine(state, feature_dim=F, device=dev)(x).cpu()
    assert got.shape == want.shape == (B, 1)
    assert float((got - want).abs().max()) <= TOL
    assert B == 1 or float(want.std()) > 1e-3                                      # scores are not saturated


def test_strided_input_reads_the_first_63_frames():
    """Features of longer clips ((B, F, T > 63) with a row stride) go in as they are: frames 0..62, like x[:, :, :63]."""
    from audioanalysisdetector_b200 import AadError, DetectorEngine
    dev = torch.device("cuda:0")
    state = random_state(7)
    g = torch.Generator(device=dev).manual_seed(3)
    big = torch.randn((50, 13, 126), generator=g, device=dev)
    eng = DetectorEngine(state, feature_dim=13, device=dev)
    with torch.no_grad():
        want = consumer_ref.forward(state, big[:, :, :63].contiguous().cpu())
    assert float((eng(big).cpu() - want).abs().max()) <= TOL
    assert float((eng(big[10:40]).cpu() - want[10:40]).abs().max()) <= TOL          # a batch slice (offset base pointer)
    with pytest.raises(AadError):
        eng(big[:, :, :40])                                                        # fewer than 63 frames
    with pytest.raises(AadError):
        eng(big[:, :12, :])                                                        # wrong feature count
    assert eng(big[:0]).shape == (0, 1)


def test_score_files_equals_the_reference_flow_chunk_by_chunk(tmp_path):
    """wav files -> 2-s chunk rows -> MFCC-13 -> model, all on the device, against the same flow done by hand with
    the oracle: per-chunk slicing (ASV_dl_func.py:407-410), librosa MFCC restatement, the model's restatement."""
    import wave
    import oracle
    from audioanalysisdetector_b200 import audio_io, score_files
    from helpers import noise, speech
    sr = 16000
    paths = []
    for i, n in enumerate((5 * sr + 100, 2 * sr, sr)):                     # 2 chunks, 1 chunk, none
        p = tmp_path / f"s{i}.wav"
        with wave.open(str(p), "wb") as w:
            w.setnchannels(1); w.setsampwidth(2); w.setframerate(sr)
            y = speech(40 + i, n) if i % 2 else noise(40 + i, n)
            w.writeframes(np.round(y * 32767).astype("<i2").tobytes())
        paths.append(str(p))
    _, weights = load_fixture()
    scores, rows = score_files(paths, weights)
    assert rows == [(0, 0.0, 2.0), (0, 2.0, 4.0), (1, 0.0, 2.0)] and scores.shape == (3,)
    feats = []
    for i, cs, ce in rows:
        y, _ = audio_io.load(paths[i])
        feats.append(oracle.extract_mfcc_ref(y, sr, chunk_start=cs, chunk_end=ce))
    with torch.no_grad():
        want = consumer_ref.forward(weights, torch.from_numpy(np.stack(feats)))[:, 0].numpy()
    assert np.abs(scores - want).max() <= 5e-5       # 1e-3 feature tolerance through the model
    assert score_files([paths[2]], weights)[0].shape == (0,)


def test_score_files_with_cqcc_the_feature_the_reference_trains_on(tmp_path):
    """The reference's own pipeline: FLAC files -> 2-s chunk rows -> CQCC-19 (cnn_bilstm_hybrid.py:6,21) -> model, on
    the device through the chunk table, against the oracle's CQCC (per-chunk slicing) through the model's restatement."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import flac_writer as FW
    from oracle import cqcc_ref
    from audioanalysisdetector_b200 import audio_io, score_files
    from helpers import noise, speech
    sr = 16000
    paths = []
    for i, n in enumerate((4 * sr + 77, 2 * sr + 1)):                      # 2 chunks + 1 chunk
        y = speech(50 + i, n) if i % 2 else noise(50 + i, n)
        p = tmp_path / f"LA_E_{i}.flac"
        p.write_bytes(FW.encode(np.round(y * 32767).astype(np.int64), sr, seed=i))
        paths.append(str(p))
    _, weights = load_fixture()
    scores, rows = score_files(paths, weights, feature="cqcc", n_features=19)
    assert rows == [(0, 0.0, 2.0), (0, 2.0, 4.0), (1, 0.0, 2.0)] and scores.shape == (3,)
    feats = []
    for i, cs, ce in rows:
        y, _ = audio_io.load(paths[i])
        f = cqcc_ref.extract_cqcc_ref(y, sr, chunk_start=cs, chunk_end=ce)
        assert f.shape == (19, 63)
        feats.append(f)
    with torch.no_grad():
        want = consumer_ref.forward(weights, torch.from_numpy(np.stack(feats)))[:, 0].numpy()
    assert np.abs(scores - want).max() <= 2e-3       # CQCC cells near 0 dB are ill-conditioned (oracle/cqcc_ref.py)
