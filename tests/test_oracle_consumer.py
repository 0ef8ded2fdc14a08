"""CPU: the functional restatement of the reference's CNN-BiLSTM consumer against the fixture produced
by the REAL reference class (tests/golden/make_consumer_golden.py imports cnn_bilstm_hybrid.py)."""
import numpy as np
import torch

import oracle
from oracle import consumer_ref
from helpers import golden


def load_fixture():
    g = golden("consumer.npz")
    weights = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w::")}
    return g, weights


def consumer_clip(seed, n=32000):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 16000.0
    return (0.2 * np.sin(2 * np.pi * (120 + 15 * seed) * t) * (1 + 0.4 * np.sin(2 * np.pi * 2.5 * t))
            + 0.02 * rng.standard_normal(n)).astype(np.float32)


def test_restatement_reproduces_reference_scores():
    g, weights = load_fixture()
    with torch.no_grad():
        got = consumer_ref.forward(weights, torch.from_numpy(g["features"])).numpy()
    assert got.shape == (8, 1)
    np.testing.assert_allclose(got, g["scores"], rtol=0, atol=1e-6)


def test_fixture_features_are_the_oracle_mfcc():
    g, _ = load_fixture()
    for s, want in zip(g["seeds"], g["features"]):
        got = oracle.extract_mfcc_ref(consumer_clip(int(s)), 16000)
        assert got.shape == (13, 63)           # 2-s chunk -> the T = 63 the model's Conv1d requires
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-4)


def test_model_only_accepts_63_frames():
    _, weights = load_fixture()
    with torch.no_grad():
        try:
            consumer_ref.forward(weights, torch.zeros(2, 13, 64))
            raised = False
        except RuntimeError:
            raised = True
    assert raised
