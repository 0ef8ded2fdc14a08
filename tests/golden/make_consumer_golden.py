"""Generates tests/golden/consumer.npz (run once in the authoring container; committed).

The consumer the front-end drops in ahead of is the reference's CNN-BiLSTM
(`/root/reference/cnn_bilstm_hybrid.py:20-68`, `AudioDeepfakeDetector`), which imports with torch
alone.  This script instantiates the REAL reference model with a fixed seed, puts it in eval mode,
sets `layer_norm.bias` to 1 (at init the LayerNorm(1) output is its bias, i.e. zero, which makes the
network input-independent -- SURVEY.md 3.2), feeds it the oracle's MFCC-13 features of eight seeded
2-second clips ((13, 63) each, the reference's chunk geometry) and stores weights, inputs and scores.
The GPU box cannot read /root/reference: tests compare against this fixture.

    python tests/golden/make_consumer_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root/reference")

import oracle  # noqa: E402
from cnn_bilstm_hybrid import AudioDeepfakeDetector  # noqa: E402  (the reference's own class)


def clip(seed, n=32000):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 16000.0
    return (0.2 * np.sin(2 * np.pi * (120 + 15 * seed) * t) * (1 + 0.4 * np.sin(2 * np.pi * 2.5 * t))
            + 0.02 * rng.standard_normal(n)).astype(np.float32)


def main():
    torch.manual_seed(20261018)
    model = AudioDeepfakeDetector(feature_dim=13).eval()
    with torch.no_grad():
        model.layer_norm.bias.fill_(1.0)
        # non-trivial BatchNorm statistics so that the eval-mode normalisation is exercised
        bn = model.feature_extractor[1]
        bn.running_mean.copy_(torch.linspace(-5.0, 5.0, 64))
        bn.running_var.copy_(torch.linspace(50.0, 400.0, 64))
    seeds = np.arange(300, 308)
    feats = np.stack([oracle.extract_mfcc_ref(clip(int(s)), 16000) for s in seeds]).astype(np.float32)
    assert feats.shape == (8, 13, 63)
    with torch.no_grad():
        scores = model(torch.from_numpy(feats)).numpy()
    out = {"seeds": seeds, "features": feats, "scores": scores}
    for k, v in model.state_dict().items():
        out["w::" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "consumer.npz"), **out)
    print("scores", scores.ravel())


if __name__ == "__main__":
    main()
