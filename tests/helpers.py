import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def noise(seed, n):
    rng = np.random.default_rng(seed)
    return np.clip(0.1 * rng.standard_normal(n), -1, 1).astype(np.float32)


def speech(seed, n, sr=16000):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / sr
    return (0.3 * np.sin(2 * np.pi * 140 * t) * (1 + 0.5 * np.sin(2 * np.pi * 3 * t))
            + 0.05 * np.sin(2 * np.pi * 2300 * t) + 0.003 * rng.standard_normal(n)).astype(np.float32)


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def pad_batch(clips, dtype=np.float32, align=4):
    lmax = (max(max(len(c) for c in clips), 1) + align - 1) // align * align
    w = np.zeros((len(clips), lmax), dtype=dtype)
    for i, c in enumerate(clips):
        w[i, :len(c)] = c
    return w, np.array([len(c) for c in clips], dtype=np.int32)
