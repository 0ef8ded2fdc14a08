"""Oracle: delta / delta-delta features (librosa.feature.delta semantics).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

The reference has no delta call site; BASELINE.json's
north_star mandates them, and the semantics are those of the library the path
already uses: librosa.feature.delta(data, width=9, order, axis=-1, mode='interp')
== scipy.signal.savgol_filter(data, width, deriv=order, polyorder=order,
axis=axis, mode='interp').  Requires T >= width.
"""
from __future__ import annotations

import numpy as np
import scipy.signal


def delta(data, width=9, order=1, axis=-1, mode="interp"):
    data = np.atleast_1d(data)
    if mode == "interp" and width > data.shape[axis]:
        raise ValueError(
            f"when mode='interp', width={width} cannot exceed data.shape[axis]={data.shape[axis]}")
    if width < 3 or width % 2 != 1:
        raise ValueError("width must be an odd integer >= 3")
    if order <= 0 or not isinstance(order, (int, np.integer)):
        raise ValueError("order must be a positive integer")
    return scipy.signal.savgol_filter(data, width, deriv=order, polyorder=order,
                                      axis=axis, mode=mode)


def stack_deltas(feat, n_delta=2, width=9):
    """[static; delta; delta-delta] stacked on the coefficient axis of a (C, T) array.

    Delta-delta is the order-2 filter applied to the STATIC features."""
    parts = [feat]
    for order in range(1, n_delta + 1):
        parts.append(delta(feat, width=width, order=order, axis=-1).astype(feat.dtype))
    return np.concatenate(parts, axis=0)


def savgol_taps(width, order):
    """Interior FIR taps (applied as sum_k w[k] * x[t + k - width//2])."""
    c = scipy.signal.savgol_coeffs(width, polyorder=order, deriv=order, use="dot")
    return np.asarray(c, dtype=np.float64)
