"""CPU emulation of k_stft_fb's per-warp index math (lanes, transpose, partner shuffles).

Development aid: checks the 32 x L decomposition + real-input split used by the CUDA
kernel against np.fft.rfft for every supported L, without a GPU.
"""
import numpy as np


def bitrev(x, bits):
    r = 0
    for i in range(bits):
        r |= ((x >> i) & 1) << (bits - 1 - i)
    return r


def emulate(L, frames):
    """frames: (Q, N) real (already windowed, unscaled). returns (Q, M+1) power."""
    Q, M = 32 // L, 32 * L
    N = 2 * M
    log2L = L.bit_length() - 1
    v = np.zeros((32, 32), dtype=np.complex128)  # [lane][reg]
    # pass 1 load: lane (g, j): z[L*a + j] at reg a (natural; DFT done by numpy here)
    for lane in range(32):
        g, j = divmod(lane, L)
        x = 0.5 * frames[g]
        z = x[0::2] + 1j * x[1::2]
        col = np.array([z[L * a + j] for a in range(32)])
        Y = np.fft.fft(col)  # natural order kA
        tw = np.exp(-2j * np.pi * j * np.arange(32) / M)
        v[lane] = Y * tw
    # transpose via scratch [kA*33 + lane]
    scr = np.zeros(32 * 33, dtype=np.complex128)
    for lane in range(32):
        for ka in range(32):
            scr[ka * 33 + lane] = v[lane, ka]
    v2 = np.zeros_like(v)
    for lane in range(32):
        g, j = divmod(lane, L)
        for q in range(Q):
            for bb in range(L):
                v2[lane, q * L + bb] = scr[(j + L * q) * 33 + g * L + bb]  # natural b order
    # pass 2: DFT_L over b for each q
    for lane in range(32):
        for q in range(Q):
            v2[lane, q * L:(q + 1) * L] = np.fft.fft(v2[lane, q * L:(q + 1) * L])
    # post-process
    P = np.full((Q, M + 1), np.nan)
    for q in range(Q):
        for s in range(L // 2):
            GEN = (Q - 1 - q) * L + (L - 1 - s)
            ALT = ((L - s) % L) if q == 0 else (Q - q) * L + (L - 1 - s)
            snd = np.array([v2[lane, ALT] if lane % L == 0 else v2[lane, GEN] for lane in range(32)])
            for lane in range(32):
                g, j = divmod(lane, L)
                partner = g * L + ((L - j) & (L - 1))
                r = snd[partner]
                A = v2[lane, q * L + s]
                Bc = np.conj(r)
                E, O = A + Bc, A - Bc
                k = j + L * q + 32 * s
                w = np.exp(-2j * np.pi * k / N)
                T = 1j * w * O
                X1, X2 = E - T, E + T
                assert np.isnan(P[g, k]) or k == M - k, (L, lane, q, s, k)
                P[g, k] = abs(X1) ** 2
                assert np.isnan(P[g, M - k]) or k == M - k or (M - k) == k, (L, lane, q, s, M - k)
                P[g, M - k] = abs(X2) ** 2
    for lane in range(32):
        g, j = divmod(lane, L)
        if j == 0:
            A = v2[lane, L // 2]
            assert np.isnan(P[g, M // 2])
            P[g, M // 2] = 4 * abs(A) ** 2
    return P


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for L in (4, 8, 16, 32):
        Q, N = 32 // L, 64 * L
        fr = rng.standard_normal((Q, N))
        P = emulate(L, fr)
        ref = np.abs(np.fft.rfft(fr, axis=1)) ** 2
        assert not np.isnan(P).any(), L
        err = np.abs(P - ref).max() / ref.max()
        print(f"L={L:2d} n_fft={N:4d} rel err {err:.2e}")
        assert err < 1e-12
    print("warp FFT index math OK")
