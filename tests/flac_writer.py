"""TEST INFRASTRUCTURE: a small FLAC encoder (RFC 9639) used to exercise the library's decoder (aad_flac_decode).

No FLAC file and no other FLAC codec exists in this image, so the round trip is the check: this writer is
independent code (Python, bit writer, encoder-side decisions) that emits every construct the decoder handles --
CONSTANT / VERBATIM / FIXED (orders 0-4) / LPC subframes, wasted bits, Rice and Rice2 partitions with escape
partitions, independent / left-side / side-right / mid-side stereo, a short last block, UTF-8 coded frame
numbers, CRC-8 / CRC-16, and the MD5 of the PCM in STREAMINFO (checked by audio_io after decoding).
"""
import hashlib
import struct

import numpy as np


class BitWriter:
    def __init__(self):
        self.buf = bytearray()
        self.acc = 0
        self.n = 0

    def write(self, value, bits):
        if bits == 0:
            return
        value &= (1 << bits) - 1
        self.acc = (self.acc << bits) | value
        self.n += bits
        while self.n >= 8:
            self.n -= 8
            self.buf.append((self.acc >> self.n) & 0xff)
        self.acc &= (1 << self.n) - 1

    def unary(self, q):
        while q >= 32:
            self.write(0, 32)
            q -= 32
        self.write(1, q + 1)

    def align(self):
        if self.n:
            self.write(0, 8 - self.n)

    def bytes(self):
        assert self.n == 0
        return bytes(self.buf)


def crc8(data):
    c = 0
    for b in data:
        c ^= b
        for _ in range(8):
            c = ((c << 1) ^ 0x07) & 0xff if c & 0x80 else (c << 1) & 0xff
    return c


def crc16(data):
    c = 0
    for b in data:
        c ^= b << 8
        for _ in range(8):
            c = ((c << 1) ^ 0x8005) & 0xffff if c & 0x8000 else (c << 1) & 0xffff
    return c


def utf8_number(v):
    if v < 0x80:
        return bytes([v])
    out, n = [], 0
    while True:
        n += 1
        lim = 1 << (6 - n)          # payload bits left in the leading byte for n continuation bytes
        out.append(0x80 | (v & 0x3f))
        v >>= 6
        if v < lim:
            lead = (0xff << (7 - n)) & 0xff
            return bytes([lead | v] + out[::-1])


def write_residual(bw, res, order, blocksize, porder, rice2, escape_first):
    bw.write(1 if rice2 else 0, 2)
    bw.write(porder, 4)
    pbits, esc = (5, 31) if rice2 else (4, 15)
    idx = 0
    for p in range(1 << porder):
        count = blocksize - order if porder == 0 else ((blocksize >> porder) - order if p == 0 else blocksize >> porder)
        part = res[idx:idx + count]
        idx += count
        u = [(2 * int(r)) if r >= 0 else (-2 * int(r) - 1) for r in part]
        if escape_first and p == 0 and count > 0:
            nb = max(int(np.max(np.abs(part))).bit_length() + 1, 1)
            bw.write(esc, pbits)
            bw.write(nb, 5)
            for r in part:
                bw.write(int(r), nb)
            continue
        mean = (sum(u) / max(count, 1)) if count else 0
        k = min(max(int(mean).bit_length() - 1, 0), esc - 1)
        bw.write(k, pbits)
        for x in u:
            bw.unary(x >> k)
            bw.write(x & ((1 << k) - 1), k)
    assert idx == len(res)


def lpc_coefficients(x, order, precision=12):
    """Quantised forward-prediction coefficients from the autocorrelation (Levinson-Durbin)."""
    xf = x.astype(np.float64) * np.hanning(len(x))
    r = np.array([np.dot(xf[:len(xf) - k], xf[k:]) for k in range(order + 1)])
    if r[0] <= 0:
        return None
    a, e = np.zeros(order + 1), r[0]
    a[0] = 1.0
    for i in range(1, order + 1):
        acc = r[i] + np.dot(a[1:i], r[i - 1:0:-1])
        k = -acc / e
        a[1:i + 1] = a[1:i + 1] + k * np.concatenate([a[i - 1:0:-1], [1.0]])
        e *= 1.0 - k * k
        if e <= 0:
            return None
    coef = -a[1:]
    cmax = np.max(np.abs(coef))
    if cmax == 0:
        return None
    shift = min(max(precision - 1 - int(np.floor(np.log2(cmax))) - 1, 0), 15)
    q = np.clip(np.round(coef * (1 << shift)), -(1 << (precision - 1)), (1 << (precision - 1)) - 1).astype(np.int64)
    return q, shift, precision


def write_subframe(bw, x, bps, kind, rng):
    """x: int64 samples of one channel; kind in constant / verbatim / fixedN / lpcN / auto"""
    n = len(x)
    wasted = 0
    if np.any(x != 0):
        low = int(np.bitwise_or.reduce(x.astype(np.int64)))
        while low & 1 == 0 and wasted < bps - 1:
            low >>= 1
            wasted += 1
    if kind == "auto":
        kind = "constant" if np.all(x == x[0]) else rng.choice(["fixed0", "fixed1", "fixed2", "fixed3", "fixed4", "lpc4", "lpc8",
                                                                 "lpc12", "verbatim"])
    if kind == "constant" and not np.all(x == x[0]):
        kind = "fixed2"
    xs = x >> wasted
    b = bps - wasted
    bw.write(0, 1)

    def header(type_code):
        bw.write(type_code, 6)
        if wasted:
            bw.write(1, 1)
            bw.unary(wasted - 1)
        else:
            bw.write(0, 1)

    porder = int(rng.integers(0, 4))
    while porder > 0 and (n % (1 << porder) != 0 or (n >> porder) <= 12):
        porder -= 1
    rice2, escape_first = bool(rng.integers(0, 2)), bool(rng.integers(0, 6) == 0)
    if kind == "constant":
        header(0)
        bw.write(int(xs[0]), b)
    elif kind == "verbatim":
        header(1)
        for v in xs:
            bw.write(int(v), b)
    elif kind.startswith("fixed"):
        order = min(int(kind[5:]), n)
        header(8 + order)
        res = xs.copy()
        for _ in range(order):
            res = np.concatenate([res[:1] * 0, np.diff(res)])   # repeated differencing = the fixed predictors
        for v in xs[:order]:
            bw.write(int(v), b)
        write_residual(bw, res[order:], order, n, porder, rice2, escape_first)
    else:
        order = min(int(kind[3:]), n - 1, 32)
        q = lpc_coefficients(xs, order) if order > 0 else None
        if q is None:
            return write_subframe_retry(bw, x, bps, rng)
        coef, shift, prec = q
        header(32 + order - 1)
        for v in xs[:order]:
            bw.write(int(v), b)
        bw.write(prec - 1, 4)
        bw.write(shift, 5)
        for c in coef:
            bw.write(int(c), prec)
        res = np.empty(n - order, dtype=np.int64)
        for i in range(order, n):
            pred = int(np.dot(coef, xs[i - order:i][::-1])) >> shift
            res[i - order] = int(xs[i]) - pred
        write_residual(bw, res, order, n, porder, rice2, escape_first)


def write_subframe_retry(bw, x, bps, rng):
    # the "0" padding bit was already written by the caller's attempt: emit a FIXED-2 body
    n = len(x)
    wasted = 0
    bw.write(8 + min(2, n), 6)
    bw.write(0, 1)
    order = min(2, n)
    res = x.copy()
    for _ in range(order):
        res = np.concatenate([res[:1] * 0, np.diff(res)])
    for v in x[:order]:
        bw.write(int(v), bps)
    write_residual(bw, res[order:], order, n, 0, False, False)


def encode(pcm, sample_rate, bps=16, blocksize=4096, seed=0, force=None):
    """pcm: int array [n] or [n, channels] -> bytes of a FLAC stream.  `force`: subframe kind for every subframe."""
    rng = np.random.default_rng(seed)
    pcm = np.asarray(pcm)
    if pcm.ndim == 1:
        pcm = pcm[:, None]
    n, nch = pcm.shape
    x = pcm.astype(np.int64)
    width = (bps + 7) // 8
    raw = b"".join(int(v).to_bytes(width, "little", signed=True) for v in x.reshape(-1)) if width != 2 else \
        x.astype("<i2").tobytes()
    md5 = hashlib.md5(raw).digest()
    frames = bytearray()
    fmin, fmax, frame_no = 1 << 24, 0, 0
    for start in range(0, n, blocksize):
        blk = x[start:start + blocksize]
        bs = len(blk)
        bw = BitWriter()
        bw.write(0b11111111111110, 14)
        bw.write(0, 1)
        bw.write(0, 1)                                   # fixed block size stream: frame number follows
        if bs == 4096:
            bw.write(0b1100, 4)
        elif bs == 192:
            bw.write(0b0001, 4)
        elif bs <= 256:
            bw.write(0b0110, 4)
        else:
            bw.write(0b0111, 4)
        sr_codes = {8000: 4, 16000: 5, 22050: 6, 24000: 7, 32000: 8, 44100: 9, 48000: 10, 96000: 11}
        sr_code = sr_codes.get(sample_rate, 13 if sample_rate < 65536 else 0)
        bw.write(sr_code, 4)
        mode = 0
        if nch == 2:
            mode = [1, 8, 9, 10][frame_no % 4]           # independent, left/side, side/right, mid/side
            bw.write(mode, 4)
        else:
            bw.write(nch - 1, 4)
        bw.write({8: 1, 12: 2, 16: 4, 20: 5, 24: 6}[bps] if frame_no % 2 else 0, 3)   # alternately "see STREAMINFO"
        bw.write(0, 1)
        for byte in utf8_number(frame_no):
            bw.write(byte, 8)
        if bs not in (4096, 192):
            bw.write(bs - 1, 8 if bs <= 256 else 16)
        if sr_code == 13:
            bw.write(sample_rate, 16)
        bw.write(crc8(bytes(bw.buf)), 8)
        chans = [blk[:, c] for c in range(nch)]
        widths = [bps] * nch
        if nch == 2 and mode == 8:
            chans, widths = [chans[0], chans[0] - chans[1]], [bps, bps + 1]
        elif nch == 2 and mode == 9:
            chans, widths = [chans[0] - chans[1], chans[1]], [bps + 1, bps]
        elif nch == 2 and mode == 10:
            chans, widths = [(chans[0] + chans[1]) >> 1, chans[0] - chans[1]], [bps, bps + 1]
        for ch, w in zip(chans, widths):
            write_subframe(bw, ch, w, force or "auto", rng)
        bw.align()
        body = bw.bytes()
        frame = body + struct.pack(">H", crc16(body))
        fmin, fmax = min(fmin, len(frame)), max(fmax, len(frame))
        frames += frame
        frame_no += 1
    si = BitWriter()
    si.write(blocksize, 16)
    si.write(blocksize, 16)
    si.write(fmin if frames else 0, 24)
    si.write(fmax, 24)
    si.write(sample_rate, 20)
    si.write(nch - 1, 3)
    si.write(bps - 1, 5)
    si.write(n, 36)
    info = si.bytes() + md5
    pad = b"\x81" + (8).to_bytes(3, "big") + b"\x00" * 8        # a PADDING block after STREAMINFO (last = 1)
    return b"fLaC" + b"\x00" + len(info).to_bytes(3, "big") + info + pad + bytes(frames)
