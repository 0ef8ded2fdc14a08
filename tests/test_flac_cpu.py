"""CPU: the library's FLAC decoder (aad_flac_decode, host code) against the independent test encoder
(tests/flac_writer.py): every subframe type, wasted bits, Rice / Rice2 / escape partitions, the four stereo
modes, short last block, CRC and MD5 checks."""
import numpy as np
import pytest

import flac_writer as FW


@pytest.fixture(scope="module")
def aio(built_lib):
    from audioanalysisdetector_b200 import audio_io
    return audio_io


def _speechlike(n, seed, sr=16000):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / sr
    y = 0.3 * np.sin(2 * np.pi * 140 * t) * (1 + 0.5 * np.sin(2 * np.pi * 3 * t)) + 0.05 * np.sin(2 * np.pi * 2300 * t)
    return np.round((y + 0.01 * rng.standard_normal(n)) * 32767 * 0.8).astype(np.int64)


@pytest.mark.parametrize("force", [None, "verbatim", "fixed0", "fixed1", "fixed2", "fixed3", "fixed4", "lpc1", "lpc8", "lpc32"])
def test_mono_16bit_round_trip(aio, force):
    x = _speechlike(3 * 4096 + 777, 1)
    x[5000:5000 + 4096 * 0 + 300] = 123                                  # a constant stretch inside a block
    data = FW.encode(x, 16000, bps=16, blocksize=4096, seed=3, force=force)
    pcm, sr, bps = aio.decode_flac(data)
    assert (sr, bps) == (16000, 16) and pcm.shape == (len(x), 1)
    np.testing.assert_array_equal(pcm[:, 0], x)


def test_constant_blocks_wasted_bits_and_small_blocks(aio):
    x = np.concatenate([np.zeros(192, np.int64), np.full(192, -7, np.int64), _speechlike(1000, 2) // 8 * 8,
                        np.array([32767, -32768, 0, 1, -1], np.int64)])
    for bs in (192, 256, 1000):
        pcm, _, _ = aio.decode_flac(FW.encode(x, 22050, blocksize=bs, seed=bs))
        np.testing.assert_array_equal(pcm[:, 0], x)


@pytest.mark.parametrize("bps", [16, 24])
def test_stereo_modes_and_bit_depths(aio, bps):
    n = 4 * 1152 + 100
    l = _speechlike(n, 4) * (1 if bps == 16 else 200)
    r = (0.7 * l).astype(np.int64) + _speechlike(n, 5) // 16
    x = np.stack([l, r], axis=1)
    pcm, sr, got_bps = aio.decode_flac(FW.encode(x, 44100, bps=bps, blocksize=1152, seed=7))   # 4 frames: all four modes
    assert got_bps == bps and sr == 44100
    np.testing.assert_array_equal(pcm, x)


def test_corruption_is_detected(aio, built_lib):
    from audioanalysisdetector_b200 import _lib as L
    x = _speechlike(9000, 6)
    data = bytearray(FW.encode(x, 16000, seed=1))
    bad = bytearray(data)
    bad[len(bad) // 2] ^= 0x10                                           # a bit flip inside a frame: CRC-16 fails
    with pytest.raises(L.AadError, match="malformed"):
        aio.decode_flac(bytes(bad))
    with pytest.raises(L.AadError):
        aio.decode_flac(b"RIFF" + bytes(data[4:]))
    md5_off = 4 + 4 + 18
    bad = bytearray(data)
    bad[md5_off] ^= 0xff                                                 # frames intact, STREAMINFO MD5 wrong
    with pytest.raises(ValueError, match="MD5"):
        aio.decode_flac(bytes(bad))


def test_load_and_info_like_librosa_and_soundfile(aio, tmp_path):
    x = _speechlike(40000, 8)
    p = tmp_path / "LA_T_1000001.flac"
    p.write_bytes(FW.encode(x, 16000, seed=2))
    assert aio.info(str(p)) == (40000, 16000)                            # soundfile.info (ASV_dl_func.py:280)
    y, sr = aio.load(str(p))                                             # librosa.load(sr=None)
    assert sr == 16000 and y.dtype == np.float32
    np.testing.assert_array_equal(y, (x / 32768.0).astype(np.float32))
    pcm, sr = aio.load_pcm(str(p))
    assert pcm.dtype == np.int16
    np.testing.assert_array_equal(pcm, x.astype(np.int16))
    st = np.stack([x, x[::-1]], axis=1)
    q = tmp_path / "stereo.flac"
    q.write_bytes(FW.encode(st, 16000, seed=3))
    y2, _ = aio.load(str(q))
    np.testing.assert_allclose(y2, ((st[:, 0] + st[:, 1]) / 2 / 32768.0).astype(np.float32), atol=1e-7)


@pytest.mark.parametrize("force", [None, "lpc8", "lpc12", "lpc32", "fixed3"])
def test_pcm16_one_call_path_with_md5_in_the_library(aio, force):
    """aad_flac_decode_pcm16 (what the corpus upload fans out over threads): same samples as the generic decoder, the
    MD5 of STREAMINFO checked in C (hashlib is the independent reference for the digest), unrolled LPC orders and the
    generic one (32)."""
    import hashlib
    x = _speechlike(2 * 4096 + 1234, 9)
    data = FW.encode(x, 22050, bps=16, blocksize=4096, seed=5, force=force)
    y, sr = aio.decode_flac_pcm16(data)
    assert sr == 22050 and y.dtype == np.int16
    np.testing.assert_array_equal(y, x)
    assert hashlib.md5(y.astype("<i2").tobytes()).digest() == data[8 + 18:8 + 34]       # STREAMINFO's MD5 field
    bad = bytearray(data)
    bad[8 + 20] ^= 0x5a                                                               # a wrong stored checksum
    with pytest.raises(ValueError, match="MD5"):
        aio.decode_flac_pcm16(bytes(bad))
    aio.set_flac_md5(False)
    try:
        np.testing.assert_array_equal(aio.decode_flac_pcm16(bytes(bad))[0], x)        # CRCs still pass: accepted
    finally:
        aio.set_flac_md5(True)
    stereo = FW.encode(np.stack([x, -x], axis=1), 16000, bps=16, blocksize=4096, seed=6)
    assert aio.decode_flac_pcm16(stereo) is None                                      # not the corpus format: generic path
    assert aio.decode_flac_pcm16(FW.encode(x // 4, 16000, bps=12, blocksize=1024, seed=7)) is None


def test_crc16_slices_agree_with_the_bytewise_definition(aio):
    """Frames of every length modulo 8 pass their CRC-16 (slicing-by-8 with a byte-wise tail) -- and a flipped bit in
    the last byte of a frame body is caught."""
    for n in range(4100, 4116):
        x = _speechlike(n, n)
        data = FW.encode(x, 16000, bps=16, blocksize=4096, seed=n, force="verbatim")
        np.testing.assert_array_equal(aio.decode_flac_pcm16(data)[0], x)
    bad = bytearray(data)
    bad[-3] ^= 1
    with pytest.raises(Exception):
        aio.decode_flac_pcm16(bytes(bad))
