"""CPU: host-side logic -- frame arithmetic, sharding, WAV decode, bench accounting, gloo gather."""
import os
import sys
import wave

import numpy as np
import pytest
import torch

from audioanalysisdetector_b200.frontend import FrontendParams
from audioanalysisdetector_b200 import audio_io, sharding
from oracle import librosa_ref as LR, spafe_ref as SR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frame_counts_match_oracle():
    mf, lf = FrontendParams.mfcc(16000), FrontendParams.lfcc(16000)
    for n in (0, 1, 159, 160, 399, 400, 401, 511, 512, 513, 2047, 2048, 2049, 32000, 64000):
        assert mf.n_frames(n) == (LR.n_frames_centered(n, 512) if n > 0 else 0)
        assert lf.n_frames(n) == SR.n_frames_uncentered(n, 400, 160)
    assert FrontendParams.mfcc(16000, n_mfcc=40, n_delta=2).c_out == 120
    assert FrontendParams.logmel(16000).c_out == 64
    assert FrontendParams.lfcc(16000, n_ceps=20, nfilts=20, n_delta=2).c_out == 60


def test_partition_by_frames_balanced_and_complete():
    rng = np.random.default_rng(0)
    nf = rng.integers(99, 800, size=4096)
    for ws in (1, 2, 4, 8):
        parts = sharding.partition_by_frames(nf, ws)
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(4096))
        loads = np.array([nf[p].sum() for p in parts])
        assert loads.max() - loads.min() <= nf.max()
    sl = [sharding.contiguous_shard(10, r, 4) for r in range(4)]
    assert sum(s.stop - s.start for s in sl) == 10 and sl[0] == slice(0, 3) and sl[3] == slice(9, 10)


def test_wav_loader_round_trip(tmp_path):
    y = (np.sin(np.arange(8000) * 0.05) * 0.5).astype(np.float32)
    path = str(tmp_path / "a.wav")
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(16000)
        w.writeframes((y * 32767).astype("<i2").tobytes())
    z, sr = audio_io.load(path)
    assert sr == 16000 and z.dtype == np.float32 and np.abs(z - y).max() < 1e-4
    with pytest.raises(ValueError):
        audio_io.load(path, sr=8000)
    z2, sr2 = audio_io.load((y, 22050))
    assert sr2 == 22050 and z2 is not None


def test_bench_algorithmic_work_matches_survey():
    sys.path.insert(0, ROOT)
    import bench
    w = bench.algorithmic_work(2048, 512, 128, 40, 2, 2020, 120)
    assert round(w["flops"]) == 77291 and w["bytes"] == 2528
    w = bench.algorithmic_work(512, 160, 80, 0, 0, 500, 80)
    assert round(w["flops"]) == 13883 and w["bytes"] == 960           # config C1


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_total = 37
    nf = (np.arange(n_total) * 7919 % 500) + 9
    parts = sharding.partition_by_frames(nf, world)
    idx = torch.from_numpy(parts[rank])
    local = torch.stack([torch.full((3, 5), float(i)) for i in parts[rank]]) if len(idx) else torch.zeros((0, 3, 5))
    full = sharding.gather_features(local, idx, n_total)
    ok = all(bool((full[i] == float(i)).all()) for i in range(n_total))
    q.put((rank, ok, int(nf[parts[rank]].sum())))
    dist.destroy_process_group()


def test_sharded_gather_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    loads = [l for _, _, l in res]
    assert abs(loads[0] - loads[1]) <= 509


def test_scaler_merge_matches_sklearn():
    """merge_stats == sklearn StandardScaler on np.vstack(...) (ASV_dl_func.py:1113-1129), including a
    constant column (scale 1) -- the host half of DeviceStandardScaler.fit."""
    from sklearn.preprocessing import StandardScaler
    from audioanalysisdetector_b200.scaler import merge_stats
    rng = np.random.default_rng(5)
    utts = [(-60 + 15 * rng.standard_normal((64, 63))).astype(np.float32) for _ in range(7)]
    for u in utts:
        u[:, 10] = -80.0
    X = np.vstack(utts).astype(np.float64)
    ref = StandardScaler().fit(X)
    mean, var, scale = merge_stats(X.shape[0], X.sum(0), (X * X).sum(0))
    np.testing.assert_allclose(mean, ref.mean_, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(var, ref.var_, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(scale, ref.scale_, rtol=1e-9, atol=1e-9)
    assert scale[10] == 1.0


def _gloo_scaler_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(11)
    X = rng.standard_normal((1000, 13))
    part = X[sharding.contiguous_shard(1000, rank, world)]
    stats = torch.from_numpy(np.concatenate([part.sum(0), (part * part).sum(0), [float(len(part))]]))
    dist.all_reduce(stats, op=dist.ReduceOp.SUM)         # the exchange step of DeviceStandardScaler.fit
    from audioanalysisdetector_b200.scaler import merge_stats
    h = stats.numpy()
    mean, var, scale = merge_stats(int(round(h[26])), h[:13], h[13:26])
    q.put((rank, float(np.abs(mean - X.mean(0)).max()), float(np.abs(var - X.var(0)).max())))
    dist.destroy_process_group()


def test_scaler_statistics_all_reduce_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_gloo_scaler_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] < 1e-12 and r[2] < 1e-12 for r in res)


def test_chunk_table_follows_the_reference_slicing():
    """chunk_bounds == y[int(cs*sr):min(int(ce*sr), len(y))] (ASV_dl_func.py:407-410) for every case,
    two_second_chunks == prepare_dataframe's rows (:281-293), layout_files keeps files aligned."""
    from audioanalysisdetector_b200.corpus import FILE_ALIGN, chunk_bounds, layout_files, two_second_chunks
    y = np.arange(70001)
    for sr in (16000, 22050, 44100):
        for cs, ce in [(0.0, 2.0), (2.0, 4.0), (1.3, 2.9), (4.0, 6.0), (100.0, 102.0), (0.0, 0.0), (3.0, 2.0),
                       (-1.0, 1.0), (None, None), (float("nan"), float("nan"))]:
            if cs is None or cs != cs:
                want = y
            else:
                want = y[int(cs * sr):min(int(ce * sr), len(y))]
            s, e = chunk_bounds(len(y), sr, cs, ce)
            assert e - s == len(want) and (len(want) == 0 or (y[s] == want[0] and y[e - 1] == want[-1]))
    assert two_second_chunks(5 * 16000 + 900, 16000) == [(0.0, 2.0), (2.0, 4.0)]
    assert two_second_chunks(2 * 16000, 16000) == [(0.0, 2.0)]
    assert two_second_chunks(2 * 16000 - 1, 16000) == []
    base, total = layout_files([5, 16, 0, 17])
    assert base == [0, 8, 24, 24] and total == 48 and all(b % FILE_ALIGN == 0 for b in base)
    assert layout_files([])[1] == FILE_ALIGN


def test_load_pcm_keeps_int16_and_equals_the_float_decode(tmp_path):
    import wave
    from audioanalysisdetector_b200 import audio_io
    rng = np.random.default_rng(3)
    pcm = rng.integers(-32768, 32767, size=5000, dtype=np.int16)
    p = tmp_path / "a.wav"
    with wave.open(str(p), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(22050)
        w.writeframes(pcm.astype("<i2").tobytes())
    raw, sr = audio_io.load_pcm(str(p))
    y, sr2 = audio_io.load(str(p))
    assert raw.dtype == np.int16 and sr == sr2 == 22050 and np.array_equal(raw, pcm)
    assert y.dtype == np.float32 and np.array_equal(raw.astype(np.float32) / 32768.0, y)
    arr, sr3 = audio_io.load_pcm((y, 22050))                              # in-memory clips stay float32
    assert arr.dtype == np.float32 and sr3 == 22050


def test_time_split_covers_every_frame_with_its_own_samples():
    """sharding.time_split (long-form audio over several GPUs): the pieces own disjoint frame
    ranges that cover the utterance, and local frame skip + i of a piece reads exactly the samples global
    frame t0 + i reads (windows clipped at the true signal ends only)."""
    for length, n_fft, hop in [(28_800_000, 2048, 480), (64000, 2048, 512), (64000, 512, 160), (5000, 2048, 512),
                               (479, 2048, 480), (100_001, 1024, 256)]:
        T = 1 + length // hop
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                p = sharding.time_split(length, n_fft, hop, r, world)
                seen += list(range(p["t0"], p["t1"]))
                if p["t1"] == p["t0"]:
                    continue
                assert 0 <= p["s0"] <= p["s1"] <= length and p["s0"] % hop == 0
                T_local = 1 + (p["s1"] - p["s0"]) // hop
                assert p["skip"] + (p["t1"] - p["t0"]) <= T_local
                for i in (0, p["t1"] - p["t0"] - 1):
                    g_lo, g_hi = (p["t0"] + i) * hop - n_fft // 2, (p["t0"] + i) * hop + n_fft // 2
                    l_lo = p["s0"] + (p["skip"] + i) * hop - n_fft // 2
                    assert l_lo == g_lo                                       # same window position
                    # the part of the window inside the true signal is inside the slice
                    assert max(g_lo, 0) >= p["s0"] and min(g_hi, length) <= p["s1"]
            assert seen == list(range(T))


def test_detector_weight_packing_validates_the_state_dict():
    """DetectorEngine's host half: the reference's parameter names and default shapes (cnn_bilstm_hybrid.py:20-52),
    pointers into float32 host copies that stay alive; anything else is rejected before the C ABI sees it."""
    import ctypes as C
    from audioanalysisdetector_b200 import AadError
    from audioanalysisdetector_b200.detector import _NAMES, _SHAPES, pack_weights
    from test_oracle_consumer import load_fixture
    _, weights = load_fixture()
    w, keep = pack_weights(weights, 13)
    assert w.struct_size == C.sizeof(type(w)) and w.feature_dim == 13 and len(keep) == len(_NAMES)
    assert all(a.dtype == np.float32 and a.flags["C_CONTIGUOUS"] for a in keep)
    assert w.conv_w[5] == np.float32(np.asarray(weights["feature_extractor.0.weight"]).reshape(-1)[5])
    bad = dict(weights)
    del bad["bilstm.weight_hh_l0_reverse"]
    with pytest.raises(AadError):
        pack_weights(bad, 13)
    bad = dict(weights)
    bad["classifier.0.weight"] = np.zeros((32, 64), np.float32)                    # dense_units != 64
    with pytest.raises(AadError):
        pack_weights(bad, 13)
    assert set(_SHAPES) == set(_NAMES)


def test_device_feature_loader_feeds_a_train_loop_like_the_reference():
    """DeviceFeatureLoader stands in for CQCCDataset + DataLoader (cnn_bilstm_hybrid.py:4-15) in train_loop
    (ASV_dl_func.py:751-829): (X [n, F, 63], y [n, 1]) batches, every kept item once per epoch, failed items dropped;
    one epoch of the reference's loop body (restated: BCE on a sigmoid output) runs on them and learns."""
    import torch
    from audioanalysisdetector_b200 import DeviceFeatureLoader
    from oracle import consumer_ref
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "consumer.npz"))
    weights = {k[3:]: torch.from_numpy(g[k]).clone() for k in g.files if k.startswith("w::")}
    for k, w in weights.items():
        if w.dtype.is_floating_point and "running" not in k:
            w.requires_grad_(True)
    rng = np.random.default_rng(0)
    N = 45
    labels = rng.integers(0, 2, size=N)
    feats = torch.from_numpy(rng.standard_normal((N, 13, 63)).astype(np.float32))
    feats[labels == 1] += 0.8                                       # separable on purpose
    status = torch.zeros(N, dtype=torch.int32)
    status[[3, 17]] = 2                                             # extractor returned None for these
    loader = DeviceFeatureLoader(feats, labels, batch_size=16, shuffle=True, status=status, seed=1)
    assert len(loader) == 3 and loader.dataset_size == N - 2
    seen = []
    for X, y in loader:
        assert X.shape[1:] == (13, 63) and X.dtype == torch.float32 and y.shape == (X.shape[0], 1) and y.dtype == torch.float32
        seen.append(X[:, 0, 0])
    got = torch.sort(torch.cat(seen)).values
    keep = [i for i in range(N) if i not in (3, 17)]
    assert torch.equal(got, torch.sort(feats[keep, 0, 0]).values)   # every kept item exactly once
    flat = DeviceFeatureLoader(feats, labels, batch_size=50, label_shape="flat")
    (Xa, ya), = list(flat)
    assert ya.dtype == torch.int64 and ya.shape == (N,) and torch.equal(Xa, feats)
    # the loop body of train_loop: forward, BCE (the model ends in a sigmoid), backward, step
    params = [w for w in weights.values() if w.requires_grad]
    opt = torch.optim.Adam(params, lr=3e-3)
    crit = torch.nn.BCELoss()

    def epoch():
        total, n = 0.0, 0
        for X, y in loader:
            opt.zero_grad()
            out = consumer_ref.forward(weights, X)
            loss = crit(out.reshape(-1, 1), y)
            loss.backward()
            opt.step()
            total, n = total + float(loss) * len(y), n + len(y)
        return total / n
    first = epoch()
    for _ in range(6):
        last = epoch()
    assert last < first
