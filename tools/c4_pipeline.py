#!/usr/bin/env python
"""BASELINE.json configs[3]: ASVspoof-2019-LA-sized synthetic corpus (25 380 two-second chunks @16 kHz)
-> MFCC-13 and log-mel-64 on the GPU -> CNN-BiLSTM inference on the same device, sharded by utterance
across the GPUs of one box (no collective in the extraction; one all-gather of the scores at the end).

    python tools/c4_pipeline.py                          # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/c4_pipeline.py

The consumer is the reference's AudioDeepfakeDetector (cnn_bilstm_hybrid.py:20-68) with the committed
fixture weights, run by the hand-written kernels of DetectorEngine (csrc/aad_detector.cu); the same model as
stock PyTorch ops (oracle/consumer_ref.py, the restatement the tests pin to the real reference class) is timed
beside it and checks the scores.  This script is an example / measurement tool, not a product entry point.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist

import audioanalysisdetector_b200 as aad
from audioanalysisdetector_b200 import Frontend, FrontendParams
from oracle import consumer_ref  # example only: the product package never imports oracle/

N_CHUNKS, SR, CHUNK = 25380, 16000, 32000


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    sl = aad.contiguous_shard(N_CHUNKS, rank, world)
    n_local = sl.stop - sl.start
    gen = torch.Generator(device=dev)
    gen.manual_seed(4242 + rank)
    wav = (0.1 * torch.randn((n_local, CHUNK), generator=gen, device=dev)).clamp_(-1, 1)
    g = np.load(os.path.join(ROOT, "tests", "golden", "consumer.npz"))
    weights = {k[3:]: torch.from_numpy(g[k]).to(dev) for k in g.files if k.startswith("w::")}
    fe_mfcc = Frontend(FrontendParams.mfcc(SR, n_mfcc=13), dev)
    fe_mel = Frontend(FrontendParams.logmel(SR, n_mels=64), dev)
    engine = aad.DetectorEngine(weights, feature_dim=13, device=dev)

    def step():
        (feats, mel), nf, st = fe_mfcc.extract_pair(fe_mel, wav)   # (n, 13, 63) and (n, 64, 63) from ONE STFT
        scores = engine(feats)                       # hand-written CNN-BiLSTM inference, features stay where they are
        with torch.no_grad():
            ref = torch.cat([consumer_ref.forward(weights, feats[i:i + 4096]) for i in range(0, n_local, 4096)])
        return feats, mel, scores, st, ref

    for _ in range(2):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    f_sep, _, _ = fe_mfcc(wav)                       # the same two features as two separate extractions
    m_sep, _, _ = fe_mel(wav)
    eb.record()
    e0.record()
    (feats, mel), nf, st = fe_mfcc.extract_pair(fe_mel, wav)
    e1.record()
    scores = engine(feats)
    e2.record()
    with torch.no_grad():
        ref = torch.cat([consumer_ref.forward(weights, feats[i:i + 4096]) for i in range(0, n_local, 4096)])
    e3.record()
    torch.cuda.synchronize(dev)
    t_feat, t_model, t_torch = e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3)
    t_sep = ea.elapsed_time(eb)
    pair_equal = float(torch.equal(feats, f_sep) and torch.equal(mel, m_sep))
    max_diff = float((scores - ref).abs().max().item())
    idx = torch.arange(sl.start, sl.stop, device=dev)
    all_scores = aad.gather_features(scores, idx, N_CHUNKS)          # one all-gather, outside the extraction
    t = torch.tensor([t_feat, t_model, t_torch, max_diff, t_sep, -pair_equal], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        hours = N_CHUNKS * CHUNK / SR / 3600.0
        print(json.dumps({
            "config": "configs[3]: 25 380 x 2 s @16 kHz -> MFCC-13 + log-mel-64 -> CNN-BiLSTM scores",
            "n_gpus": world, "chunks_per_gpu": n_local, "status_nonzero": int(st.ne(0).sum().item()),
            "features_ms": float(t[0]), "features_as_two_extractions_ms": float(t[4]),
            "paired_features_equal_separate_bitwise": bool(float(t[5]) == -1.0),
            "model_ms": float(t[1]), "model_stock_pytorch_ms": float(t[2]),
            "scores_max_abs_diff_vs_stock_pytorch": float(t[3]),
            "features_audio_hours_per_s": hours / (float(t[0]) * 1e-3),
            "pipeline_audio_hours_per_s": hours / ((float(t[0]) + float(t[1])) * 1e-3),
            "scores_shape": list(all_scores.shape), "scores_mean": float(all_scores.mean().item()),
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
