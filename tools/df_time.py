"""Dev tool: the reference-facing call itself -- extract_features(DataFrame of chunk rows, the notebook's five-extractor
map) -- on a synthetic corpus of N 4-second FLAC files (two 2-second chunk rows each).  Wall time, split by stage."""
import cProfile, io, json, os, pstats, sys, tempfile, shutil, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, pandas as pd, torch
import audioanalysisdetector_b200 as aad
import flac_writer
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
sr = 16000
rng = np.random.default_rng(0)
t = np.arange(4 * sr) / sr
y = 0.3 * np.sin(2 * np.pi * 220 * t) * (1 + 0.5 * np.sin(2 * np.pi * 3 * t)) + 0.01 * rng.standard_normal(len(t))
data = flac_writer.encode(np.round(y * 32767).astype(np.int64), sr, force="lpc8")
d = tempfile.mkdtemp(prefix="aad_df_")
try:
    rows = []
    for i in range(N):
        p = os.path.join(d, f"LA_{i:06d}.flac")
        with open(p, "wb") as f:
            f.write(data)
        rows += [{"filepath": p, "chunk_start": 0.0, "chunk_end": 2.0}, {"filepath": p, "chunk_start": 2.0, "chunk_end": 4.0}]
    df = pd.DataFrame(rows)
    fmap = {"cqcc": aad.extract_cqcc, "gtcc": aad.extract_gtcc, "mel-spect": aad.extract_mel_spectrogram,
            "mfcc": aad.extract_mfcc, "lfcc": aad.extract_lfcc}
    aad.extract_features(df.iloc[:64].copy(), fmap)          # warm-up: plans, kernels, pinned buffers
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    t0 = time.time()
    pr.enable()
    out = aad.extract_features(df.copy(), fmap)
    pr.disable()
    wall = time.time() - t0
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
    print(s.getvalue()[:3500])
    print(json.dumps({"files": N, "rows": len(df), "features": list(fmap), "wall_s": wall,
                      "audio_hours_per_s": len(df) * 2 / 3600 / wall,
                      "shapes": {k: list(np.shape(out[k].iloc[0])) for k in fmap}}))
finally:
    shutil.rmtree(d, ignore_errors=True)
