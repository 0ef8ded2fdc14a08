"""Dev tool: files in, features out.  A synthetic corpus of N mono 16-bit FLAC files of 4 s (one encoded file copied N
times: the decoder's work is the same for every copy) -> DeviceCorpus (header scan, threaded decode, streamed upload)
-> the notebook's 5-feature map over its 2-second chunks.  Reports where the wall time goes."""
import json, os, sys, time, tempfile, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import audioanalysisdetector_b200 as aad
from audioanalysisdetector_b200.frontend import Frontend, FrontendParams
import flac_writer
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
sr = 16000
rng = np.random.default_rng(0)
t = np.arange(4 * sr) / sr
y = 0.3 * np.sin(2 * np.pi * 220 * t) * (1 + 0.5 * np.sin(2 * np.pi * 3 * t)) + 0.01 * rng.standard_normal(len(t))
data = flac_writer.encode(np.round(y * 32767).astype(np.int64), sr, force="lpc8")
d = tempfile.mkdtemp(prefix="aad_files_")
try:
    paths = []
    for i in range(N):
        p = os.path.join(d, f"LA_{i:06d}.flac")
        with open(p, "wb") as f:
            f.write(data)
        paths.append(p)
    dev = torch.device("cuda:0")
    torch.zeros(1, device=dev)
    res = {"files": N, "seconds_of_audio": 4 * N, "flac_bytes_per_pcm_byte": len(data) / (8 * sr), "host_cores": os.cpu_count()}
    for threads in (1, None):
        t0 = time.time()
        corpus = aad.DeviceCorpus(dev)
        for p in paths:
            corpus.add(p)
        t1 = time.time()
        pcm = corpus.upload(decode_threads=threads)
        torch.cuda.synchronize()
        t2 = time.time()
        res[f"decode_threads={threads}"] = {"header_scan_s": t1 - t0, "decode_and_upload_s": t2 - t1,
                                            "audio_hours_per_s": 4 * N / 3600 / (t2 - t0)}
    rows = [(i, s, s + 2.0) for i in range(N) for s in (0.0, 2.0)]
    off, ln = corpus.table(rows)
    cq = aad.CqccFrontend(sr, device=dev)
    fes = {"gtcc": Frontend(FrontendParams.gtcc(sr), dev), "lfcc": Frontend(FrontendParams.lfcc(sr), dev)}
    mf, ml = Frontend(FrontendParams.mfcc(sr, n_mfcc=13), dev), Frontend(FrontendParams.logmel(sr, n_mels=64), dev)
    o = torch.from_numpy(off).to(dev); l = torch.from_numpy(ln).to(dev)
    def features():
        cq.extract_indexed(pcm, o, l, max_len=2 * sr)
        for fe in fes.values():
            corpus.extract(fe, off, ln)
        corpus.extract_pair(mf, ml, off, ln)
    features(); torch.cuda.synchronize()
    t0 = time.time(); features(); torch.cuda.synchronize()
    res["features_5_map_s"] = time.time() - t0
    res["chunks"] = len(rows)
    print(json.dumps(res))
finally:
    shutil.rmtree(d, ignore_errors=True)
