"""PCIe / host-memory ceiling probe (dev tool).

  python tools/h2d_probe.py                                   # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_probe.py

Every rank copies between its own pinned host buffers and its GPU at the same time as the others (barrier in
front of every measurement); rank 0 prints per-GPU and aggregate GB/s for H2D only, D2H only and both directions,
with ordinary pinned memory and with write-combined pinned memory as the H2D source.  Says whether the e2e path of
bench.py is limited by one GPU's PCIe link or by what the host can feed to N links at once."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import audioanalysisdetector_b200 as aad

world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
n = 1 << 29                                   # 512 MiB in, 128 MiB out: the proportions of the C2 step with PCM input
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_in.fill_(1)
h_out = torch.empty(n // 4, dtype=torch.uint8, pin_memory=True)
wc = aad.pinned_empty((n,), np.uint8, write_combined=True); wc[...] = 1
h_wc = torch.from_numpy(np.asarray(wc))       # a CPU tensor over the write-combined block (pinned by CUDA, not by torch)
d_in = torch.empty(n, dtype=torch.uint8, device=dev)
d_out = torch.empty(n // 4, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return float(dt)


def both(src):
    def f():
        with torch.cuda.stream(s1):
            d_in.copy_(src, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
    return f


res = {
    "h2d": n / t(lambda: d_in.copy_(h_in, non_blocking=True)),
    "d2h": (n // 4) / t(lambda: h_out.copy_(d_out, non_blocking=True)),
    "both (h2d-equivalent)": n / t(both(h_in)),
}
try:
    rt = C.CDLL("libcudart.so.12")
    rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]

    def h2d_wc():
        rt.cudaMemcpyAsync(C.c_void_p(d_in.data_ptr()), C.c_void_p(h_wc.data_ptr()), n, 1,
                           C.c_void_p(torch.cuda.current_stream().cuda_stream))

    def both_wc():
        with torch.cuda.stream(s1):
            h2d_wc()
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
    res["h2d write-combined source"] = n / t(h2d_wc)
    res["both, write-combined source (h2d-equivalent)"] = n / t(both_wc)
except Exception as e:  # no libcudart by that name: skip the write-combined lines
    res["write-combined"] = f"unavailable ({e})"
if rank == 0:
    print(f"{world} GPU(s) copying concurrently, per GPU (slowest rank) and aggregate:")
    for k, v in res.items():
        if isinstance(v, str):
            print(f"  {k}: {v}")
        else:
            print(f"  {k:48s} {v / 1e9:7.1f} GB/s per GPU   {world * v / 1e9:8.1f} GB/s aggregate")
if world > 1:
    dist.destroy_process_group()
