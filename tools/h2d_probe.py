"""PCIe ceiling probe (dev tool): pinned H2D / D2H bandwidth alone and concurrently."""
import torch, time
dev = torch.device("cuda:0")
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_in.fill_(1)
h_out = torch.empty(n // 4, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n, dtype=torch.uint8, device=dev)
d_out = torch.empty(n // 4, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
a = t(lambda: d_in.copy_(h_in, non_blocking=True))
b = t(lambda: h_out.copy_(d_out, non_blocking=True))
def both():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
c = t(both)
print(f"H2D 1 GiB: {n/a/1e9:.1f} GB/s   D2H 256 MiB: {n/4/b/1e9:.1f} GB/s   concurrent: {c*1e3:.2f} ms (H2D-equivalent {n/c/1e9:.1f} GB/s)")
for chunk_mb in (8, 32, 128):
    cb = chunk_mb << 20
    def chunks():
        for o in range(0, n, cb): d_in[o:o+cb].copy_(h_in[o:o+cb], non_blocking=True)
    d = t(chunks, 3)
    print(f"H2D in {chunk_mb} MiB chunks: {n/d/1e9:.1f} GB/s")
