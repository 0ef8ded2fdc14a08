"""BASELINE config 5 with ONE utterance: 10 min of 48 kHz audio, n_fft 2048, hop 480, 128 mels, split along time
over the ranks (sharding.long_form_logmel; one MAX all-reduce of a float is the only exchange).
  python tools/c5_time_split.py                      # 1 GPU, unsplit
  torchrun --nproc-per-node N tools/c5_time_split.py # N GPUs, checks the pieces against the unsplit result
Prints one JSON line on rank 0 (device time of the slowest rank, CUDA events, inputs resident in HBM)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from audioanalysisdetector_b200 import FrontendParams, get_frontend, long_form_logmel

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
sr, L = 48000, 48000 * 600
g = torch.Generator(device=dev).manual_seed(5)
wav = (0.1 * torch.randn(L, generator=g, device=dev)).clamp_(-1, 1)
wav[: 10 * sr] *= 1e-5                      # a quiet stretch so that the top_db floor is exercised
params = FrontendParams.logmel(sr, n_mels=128, n_fft=2048, hop_length=480)

def step():
    return long_form_logmel(params, wav, rank, world, check_status=False)

for _ in range(3):
    feats, (t0, t1) = step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 20
e0.record()
for _ in range(iters):
    feats, (t0, t1) = step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
# parity against the unsplit extraction of the same signal (every rank checks its own piece)
want, nf, st = get_frontend(params, dev)(wav[None, :])
ok = torch.tensor([int(torch.equal(feats, want[0, :, t0:t1]))], device=dev)
if world > 1:
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"workload": "configs[4]: 1 x 10 min @48 kHz log-mel 128 (n_fft 2048, hop 480), time-axis split",
                      "n_gpus": world, "ms_per_step": float(ms), "audio_hours_per_s": 600 / 3600 / (float(ms) * 1e-3),
                      "frames": int(nf[0]), "pieces_equal_unsplit_bitwise": bool(int(ok))}))
if world > 1:
    dist.destroy_process_group()
