#!/bin/bash
# Round-2 multi-GPU measurements (row g of VERDICT r1): usage  tools/r2_multigpu.sh N "what"   what in: probe c2 c3 c4 c5
cd "$(dirname "$0")/.."
N=$1; shift
PORT=29512
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT "$@"; PORT=$((PORT+1)); }
for w in "$@"; do
  case $w in
    probe) run tools/h2d_probe.py 2>/dev/null | tail -7 | tee gpurun_out/r2_pcie_probe_${N}gpu.log ;;
    c2) run bench.py --gpus $N --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err || tail -5 gpurun_out/r2_bench_${N}gpu.err ;;
    c3|c4|c5) run bench.py --gpus $N --workload $w --steps 10 > gpurun_out/r2_bench_${w}_${N}gpu.json 2> gpurun_out/r2_${w}_${N}.err || tail -5 gpurun_out/r2_${w}_${N}.err ;;
  esac
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench_*${N}gpu.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.1f"%d["value"], "ms %.3f"%d["ms_per_step"], "e2e", {k:round(d[k]["value"],1) for k in ("e2e","e2e_wc","e2e_f32") if d.get(k)})
    except Exception as e:
        print(f, "unreadable", e)
PY
