#!/bin/bash
# usage: tools/ncu_list.sh <regex> <skip> <count> <script...>   -> per-launch time and pipe utilisation
cd "$(dirname "$0")/.."
re=$1; skip=$2; cnt=$3; shift 3
timeout 600 ncu --metrics gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:$re -s $skip -c $cnt --csv --log-file gpurun_out/ncu_list.csv "$@" > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/ncu_list.csv")) if len(r)>10]
hdr=rows[0]; i_k=hdr.index("Kernel Name"); i_m=hdr.index("Metric Name"); i_v=hdr.index("Metric Value"); i_id=hdr.index("ID")
from collections import OrderedDict
d=OrderedDict()
for r in rows[1:]:
    d.setdefault((r[i_id], r[i_k].split("(")[0][-28:]), {})[r[i_m]]=r[i_v]
short={"gpu__time_duration.sum":"ns","sm__throughput.avg.pct_of_peak_sustained_elapsed":"sm%","l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed":"l1%","sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active":"fma%","dram__throughput.avg.pct_of_peak_sustained_elapsed":"dram%","smsp__issue_active.avg.pct_of_peak_sustained_active":"issue%","sm__warps_active.avg.pct_of_peak_sustained_active":"warps%"}
for k,v in d.items(): print(k[0], k[1], " ".join(f"{short.get(a,a)}={b}" for a,b in v.items()))
PY
