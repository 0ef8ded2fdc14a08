"""Regenerate the tracked summaries under profiles/ from the scratch files of the last GPU calls.
usage: refresh_profiles.py <ncu-rep> <launches.csv> <bench.json> <bench_ref.json>"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, launches, bench, bench_ref = sys.argv[1:5]
P = lambda *a: os.path.join(ROOT, "profiles", *a)
R = os.environ.get("AAD_ROUND", "r2")   # file-name prefix of the round
FRAMES = 516096

# 1. ncu summary
with open(P(f"{R}_ncu_summary.md"), "w") as f:
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], stdout=f, check=True)

# 2. DRAM traffic of k_stft_fb
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr, units = rr[0], rr[1]
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
for r in rr[2:]:
    d = dict(zip(hdr, r))
    if "k_stft_fb" in d["Kernel Name"]:
        rd = float(d["dram__bytes_read.sum"]) * scale[units[hdr.index("dram__bytes_read.sum")]]
        wr = float(d["dram__bytes_write.sum"]) * scale[units[hdr.index("dram__bytes_write.sum")]]
        json.dump({"kernel": d["Kernel Name"], "source": f"profiles/{R}_ncu_summary.md (ncu --set full, one launch of the C2 workload: 4096 x 4 s)",
                   "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
                   "algorithmic_bytes_per_launch": FRAMES * (4 * 512 + 4 * 128)}, open(P("stft_fb_traffic.json"), "w"), indent=1)

# 3. SASS stall profile per kernel
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
seen, out = set(), []
for n, hi in enumerate(his):
    name = rows[hi - 1][1] if rows[hi - 1] and rows[hi - 1][0] == "Kernel Name" else "?"
    if name in seen or "k_prepare" in name:
        continue
    seen.add(name)
    end = his[n + 1] - 1 if n + 1 < len(his) else len(rows)
    tmp = f"/tmp/_src_{n}.csv"
    with open(tmp, "w", newline="") as f:
        csv.writer(f).writerows(rows[hi:end])
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_profile.py"), tmp, str(FRAMES), "20"],
                         capture_output=True, text=True).stdout
    out.append(f"# {name}\n# SASS opcode mix, per-section cost (sections split at BAR) and warp-stall reasons; "
               f"ncu --set full --import-source on, C2 workload ({FRAMES} frames per launch); tools/sass_profile.py\n" + txt)
open(P(f"{R}_sass_stalls.txt"), "w").write("\n".join(out))

# 4. launch list
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
h = rows[0]
ki, vi, bi, gi, ii = (h.index(x) for x in ("Kernel Name", "Metric Value", "Block Size", "Grid Size", "ID"))
tot = {}
with open(P(f"{R}_launches.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["id", "kernel", "block", "grid", "gpu__time_duration.sum [ns]"])
    for r in rows[1:]:
        n = r[ki]
        if "aad::" not in n:
            n = n.split("<")[0].replace("void ", "") + "<...> (torch, bench set-up)"
        else:
            k = n.split("(")[0].split("<")[0]
            tot[k] = tot.get(k, 0) + float(r[vi].replace(",", ""))
        w.writerow([r[ii], n, r[bi], r[gi], r[vi]])
s = sum(tot.values())
print("launch-list shares:", {k: f"{100 * v / s:.1f}%" for k, v in tot.items()})

# 5. bench lines
for src_, dst in ((bench, f"{R}_bench.json"), (bench_ref, f"{R}_bench_reference.json")):
    line = [l for l in open(src_).read().splitlines() if l.startswith("{")][-1]
    open(P(dst), "w").write(line + "\n")
d = json.loads(open(P(f"{R}_bench.json")).read())
print("bench:", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"])
