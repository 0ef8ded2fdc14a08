"""Aggregate an `ncu --page source --csv` dump: opcode mix, per-section (between BARs) cost and
stall reasons, top stalled instructions.  usage: sass_profile.py file.csv frames [top_n]"""
import csv, collections, re, sys
path, frames = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(open(path)))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[0]
end = next((i for i in range(hi + 1, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:end] if len(r) > ci["Instructions Executed"] and r[0].startswith("0x")]
ex = lambda r: int(r[ci["Instructions Executed"]] or 0)
sm = lambda r: int(r[ci["# Samples"]] or 0)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_e, tot_s = sum(map(ex, data)), sum(map(sm, data))
print(f"{len(data)} SASS instr; executed warp-inst {tot_e} ({tot_e/frames:.1f}/frame); samples {tot_s}")
op_e, op_s = collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)", r[ci["Source"]])
    o = m.group(2) if m else "?"
    op_e[o] += ex(r); op_s[o] += sm(r)
for o, c in op_e.most_common(22):
    print(f"  {o:10s} {c/frames:8.1f}/frame {100*c/tot_e:5.1f}% of inst   {100*op_s[o]/max(tot_s,1):5.1f}% of samples")
def stall_mix(rs):
    tot = collections.Counter()
    for r in rs:
        for h in stalls:
            tot[h] += int(r[ci[h]] or 0)
    t = sum(tot.values()) or 1
    return " ".join(f"{h[6:]}={100*v/t:.0f}" for h, v in tot.most_common(7))
bars = [i for i, r in enumerate(data) if re.search(r"\bBAR\.", r[ci["Source"]])]
prev = 0
for b in bars + [len(data)]:
    e = sum(map(ex, data[prev:b])); s = sum(map(sm, data[prev:b]))
    print(f"section [{prev:5d},{b:5d}) {e/frames:8.1f} inst/frame  {100*s/max(tot_s,1):5.1f}% samples | {stall_mix(data[prev:b])}")
    prev = b
print("overall:", stall_mix(data))
if len(sys.argv) > 3:
    idx = {id(r): i for i, r in enumerate(data)}
    top = sorted(data, key=sm, reverse=True)[:int(sys.argv[3])]
    for r in top:
        print(f"  [{idx[id(r)]:5d}] {sm(r):6d} {ex(r):9d}  {r[ci['Source']].strip()[:70]:70s} | {stall_mix([r])}")
