"""Static SASS opcode histogram of one kernel in libaad_b200.so (dev tool).
usage: sass_static.py <mangled-substring> [--dump]"""
import collections, os, re, subprocess, sys
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "audioanalysisdetector_b200", "libaad_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
parts = re.split(r"\n\s*Function : ", txt)
for p in parts[1:]:
    name = p.split("\n", 1)[0].strip()
    if sys.argv[1] not in name:
        continue
    ins = re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)[^;]*;", p, re.M)
    c = collections.Counter(ins)
    print(name, len(ins), "instructions")
    print("  " + "  ".join(f"{k}:{v}" for k, v in c.most_common(28)))
    if "--dump" in sys.argv:
        for m in re.finditer(r"^\s+/\*([0-9a-f]{4,})\*/\s+([^;]*;)", p, re.M):
            print(m.group(1), m.group(2))
    break
