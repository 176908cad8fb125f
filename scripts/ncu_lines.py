#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel of an .ncu-rep: joins the SASS page of the report (instructions executed per
SASS instruction) with nvdisasm's line info of the same kernel in librt_b200.so (the instruction order is the same).
usage: ncu_lines.py rep kernel-regex [mangled-substring] [top]"""
import csv, io, os, re, subprocess, sys, collections, tempfile

rep, kre = sys.argv[1], sys.argv[2]
mangled = sys.argv[3] if len(sys.argv) > 3 else kre
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
# several launches may match: keep the first block
blocks = raw.split('"Kernel Name",')
rows = list(csv.reader(io.StringIO('"Kernel Name",' + blocks[1])))
print("kernel:", rows[0][1][:100])
h = rows[1]
ie, te, sm, src = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples"), h.index("Source")
ncu_rows = [(r[src].strip(), int(r[ie]), int(r[te]), int(r[sm])) for r in rows[2:] if len(r) > te and r[ie].isdigit()]

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(REPO, "simd-raytracer_b200", "librt_b200.so")], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.startswith("rt_api")][0]
sass = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(sass) if l.startswith(".text.") and mangled in l and l.rstrip().endswith(":")]
if not start:
    sys.exit("no such kernel in the cubin: " + mangled)
i = start[0] + 1
line_of = []          # (file:line of the innermost frame, inline chain) per instruction
cur = ("?", 0); cur_ours = False; fresh = True     # fresh: the next "//## File" line starts a new group (innermost frame first)
while i < len(sass) and not sass[i].startswith("\t.section"):
    l = sass[i]
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        ours = "/simd-raytracer_b200/" in m.group(1)
        if fresh or (not cur_ours and ours):          # innermost frame that is in this repo
            cur = (os.path.basename(m.group(1)), int(m.group(2))); cur_ours = ours
        fresh = False
    elif re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
        line_of.append(cur)
        fresh = True
    i += 1
print(f"ncu instructions {len(ncu_rows)}, nvdisasm instructions {len(line_of)}")
n = min(len(ncu_rows), len(line_of))
by = collections.defaultdict(lambda: [0, 0, 0])
for (s, a, b, c), loc in zip(ncu_rows[:n], line_of[:n]):
    by[loc][0] += a; by[loc][1] += b; by[loc][2] += c
tot = sum(v[0] for v in by.values()); tots = sum(v[2] for v in by.values())
print(f"total warp instructions {tot}, samples {tots}")
cache = {}
def text(f, ln):
    for d in ("simd-raytracer_b200/csrc", "simd-raytracer_b200/host"):
        p = os.path.join(REPO, d, f)
        if os.path.exists(p):
            if p not in cache: cache[p] = open(p).read().splitlines()
            return cache[p][ln - 1].strip()[:100] if ln - 1 < len(cache[p]) else ""
    return ""
byfile = collections.defaultdict(int)
for (f, ln), v in by.items(): byfile[f] += v[0]
print({k: f"{100 * v / tot:.1f}%" for k, v in sorted(byfile.items(), key=lambda kv: -kv[1])})
for (f, ln), v in sorted(by.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * v[0] / tot:5.1f}% inst {100 * v[2] / max(tots, 1):5.1f}% smp  thr/inst {v[1] / max(v[0], 1):5.1f}  {f}:{ln}  {text(f, ln)}")
