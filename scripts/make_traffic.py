#!/usr/bin/env python
"""profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch of every kernel in an ncu --set full capture
(mean over the captured launches of that kernel).  bench.py quotes it as roofline.traffic.
usage: make_traffic.py rep workload-key mode"""
import csv, io, json, os, subprocess, sys, collections
rep, key, mode = sys.argv[1:4]
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
def col(r, name):
    i = hdr.index(name); v = float(r[i].replace(",", "")); u = units[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
acc = collections.defaultdict(list)
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("rtb::", "").replace("<unnamed>::", "").split("<")[0]
    acc[name].append(col(r, "dram__bytes_read.sum") + col(r, "dram__bytes_write.sum"))
path = os.path.join(REPO, "profiles", "traffic.json")
try: table = json.load(open(path))
except OSError: table = {}
table.setdefault(key, {})[mode] = {k: int(sum(v) / len(v)) for k, v in acc.items()}
table[key][mode]["_source"] = os.path.basename(rep) + " (ncu --set full --clock-control none; bytes per launch, mean over captured launches)"
json.dump(table, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(table[key][mode], indent=1))
