#!/usr/bin/env python
"""profiles/kernel_counts.json: what every kernel of ONE frame executed and moved, from an ncu metrics pass over a bench.py run.

    ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        --clock-control none --csv --log-file gpurun_out/<tag>.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --ns-tris 0 [...]
    python scripts/make_kernel_counts.py gpurun_out/<tag>.csv <workload-key> <mode>/<accel>/spp<S> [passes-per-frame]

Every frame of such a run is the same frame (warm-up, timed, end-to-end and verification frames all render the workload), so the
per-frame figure of a kernel is its total over the run divided by the number of frames = launches of k_pass_commit / passes per
frame.  bench.py divides these counts by the kernel time it measures live (roofline, roofline_by_kind)."""
import collections
import csv
import json
import os
import sys

src, key, variant = sys.argv[1:4]
passes = int(sys.argv[4]) if len(sys.argv) > 4 else 1
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = [r for r in csv.reader(open(src)) if len(r) > 14 and r[0].isdigit()]
per = collections.defaultdict(lambda: collections.defaultdict(float))
launch_ids = collections.defaultdict(set)
for r in rows:
    name = r[4].replace("(anonymous namespace)::", "").replace("<unnamed>::", "").split("(")[0].replace("void ", "").replace("rtb::", "").split("<")[0].strip()
    val = float(r[14].replace(",", ""))
    unit = r[13].lower()
    val *= {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1.0)
    per[name][r[12]] += val
    launch_ids[name].add(r[0])
frames = len(launch_ids.get("k_pass_commit", ())) / passes
if not frames:
    raise SystemExit("no k_pass_commit launch in the capture")
kernels = {}
for name, m in per.items():
    kernels[name] = {"launches": len(launch_ids[name]) / frames,
                     "warp_inst": m.get("smsp__inst_executed.sum", 0.0) / frames,
                     "thread_inst": m.get("smsp__thread_inst_executed.sum", 0.0) / frames,
                     "dram_bytes": (m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)) / frames,
                     "ncu_us": m.get("gpu__time_duration.sum", 0.0) / 1e3 / frames}
path = os.path.join(REPO, "profiles", "kernel_counts.json")
try:
    table = json.load(open(path))
except OSError:
    table = {}
table.setdefault(key, {})[variant] = {"_source": f"{os.path.basename(src)}: ncu --metrics (inst_executed, thread_inst_executed, dram bytes, duration) "
                                                 f"--clock-control none over {frames:g} identical frames of one bench.py run; per-frame means",
                                      "frames": frames, "kernels": kernels}
json.dump(table, open(path, "w"), indent=1, sort_keys=True)
tot = sum(k["ncu_us"] for k in kernels.values())
for name, k in sorted(kernels.items(), key=lambda kv: -kv[1]["ncu_us"]):
    print(f"{name:28s} launches/frame {k['launches']:5.1f}  us/frame {k['ncu_us']:9.1f} ({100 * k['ncu_us'] / tot:4.1f} %)  warp-inst {k['warp_inst'] / 1e6:9.2f} M  "
          f"lanes/inst {k['thread_inst'] / max(k['warp_inst'], 1):5.2f}  dram {k['dram_bytes'] / 1e6:9.2f} MB")
