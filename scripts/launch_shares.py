#!/usr/bin/env python
"""Per-kernel totals and shares of an ncu launch list (--metrics gpu__time_duration.sum --csv).  usage: launch_shares.py csv [frames]"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 1
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows:
    name = r[4].split("(")[0].replace("void ", "")
    if name.startswith("at::"): name = "torch fill (L2 flush / zero)"
    tot[name] += float(r[14]) / 1e3; cnt[name] += 1
ours = sum(v for k, v in tot.items() if not k.startswith("torch"))
print(f"{'kernel':44s} {'launches':>8s} {'total us':>10s} {'us/launch':>10s} {'share of own kernels':>8s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    share = "" if k.startswith("torch") else f"{100 * v / ours:6.1f} %"
    print(f"{k[:44]:44s} {cnt[k]:8d} {v:10.1f} {v / cnt[k]:10.2f} {share}")
print(f"own kernels total {ours:.1f} us over {len(rows)} launches")
