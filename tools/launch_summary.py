"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel totals and one step's
per-launch timeline.  usage: python tools/launch_summary.py gpurun_out/launches.csv [detail-regex]"""
import csv, collections, re, sys
path = sys.argv[1]
detail = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
for r in rows:
    t = float(r["Metric Value"])
    r["us"] = t / 1e3 if r["Metric Unit"] == "ns" else (t * 1e3 if r["Metric Unit"] == "ms" else t)
    r["k"] = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
idx = [i for i, r in enumerate(rows) if "sgd_ema" in r["k"]]
step = rows[idx[-2] + 1: idx[-1] + 1] if len(idx) >= 2 else rows
tot = sum(r["us"] for r in step)
agg = collections.defaultdict(lambda: [0, 0.0])
for r in step:
    agg[r["k"]][0] += 1
    agg[r["k"]][1] += r["us"]
print(f"one step: {len(step)} kernels, {tot:.1f} us (serialised, cold cache)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:20]:
    print(f"{v[1]:9.1f} us {v[0]:4d}x {v[1] / tot * 100:5.1f}%  {k[:90]}")
if detail:
    for r in step:
        if detail.search(r["k"]):
            print(f"{r['us']:8.1f} us grid {r['Grid Size']:>16} {r['k'][:70]}")
