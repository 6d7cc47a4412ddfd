"""Regenerates the headline table of DESIGN.md section 8 (between the S8TABLE markers) from profiles/rN_bench_config*.json.
usage: python tools/fill_design.py r2"""
import json
import os
import re
import sys

R = sys.argv[1] if len(sys.argv) > 1 else "r2"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROUND1 = {2: "6.43 ms", 3: "12.2 ms", 4: "37.8 ms", 5: "19.5 ms"}
NAMES = {2: "config 2: MT-UNet 256² bs24", 3: "config 3: Cross-Teaching UNet↔Swin-UNet 224² bs16",
         4: "config 4: UAMT-VNet 96³ bs4 T=8", 5: "config 5: fully-supervised UNETR 96³ bs2"}
rows = ["| workload (1 GPU) | round 1 | round 2 (CUDA graph, batch resident) | end to end (pipelined `submit` / blocking `step`) | torch-eager, same B200 (TF32 cuDNN/cuBLAS) | reference CPU path (oracle port, host cores) |",
        "|---|---|---|---|---|---|"]
for c in (2, 3, 4, 5):
    f = os.path.join(root, "profiles", f"{R}_bench_config{c}.json")
    if not os.path.exists(f):
        continue
    d = json.loads(open(f).read().strip().splitlines()[-1])
    e, cb, e2e = d.get("gpu_eager_baseline") or {}, d.get("cpu_baseline") or {}, d["e2e"]
    unit = d["unit"]
    eager = f"{e['ms_per_step']:.1f} ms ({e['ours_over_eager']:.1f}× slower than ours)" if "ms_per_step" in e else "n/a"
    rows.append(f"| {NAMES[c]} | {ROUND1[c]} | **{d['ms_per_step']:.2f} ms = {d['value']:.0f} {unit}** | {e2e['ms_per_step']:.2f} / "
                f"{e2e.get('blocking_api_ms_per_step', float('nan')):.2f} ms | {eager} | {cb.get('value', float('nan')):.2f} {unit} |")
for n in (2, 4, 8):
    f = os.path.join(root, "profiles", f"{R}_bench_config2_{n}gpu.json")
    if os.path.exists(f):
        d = json.loads(open(f).read().strip().splitlines()[-1])
        rows.append(f"| config 2 on {n} GPUs (weak scaling) | | {d['ms_per_step']:.2f} ms = {d['value']:.0f} {d['unit']} | {d['e2e']['ms_per_step']:.2f} ms | | |")
p = os.path.join(root, "DESIGN.md")
s = open(p).read()
s = re.sub(r"<!-- S8TABLE -->.*?<!-- /S8TABLE -->", "<!-- S8TABLE -->\n" + "\n".join(rows) + "\n<!-- /S8TABLE -->", s, flags=re.S)
open(p, "w").write(s)
print("\n".join(rows))
