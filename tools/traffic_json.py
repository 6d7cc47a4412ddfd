"""ncu --set full raw CSV of the BatchNorm-backward kernels -> profiles/rN_traffic_config2.json (DRAM bytes per call of the
C-ABI entry point b200_bn_act_bwd = bn_reduce_kernel<1> + bn_bwd_finalize_kernel + bn_act_bwd_kernel).
usage: python tools/traffic_json.py r2"""
import csv
import json
import os
import sys

R = sys.argv[1] if len(sys.argv) > 1 else "r2"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(root, "profiles", f"{R}_ncu_full_bn_act_bwd_entry.csv")
rows = list(csv.reader(open(src)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def val(r, key):
    return float(r[ix[key]].replace(",", "")) * SCALE[units[ix[key]]]


calls = 18                                   # BatchNorm layers with a backward pass in one config-2 step
groups = {"bn_reduce_kernel<1>": [], "bn_bwd_finalize_kernel": [], "bn_act_bwd_kernel": []}
for r in data:
    name = r[ix["Kernel Name"]]
    for k in groups:
        if k in name:
            groups[k].append(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"))
            break
per = {k: v[-calls:] for k, v in groups.items()}
assert all(len(v) == calls for v in per.values()), {k: len(v) for k, v in groups.items()}
total = sum(sum(v) for v in per.values())
out = {"kernel": "b200_bn_act_bwd", "launches": calls, "dram_bytes_per_launch": total / calls,
       "by_kernel_per_step": {k: sum(v) for k, v in per.items()},
       "source": f"profiles/{R}_ncu_full_bn_act_bwd_entry.csv (ncu --set full --clock-control none -k regex:bn_reduce_kernel|bn_bwd_finalize_kernel|"
                 "^bn_act_bwd_kernel python bench.py --config 2 --steps 1 --warmup 3 --no-graph --no-cpu --no-eager): the 18 calls of one "
                 "step, dram__bytes_read.sum + dram__bytes_write.sum of the three kernels each call launches"}
json.dump(out, open(os.path.join(root, "profiles", f"{R}_traffic_config2.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
