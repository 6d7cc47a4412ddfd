"""BASELINE config 4: 3D VNet uncertainty-aware Mean-Teacher step, synthetic BraTS-shape 96^3, bs4 (2/2), T = 8.
Prints one JSON line (patches/s; device-resident and end-to-end) -- the secondary workload next to bench.py's config 2.
usage: python tools/bench_uamt.py [--steps K] [--warmup W] [--cpu]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--cpu", action="store_true", help="also time the oracle port of the reference step on the host")
ap.add_argument("--profile", action="store_true")
args = ap.parse_args()
B, Lb, P, T = 4, 2, 96, 8

from cv_ssl_mis_b200 import _lib
from cv_ssl_mis_b200.networks.net_factory_3d import net_factory_3d
from cv_ssl_mis_b200.trainers import MeanTeacherTrainer

g = torch.Generator().manual_seed(1337)
x = torch.randn(B, 1, P, P, P, generator=g).pin_memory()
low = torch.randint(0, 2, (B, P // 8, P // 8, P // 8), generator=g)
y = low.repeat_interleave(8, 1).repeat_interleave(8, 2).repeat_interleave(8, 3).long().pin_memory()
torch.manual_seed(1337)
s, t = net_factory_3d("vnet", 1, 2), net_factory_3d("vnet", 1, 2)
tr = MeanTeacherTrainer(s, t, batch_size=B, labeled_bs=Lb, patch_size=(P, P, P), num_classes=2, start_iter=3000,
                        consistency_gate_iters=0, uncertainty_T=T, use_cuda_graph=True)
xd, yd = x.cuda(), y.cuda()
for _ in range(args.warmup):
    tr.step(xd, yd)


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.steps


ms = timed(lambda: tr.step(xd, yd))
ms_e2e = timed(lambda: tr.step(x, y, read_loss=True))
out = {"metric": "train-step patches/sec (BraTS 96^3 bs4 UAMT-VNet, T=8)", "value": B / ms * 1e3, "unit": "patches/s",
       "ms_per_step": ms, "e2e": {"value": B / ms_e2e * 1e3, "ms_per_step": ms_e2e}, "n_gpus": 1, "steps": args.steps,
       "gpu_launches_per_step": tr.kernel_launches_per_step, "algorithmic_tflop_per_step": 2.12,
       "achieved_tflops": 2.12 / ms * 1e3, "loss": tr.lossbuf[:4].tolist(), "dtype": "tf32", "data": "synthetic"}
if args.profile:
    tr.use_graph = False
    tr.step(xd, yd)
    torch.cuda.synchronize()
    _lib.profile = []
    tr.step(xd, yd)
    torch.cuda.synchronize()
    rec, _lib.profile = _lib.profile, None
    agg = {}
    for name, tag, e0, e1 in rec:
        agg[name] = agg.get(name, 0.0) + e0.elapsed_time(e1)
    tot = sum(agg.values())
    out["step_time_shares"] = {k: round(v / tot, 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]}
    out["profiled_eager_ms_per_step"] = tot
if args.cpu:
    from oracle import ssl_oracle as O
    from cv_ssl_mis_b200.networks.vnet import VNet
    torch.manual_seed(1337)
    s_sd = {k: v.clone() for k, v in VNet(1, 2, has_dropout=True).state_dict().items()}
    t_sd = {k: v.clone() for k, v in VNet(1, 2, has_dropout=True).state_dict().items()}
    bufs = {k: torch.zeros_like(s_sd[k]) for k in O.param_keys(s_sd)}
    xc, yc = x.clone(), y.clone()
    U = B - Lb
    ts = []
    for i in range(2):
        noises = [O.clamp_noise(torch.empty(U if k == 0 else 2 * U, 1, P, P, P), g) for k in range(1 + T // 2)]
        drops = lambda n: ((torch.rand(n, 256, generator=g) >= 0.5).float(), (torch.rand(n, 16, generator=g) >= 0.5).float())
        t0 = time.perf_counter()
        O.uamt3d_step(s_sd, t_sd, bufs, xc, yc, noises, 3000 + i, labeled_bs=Lb, T=T, student_drops=drops(B),
                      teacher_drops=[drops(U if k == 0 else 2 * U) for k in range(1 + T // 2)])
        ts.append(time.perf_counter() - t0)
    out["cpu_baseline"] = {"value": B / ts[-1], "unit": "patches/s", "cores": os.cpu_count(), "kind": "port",
                           "sample": "2nd of 2 full UAMT steps, oracle port, torch CPU fp32"}
print(json.dumps(out))
