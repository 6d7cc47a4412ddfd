"""BASELINE config 5: 3D UNETR (ViT-B encoder) fully supervised, synthetic BraTS-shape 96^3, 2 classes, bs2.
Prints one JSON line (patches/s; device-resident and end-to-end) -- a secondary workload next to bench.py's config 2.
usage: python tools/bench_unetr.py [--steps K] [--warmup W] [--cpu] [--profile]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--cpu", action="store_true", help="also time the oracle port of the reference step on the host")
ap.add_argument("--profile", action="store_true")
args = ap.parse_args()
B, P = 2, 96

from cv_ssl_mis_b200 import _lib
from cv_ssl_mis_b200.networks.net_factory_3d import net_factory_3d
from cv_ssl_mis_b200.trainers import MeanTeacherTrainer

g = torch.Generator().manual_seed(1337)
x = torch.randn(B, 1, P, P, P, generator=g).pin_memory()
low = torch.randint(0, 2, (B, P // 8, P // 8, P // 8), generator=g)
y = low.repeat_interleave(8, 1).repeat_interleave(8, 2).repeat_interleave(8, 3).pin_memory()
torch.manual_seed(1337)
net = net_factory_3d("unetr", 1, 2)
tr = MeanTeacherTrainer(net, None, batch_size=B, labeled_bs=B, patch_size=(P, P, P), num_classes=2, use_cuda_graph=True)
xd, yd = x.cuda(), y.cuda()
for _ in range(max(args.warmup, 3)):
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
out = {"metric": "train-step patches/sec (BraTS 96^3 bs2 fully-supervised UNETR)", "value": B / ms * 1e3, "unit": "patches/s",
       "ms_per_step": ms, "e2e": {"value": B / ms_e2e * 1e3, "ms_per_step": ms_e2e}, "n_gpus": 1, "steps": args.steps,
       "gpu_launches_per_step": tr.kernel_launches_per_step, "loss": tr.lossbuf[:4].tolist(), "dtype": "tf32", "data": "synthetic",
       "params": sum(p.numel() for p in net.parameters())}
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
    out["step_time_shares"] = {k: round(v / tot, 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:12]}
    out["profiled_eager_ms_per_step"] = tot
if args.cpu:
    from oracle import unetr_oracle as UO
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in net.state_dict().items()}
    ts = []
    for i in range(2):
        t0 = time.perf_counter()
        loss, _ = UO.fully_supervised_loss(sd, x, y, 12, 2)
        loss.backward()
        ts.append(time.perf_counter() - t0)
    out["cpu_baseline"] = {"value": B / ts[-1], "unit": "patches/s", "cores": os.cpu_count(), "kind": "port",
                           "sample": "2nd of 2 forward+backward passes of the oracle port (torch CPU fp32), optimizer excluded"}
print(json.dumps(out))
