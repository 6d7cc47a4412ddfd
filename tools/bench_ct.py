"""BASELINE config 3: 2D Cross-Teaching UNet <-> Swin-UNet step, synthetic ACDC-shape 224x224, 4 classes, bs16 (8/8).
Prints one JSON line (slices/s; device-resident and end-to-end) -- a secondary workload next to bench.py's config 2.
usage: python tools/bench_ct.py [--steps K] [--warmup W] [--cpu] [--profile]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--cpu", action="store_true", help="also time the oracle port of the reference step on the host")
ap.add_argument("--profile", action="store_true")
ap.add_argument("--batch", type=int, default=16)
args = ap.parse_args()
B, P = args.batch, 224
Lb = B // 2

from cv_ssl_mis_b200 import _lib
from cv_ssl_mis_b200.networks.net_factory import net_factory
from cv_ssl_mis_b200.trainers import CrossTeachingTrainer

g = torch.Generator().manual_seed(1337)
x = torch.rand(B, 1, P, P, generator=g).pin_memory()
low = torch.randint(0, 4, (B, P // 8, P // 8), generator=g)
y = low.repeat_interleave(8, 1).repeat_interleave(8, 2).to(torch.uint8).pin_memory()
torch.manual_seed(1337)
m1, m2 = net_factory("unet", 1, 4), net_factory("ViT_Seg", 1, 4)
tr = CrossTeachingTrainer(m1, m2, batch_size=B, labeled_bs=Lb, patch_size=(P, P), num_classes=4, start_iter=3000,
                          use_cuda_graph=True)
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
# Swin-UNet-lite fwd 12.17 GFLOP/img, UNet 4.52 GFLOP/img @224^2; fwd + bwd = 3x
tflop = 3 * B * (12.17 + 4.52) / 1e3
out = {"metric": "train-step slices/sec (ACDC 224x224 bs16 Cross-Teaching UNet<->SwinUNet)", "value": B / ms * 1e3,
       "unit": "slices/s", "ms_per_step": ms, "e2e": {"value": B / ms_e2e * 1e3, "ms_per_step": ms_e2e}, "n_gpus": 1,
       "steps": args.steps, "gpu_launches_per_step": tr.kernel_launches_per_step, "algorithmic_tflop_per_step": tflop,
       "achieved_tflops": tflop / ms * 1e3, "loss": [b[:4].tolist() for b in tr.lossbufs], "dtype": "tf32",
       "data": "synthetic", "swin_grad_pool_mb": tr.plans[1].grad_floats * 4 / 2 ** 20}
if args.profile:
    tr.use_graph = False
    tr.step(xd, yd)
    torch.cuda.synchronize()
    _lib.profile = []
    tr.step(xd, yd)
    torch.cuda.synchronize()
    rec, _lib.profile = _lib.profile, None
    agg, by_tag = {}, {}
    for name, tag, e0, e1 in rec:
        t = e0.elapsed_time(e1)
        agg[name] = agg.get(name, 0.0) + t
        fam = name + ":" + (".".join(tag.split(".")[-1:]) if tag else "")
        by_tag[fam] = by_tag.get(fam, 0.0) + t
    tot = sum(agg.values())
    out["step_time_shares"] = {k: round(v / tot, 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:14]}
    out["top_layers_ms"] = {k: round(v, 3) for k, v in sorted(by_tag.items(), key=lambda kv: -kv[1])[:16]}
    out["profiled_eager_ms_per_step"] = tot
if args.cpu:
    from oracle import ssl_oracle as O, swin_oracle as SO
    from cv_ssl_mis_b200.networks.unet import UNet
    from cv_ssl_mis_b200.networks.swin_unet import SwinUnet
    torch.manual_seed(1337)
    sd1 = {k: v.clone() for k, v in UNet(1, 4).state_dict().items()}
    sd2 = {k[len("swin_unet."):]: v.clone() for k, v in SwinUnet(None, num_classes=4).state_dict().items()}
    cfg = SO.swin_config(sd2, 224, 7, 0.2)
    bufs1 = {k: torch.zeros_like(sd1[k]) for k in O.param_keys(sd1)}
    bufs2 = {k: torch.zeros_like(v) for k, v in sd2.items() if v.dtype.is_floating_point}
    ts = []
    for i in range(2):
        keeps = [tuple((torch.rand(B, generator=g) >= p).float() for _ in range(2)) for p in cfg["dpr"] + cfg["dpr"][:6]]
        t0 = time.perf_counter()
        SO.ct2d_step(sd1, sd2, bufs1, bufs2, x.clone(), y.clone(), 3000 + i, cfg, labeled_bs=Lb, drop_keep=keeps)
        ts.append(time.perf_counter() - t0)
    out["cpu_baseline"] = {"value": B / ts[-1], "unit": "slices/s", "cores": os.cpu_count(), "kind": "port",
                           "sample": "2nd of 2 full Cross-Teaching steps, oracle port, torch CPU fp32"}
print(json.dumps(out))
