"""Diagnostic: per-parameter gradient error of the CUDA UNet (exact / TF32) and of the fp32 oracle, all
relative to an fp64 run of the oracle (conditioning check for tiny-batch BatchNorm)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ssl_oracle as O
from cv_ssl_mis_b200.networks import unet as unet_mod
from tests.test_host_logic import unet_masks

B, H, W = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (3, 48, 32)
g = torch.Generator().manual_seed(1)
x = torch.rand(B, 1, H, W, generator=g)
y = torch.randint(0, 4, (B, H, W), generator=g).to(torch.uint8)
res = {}
for exact in (True, False):
    torch.manual_seed(5)
    net = unet_mod.UNet(1, 4, seed=77, exact=exact)
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.cuda()
    logits = net(x.cuda())
    loss, _, _ = O.supervised_loss(logits, y.cuda(), 4)
    loss.backward()
    res[exact] = ({n: p.grad.cpu().double() for n, p in net.named_parameters()}, logits.detach().cpu().double())
keys = O.param_keys(sd0)
masks = unet_masks(78, B, H, W)
out = {}
for dt in (torch.float64, torch.float32):
    leaf = {k: (v.clone().to(dt).requires_grad_(True) if k in keys else v.clone().to(dt) if v.dtype.is_floating_point else v.clone()) for k, v in sd0.items()}
    ref = O.unet_forward(leaf, x.to(dt), True, [m.to(dt) for m in masks], update_running=False)
    l, _, _ = O.supervised_loss(ref, y, 4)
    out[dt] = (dict(zip(keys, [t.double() for t in torch.autograd.grad(l, [leaf[k] for k in keys])])), ref.detach().double())
truth, tl = out[torch.float64]
def rel(a, b): return float((a - b).norm() / (b.norm() + 1e-30))
print("logits rel err: exact %.2e tf32 %.2e oracle32 %.2e" % (rel(res[True][1], tl), rel(res[False][1], tl), rel(out[torch.float32][1], tl)))
print("%-45s %10s %10s %10s" % ("param", "exact", "tf32", "oracle32"))
for k in keys:
    print("%-45s %10.2e %10.2e %10.2e" % (k, rel(res[True][0][k], truth[k]), rel(res[False][0][k], truth[k]), rel(out[torch.float32][0][k], truth[k])))
