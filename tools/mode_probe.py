"""Losses of two full-size Mean-Teacher steps (config 2) in exact (3xTF32) mode and in the production TF32 mode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cv_ssl_mis_b200.networks import unet as unet_mod
from cv_ssl_mis_b200.trainers import MeanTeacherTrainer
B, Lb, H, W = 24, 12, 256, 256
g = torch.Generator().manual_seed(3)
x = torch.rand(B, 1, H, W, generator=g).pin_memory()
low = torch.randint(0, 4, (B, H // 16, W // 16), generator=g)
y = low.repeat_interleave(16, 1).repeat_interleave(16, 2).to(torch.uint8).pin_memory()
exact = len(sys.argv) > 1 and sys.argv[1] == "exact"
torch.manual_seed(9)
s, t = unet_mod.UNet(1, 4, seed=1, exact=exact).cuda(), unet_mod.UNet(1, 4, seed=2, exact=exact).cuda()
tr = MeanTeacherTrainer(s, t, batch_size=B, labeled_bs=Lb, patch_size=(H, W), start_iter=2000, use_cuda_graph=False)
for _ in range(2):
    print(["%.7f" % v for v in tr.step(x, y, read_loss=True)])
print("param norm %.7f" % float(tr.flat.data.norm()))
