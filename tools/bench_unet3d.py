"""Step time of Mean Teacher over two unet_3Ds (code/train_mean_teacher_3D.py defaults: 96^3 patches, batch 4 = 2 labeled +
2 unlabeled) on one GPU, CUDA-graph replay, batch resident.  usage: python tools/bench_unet3d.py [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cv_ssl_mis_b200.networks.net_factory_3d import net_factory_3d
from cv_ssl_mis_b200.trainers import MeanTeacherTrainer

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B, Lb, P = 4, 2, 96
s, t = net_factory_3d("unet_3D", 1, 2, seed=1), net_factory_3d("unet_3D", 1, 2, seed=2)
for p in t.parameters():
    p.detach_()
tr = MeanTeacherTrainer(s, t, batch_size=B, labeled_bs=Lb, patch_size=(P, P, P), num_classes=2, start_iter=1500, use_cuda_graph=True)
g = torch.Generator().manual_seed(0)
x = torch.randn(B, 1, P, P, P, generator=g).cuda()
y = (torch.rand(B, P, P, P, generator=g) > 0.5).long().cuda()
for _ in range(3):
    tr.step(x, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    tr.step(x, y)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"unet_3D Mean Teacher 96^3 bs{B}: {ms:.2f} ms/step = {B / ms * 1e3:.1f} patches/s, {tr.kernel_launches_per_step} launches/step, "
      f"loss {tr.lossbuf[:4].tolist()}")
