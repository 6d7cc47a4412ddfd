"""Per-tensor weight-gradient error of one full-size config-2 step against the fp32 CPU oracle, production TF32 path vs the
exact (3xTF32) path: separates accumulated TF32 round-off from indexing bugs.  usage: python tools/grad_fullsize.py [exact]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import philox, ssl_oracle as O
from cv_ssl_mis_b200.networks import unet as unet_mod
from cv_ssl_mis_b200.trainers import MeanTeacherTrainer
from tests.test_host_logic import unet_masks
exact = len(sys.argv) > 1 and sys.argv[1] == "exact"
B, Lb, H, W = 24, 12, 256, 256
torch.manual_seed(33)
student, teacher = unet_mod.UNet(1, 4, seed=301, exact=exact), unet_mod.UNet(1, 4, seed=302, exact=exact)
s_sd = {k: v.clone() for k, v in student.state_dict().items()}
t_sd = {k: v.clone() for k, v in teacher.state_dict().items()}
student, teacher = student.cuda(), teacher.cuda()
tr = MeanTeacherTrainer(student, teacher, batch_size=B, labeled_bs=Lb, patch_size=(H, W), num_classes=4, start_iter=2000, noise_seed=77,
                        use_cuda_graph=False)
lr = tr.lr
g = torch.Generator().manual_seed(4)
x = torch.rand(B, 1, H, W, generator=g)
low = torch.randint(0, 4, (B, H // 16, W // 16), generator=g)
y = low.repeat_interleave(16, 1).repeat_interleave(16, 2).to(torch.uint8)
losses = tr.step(x.pin_memory(), y.pin_memory(), read_loss=True)
grads = {n: p.grad.detach().cpu().clone() for n, p in student.named_parameters()}
noise = torch.from_numpy(philox.clamp_noise(77 + 1, 1000, (B - Lb) * H * W)).reshape(B - Lb, 1, H, W)
bufs = {k: torch.zeros_like(s_sd[k]) for k in O.param_keys(s_sd)}
r = O.mt2d_step(s_sd, t_sd, bufs, x, y, noise, 2000, labeled_bs=Lb, lr=lr,
                student_masks=unet_masks(301 + 1, B, H, W), teacher_masks=unet_masks(302 + 1, B - Lb, H, W))
print("mode", "exact" if exact else "tf32", "losses", losses, "oracle", float(r["loss"]), float(r["ce"]), float(r["dice"]), float(r["cons"]))
for k in O.param_keys(s_sd):
    ref = r["grads"][k]
    if k.endswith("weight") and ref.dim() == 4:
        print(f"{k:55s} |g| {float(ref.norm()):.3e} rel {float((grads[k] - ref).norm() / ref.norm()):.3e}")
