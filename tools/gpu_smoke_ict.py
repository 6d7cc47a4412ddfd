"""GPU check of ICTTrainer (CUDA graph replay, torch glue ops inside the capture)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cv_ssl_mis_b200.networks.net_factory import net_factory
from cv_ssl_mis_b200.trainers import ICTTrainer
torch.manual_seed(0)
tr = ICTTrainer(net_factory("unet", 1, 4), net_factory("unet", 1, 4), batch_size=8, labeled_bs=4, patch_size=(64, 64), num_classes=4,
                start_iter=0, mix_seed=1, use_cuda_graph=True)
g = torch.Generator().manual_seed(1)
x = torch.rand(8, 1, 64, 64, generator=g).pin_memory()
y = torch.randint(0, 4, (8, 64, 64), generator=g).to(torch.uint8).pin_memory()
out = [tr.step(x, y, read_loss=True) for _ in range(4)]
print("ict", [round(o[3], 5) for o in out], [round(o[2], 6) for o in out])
assert all(v == v for o in out for v in o) and out[-1][3] < out[0][3] + 0.1
