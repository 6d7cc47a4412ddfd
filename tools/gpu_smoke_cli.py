"""One-shot GPU check of the Python-level paths (command line with a CUDA graph, batched validation, sliding window)."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
os.chdir(tempfile.mkdtemp())
from cv_ssl_mis_b200.cli import train_mean_teacher_2D as cli
print(cli.main(["--batch_size", "4", "--labeled_bs", "2", "--patch_size", "64", "64", "--max_iterations", "4", "--log_every", "2",
                "--save_every", "0", "--exp", "smoke/MT"]))
from cv_ssl_mis_b200 import val_2D, val_3D
from cv_ssl_mis_b200.networks.net_factory import net_factory
from cv_ssl_mis_b200.networks.net_factory_3d import net_factory_3d
net = net_factory("unet", 1, 4)
g = torch.Generator().manual_seed(0)
m = val_2D.test_single_volume(torch.rand(1, 6, 80, 72, generator=g), torch.randint(0, 4, (1, 6, 80, 72), generator=g), net, 4, [64, 64])
print("val_2D", [tuple(round(float(v), 4) for v in t) for t in m])
v = net_factory_3d("vnet", 1, 2)
v.eval()
lm = val_3D.test_single_case(v, np.random.default_rng(0).standard_normal((40, 36, 28)).astype(np.float32), 16, 16, (32, 32, 32), 2)
print("val_3D", lm.shape, int(lm.sum()))
