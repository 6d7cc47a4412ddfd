"""One-shot GPU check of the reference-named command lines that build Swin-UNets / unet_3D (too heavy for the CPU suite):
four iterations each on synthetic batches, CUDA graph on."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.chdir(tempfile.mkdtemp())
from cv_ssl_mis_b200.cli import (train_mean_teacher_ViT, train_fully_supervised_2D_ViT, train_uncertainty_aware_mean_teacher_ViT_2D,
                                 train_interpolation_consistency_training_2D_ViT, train_cross_pseudo_supervision_2D_ViT,
                                 train_mean_teacher_3D, train_fully_supervised_3D)
common = ["--max_iterations", "4", "--log_every", "2", "--save_every", "0"]
small2d = ["--batch_size", "4", "--labeled_bs", "2"]
for mod, extra in ((train_mean_teacher_ViT, small2d), (train_fully_supervised_2D_ViT, ["--batch_size", "2"]),
                   (train_uncertainty_aware_mean_teacher_ViT_2D, small2d), (train_interpolation_consistency_training_2D_ViT, small2d),
                   (train_cross_pseudo_supervision_2D_ViT, small2d),
                   (train_mean_teacher_3D, ["--batch_size", "2", "--labeled_bs", "1", "--patch_size", "64", "64", "64"]),
                   (train_fully_supervised_3D, ["--batch_size", "1", "--patch_size", "64", "64", "64"])):
    print(mod.__name__.split(".")[-1], mod.main(extra + common), flush=True)
    log = [os.path.join(r, f) for r, _, fs in os.walk("../model") for f in fs if f == "log.txt"]
for l in sorted(log):
    print(l, open(l).read().strip().splitlines()[-2][-90:])
