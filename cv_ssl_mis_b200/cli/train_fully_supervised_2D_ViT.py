"""Command line of code/train_fully_supervised_2D_ViT.py (fully supervised Swin-UNet): the loop of cli/train_mean_teacher_2D.py with the defaults of this script."""
import sys

from . import train_mean_teacher_2D as _impl

DEFAULTS = dict(exp='ACDC/Fully_Supervised_ViT', patch_size=[224, 224], labeled_num=7, supervised=1, vit=1)


def main(argv=None, loader=None, val_loader=None):
    return _impl.main(argv, loader, defaults=DEFAULTS, val_loader=val_loader)


if __name__ == "__main__":
    print(main(sys.argv[1:]))
