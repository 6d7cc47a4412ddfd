"""Command line of code/train_mean_teacher_ViT.py (Mean Teacher over two Swin-UNets; --model stays 'unet' for the snapshot path, like the reference): the loop of cli/train_mean_teacher_2D.py with the defaults of this script."""
import sys

from . import train_mean_teacher_2D as _impl

DEFAULTS = dict(exp='ACDC/Mean_Teacher_ViT', patch_size=[224, 224], labeled_num=7, vit=1)


def main(argv=None, loader=None, val_loader=None):
    return _impl.main(argv, loader, defaults=DEFAULTS, val_loader=val_loader)


if __name__ == "__main__":
    print(main(sys.argv[1:]))
