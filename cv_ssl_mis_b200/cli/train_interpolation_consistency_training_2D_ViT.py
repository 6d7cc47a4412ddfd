"""Command line of code/train_interpolation_consistency_training_2D_ViT.py (ICT over two Swin-UNets): the loop of cli/train_interpolation_consistency_training_2D.py with the defaults of this script."""
import sys

from . import train_interpolation_consistency_training_2D as _impl

DEFAULTS = dict(exp='ACDC/Interpolation_Consistency_Training_ViT', patch_size=[224, 224], labeled_num=7, vit=1)


def main(argv=None, loader=None, val_loader=None):
    return _impl.main(argv, loader, defaults=DEFAULTS, val_loader=val_loader)


if __name__ == "__main__":
    print(main(sys.argv[1:]))
