"""Command line of code/train_fully_supervised_2D.py: the loop of cli/train_mean_teacher_2D.py with the defaults of this script."""
import sys

from . import train_mean_teacher_2D as _impl

DEFAULTS = dict(exp='ACDC/Fully_Supervised', patch_size=[256, 256], labeled_num=50, supervised=1)


def main(argv=None, loader=None, val_loader=None):
    return _impl.main(argv, loader, defaults=DEFAULTS, val_loader=val_loader)


if __name__ == "__main__":
    print(main(sys.argv[1:]))
