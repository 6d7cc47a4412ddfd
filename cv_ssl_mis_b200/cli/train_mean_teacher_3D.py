"""Command line of code/train_mean_teacher_3D.py (default --model unet_3D, like the reference): the loop of cli/train_uncertainty_aware_mean_teacher_3D.py with the defaults of this script."""
import sys

from . import train_uncertainty_aware_mean_teacher_3D as _impl

DEFAULTS = dict(exp='BraTs2019_Mean_Teacher', model='unet_3D', uncertainty_T=0)


def main(argv=None, loader=None):
    return _impl.main(argv, loader, defaults=DEFAULTS)


if __name__ == "__main__":
    print(main(sys.argv[1:]))
