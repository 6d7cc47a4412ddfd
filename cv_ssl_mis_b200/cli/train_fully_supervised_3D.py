"""Command line of code/train_fully_supervised_3D.py (default --model unet_3D): the loop of cli/train_fully_supervised_3D_ViT.py with the defaults of this script."""
import sys

from . import train_fully_supervised_3D_ViT as _impl

DEFAULTS = dict(exp='BraTS2019/Fully_Supervised', model='unet_3D')


def main(argv=None, loader=None):
    return _impl.main(argv, loader, defaults=DEFAULTS)


if __name__ == "__main__":
    print(main(sys.argv[1:]))
