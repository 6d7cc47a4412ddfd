"""Command line of code/train_cross_pseudo_supervision_2D.py (two CNNs, CE on the pseudo labels of the other network): the loop of cli/train_cross_teaching_between_cnn_transformer_2D.py with the defaults of this script."""
import sys

from . import train_cross_teaching_between_cnn_transformer_2D as _impl

DEFAULTS = dict(exp='ACDC/Cross_Pseudo_Supervision', batch_size=24, labeled_bs=12, patch_size=[256, 256], labeled_num=1, model2='unet', pseudo_loss='ce')


def main(argv=None, loader=None, val_loader=None):
    return _impl.main(argv, loader, defaults=DEFAULTS, val_loader=val_loader)


if __name__ == "__main__":
    print(main(sys.argv[1:]))
