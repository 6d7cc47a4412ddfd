"""Command line of code/train_cross_pseudo_supervision_2D_ViT.py (cross pseudo supervision between two Swin-UNets; the reference keeps the Cross_Teaching experiment name): the loop of cli/train_cross_teaching_between_cnn_transformer_2D.py with the defaults of this script."""
import sys

from . import train_cross_teaching_between_cnn_transformer_2D as _impl

DEFAULTS = dict(exp='ACDC/Cross_Teaching_Between_CNN_Transformer', pseudo_loss='ce', vit1=1)


def main(argv=None, loader=None, val_loader=None):
    return _impl.main(argv, loader, defaults=DEFAULTS, val_loader=val_loader)


if __name__ == "__main__":
    print(main(sys.argv[1:]))
