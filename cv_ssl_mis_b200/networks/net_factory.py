"""`net_factory(net_type, in_chns, class_num)` -- the reference's 2D model switch
(code/networks/net_factory.py:77-107): a model name -> a constructed module on the GPU; an unknown name
returns None exactly like the reference (its final `else: net = None`).  Unlike the reference, importing this
module has no side effects (no argv parsing, no yaml read -- code/networks/net_factory.py:13-74)."""
from .unet import UNet
from .swin_unet import SwinUnet as ViT_seg


def net_factory(net_type="unet", in_chns=1, class_num=3, config=None, img_size=224, **kw):
    if net_type == "unet":
        return UNet(in_chns=in_chns, class_num=class_num, **kw).cuda()
    if net_type == "ViT_Seg":
        # code/networks/net_factory.py:91-93: ViT_seg(config, img_size=args.patch_size, num_classes=args.num_classes);
        # config None == the yaml-lite defaults the reference parses at import time
        return ViT_seg(config, img_size=img_size, num_classes=class_num, **kw).cuda()
    # TODO(next rows of SURVEY.md 8f): the alternative CNN backbones (enet, pnet, nnUNet, unet_ds/cct/urpc)
    return None
