"""`net_factory(net_type, in_chns, class_num)` -- the reference's 2D model switch
(code/networks/net_factory.py:77-107): a model name -> a constructed module on the GPU; an unknown name
returns None exactly like the reference (its final `else: net = None`).  Unlike the reference, importing this
module has no side effects (no argv parsing, no yaml read -- code/networks/net_factory.py:13-74)."""
from .unet import UNet


def net_factory(net_type="unet", in_chns=1, class_num=3, **kw):
    if net_type == "unet":
        return UNet(in_chns=in_chns, class_num=class_num, **kw).cuda()
    # TODO(next rows of SURVEY.md 8f): ViT_Seg (SwinUNet) and the alternative backbones
    return None
