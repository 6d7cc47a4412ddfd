"""`net_factory_3d(net_type, in_chns, class_num)` -- the reference's 3D model switch
(code/networks/net_factory_3d.py:10-41).  "vnet" builds VNet(normalization='batchnorm', has_dropout=True) exactly like
the reference (:18-20); names whose kernels are not built yet return None, the reference's own answer for unknown names
(:39-40)."""
from .vnet import VNet


def net_factory_3d(net_type="unet_3D", in_chns=1, class_num=2, **kw):
    if net_type == "vnet":
        return VNet(n_channels=in_chns, n_classes=class_num, normalization="batchnorm", has_dropout=True, **kw).cuda()
    # TODO(SURVEY.md 8f): unet_3D (the 3D trainers' default --model), unetr (needs MONAI semantics), attention_unet, ...
    return None
