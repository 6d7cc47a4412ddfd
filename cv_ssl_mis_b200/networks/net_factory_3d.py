"""`net_factory_3d(net_type, in_chns, class_num)` -- the reference's 3D model switch
(code/networks/net_factory_3d.py:10-41).  "vnet" builds VNet(normalization='batchnorm', has_dropout=True) exactly like
the reference (:18-20), "unetr" the UNETR of :27-39, "unet_3D" (the default) the 3D U-Net of :12-13; names whose kernels are not built yet return None, the reference's own answer for unknown names
(:39-40)."""
from .unet_3d import unet_3D
from .unetr import UNETR
from .vnet import VNet


def net_factory_3d(net_type="unet_3D", in_chns=1, class_num=2, **kw):
    if net_type == "unet_3D":
        return unet_3D(n_classes=class_num, in_channels=in_chns, **kw).cuda()
    if net_type == "vnet":
        return VNet(n_channels=in_chns, n_classes=class_num, normalization="batchnorm", has_dropout=True, **kw).cuda()
    if net_type == "unetr":              # :27-39 (fixed 96^3 patches, single input channel, exactly as the reference)
        return UNETR(in_channels=1, out_channels=class_num, img_size=(96, 96, 96), feature_size=16, hidden_size=768,
                     mlp_dim=3072, num_heads=12, pos_embed="perceptron", norm_name="instance", conv_block=True,
                     res_block=True, dropout_rate=0.0, **kw).cuda()
    # (attention_unet, voxresnet, nnUNet, swinunetr: not on the path)
    return None
