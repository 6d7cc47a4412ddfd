"""ORACLE (test infrastructure only -- never imported by the product): CPU restatement of the reference's 3D U-Net,
code/networks/unet_3D.py:73-94 with the blocks of code/networks/utils.py:99-123 (UnetConv3) and :260-276 (UnetUp3_CT), as
net_factory_3d builds it (code/networks/net_factory_3d.py:12-13; feature_scale 4, is_batchnorm=True -> InstanceNorm3d).

PARITY PINNED: tests/golden/unet3d.pt holds logits / loss / gradient norms produced by the reference's own `unet_3D`
module (tests/golden/make_golden.py:unet3d_fixture imports it from /root/reference/code);
tests/test_oracle_golden.py::test_unet3d_oracle_matches_reference_fixture checks this restatement against it.
"""
import torch
import torch.nn.functional as F

EPS = 1e-5          # nn.InstanceNorm3d default (affine=False, track_running_stats=False)


def unet_conv3(sd, prefix, x):
    """UnetConv3.forward (utils.py:120-123): 2 x [Conv3d 3x3x3 pad 1, InstanceNorm3d, ReLU]"""
    for k in ("conv1", "conv2"):
        x = F.conv3d(x, sd[f"{prefix}.{k}.0.weight"], sd[f"{prefix}.{k}.0.bias"], padding=1)
        x = F.relu(F.instance_norm(x, eps=EPS))
    return x


def unet3d_forward(sd, x, drop_masks=None):
    """unet_3D.forward (unet_3D.py:73-94).  drop_masks: None (dropout off) or (mask_center, mask_up1) keep masks already
    scaled by 1 / (1 - p) -- the element-wise nn.Dropout(p=0.3) of :86 and :91."""
    skips = []
    cur = x
    for name in ("conv1", "conv2", "conv3", "conv4"):
        cur = unet_conv3(sd, name, cur)
        skips.append(cur)
        cur = F.max_pool3d(cur, 2)
    cur = unet_conv3(sd, "center", cur)
    if drop_masks is not None:
        cur = cur * drop_masks[0]
    for name, skip in zip(("up_concat4", "up_concat3", "up_concat2", "up_concat1"), reversed(skips)):
        up = F.interpolate(cur, scale_factor=2, mode="trilinear", align_corners=False)      # nn.Upsample(scale_factor=(2,2,2))
        cur = unet_conv3(sd, f"{name}.conv", torch.cat([skip, up], 1))                      # offset padding is 0 for even sizes
    if drop_masks is not None:
        cur = cur * drop_masks[1]
    return F.conv3d(cur, sd["final.weight"], sd["final.bias"])


FILTERS = (16, 32, 64, 128, 256)


def param_shapes(in_channels=1, n_classes=2, filters=FILTERS):
    """state_dict keys and shapes of unet_3D in registration order (unet_3D.py:32-58)."""
    f = filters
    out = []

    def block(prefix, cin, cout):
        out.extend([(f"{prefix}.conv1.0.weight", (cout, cin, 3, 3, 3)), (f"{prefix}.conv1.0.bias", (cout,)),
                    (f"{prefix}.conv2.0.weight", (cout, cout, 3, 3, 3)), (f"{prefix}.conv2.0.bias", (cout,))])

    block("conv1", in_channels, f[0])
    block("conv2", f[0], f[1])
    block("conv3", f[1], f[2])
    block("conv4", f[2], f[3])
    block("center", f[3], f[4])
    for name, lo, sk in (("up_concat4", f[4], f[3]), ("up_concat3", f[3], f[2]), ("up_concat2", f[2], f[1]), ("up_concat1", f[1], f[0])):
        block(f"{name}.conv", lo + sk, sk)
    out.extend([("final.weight", (n_classes, f[0], 1, 1, 1)), ("final.bias", (n_classes,))])
    return out


def fixture_state_dict(seed, in_channels=1, n_classes=2, filters=FILTERS):
    """Deterministic weights for the fixtures (23 MB of parameters are regenerated from the seed instead of being stored):
    kaiming-scaled normal weights, small normal biases, drawn key by key from one CPU generator."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shape in param_shapes(in_channels, n_classes, filters):
        if k.endswith("weight"):
            fan_in = shape[1] * shape[2] * shape[3] * shape[4]
            sd[k] = torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5
        else:
            sd[k] = torch.randn(shape, generator=g) * 0.05
    return sd


def fixture_inputs(seed, B, P, n_classes=2):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 1, P, P, P, generator=g)
    low = torch.randint(0, n_classes, (B, P // 8, P // 8, P // 8), generator=g)
    y = low.repeat_interleave(8, 1).repeat_interleave(8, 2).repeat_interleave(8, 3).long()
    return x, y
