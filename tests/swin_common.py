"""Shared set-up of the Swin-UNet / Cross-Teaching tests: rebuilds, from seeds, exactly the models, inputs and DropPath
draws that tests/golden/make_golden.py:swin_fixture() fed to the reference (nothing under /root/reference is read)."""
import numpy as np
import torch

from oracle import philox
from cv_ssl_mis_b200.networks.unet import UNet
from cv_ssl_mis_b200.networks.swin_unet import SwinUnet, DROPPATH_STREAM


def checksum(sd):
    return float(sum(v.double().abs().sum() for k, v in sd.items() if v.dtype.is_floating_point))


def blocky_labels(gen, B, H, W, ncls, dtype=torch.uint8):
    low = torch.randint(0, ncls, (B, H // 8, W // 8), generator=gen)
    return low.repeat_interleave(8, 1).repeat_interleave(8, 2).to(dtype)


def build_models(g, unet_seed=None, swin_seed=None):
    """(unet, swin) constructed in the fixture's order from its seed, with its seeded perturbation of the Swin's
    1-D parameters and relative-position tables."""
    c = dict(g["cfg"])
    torch.manual_seed(g["seed"])
    unet = UNet(1, 4, seed=unet_seed)
    swin = SwinUnet(c, img_size=c["img_size"], num_classes=4, seed=swin_seed)
    gen = torch.Generator().manual_seed(g["perturb_seed"])
    with torch.no_grad():
        for k, p in swin.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn(p.shape, generator=gen))
            elif k.endswith("relative_position_bias_table"):
                p.add_(0.5 * torch.randn(p.shape, generator=gen))
    return unet, swin


def build_inputs(g):
    gen = torch.Generator().manual_seed(g["gen_seed"])
    P = g["cfg"]["img_size"]
    x = torch.rand(g["B"], 1, P, P, generator=gen)
    y = blocky_labels(gen, g["B"], P, P, 4)
    return x, y


def swin_sd(swin):
    return {k[len("swin_unet."):]: v.clone() for k, v in swin.state_dict().items()}


def block_rates(depths, drop_path_rate):
    """DropPath rate of every block in construction order (encoder, then decoder stages reusing the encoder's rates)."""
    dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]
    enc = list(dpr)
    dec = []
    for inx in range(1, 4):
        j = 3 - inx
        dec += dpr[sum(depths[:j]):sum(depths[:j + 1])]
    return enc + dec


def drop_keeps(seed_plus_off, B, depths, drop_path_rate):
    """Per block (keep_attn, keep_mlp) vectors as the CUDA path draws them: Philox(seed + offset, stream, sample)."""
    keeps = []
    for bi, p in enumerate(block_rates(depths, drop_path_rate)):
        pair = []
        for which in range(2):
            w0 = philox.philox4x32_10(seed_plus_off, DROPPATH_STREAM + 2 * bi + which, np.arange(B, dtype=np.uint64))[0]
            u = (w0 >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
            pair.append(torch.from_numpy((u >= np.float32(p)).astype(np.float32)))
        keeps.append(tuple(pair))
    return keeps
