"""Validation path: metrics restated from medpy against brute force, batched `test_single_volume`
and device-resident sliding-window inference against line-by-line transcriptions of code/val_2D.py:18-39 and
code/val_3D.py:14-79 (CPU, stand-in ops)."""
import math

import numpy as np
import pytest
import torch
from scipy.ndimage import zoom

from tests import fake_ops


@pytest.fixture()
def fake(monkeypatch):
    fake_ops.install(monkeypatch)
    return fake_ops


def _brute_hd95(a, b):
    """Surface voxels by definition (an object voxel with a background 4-/6-neighbour or on the array border)."""
    def border(m):
        p = np.pad(m, 1, constant_values=False)
        inner = np.ones_like(m, dtype=bool)
        for ax in range(m.ndim):
            for sh in (-1, 1):
                inner &= np.roll(p, sh, ax)[tuple(slice(1, -1) for _ in range(m.ndim))]
        return m & ~inner
    pa, pb = np.argwhere(border(a)), np.argwhere(border(b))
    d = np.sqrt(((pa[:, None, :] - pb[None, :, :]) ** 2).sum(-1))
    return np.percentile(np.hstack((d.min(1), d.min(0))), 95)


@pytest.mark.parametrize("shape", [(12, 14), (6, 8, 7)])
def test_dc_and_hd95_against_brute_force(shape):
    from cv_ssl_mis_b200.utils import metrics as M
    rng = np.random.default_rng(len(shape))
    for _ in range(5):
        a = rng.random(shape) > 0.55
        b = rng.random(shape) > 0.55
        assert abs(M.dc(a, b) - 2 * (a & b).sum() / (a.sum() + b.sum())) < 1e-12
        assert abs(M.hd95(a, b) - _brute_hd95(a, b)) < 1e-9
    assert M.dc(np.zeros(shape, bool), np.zeros(shape, bool)) == 0.0
    with pytest.raises(RuntimeError):
        M.hd95(np.zeros(shape, bool), np.ones(shape, bool))
    assert M.calculate_metric_percase(np.zeros(shape), np.ones(shape)) == (0, 0)
    assert (M.cal_metric(np.ones(shape), np.zeros(shape)) == 0).all()


def test_single_volume_matches_the_reference_loop(fake):
    from cv_ssl_mis_b200 import val_2D
    from cv_ssl_mis_b200.networks.unet import UNet
    from cv_ssl_mis_b200.utils.metrics import calculate_metric_percase
    torch.manual_seed(2)
    net = UNet(1, 4)
    g = torch.Generator().manual_seed(3)
    image = torch.rand(1, 5, 40, 36, generator=g)
    label = torch.randint(0, 4, (1, 5, 40, 36), generator=g)
    patch = [32, 32]
    got = val_2D.test_single_volume(image, label, net, 4, patch)
    # code/val_2D.py:18-39, slice by slice
    im, lab = image.squeeze(0).numpy(), label.squeeze(0).numpy()
    pred = np.zeros_like(lab)
    net.eval()
    for ind in range(im.shape[0]):
        sl = im[ind]
        x, y = sl.shape
        sl = zoom(sl, (patch[0] / x, patch[1] / y), order=0)
        inp = torch.from_numpy(sl).unsqueeze(0).unsqueeze(0).float()
        with torch.no_grad():
            out = torch.argmax(torch.softmax(net(inp), dim=1), dim=1).squeeze(0).numpy()
        pred[ind] = zoom(out, (x / patch[0], y / patch[1]), order=0)
    want = [calculate_metric_percase(pred == i, lab == i) for i in range(1, 4)]
    assert len(got) == 3
    np.testing.assert_allclose(np.array(got, dtype=float), np.array(want, dtype=float), rtol=1e-9, atol=1e-12)


def test_sliding_window_matches_the_reference_loop(fake):
    from cv_ssl_mis_b200 import val_3D
    from cv_ssl_mis_b200.networks.vnet import VNet
    torch.manual_seed(4)
    net = VNet(n_channels=1, n_classes=2, normalization="batchnorm", has_dropout=True)
    net.eval()
    rng = np.random.default_rng(5)
    for shape in [(20, 24, 18), (12, 20, 16)]:                      # larger than / smaller than the 16^3 patch in one dim
        image = rng.standard_normal(shape).astype(np.float32)
        patch, sxy, sz = (16, 16, 16), 8, 8
        got = val_3D.test_single_case(net, image, sxy, sz, patch, num_classes=2)
        # code/val_3D.py:14-79
        w, h, d = image.shape
        pads = [max(p - s, 0) for p, s in zip(patch, shape)]
        img = np.pad(image, [(p // 2, p - p // 2) for p in pads], mode='constant') if any(pads) else image
        ww, hh, dd = img.shape
        score = np.zeros((2,) + img.shape, np.float32)
        cnt = np.zeros(img.shape, np.float32)
        for x in range(math.ceil((ww - patch[0]) / sxy) + 1):
            xs = min(sxy * x, ww - patch[0])
            for y in range(math.ceil((hh - patch[1]) / sxy) + 1):
                ys = min(sxy * y, hh - patch[1])
                for z in range(math.ceil((dd - patch[2]) / sz) + 1):
                    zs = min(sz * z, dd - patch[2])
                    tp = torch.from_numpy(img[xs:xs + 16, ys:ys + 16, zs:zs + 16][None, None].astype(np.float32))
                    with torch.no_grad():
                        yv = torch.softmax(net(tp), dim=1).numpy()[0]
                    score[:, xs:xs + 16, ys:ys + 16, zs:zs + 16] += yv
                    cnt[xs:xs + 16, ys:ys + 16, zs:zs + 16] += 1
        lm = np.argmax(score / cnt[None], axis=0)
        if any(pads):
            lm = lm[pads[0] // 2:pads[0] // 2 + w, pads[1] // 2:pads[1] // 2 + h, pads[2] // 2:pads[2] // 2 + d]
        assert got.shape == image.shape
        assert (got == lm).mean() > 0.999                              # ties of the averaged scores may break differently
