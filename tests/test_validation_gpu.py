"""Validation / inference path on the GPU (code/val_2D.py:18-39, code/val_3D.py:14-79): the batched eval-mode forward and
the device-resident sliding window against the reference's slice-by-slice / patch-by-patch loops run through the same
network, plus a throughput print (slices/s, patches/s)."""
import math
import time

import numpy as np
import pytest
import torch
from scipy.ndimage import zoom

from cv_ssl_mis_b200 import val_2D, val_3D
from cv_ssl_mis_b200.networks.net_factory import net_factory
from cv_ssl_mis_b200.networks.net_factory_3d import net_factory_3d
from cv_ssl_mis_b200.utils.metrics import calculate_metric_percase

pytestmark = pytest.mark.gpu


def test_single_volume_batched_equals_slice_loop():
    torch.manual_seed(2)
    net = net_factory("unet", 1, 4, exact=True)
    g = torch.Generator().manual_seed(3)
    image = torch.rand(1, 12, 200, 180, generator=g)
    label = torch.randint(0, 4, (1, 12, 200, 180), generator=g)
    patch = [256, 256]
    torch.cuda.synchronize()
    t0 = time.time()
    got = val_2D.test_single_volume(image, label, net, 4, patch)
    t_batched = time.time() - t0
    im, lab = image.squeeze(0).numpy(), label.squeeze(0).numpy()
    pred = np.zeros_like(lab)
    net.eval()
    t0 = time.time()
    for ind in range(im.shape[0]):                       # code/val_2D.py:22-33, slice by slice
        sl = im[ind]
        x, y = sl.shape
        sl = zoom(sl, (patch[0] / x, patch[1] / y), order=0)
        inp = torch.from_numpy(sl).unsqueeze(0).unsqueeze(0).float().cuda()
        with torch.no_grad():
            out = torch.argmax(torch.softmax(net(inp), dim=1), dim=1).squeeze(0).cpu().numpy()
        pred[ind] = zoom(out, (x / patch[0], y / patch[1]), order=0)
    t_loop = time.time() - t0
    full = val_2D.predict_volume(im, net, patch)
    assert (full == pred).mean() > 0.999                 # argmax ties of near-equal logits may break differently
    want = [calculate_metric_percase(pred == i, lab == i) for i in range(1, 4)]
    np.testing.assert_allclose(np.array(got, dtype=float)[:, 0], np.array(want, dtype=float)[:, 0], rtol=0, atol=2e-3)
    print(f"val_2D: batched {t_batched * 1e3:.1f} ms (incl. metrics), slice loop {t_loop * 1e3:.1f} ms for {im.shape[0]} slices")


def test_sliding_window_device_equals_patch_loop():
    torch.manual_seed(4)
    net = net_factory_3d("vnet", 1, 2, exact=True)
    net.eval()
    rng = np.random.default_rng(5)
    image = rng.standard_normal((72, 80, 48)).astype(np.float32)
    patch, sxy, sz = (32, 32, 32), 16, 16
    torch.cuda.synchronize()
    t0 = time.time()
    got = val_3D.test_single_case(net, image, sxy, sz, patch, num_classes=2)
    t_dev = time.time() - t0
    ww, hh, dd = image.shape
    score = np.zeros((2,) + image.shape, np.float32)
    cnt = np.zeros(image.shape, np.float32)
    n = 0
    t0 = time.time()
    for x in range(math.ceil((ww - patch[0]) / sxy) + 1):               # code/val_3D.py:39-66
        xs = min(sxy * x, ww - patch[0])
        for y in range(math.ceil((hh - patch[1]) / sxy) + 1):
            ys = min(sxy * y, hh - patch[1])
            for z in range(math.ceil((dd - patch[2]) / sz) + 1):
                zs = min(sz * z, dd - patch[2])
                tp = torch.from_numpy(image[xs:xs + 32, ys:ys + 32, zs:zs + 32][None, None].copy()).cuda()
                with torch.no_grad():
                    yv = torch.softmax(net(tp), dim=1).cpu().numpy()[0]
                score[:, xs:xs + 32, ys:ys + 32, zs:zs + 32] += yv
                cnt[xs:xs + 32, ys:ys + 32, zs:zs + 32] += 1
                n += 1
    t_loop = time.time() - t0
    lm = np.argmax(score / cnt[None], axis=0)
    assert got.shape == image.shape and (got == lm).mean() > 0.999
    print(f"val_3D: device-resident window {t_dev * 1e3:.1f} ms, patch loop {t_loop * 1e3:.1f} ms for {n} patches")
