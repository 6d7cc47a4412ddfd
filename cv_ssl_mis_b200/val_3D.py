"""Sliding-window inference of the reference (code/val_3D.py:14-79) with the score map kept on the device.

Same window positions (stride_xy / stride_z, the last window clamped to the border), same zero padding of volumes
smaller than the patch, same overlap averaging and argmax; the softmax scores are accumulated in a device tensor instead
of being copied to the host after every window."""
import math

import numpy as np
import torch

from .utils.metrics import cal_metric


def test_single_case(net, image, stride_xy, stride_z, patch_size, num_classes=1):
    w, h, d = image.shape
    pads = [max(p - s, 0) for p, s in zip(patch_size, (w, h, d))]
    add_pad = any(pads)
    lo = [p // 2 for p in pads]
    if add_pad:
        image = np.pad(image, [(l, p - l) for l, p in zip(lo, pads)], mode='constant', constant_values=0)
    ww, hh, dd = image.shape
    sx = math.ceil((ww - patch_size[0]) / stride_xy) + 1
    sy = math.ceil((hh - patch_size[1]) / stride_xy) + 1
    sz = math.ceil((dd - patch_size[2]) / stride_z) + 1
    dev = next(net.parameters()).device
    vol = torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32)).to(dev)
    score_map = torch.zeros((num_classes,) + image.shape, dtype=torch.float32, device=dev)
    cnt = torch.zeros(image.shape, dtype=torch.float32, device=dev)
    with torch.no_grad():
        for x in range(sx):
            xs = min(stride_xy * x, ww - patch_size[0])
            for y in range(sy):
                ys = min(stride_xy * y, hh - patch_size[1])
                for z in range(sz):
                    zs = min(stride_z * z, dd - patch_size[2])
                    sl = (slice(xs, xs + patch_size[0]), slice(ys, ys + patch_size[1]), slice(zs, zs + patch_size[2]))
                    patch = vol[sl].contiguous()[None, None]
                    prob = torch.softmax(net(patch), dim=1)[0]
                    score_map[(slice(None),) + sl] += prob
                    cnt[sl] += 1
    score_map = score_map / cnt.unsqueeze(0)
    label_map = torch.argmax(score_map, dim=0).cpu().numpy()
    if add_pad:
        label_map = label_map[lo[0]:lo[0] + w, lo[1]:lo[1] + h, lo[2]:lo[2] + d]
    return label_map


def evaluate_cases(net, cases, num_classes=4, patch_size=(48, 160, 160), stride_xy=32, stride_z=24):
    """`test_all_case` (code/val_3D.py:91-110) over an iterable of (image, label) numpy volumes instead of h5 paths."""
    total_metric = np.zeros((num_classes - 1, 2))
    n = 0
    for image, label in cases:
        prediction = test_single_case(net, image, stride_xy, stride_z, patch_size, num_classes=num_classes)
        for i in range(1, num_classes):
            total_metric[i - 1, :] += cal_metric(label == i, prediction == i)
        n += 1
    return total_metric / max(n, 1)
