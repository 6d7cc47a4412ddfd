"""`test_single_volume` of the reference (code/val_2D.py:18-39) with the per-slice network calls batched.

The reference loops over the slices of a volume: nearest-neighbour zoom to the patch size, `net.eval()` forward of ONE
slice, argmax, zoom back -- one tiny launch sequence and two host<->device copies per slice.  Here the slices are zoomed
with the same scipy call (results identical by construction), pushed through the network in eval mode as ONE batch on
our kernels, arg-maxed on the device and zoomed back; metrics as in the reference (utils/metrics.py)."""
import numpy as np
import torch
from scipy.ndimage import zoom

from .utils.metrics import calculate_metric_percase


def predict_volume(image, net, patch_size=(256, 256), max_batch=64):
    """image: numpy [slices, x, y] -> integer label volume of the same shape."""
    n, x, y = image.shape
    slices = np.stack([zoom(image[i], (patch_size[0] / x, patch_size[1] / y), order=0) for i in range(n)])
    dev = next(net.parameters()).device
    was_training = net.training
    net.eval()
    outs = []
    with torch.no_grad():
        for s in range(0, n, max_batch):
            inp = torch.from_numpy(slices[s:s + max_batch]).unsqueeze(1).float().to(dev)
            outs.append(torch.argmax(torch.softmax(net(inp), dim=1), dim=1).cpu().numpy())
    net.train(was_training)
    out = np.concatenate(outs, 0)
    return np.stack([zoom(out[i], (x / patch_size[0], y / patch_size[1]), order=0) for i in range(n)])


def test_single_volume(image, label, net, classes, patch_size=[256, 256]):
    """image, label: tensors [1, slices, x, y] (the reference's validation batch).  Returns [(dice, hd95)] per class."""
    image, label = image.squeeze(0).cpu().detach().numpy(), label.squeeze(0).cpu().detach().numpy()
    prediction = np.zeros_like(label)
    prediction[...] = predict_volume(image, net, patch_size)
    metric_list = []
    for i in range(1, classes):
        metric_list.append(calculate_metric_percase(prediction == i, label == i))
    return metric_list
