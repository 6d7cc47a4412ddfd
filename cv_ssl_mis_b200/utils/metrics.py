"""Dice and 95th-percentile Hausdorff distance as the reference's validation uses them (code/val_2D.py:7-15,
code/val_3D.py:82-88, code/utils/metrics.py:13-33 -- all through `medpy.metric.binary.dc / hd95`).

medpy is a third-party dependency that is neither vendored in the reference nor installed here; `dc` and `hd95` restate
its published algorithm (medpy/metric/binary.py: `dc`, `hd95`, `__surface_distances`) on scipy.ndimage.  CPU code: the
surface-distance transform is not on the training hot path."""
import numpy as np
from scipy.ndimage import binary_erosion, distance_transform_edt, generate_binary_structure


def dc(result, reference):
    """Dice coefficient 2 |A and B| / (|A| + |B|) of two binary objects (0.0 when both are empty)."""
    result = np.atleast_1d(np.asarray(result).astype(bool))
    reference = np.atleast_1d(np.asarray(reference).astype(bool))
    intersection = np.count_nonzero(result & reference)
    size = np.count_nonzero(result) + np.count_nonzero(reference)
    return 2.0 * intersection / float(size) if size else 0.0


def _surface_distances(result, reference, voxelspacing=None, connectivity=1):
    result = np.atleast_1d(np.asarray(result).astype(bool))
    reference = np.atleast_1d(np.asarray(reference).astype(bool))
    footprint = generate_binary_structure(result.ndim, connectivity)
    if 0 == np.count_nonzero(result):
        raise RuntimeError('The first supplied array does not contain any binary object.')
    if 0 == np.count_nonzero(reference):
        raise RuntimeError('The second supplied array does not contain any binary object.')
    result_border = result ^ binary_erosion(result, structure=footprint, iterations=1)
    reference_border = reference ^ binary_erosion(reference, structure=footprint, iterations=1)
    dt = distance_transform_edt(~reference_border, sampling=voxelspacing)
    return dt[result_border]


def hd95(result, reference, voxelspacing=None, connectivity=1):
    """95th percentile of the symmetric surface distances between two binary objects."""
    hd1 = _surface_distances(result, reference, voxelspacing, connectivity)
    hd2 = _surface_distances(reference, result, voxelspacing, connectivity)
    return np.percentile(np.hstack((hd1, hd2)), 95)


def calculate_metric_percase(pred, gt):
    """code/val_2D.py:7-15"""
    pred = np.asarray(pred).copy()
    gt = np.asarray(gt).copy()
    pred[pred > 0] = 1
    gt[gt > 0] = 1
    if pred.sum() > 0:
        return dc(pred, gt), hd95(pred, gt)
    return 0, 0


def cal_metric(gt, pred):
    """code/val_3D.py:82-88"""
    if pred.sum() > 0 and gt.sum() > 0:
        return np.array([dc(pred, gt), hd95(pred, gt)])
    return np.zeros(2)
