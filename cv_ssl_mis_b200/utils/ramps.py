"""Host-side ramp schedules, bit-identical to the reference (code/utils/ramps.py:20-61): the same numpy
float64 expressions, so `sigmoid_rampup(c, l)` returns exactly the reference's python float."""
import numpy as np


def sigmoid_rampup(current, rampup_length):
    """exp(-5 (1 - clip(current, 0, L)/L)^2)   -- code/utils/ramps.py:20-27"""
    if rampup_length == 0:
        return 1.0
    current = np.clip(current, 0.0, rampup_length)
    phase = 1.0 - current / rampup_length
    return float(np.exp(-5.0 * phase * phase))


def linear_rampup(current, rampup_length):
    """code/utils/ramps.py:47-53"""
    assert current >= 0 and rampup_length >= 0
    if current >= rampup_length:
        return 1.0
    return current / rampup_length


def cosine_rampdown(current, rampdown_length):
    """code/utils/ramps.py:56-61"""
    assert 0 <= current <= rampdown_length
    return float(.5 * (np.cos(np.pi * current / rampdown_length) + 1))
