"""TwoStreamBatchSampler against a transcription of the reference's generator pipeline
(code/dataloaders/dataset.py:263-294: iterate_once / iterate_eternally / grouper) under the same numpy seed."""
import itertools

import numpy as np

from cv_ssl_mis_b200.dataloaders import TwoStreamBatchSampler, patients_to_slices


def _reference_batches(primary, secondary, batch_size, secondary_bs):
    def iterate_eternally(indices):
        def infinite_shuffles():
            while True:
                yield np.random.permutation(indices)
        return itertools.chain.from_iterable(infinite_shuffles())

    def grouper(iterable, n):
        return zip(*([iter(iterable)] * n))

    p_iter = np.random.permutation(primary)
    s_iter = iterate_eternally(secondary)
    return [a + b for a, b in zip(grouper(p_iter, batch_size - secondary_bs), grouper(s_iter, secondary_bs))]


def test_two_stream_batches_equal_the_reference_pipeline():
    primary, secondary = list(range(0, 136)), list(range(136, 1312))
    for bs, sbs in ((24, 12), (16, 8), (5, 3)):
        np.random.seed(1337)
        want = _reference_batches(primary, secondary, bs, sbs)
        np.random.seed(1337)
        sampler = TwoStreamBatchSampler(primary, secondary, bs, sbs)
        got = list(sampler)
        assert len(got) == len(sampler) == len(primary) // (bs - sbs)
        assert [tuple(int(i) for i in b) for b in got] == [tuple(int(i) for i in b) for b in want]
        for b in got:                       # labeled first, unlabeled after: the trainers slice [:labeled_bs] / [labeled_bs:]
            assert all(i < 136 for i in b[:bs - sbs]) and all(i >= 136 for i in b[bs - sbs:])


def test_secondary_stream_wraps_around():
    np.random.seed(0)
    got = list(TwoStreamBatchSampler(list(range(10)), [100, 101, 102], 4, 2))
    assert len(got) == 5
    tail = [i for b in got for i in b[2:]]
    assert sorted(tail[:3]) == [100, 101, 102] and sorted(tail[3:6]) == [100, 101, 102]


def test_patients_to_slices():
    assert patients_to_slices("../data/ACDC", 7) == 136 and patients_to_slices("../data/Prostate", 8) == 120
