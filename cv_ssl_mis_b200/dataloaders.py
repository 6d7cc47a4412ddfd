"""Batch composition of the semi-supervised trainers (code/dataloaders/dataset.py:247-294, code/train_mean_teacher_2D.py:108-118).

Only the pieces the training step's contract depends on are restated: every batch is `labeled_bs` labeled indices
followed by `batch_size - labeled_bs` unlabeled ones (the trainers slice `[:labeled_bs]` / `[labeled_bs:]`), one epoch is
one pass over the labeled ("primary") indices while the unlabeled ("secondary") ones are reshuffled for ever.  The h5
readers and CPU augmentations of the reference stay where they are: pass
`DataLoader(dataset, batch_sampler=TwoStreamBatchSampler(...), pin_memory=True)` to `cli.*.main(argv, loader=...)`."""
import numpy as np
from torch.utils.data.sampler import Sampler


class TwoStreamBatchSampler(Sampler):
    """Same constructor, draws and batches as the reference class: `batch_size - secondary_batch_size` primary indices
    (one permutation per epoch) + `secondary_batch_size` secondary indices (an endless chain of permutations)."""

    def __init__(self, primary_indices, secondary_indices, batch_size, secondary_batch_size):
        self.primary_indices = primary_indices
        self.secondary_indices = secondary_indices
        self.secondary_batch_size = secondary_batch_size
        self.primary_batch_size = batch_size - secondary_batch_size
        assert len(self.primary_indices) >= self.primary_batch_size > 0
        assert len(self.secondary_indices) >= self.secondary_batch_size > 0

    def __iter__(self):
        primary = np.random.permutation(self.primary_indices)        # drawn when the epoch starts, like iterate_once()

        def batches():
            pool = iter(())
            for b in range(len(primary) // self.primary_batch_size):
                head = tuple(primary[b * self.primary_batch_size:(b + 1) * self.primary_batch_size])
                tail = []
                while len(tail) < self.secondary_batch_size:
                    try:
                        tail.append(next(pool))
                    except StopIteration:
                        pool = iter(np.random.permutation(self.secondary_indices))
                yield head + tuple(tail)

        return batches()

    def __len__(self):
        return len(self.primary_indices) // self.primary_batch_size


def patients_to_slices(dataset, patiens_num):
    """Labeled-slice counts per number of labeled patients (code/train_mean_teacher_2D.py:108-118)."""
    if "ACDC" in dataset:
        ref_dict = {"3": 68, "7": 136, "14": 256, "21": 396, "28": 512, "35": 664, "140": 1312}
    else:                                   # the reference's `elif "Prostate":` is always true
        ref_dict = {"2": 27, "4": 53, "8": 120, "12": 179, "16": 256, "21": 312, "42": 623}
    return ref_dict[str(patiens_num)]
