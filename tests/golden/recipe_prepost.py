"""Seed recipe for the label-map / one-hot fixtures (shared by the generator and the tests; inputs are not stored)."""
import torch

# name -> (label-map shape (n,1,d,h,w), values present, n_labels, labels, seed)
ONE_HOT_CASES = {
    "default_1_to_3": ((1, 1, 6, 7, 8), [0, 1, 2, 3], 3, None, 1),
    "brats_hierarchy": ((1, 1, 8, 8, 8), [0, 1, 2, 4], 3, [[1, 2, 4], [1, 4], 4], 2),       # WT / TC / ET groups
    "explicit_labels_n2": ((2, 1, 5, 6, 7), [0, 2, 4, 7], 3, [2, 4, 7], 3),
    "noisy_values_round": ((1, 1, 6, 6, 6), [0, 1, 2], 2, [1, 2], 4),
}

# name -> (prediction shape (L,d,h,w), labels, kwargs, seed)
LABEL_MAP_CASES = {
    "hierarchy_brats": ((3, 8, 9, 10), [2, 1, 4], dict(label_hierarchy=True, threshold=0.5), 11),
    "argmax_any": ((3, 8, 9, 10), [1, 2, 4], dict(threshold=0.5), 12),
    "argmax_sum": ((3, 8, 9, 10), [1, 2, 4], dict(threshold=0.9, sum_then_threshold=True), 13),
    "grouped_volumes": ((4, 6, 7, 8), [[1, 2], [3, 5]], dict(threshold=0.4), 14),
    "single_channel": ((1, 6, 7, 8), [1], dict(threshold=0.5), 15),
}


def label_map_input(shape, values, seed):
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(0, len(values), shape, generator=g)
    data = torch.tensor(values, dtype=torch.float32)[idx]
    # interpolation-style noise that torch.round removes (one_hot.py:19-20), except exact .5 which nothing produces here
    return data + (torch.rand(shape, generator=g) - 0.5) * 0.6


def prediction_input(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(shape, generator=g)
