"""Seed recipe shared by the golden generator and the tests (inputs are regenerated, never stored)."""
import torch

CASES = {
    "c1_bw8_32": (dict(n_features=4, n_outputs=3, base_width=8), (1, 4, 32, 32, 32)),
    "c1_bw8_64": (dict(n_features=4, n_outputs=3, base_width=8), (1, 4, 64, 64, 64)),
    "bw16_n2_32": (dict(n_features=4, n_outputs=3, base_width=16), (2, 4, 32, 32, 32)),
    "bw8_convT_32": (dict(n_features=4, n_outputs=3, base_width=8, use_transposed_convolutions=True), (1, 4, 32, 32, 32)),
    "c5like_1ch_5lev_32": (dict(n_features=1, n_outputs=1, base_width=8, encoder_blocks=[1, 2, 2, 4, 4]), (1, 1, 32, 32, 32)),
    "bw8_nonpow2_24x32x40": (dict(n_features=4, n_outputs=3, base_width=8), (1, 4, 24, 32, 40)),
}


def golden_inputs(shape, n_outputs, seed=1):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(shape, generator=g, dtype=torch.float32)
    g2 = torch.Generator().manual_seed(seed + 1)
    t = (torch.rand((shape[0], n_outputs) + tuple(shape[2:]), generator=g2) > 0.7).to(torch.uint8)
    g3 = torch.Generator().manual_seed(seed + 2)
    return x, t, g3


def dropout_mask(n, c, p, gen):
    keep = (torch.rand((n, c), generator=gen) >= p).to(torch.float32)
    return keep / (1.0 - p)
