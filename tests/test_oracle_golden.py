"""The oracle (CPU restatement) against the committed golden fixtures, which were produced by the UNMODIFIED
reference UNet3D (tests/golden/make_golden.py).  Also, when /root/reference is present (build container), pins the
restatement against the live reference class directly."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import UNetConfig, make_state_dict, unet3d_forward, unet3d_state_dict_spec, dice_loss
from oracle.ref_loader import reference_available, reference_unet3d

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from recipe import CASES, golden_inputs, dropout_mask  # noqa: E402

FAST = ["c1_bw8_32", "bw16_n2_32", "bw8_convT_32", "c5like_1ch_5lev_32", "bw8_nonpow2_24x32x40"]


def _oracle_run(name, dtype=torch.float64):
    kw, shape = CASES[name]
    cfg = UNetConfig(**kw)
    sd = {k: v.requires_grad_(True) for k, v in make_state_dict(cfg, seed=0, dtype=dtype).items()}
    x, t, g3 = golden_inputs(shape, cfg.n_outputs)
    mask = dropout_mask(shape[0], cfg.enc_widths()[0], cfg.dropout, g3)
    logits = unet3d_forward(sd, x.to(dtype), cfg, dropout_mask=mask)
    loss = dice_loss(logits, t)
    loss.backward()
    with torch.no_grad():
        logits_eval = unet3d_forward(sd, x.to(dtype), cfg)
    return cfg, sd, logits.detach(), loss.detach(), logits_eval


@pytest.mark.parametrize("name", FAST)
def test_oracle_matches_golden(name, golden_dir):
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    cfg, sd, logits, loss, logits_eval = _oracle_run(name)
    sub = (slice(None), slice(None), slice(None, None, 4), slice(None, None, 4), slice(None, None, 4))
    np.testing.assert_allclose(logits[sub].numpy(), gold["logits_sub4"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(logits_eval[sub].numpy(), gold["logits_eval_sub4"], rtol=0, atol=2e-6)
    assert abs(float(logits.norm()) - float(gold["logits_norm"])) <= 1e-9 * float(gold["logits_norm"])
    assert abs(float(loss) - float(gold["dice"])) < 1e-12
    keys = [str(k) for k in gold["grad_keys"]]
    norms = {k: float(n) for k, n in zip(keys, gold["grad_norms"])}
    assert sorted(sd) == keys
    for k, p in sd.items():
        assert abs(float(p.grad.norm()) - norms[k]) <= 1e-8 * max(norms[k], 1e-12), k
        gk = "grad::" + k
        if gk in gold:
            np.testing.assert_allclose(p.grad.numpy(), gold[gk], rtol=0, atol=1e-6 * max(1.0, float(np.abs(gold[gk]).max())))


def test_oracle_fp32_within_documented_noise(golden_dir):
    """fp32 CPU vs fp64: the noise floor the parity tolerance (1e-3) is quoted against (8.7e-7)."""
    gold = np.load(os.path.join(golden_dir, "c1_bw8_32.npz"))
    assert float(gold["fp32_vs_fp64_logits_rel"]) < 5e-6
    _, _, logits32, loss32, _ = _oracle_run("c1_bw8_32", torch.float32)
    _, _, logits64, loss64, _ = _oracle_run("c1_bw8_32", torch.float64)
    assert float((logits32.double() - logits64).norm() / logits64.norm()) < 5e-6
    assert abs(float(loss32) - float(loss64)) < 1e-6


def test_state_dict_spec_counts():
    spec = unet3d_state_dict_spec(UNetConfig(n_features=4, n_outputs=3, base_width=32))
    assert len(spec) == 90
    n_params = sum(int(np.prod(s)) for _, s in spec)
    assert abs(n_params - 23.97e6) < 0.02e6                      # 23.97 M (trilinear)
    spec_t = unet3d_state_dict_spec(UNetConfig(n_features=4, n_outputs=3, base_width=32, use_transposed_convolutions=True))
    assert abs(sum(int(np.prod(s)) for _, s in spec_t) - 25.35e6) < 0.02e6


@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted (GPU box)")
@pytest.mark.parametrize("kw,shape", [
    (dict(n_features=4, n_outputs=3, base_width=8), (1, 4, 16, 16, 16)),
    (dict(n_features=2, n_outputs=2, base_width=8, encoder_blocks=[1, 1, 2]), (2, 2, 16, 24, 16)),
    (dict(n_features=4, n_outputs=3, base_width=8, use_transposed_convolutions=True), (1, 4, 16, 16, 16)),
])
def test_oracle_matches_live_reference(kw, shape):
    cfg = UNetConfig(**kw)
    ref = reference_unet3d(**kw).double()
    assert [(k, tuple(v.shape)) for k, v in ref.state_dict().items()] == unet3d_state_dict_spec(cfg)
    sd = make_state_dict(cfg, seed=3, dtype=torch.float64)
    ref.load_state_dict(sd, strict=True)
    ref.eval()
    x = torch.randn(shape, dtype=torch.float64, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        a = ref(x)
        b = unet3d_forward(sd, x, cfg)
    assert float((a - b).abs().max()) < 1e-10
