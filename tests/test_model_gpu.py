"""GPU parity: the drop-in nn.Module + loss modules (sm_100a kernels through the C ABI) against the
oracle on identical seeded inputs, parameters and dropout masks.

Tolerances (BASELINE.json north_star): bf16 operands / fp32 accumulation ->
  max |err| / max |ref| <= 1e-2 on predictions, embeddings and loss terms, <= 2e-2 on gradients.
"""
import pytest
import torch

from tests.parity_common import run_parity

pytestmark = pytest.mark.gpu

OUT_TOL = 1e-2
GRAD_TOL = 2e-2

SMALL = dict(dims=(96, 160, 64, 160), frames=(20, 7, 13, 9), B=6)
MEDIUM = dict(dims=(1024, 4096, 1024, 4096), frames=(150, 40, 100, 37), B=8)


def _check(res):
    bad = {k: v for k, v in res.items() if v > (GRAD_TOL if k.startswith("grad/") else OUT_TOL) or v != v}
    assert not bad, f"{len(bad)} tensors out of tolerance: " + ", ".join(f"{k}={v:.3g}" for k, v in sorted(bad.items()))


@pytest.mark.parametrize("cfg", [SMALL, MEDIUM], ids=["small", "s0dims"])
def test_eval_mode_forward_loss_backward(cfg):
    _check(run_parity(cfg["dims"], cfg["frames"], cfg["B"], gain=1.3 if cfg is SMALL else 1.0, train=False))


@pytest.mark.parametrize("cfg", [SMALL, MEDIUM], ids=["small", "s0dims"])
def test_train_mode_with_kernel_dropout_masks(cfg):
    _check(run_parity(cfg["dims"], cfg["frames"], cfg["B"], gain=1.0, train=True))
