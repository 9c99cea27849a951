"""GPU parity of the drop-in nn.Module + loss modules (sm_100a kernels through the C ABI) against the
oracle on identical seeded inputs, parameters and dropout masks (the kernels' Philox masks are
materialised through the C ABI test hooks and replayed by the oracle).

Tolerances (BASELINE.json north_star: bf16 operands, fp32 accumulation):
  * vs the EXACT fp64 oracle: predictions and loss terms max|err| / max|ref| <= 1e-2, embeddings
    <= 2.5e-2 (the reference's own bf16-autocast forward differs from its fp32 forward
    by 0.5-2.1e-2 on the embeddings);
  * gradients vs the oracle evaluated with the CUDA path's rounding points (bf16 storage of H / K / the
    backward GEMM operands, tf32 forward MLP operands — tests/parity_common.emu_*): relative L2 error
    <= 2e-2 for every parameter tensor and max|err| / max|ref| <= 2e-2 for at least 85 % of them.
    Why not the exact oracle / a pure max-norm for gradients: a ReLU unit whose pre-activation is within
    forward rounding noise of 0 flips between the two implementations and changes one row of a weight
    gradient by O(1/rows) of its magnitude — a discontinuity of the model, not an error of the kernels
    (measured: with the rounding points emulated the forward agrees to 4e-4 and gradients to ~5e-3, except
    those isolated rows).
"""
import pytest
import torch

from tests.parity_common import run_parity

pytestmark = pytest.mark.gpu

OUT_TOL = 1e-2
EMB_TOL = 2.5e-2
GRAD_TOL = 2e-2

SMALL = dict(dims=(96, 160, 64, 160), frames=(20, 7, 13, 9), B=32)
S0DIMS = dict(dims=(1024, 4096, 1024, 4096), frames=(150, 40, 100, 37), B=32)


def _check_outputs(res):
    def tol(k):
        return OUT_TOL if (k.endswith("/vals") or k.startswith("term/") or k == "loss") else EMB_TOL
    bad = {k: v for k, v in res.items() if not k.startswith("grad") and not (v <= tol(k))}
    assert not bad, "outputs out of tolerance: " + ", ".join(f"{k}={v:.3g}" for k, v in sorted(bad.items()))


def _check_grads(res):
    l2 = {k: v for k, v in res.items() if k.startswith("gradl2/")}
    mx = {k: v for k, v in res.items() if k.startswith("grad/")}
    assert len(l2) == 83 and len(mx) == 83
    bad = {k: v for k, v in l2.items() if not (v <= GRAD_TOL)}
    assert not bad, "gradient L2 error: " + ", ".join(f"{k}={v:.3g}" for k, v in sorted(bad.items()))
    n_ok = sum(1 for v in mx.values() if v <= GRAD_TOL)
    assert n_ok >= 0.85 * len(mx), f"only {n_ok}/{len(mx)} gradient tensors within max-norm {GRAD_TOL}"


@pytest.mark.parametrize("cfg", [SMALL, S0DIMS], ids=["small", "s0dims"])
def test_forward_matches_exact_oracle(cfg):
    res = run_parity(cfg["dims"], cfg["frames"], cfg["B"], gain=1.0, train=False, cotangent=True, emulate=False)
    _check_outputs(res)


@pytest.mark.parametrize("train", [False, True], ids=["eval", "train"])
@pytest.mark.parametrize("cfg", [SMALL, S0DIMS], ids=["small", "s0dims"])
def test_vjp_random_cotangents(cfg, train):
    """backward of the model alone: d/dparams of sum_p <outputs_p, C_p> for fixed random C."""
    res = run_parity(cfg["dims"], cfg["frames"], cfg["B"], gain=1.0, train=train, cotangent=True, emulate=True)
    _check_outputs(res)
    _check_grads(res)


@pytest.mark.parametrize("train", [False, True], ids=["eval", "train"])
def test_distillation_loss_and_gradients(train):
    """the full 6-term loss of main_frame_val_text_missing.py:148 through the drop-in loss modules."""
    cfg = S0DIMS
    res = run_parity(cfg["dims"], cfg["frames"], cfg["B"], gain=1.0, train=train, emulate=True)
    _check_outputs(res)
    # the RMSE / RnC terms are direction-like functions of differences of nearly equal features at
    # initialisation: they amplify forward rounding noise, so only the L2 criterion is applied, at 5e-2
    l2 = {k: v for k, v in res.items() if k.startswith("gradl2/")}
    bad = {k: v for k, v in l2.items() if not (v <= 5e-2)}
    assert not bad, "gradient L2 error: " + ", ".join(f"{k}={v:.3g}" for k, v in sorted(bad.items()))
