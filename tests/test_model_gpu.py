"""GPU parity of the drop-in nn.Module + loss modules (sm_100a kernels through the C ABI) against the
oracle on identical seeded inputs, parameters and dropout masks (the kernels' Philox masks are
materialised through the C ABI test hooks and replayed by the oracle).

Tolerances (BASELINE.json north_star: bf16 operands, fp32 accumulation):
  * vs the EXACT fp64 oracle: predictions and loss terms max|err| / max|ref| <= 1e-2, embeddings
    <= 2.5e-2 (the reference's own bf16-autocast forward differs from its fp32 forward
    by 0.5-2.1e-2 on the embeddings);
  * gradients vs the oracle evaluated (a) with the CUDA path's rounding points (bf16 storage of H / K /
    the backward GEMM operands, tf32 forward MLP operands — tests/parity_common.emu_*) and (b) with every
    ReLU's on/off pattern pinned to the one the CUDA forward took (the saved activations are read back
    through a test hook; the fraction of units on which the oracle's own sign disagrees is asserted to be
    < 0.2 %): relative L2 error <= 2e-2 for every parameter tensor and max|err| / max|ref| <= 2e-2 for at
    least 85 % of them.
    Why (b): a ReLU unit whose pre-activation is within forward rounding noise of 0 flips between two
    implementations and changes a whole sample's backward signal — a discontinuity of the model, not an
    error of the kernels (measured without (b): forward agrees to 4e-4, gradients to ~5e-3 except where a
    flipped unit sits upstream).
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
    bad = {k: v for k, v in res.items() if not k.startswith(("grad", "stat/")) and not (v <= tol(k))}
    if "stat/relu_flip_frac" in res:   # units within rounding noise of 0: must be rare
        assert res["stat/relu_flip_frac"] < 2e-3, res["stat/relu_flip_frac"]
    assert not bad, "outputs out of tolerance: " + ", ".join(f"{k}={v:.3g}" for k, v in sorted(bad.items()))


def _check_grads(res):
    l2 = {k: v for k, v in res.items() if k.startswith("gradl2/")}
    mx = {k: v for k, v in res.items() if k.startswith("grad/")}
    assert len(l2) == 83 and len(mx) == 83
    # fc_att.bias (3 elements) and fc_out_v.bias (1) are plain sums of signed per-sample terms over the batch:
    # cancellation leaves a norm a few times smaller than the terms, so the same absolute bf16 noise that
    # gives <1e-2 on every other tensor reads larger on them.
    tiny = {"gradl2/fc_att.bias": 2.5 * GRAD_TOL, "gradl2/fc_out_v.bias": 2.5 * GRAD_TOL}
    bad = {k: v for k, v in l2.items() if not (v <= tiny.get(k, GRAD_TOL))}
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
