"""Gradient parity against the UNMODIFIED fp64 oracle: no rounding emulation (tests/parity_common.emu_*), no ReLU
pinning - only the dropout masks are shared (they are the random draws of the step, not a numerical approximation).
This is north_star's criterion as written ("gradients must agree within ... <= 2e-2"), beside the emulated
comparison of tests/test_model_gpu.py which isolates kernel errors from the model's ReLU discontinuities.

What is asserted, for the model's vector-Jacobian product with fixed random cotangents (eval and train mode):
  * every one of the 83 live gradient tensors: relative L2 error <= UNEMU_L2_TOL;
  * the count of ReLU units on which the two forwards disagree (units within bf16 rounding noise of 0; dropped units
    excluded in train mode) as a fraction of all units: < 2e-3, in eval AND train mode.
A per-sample exclusion of flipped units is not meaningful here: with ~12.7k ReLU units per sample-pass and a flip rate
of ~1e-4..1e-3 nearly every sample owns one (the fraction is reported as stat/sample_flip_frac by tests/parity_report.py),
so the bound below is on ALL samples, flips included.
"""
import pytest

from tests.parity_common import run_parity

pytestmark = pytest.mark.gpu

UNEMU_L2_TOL = 2e-2      # relative L2 per tensor vs the plain fp64 oracle (north_star: <= 2e-2 on gradients)
UNEMU_OUT_TOL = 1e-2     # predictions and embeddings, max|err| / max|ref|

SMALL = dict(dims=(96, 160, 64, 160), frames=(20, 7, 13, 9), B=32)
S0DIMS = dict(dims=(1024, 4096, 1024, 4096), frames=(150, 40, 100, 37), B=32)


@pytest.mark.parametrize("train", [False, True], ids=["eval", "train"])
@pytest.mark.parametrize("cfg", [SMALL, S0DIMS], ids=["small", "s0dims"])
def test_vjp_against_the_plain_fp64_oracle(cfg, train):
    res = run_parity(cfg["dims"], cfg["frames"], cfg["B"], gain=1.0, train=train, cotangent=True, emulate=False,
                     count_flips=True)
    outs = {k: v for k, v in res.items() if not k.startswith(("grad", "stat/"))}
    bad = {k: v for k, v in outs.items() if not (v <= UNEMU_OUT_TOL)}
    assert not bad, "outputs vs plain oracle: " + ", ".join(f"{k}={v:.3g}" for k, v in sorted(bad.items()))
    assert res["stat/relu_flip_frac"] < 2e-3, res["stat/relu_flip_frac"]
    l2 = {k: v for k, v in res.items() if k.startswith("gradl2/")}
    assert len(l2) == 83
    bad = {k: v for k, v in l2.items() if not (v <= UNEMU_L2_TOL)}
    assert not bad, "gradient L2 error vs plain oracle: " + ", ".join(f"{k}={v:.3g}" for k, v in sorted(bad.items()))
