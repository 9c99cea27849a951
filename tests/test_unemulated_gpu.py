"""Gradient parity against the UNMODIFIED fp64 oracle arithmetic: no rounding emulation (tests/parity_common.emu_*
is NOT used here).  Only the dropout masks are shared (they are the random draws of the step, not a numerical
approximation).  This is north_star's criterion ("gradients must agree within ... <= 2e-2") beside the emulated
comparison of tests/test_model_gpu.py.

Two variants, for the model's vector-Jacobian product with fixed random cotangents, eval and train mode:

  pinned  the exact fp64 oracle with ONE thing taken from the CUDA forward: the on/off pattern of its ReLUs.
          Every one of the 83 live gradient tensors: relative L2 <= 2e-2 at the model's real dimensions (S0 dims
          1024/4096/1024/4096; measured 0.96e-2 eval, 1.14e-2 train, round 2), <= 7.5e-2 at the toy dimensions
          (reductions of 64..160 terms: the same bf16 storage noise is relatively larger; measured 3.3e-2 / 6.1e-2).
  free    the exact fp64 oracle with its own ReLU pattern.  Units whose pre-activation lies within bf16 rounding noise
          of 0 fall on different sides in the two forwards (asserted: < 2e-3 of all units, dropped units excluded;
          measured 4-7e-4) and each such unit switches a whole (sample, unit) backward signal on or off.  That is a
          discontinuity of the model, not of the kernels: tests/test_oracle_golden.py::
          test_relu_flips_between_two_cpu_evaluations_move_gradients_as_much shows the same size of effect between
          two CPU evaluations of the oracle.  Bound asserted here: relative L2 <= 0.25 (S0 dims; measured 0.09 eval,
          0.13 train) - a regression canary, not a precision claim.
A per-sample exclusion of flipped units is vacuous: with ~12.7k ReLU units per sample-pass every sample owns one
(stat/sample_flip_frac = 1.0, tests/parity_report.py).
"""
import pytest

from tests.parity_common import run_parity

pytestmark = pytest.mark.gpu

PINNED_L2_TOL = {"s0dims": 2e-2, "small": 7.5e-2}
FREE_L2_TOL = {"s0dims": 0.25, "small": 0.5}
OUT_TOL = {"s0dims": 1e-2, "small": 6e-2}      # outputs vs the exact oracle, max|err| / max|ref| (train mode included)

CFGS = {"small": dict(dims=(96, 160, 64, 160), frames=(20, 7, 13, 9), B=32),
        "s0dims": dict(dims=(1024, 4096, 1024, 4096), frames=(150, 40, 100, 37), B=32)}


def _run(name, train, pin):
    cfg = CFGS[name]
    res = run_parity(cfg["dims"], cfg["frames"], cfg["B"], gain=1.0, train=train, cotangent=True, emulate=False,
                     count_flips=True, pin=pin)
    outs = {k: v for k, v in res.items() if not k.startswith(("grad", "stat/"))}
    bad = {k: v for k, v in outs.items() if not (v <= OUT_TOL[name])}
    assert not bad, "outputs vs exact oracle: " + ", ".join(f"{k}={v:.3g}" for k, v in sorted(bad.items()))
    assert res["stat/relu_flip_frac"] < 2e-3, res["stat/relu_flip_frac"]
    l2 = {k: v for k, v in res.items() if k.startswith("gradl2/")}
    assert len(l2) == 83
    return l2


@pytest.mark.parametrize("train", [False, True], ids=["eval", "train"])
@pytest.mark.parametrize("name", ["small", "s0dims"])
def test_vjp_against_exact_fp64_oracle_with_the_relu_pattern_pinned(name, train):
    l2 = _run(name, train, pin=True)
    bad = {k: v for k, v in l2.items() if not (v <= PINNED_L2_TOL[name])}
    assert not bad, "gradient L2 error vs exact oracle (ReLU pattern pinned): " + \
        ", ".join(f"{k}={v:.3g}" for k, v in sorted(bad.items()))


@pytest.mark.parametrize("train", [False, True], ids=["eval", "train"])
@pytest.mark.parametrize("name", ["small", "s0dims"])
def test_vjp_against_exact_fp64_oracle_unpinned(name, train):
    l2 = _run(name, train, pin=False)
    bad = {k: v for k, v in l2.items() if not (v <= FREE_L2_TOL[name])}
    assert not bad, "gradient L2 error vs exact oracle (free ReLU pattern): " + \
        ", ".join(f"{k}={v:.3g}" for k, v in sorted(bad.items()))
