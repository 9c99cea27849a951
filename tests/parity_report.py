"""Prints every parity number the GPU tests assert on (python -m tests.parity_report [--json out]): the CUDA path
against the UNMODIFIED fp64 oracle (no rounding emulation, no ReLU pinning) and against the emulated one, eval and
train mode.  Used to set / audit the tolerances written in the tests; not collected by pytest."""
from __future__ import annotations

import json
import sys

import torch

from tests.parity_common import run_parity

SMALL = dict(dims=(96, 160, 64, 160), frames=(20, 7, 13, 9), B=32)
S0DIMS = dict(dims=(1024, 4096, 1024, 4096), frames=(150, 40, 100, 37), B=32)


def summarise(res):
    l2 = {k[7:]: v for k, v in res.items() if k.startswith("gradl2/")}
    mx = {k[5:]: v for k, v in res.items() if k.startswith("grad/")}
    outs = {k: v for k, v in res.items() if not k.startswith(("grad", "stat/"))}
    worst_l2 = sorted(l2.items(), key=lambda kv: -kv[1])[:6]
    return {"outputs_max": max(outs.values()), "outputs": {k: float(f"{v:.3g}") for k, v in outs.items()},
            "grad_l2_max": max(l2.values()) if l2 else None, "grad_max_max": max(mx.values()) if mx else None,
            "grad_l2_n_over_2e-2": sum(v > 2e-2 for v in l2.values()),
            "grad_max_n_over_2e-2": sum(v > 2e-2 for v in mx.values()),
            "grad_l2_worst": [(k, float(f"{v:.3g}")) for k, v in worst_l2],
            "stats": {k: v for k, v in res.items() if k.startswith("stat/")}}


def main():
    out = {}
    for cname, cfg in (("small", SMALL), ("s0dims", S0DIMS)):
        for train in (False, True):
            for emulate, pin in ((False, False), (False, True), (True, False), (True, True)):
                for cot in (True, False):
                    if not cot and (cname == "small" or emulate != pin):
                        continue
                    key = (f"{cname}/{'train' if train else 'eval'}/{'emu' if emulate else 'plain'}"
                           f"{'+pin' if pin else ''}/{'vjp' if cot else 'loss'}")
                    res = run_parity(cfg["dims"], cfg["frames"], cfg["B"], gain=1.0, train=train, cotangent=cot,
                                     emulate=emulate, count_flips=True, pin=pin)
                    out[key] = summarise(res)
                    print(key, json.dumps(out[key]), flush=True)
    if len(sys.argv) > 2 and sys.argv[1] == "--json":
        with open(sys.argv[2], "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    torch.manual_seed(0)
    main()
