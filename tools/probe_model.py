"""GPU bring-up probe: per-tensor parity errors of the nn.Module path vs the oracle."""
import json
import sys
import traceback
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from tests.parity_common import run_parity  # noqa: E402

CASES = {
    "small_eval": dict(dims=(96, 160, 64, 160), frames=(20, 7, 13, 9), B=6, gain=1.3, train=False),
    "small_train": dict(dims=(96, 160, 64, 160), frames=(20, 7, 13, 9), B=6, gain=1.0, train=True),
    "s0_eval": dict(dims=(1024, 4096, 1024, 4096), frames=(150, 40, 100, 37), B=8, gain=1.0, train=False),
    "s0_train": dict(dims=(1024, 4096, 1024, 4096), frames=(150, 40, 100, 37), B=8, gain=1.0, train=True),
    "s0_eval_ct": dict(dims=(1024, 4096, 1024, 4096), frames=(150, 40, 100, 37), B=8, gain=1.0, train=False,
                       cotangent=True),
    "s0_train_ct": dict(dims=(1024, 4096, 1024, 4096), frames=(150, 40, 100, 37), B=8, gain=1.0, train=True,
                        cotangent=True),
    "s0_eval_ct_emu": dict(dims=(1024, 4096, 1024, 4096), frames=(150, 40, 100, 37), B=8, gain=1.0, train=False,
                           cotangent=True, emulate=True),
    "s0_train_ct_emu": dict(dims=(1024, 4096, 1024, 4096), frames=(150, 40, 100, 37), B=8, gain=1.0, train=True,
                            cotangent=True, emulate=True),
    "s0_train_emu": dict(dims=(1024, 4096, 1024, 4096), frames=(150, 40, 100, 37), B=8, gain=1.0, train=True,
                         emulate=True),
    "small_train_ct_emu": dict(dims=(96, 160, 64, 160), frames=(20, 7, 13, 9), B=6, gain=1.0, train=True,
                               cotangent=True, emulate=True),
    "small_eval_ct": dict(dims=(96, 160, 64, 160), frames=(20, 7, 13, 9), B=6, gain=1.0, train=False, cotangent=True),
    "small_train_ct": dict(dims=(96, 160, 64, 160), frames=(20, 7, 13, 9), B=6, gain=1.0, train=True, cotangent=True),
    "s0_eval_emu": dict(dims=(1024, 4096, 1024, 4096), frames=(150, 40, 100, 37), B=8, gain=1.0, train=False,
                        emulate=True),
    "s0_eval_b64": dict(dims=(1024, 4096, 1024, 4096), frames=(150, 40, 100, 37), B=64, gain=1.0, train=False),
    "s0_eval_b64_emu": dict(dims=(1024, 4096, 1024, 4096), frames=(150, 40, 100, 37), B=64, gain=1.0, train=False,
                            emulate=True),
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    out = {}
    for n in names:
        try:
            res = run_parity(**CASES[n])
            out[n] = res
            worst = sorted(res.items(), key=lambda kv: -kv[1] if kv[1] == kv[1] else -1e9)[:12]
            nbad = sum(1 for k, v in res.items() if not (v <= (2e-2 if k.startswith("grad/") else 1e-2)))
            print(f"== {n}: {len(res)} tensors, {nbad} out of tolerance")
            for k, v in worst:
                print(f"   {k:60s} {v:.3e}")
            nan = [k for k, v in res.items() if v != v]
            if nan:
                print("   NaN:", nan[:20])
        except Exception:  # noqa: BLE001
            print(f"== {n}: EXCEPTION")
            traceback.print_exc()
            try:
                torch.cuda.synchronize()
            except Exception as e:  # noqa: BLE001
                print("cuda sync:", e)
                break
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    json.dump(out, open(ROOT / "gpurun_out" / "probe_model.json", "w"), indent=1)
