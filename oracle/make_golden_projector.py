"""Golden vectors for EncoderProjectorConcat from the REFERENCE class itself.

    python oracle/make_golden_projector.py      (build container only: needs /root/reference)

feature_extraction/llm4wav/extract_wavlm_vicuna.py imports transformers models and local checkpoints at module level,
so the class (:162-185) is cut out of the file by its AST node and executed unmodified.  Recorded (small dims, fp64):
parameters, an input whose length is not a multiple of k, the output -> tests/golden/projector_small.npz.
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import ast
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

ROOT = Path(__file__).resolve().parent.parent
SRC = Path("/root/reference/feature_extraction/llm4wav/extract_wavlm_vicuna.py")


def reference_class():
    tree = ast.parse(SRC.read_text())
    ns = {"torch": torch, "nn": nn}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == "EncoderProjectorConcat":
            exec(compile(ast.Module([node], []), str(SRC), "exec"), ns)   # noqa: S102 - the reference's own code
    return ns["EncoderProjectorConcat"]


def main():
    torch.manual_seed(7)
    k, dim, llm = 5, 32, 48
    ref = reference_class()(k, dim, llm).double()
    x = torch.randn(3, 23, dim, dtype=torch.float64)           # 23 % 5 = 3 trailing frames are discarded
    with torch.no_grad():
        y = ref(x)
    out = {"x": x.numpy(), "y": y.numpy(), "k": np.int64(k)}
    for name, p in ref.state_dict().items():
        out["param/" + name] = p.numpy()
    dst = ROOT / "tests" / "golden" / "projector_small.npz"
    np.savez_compressed(dst, **out)
    print("wrote", dst, "y", y.shape)


if __name__ == "__main__":
    main()
