"""Golden vectors for the batch collate, produced by the REFERENCE's own padding code.

    python oracle/make_golden_collate.py        (build container only: needs /root/reference)

toolkit/utils/read_data.py does not import here (it needs `prefetch_generator`), so the two functions the
4-feature collater uses (feat_data.py:232-253) - func_mapping_feature_tensor (:139-162) and
pad_to_maxlen_pre_modality_tensor_4 (:223-248) - are cut out of the reference file by their AST nodes and executed
unmodified.  Recorded: ragged per-utterance inputs (values exactly representable in bf16), the stacked padded
batch per modality and the pad lengths -> tests/golden/collate_small.npz.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import ast
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
SRC = Path("/root/reference/toolkit/utils/read_data.py")
WANT = ("func_mapping_feature_tensor", "pad_to_maxlen_pre_modality_tensor_4")

DIMS = (16, 24, 8, 24)
LENGTHS = {  # per modality, per utterance (ragged; one utterance is the longest in every modality, one has 1 frame)
    "audio": (9, 3, 12, 1, 7), "text": (4, 6, 2, 1, 5), "video": (5, 5, 8, 1, 2), "feat4": (3, 6, 1, 1, 4)}


def reference_functions():
    text = SRC.read_text()
    tree = ast.parse(text)
    ns = {"torch": torch, "np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in WANT:
            exec(compile(ast.Module([node], []), str(SRC), "exec"), ns)   # noqa: S102 - the reference's own code
    assert all(w in ns for w in WANT)
    return ns


def make_inputs(seed=20241017):
    rng = np.random.RandomState(seed)
    feats = {}
    for s, D in zip(("audio", "text", "video", "feat4"), DIMS):
        feats[s] = [torch.from_numpy(rng.randint(-64, 65, size=(L, D)).astype(np.float32) / 8.0) for L in LENGTHS[s]]
    return feats


def main():
    ns = reference_functions()
    feats = make_inputs()
    lists = [[x.clone() for x in feats[s]] for s in ("audio", "text", "video", "feat4")]
    a, t, v, f4, pads = ns["pad_to_maxlen_pre_modality_tensor_4"](*lists)
    out = {}
    for s, padded in zip(("audio", "text", "video", "feat4"), (a, t, v, f4)):
        out[f"batch/{s}"] = torch.stack(padded).numpy()                   # feat_data.py:242-247
        for i, x in enumerate(feats[s]):
            out[f"in/{s}/{i}"] = x.numpy()
    out["pads"] = np.asarray(pads, dtype=np.int64)
    dst = ROOT / "tests" / "golden" / "collate_small.npz"
    np.savez_compressed(dst, **out)
    print("wrote", dst, {k: v.shape for k, v in out.items() if k.startswith("batch/")}, "pads", out["pads"].tolist())


if __name__ == "__main__":
    main()
