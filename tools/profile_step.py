"""Runs a few eager (non-graph) train steps of BASELINE config 2 for ncu / timing breakdowns."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from sdumc_b200.data import S0_DIMS, S0_FRAMES, synth_batch  # noqa: E402
from sdumc_b200.trainer import Trainer  # noqa: E402

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    dev = torch.device("cuda", 0)
    tr = Trainer(S0_DIMS, B, S0_FRAMES, dev, use_graph=False)
    b = synth_batch(B, S0_DIMS, S0_FRAMES, device=dev)
    tr.load_batch(b["audio"], b["text"], b["video"], b["feat4"], b["vals"])
    torch.cuda.synchronize()
    for _ in range(steps):
        torch.cuda.nvtx.range_push("step")
        tr.train_step()
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
    print("terms", tr.terms.tolist())
