"""Times Trainer.score() (both eval passes) at BASELINE config 2 size."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from sdumc_b200.data import S0_DIMS, S0_FRAMES, synth_batch  # noqa: E402
from sdumc_b200.trainer import Trainer  # noqa: E402

if __name__ == "__main__":
    B = 512
    dev = torch.device("cuda", 0)
    tr = Trainer(S0_DIMS, B, S0_FRAMES, dev)
    b = synth_batch(B, S0_DIMS, S0_FRAMES, device=dev)
    tr.load_batch(b["audio"], b["text"], b["video"], b["feat4"], b["vals"])
    for _ in range(3):
        tr.score()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        tr.score()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"SCORE ms_per_batch {ms:.3f} samples_per_s {B / ms * 1e3:.0f}")
