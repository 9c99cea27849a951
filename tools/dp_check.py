"""Multi-GPU check (run under torch.distributed.run): a data-parallel step over W ranks must equal a
single-process step on the concatenated batch (global RnC / RMSE semantics, summed gradients)."""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from sdumc_b200.data import synth_batch  # noqa: E402
from sdumc_b200.trainer import Trainer  # noqa: E402

if __name__ == "__main__":
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    dims, frames, B = (256, 512, 128, 512), (48, 16, 32, 12), 24
    torch.manual_seed(100)                      # same initial weights on every rank
    ref = Trainer(dims, B * world, frames, dev, use_graph=False)
    sd = ref.state_dict()
    full = synth_batch(B * world, dims, frames, seed=99, device=dev)
    sl = slice(rank * B, (rank + 1) * B)
    tr = Trainer(dims, B, frames, dev, state_dict=sd, process_group=dist.group.WORLD)
    res = {}
    for dropout in (False, True):
        tr.train_dropout = dropout
        tr.load_batch(*(full[k][sl] for k in ("audio", "text", "video", "feat4")), full["vals"][sl])
        tr.train_step()
        torch.cuda.synchronize()
        if not dropout:
            ref.train_dropout = False
            ref.load_batch(*(full[k] for k in ("audio", "text", "video", "feat4")), full["vals"])
            ref.train_step()
            torch.cuda.synchronize()
            res["terms_dp"] = tr.terms[:7].tolist()
            res["terms_single"] = ref.terms[:7].tolist()
            n = tr.layout.n_live
            d = (tr.master[:n] - ref.master[:n]).abs().max().item()
            upd = (ref.master[:n] - torch.cat([v.flatten() for v in []] or [ref.master[:n]])).abs().max().item()
            res["param_max_abs_diff"] = d
            res["update_max_abs"] = (ref.master[:n] - tr.layout.view(ref.master, "fc_att.weight").new_zeros(1)).abs().max().item()
            g = (tr.grads[:n] - ref.grads[:n]).abs().max().item() / ref.grads[:n].abs().max().item()
            res["grad_rel_diff"] = g
        else:
            res["terms_dp_dropout"] = tr.terms[:7].tolist()
    # a few replays of the captured data-parallel step (NCCL collectives as graph nodes)
    for _ in range(3):
        tr.train_step()
    torch.cuda.synchronize()
    res["graph_captured"] = tr._graph is not None
    res["terms_dp_replay"] = tr.terms[:7].tolist()
    # replicas stay identical
    n = tr.layout.n_live
    mine = tr.master[:n].clone()
    dist.broadcast(mine, src=0)
    res["replica_divergence"] = (mine - tr.master[:n]).abs().max().item()
    if rank == 0:
        print("DPCHECK " + json.dumps(res), flush=True)
    tr.close()
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} clean exit", flush=True)
