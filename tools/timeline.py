"""Kernel timeline of the CUDA-graph train step (BASELINE config 2) from CUPTI via torch.profiler - the image has no
nsys.  Writes gpurun_out/timeline_<tag>.json (every kernel of ONE replay: name, start, duration, stream) and prints:
the step span, the time during which no kernel runs, the time during which exactly one / several kernels run, and
per-kernel totals.

    python tools/timeline.py [tag] [--batch 512] [--score]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/timeline.py tag   # rank 0's view of
                                                                                                       # the data-parallel step
"""
import json
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from sdumc_b200.data import S0_DIMS, S0_FRAMES, synth_batch  # noqa: E402
from sdumc_b200.trainer import Trainer  # noqa: E402


def short(name):
    name = name.replace("void ", "").replace("sdumc::", "").replace("(anonymous namespace)::", "")
    return name.split("(")[0][:70]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "step"
    B = int(sys.argv[sys.argv.index("--batch") + 1]) if "--batch" in sys.argv else 512
    score = "--score" in sys.argv
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    pg = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    tr = Trainer(S0_DIMS, B, S0_FRAMES, dev, seed=100, process_group=pg)
    batch = synth_batch(B, S0_DIMS, S0_FRAMES, seed=1234 + rank, device=dev)
    tr.load_batch(batch["audio"], batch["text"], batch["video"], batch["feat4"], batch["vals"])
    run = tr.score if score else tr.train_step
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            run()
        torch.cuda.synchronize()
    if rank != 0:
        tr.close()
        import torch.distributed as dist
        dist.destroy_process_group()
        return
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
    ks = sorted(({"name": short(e.name), "t0": e.time_range.start, "t1": e.time_range.end,
                  "stream": getattr(e, "stream", -1) if hasattr(e, "stream") else -1} for e in evs), key=lambda k: k["t0"])
    # three identical replays were recorded (graph launches are serialised): the middle third is one step
    assert len(ks) % 3 == 0, f"{len(ks)} kernels recorded for 3 replays"
    n = len(ks) // 3
    step = ks[n:2 * n]
    t0 = step[0]["t0"]
    for k in step:
        k["t0"] -= t0
        k["t1"] -= t0
    span = max(k["t1"] for k in step)
    # occupancy profile
    pts = sorted([(k["t0"], 1) for k in step] + [(k["t1"], -1) for k in step])
    idle = one = multi = 0.0
    cur, last = 0, 0.0
    for t, d in pts:
        dt = t - last
        if cur == 0:
            idle += dt
        elif cur == 1:
            one += dt
        else:
            multi += dt
        cur += d
        last = t
    tot = defaultdict(lambda: [0, 0.0])
    for k in step:
        tot[k["name"]][0] += 1
        tot[k["name"]][1] += k["t1"] - k["t0"]
    print(f"TIMELINE {tag}: {len(step)} kernels, span {span:.1f} us, sum of durations {sum(v[1] for v in tot.values()):.1f} us, "
          f"no kernel {idle:.1f} us, one kernel {one:.1f} us, >= 2 kernels {multi:.1f} us")
    for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f"  {us:9.1f} us  x{n:<3d} {name}")
    out = ROOT / "gpurun_out" / f"timeline_{tag}.json"
    out.parent.mkdir(exist_ok=True)
    out.write_text(json.dumps({"span_us": span, "idle_us": idle, "one_us": one, "multi_us": multi, "kernels": step}))
    if world > 1:
        # the loss section of the data-parallel step in launch order: from the last pooling kernel to the first
        # attn_bwd kernel (collectives, Rank-N-Contrast, backward chain)
        t_a = max((k["t1"] for k in step if "pool_fwd" in k["name"]), default=0.0)
        t_b = min((k["t0"] for k in step if "attn_bwd" in k["name"]), default=span)
        print(f"  loss section {t_a:.1f} .. {t_b:.1f} us")
        for k in step:
            if t_a - 150 <= k["t0"] <= t_b and ("nccl" in k["name"].lower() or "rnc" in k["name"] or "loss" in k["name"]):
                print(f"    {k['t0']:8.1f} {k['t1'] - k['t0']:7.1f}  {k['name']}")
        for k in step:
            if "nccl" in k["name"].lower() and not (t_a - 150 <= k["t0"] <= t_b):
                print(f"    {k['t0']:8.1f} {k['t1'] - k['t0']:7.1f}  {k['name']}")
        tr.close()
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
