"""Times the RnC kernels at the data-parallel problem sizes: n = 2 * 512 * world rows, this rank's 1024 anchors
(one contiguous row range in the rank-major row order of sdumc_b200/dp.py)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from sdumc_b200 import ops  # noqa: E402

if __name__ == "__main__":
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    B, D = 512, 64
    n = 2 * B * world
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    feats = torch.randn(n, D, device=dev) * 0.3
    y = (torch.randn(B * world, device=dev).clamp(-3, 3)).repeat(2).contiguous()
    ws = torch.empty(ops.rnc_workspace_bytes(n, D), dtype=torch.uint8, device=dev)
    loss = torch.zeros(1, device=dev)
    df = torch.zeros(n, D, device=dev)
    ranges = [(0, 2 * B)]

    def go():
        for k, (lo, hi) in enumerate(ranges):
            ops.rnc(feats, y, loss=loss, dfeats=df, row_begin=lo, row_end=hi, grad_scale=0.8, workspace=ws,
                    reuse_sort=k > 0)
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        go()
    e1.record()
    torch.cuda.synchronize()
    print(f"RNC world={world} n={n} ms_per_step {e0.elapsed_time(e1) / 10:.3f}")
    if "--kernels" in sys.argv:       # per-kernel durations (CUPTI through torch.profiler; the image has no nsys)
        from collections import defaultdict
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(5):
                go()
            torch.cuda.synchronize()
        tot = defaultdict(lambda: [0, 0.0])
        for e in prof.events():
            if e.device_type == torch.autograd.DeviceType.CUDA:
                tot[e.name.split("(")[0][-60:]][0] += 1
                tot[e.name.split("(")[0][-60:]][1] += e.time_range.end - e.time_range.start
        for name, (cnt, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            print(f"  {us / 5:8.1f} us/step  x{cnt // 5}  {name}")
