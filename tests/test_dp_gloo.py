"""world_size-2 gloo test (CPU) of the data-parallel exchanges in sdumc_b200/dp.py: the sharded RnC / RMSE
equal the single-process values, with the oracle as the compute callback."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sdumc_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sdumc_b200 import dp
    g = torch.Generator().manual_seed(5)
    Bg = B * world
    feats = torch.randn(Bg, 2, 64, generator=g, dtype=torch.float64)
    y = (torch.randint(-9, 10, (Bg,), generator=g).double() / 3.0)            # tied labels on purpose
    a = torch.randn(Bg, 256, generator=g, dtype=torch.float64)
    b = torch.randn(Bg, 256, generator=g, dtype=torch.float64)
    sl = slice(rank * B, (rank + 1) * B)
    local = feats[sl].permute(1, 0, 2).contiguous()                            # [2,B,64] (view, sample)

    def rnc_fn(feats_g, y_g, lo, hi, loss, dfeats):
        # rows arrive rank-major ([rank, view, sample]); the oracle wants [Bg,2,D] and numbers its anchors view-major
        n = feats_g.shape[0]
        f = feats_g.clone().requires_grad_(True)
        views = f.view(world, 2, B, -1).permute(0, 2, 1, 3).reshape(Bg, 2, -1)
        yv = y_g.view(world, 2, B)[:, 0].reshape(Bg)
        r = lo // (2 * B)
        assert (lo, hi) == (r * 2 * B, (r + 1) * 2 * B)
        anchors = list(range(r * B, r * B + B)) + list(range(Bg + r * B, Bg + r * B + B))
        l = O.rnc_loss(views, yv.unsqueeze(1), anchors=anchors)
        l.backward()
        loss += l.detach()
        dfeats += f.grad

    # the sums of squares ride on the RnC all-reduce (what the trainer does) ...
    sums = torch.tensor([((a[sl] - b[sl]) ** 2).sum()], dtype=torch.float64)
    loss_g, d_local = dp.rnc_global(local, y[sl].contiguous(), dist.group.WORLD, rnc_fn, extra=sums)
    # ... and the stand-alone reduction gives the same value
    sums2 = torch.tensor([((a[sl] - b[sl]) ** 2).sum()], dtype=torch.float64)
    dp.reduce_sums(sums2, dist.group.WORLD)
    assert torch.equal(sums, sums2)
    # single-process reference on the whole batch
    fr = feats.clone().requires_grad_(True)
    lr = O.rnc_loss(fr, y.unsqueeze(1))
    lr.backward()
    ok = (abs(float(loss_g) - float(lr)) < 1e-12
          and torch.allclose(d_local, fr.grad[sl].permute(1, 0, 2), atol=1e-13)
          and abs(float(torch.sqrt(sums / a.numel())) - float(O.rmse_loss(a, b))) < 1e-13
          and dp.anchor_range(B, world, rank) == (rank * 2 * B, rank * 2 * B + 2 * B))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_sharded_rnc_and_rmse_equal_single_process():
    world, B = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}, res
