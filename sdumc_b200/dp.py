"""Data-parallel exchanges that keep the sharded step identical to one process holding the whole batch
(SURVEY.md §8e).  Backend-agnostic (NCCL on the GPUs; the CPU test suite drives the same code over gloo
with the oracle as the compute callback)."""
from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.distributed as dist


def _all_gather(t: torch.Tensor, pg) -> torch.Tensor:
    out = [torch.empty_like(t) for _ in range(dist.get_world_size(pg))]
    dist.all_gather(out, t.contiguous(), group=pg)
    return torch.stack(out, dim=0)


def global_views(rnc_local: torch.Tensor, y_local: torch.Tensor, pg) -> Tuple[torch.Tensor, torch.Tensor]:
    """rnc_local [2,B,D] (view, sample), y_local [B] -> feats_g [2*Bg,D], y_g [2*Bg] in the row order a single
    process would build (toolkit/utils/loss.py:282-283): all view-0 rows (rank-major), then all view-1 rows.
    One collective: both views and the label travel as one packed row [2*D + 1] per sample."""
    _, B, D = rnc_local.shape
    pack = torch.empty(B, 2 * D + 1, dtype=rnc_local.dtype, device=rnc_local.device)
    pack[:, :D] = rnc_local[0]
    pack[:, D:2 * D] = rnc_local[1]
    pack[:, 2 * D] = y_local.to(rnc_local.dtype)
    allp = _all_gather(pack, pg)                                 # [W,B,2D+1]
    W = allp.shape[0]
    feats_g = torch.cat((allp[:, :, :D].reshape(W * B, D), allp[:, :, D:2 * D].reshape(W * B, D)), dim=0).contiguous()
    y_g = allp[:, :, 2 * D].reshape(W * B).to(y_local.dtype).repeat(2).contiguous()
    return feats_g, y_g


def anchor_ranges(B: int, world: int, rank: int):
    """Row ranges of the global 2*Bg problem owned by `rank`: its samples in view 0 and in view 1."""
    Bg = B * world
    return [(v * Bg + rank * B, v * Bg + rank * B + B) for v in range(2)]


def rnc_global(rnc_local: torch.Tensor, y_local: torch.Tensor, pg,
               rnc_fn: Callable[[torch.Tensor, torch.Tensor, int, int, torch.Tensor, torch.Tensor], None],
               extra: torch.Tensor = None):
    """Global Rank-N-Contrast over the sharded batch.  rnc_fn(feats_g, y_g, row_begin, row_end, loss, dfeats)
    adds the share of the loss of anchors [row_begin,row_end) to loss[0] and its gradient w.r.t. ALL rows to
    dfeats.  Returns (global loss [1], d loss / d rnc_local [2,B,D]).
    `extra` (optional 1-D tensor of the same dtype, e.g. the sums of squares of the MSE / RMSE terms) is summed over
    the ranks in place by the same all-reduce: two collectives per step instead of five."""
    world, rank = dist.get_world_size(pg), dist.get_rank(pg)
    B, D = rnc_local.shape[1], rnc_local.shape[2]
    feats_g, y_g = global_views(rnc_local, y_local, pg)
    n = feats_g.shape[0]
    k = 0 if extra is None else extra.numel()
    buf = torch.zeros(n * D + 1 + k, dtype=feats_g.dtype, device=feats_g.device)   # [dfeats | loss | extra]
    dfeats, loss = buf[:n * D].view(n, D), buf[n * D:n * D + 1]
    for lo, hi in anchor_ranges(B, world, rank):
        rnc_fn(feats_g, y_g, lo, hi, loss, dfeats)
    if k:
        assert extra.dtype == buf.dtype and extra.dim() == 1
        buf[n * D + 1:] = extra
    dist.all_reduce(buf, group=pg)
    if k:
        extra.copy_(buf[n * D + 1:])
    return loss.clone(), dfeats.view(2, world, B, D)[:, rank].contiguous()


def reduce_sums(sums: torch.Tensor, pg) -> torch.Tensor:
    """Sums of squares of the MSE / RMSE terms over the global batch (RMSE = sqrt of the GLOBAL mean)."""
    dist.all_reduce(sums, group=pg)
    return sums
