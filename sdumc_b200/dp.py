"""Data-parallel exchanges that keep the sharded step identical to one process holding the whole batch
(SURVEY.md §8e).  Backend-agnostic (NCCL on the GPUs; the CPU test suite drives the same code over gloo
with the oracle as the compute callback).

Row order of the global Rank-N-Contrast problem: rank-major - rank r's B view-0 rows, then its B view-1 rows, then
rank r+1's.  The loss (toolkit/utils/loss.py:278-315) is a sum over ordered pairs of rows and does not depend on the
order of the rows (a single process builds "all view-0 rows, then all view-1 rows", :282-283; only the summation
order of floating-point terms differs).  Rank-major order makes every exchange a plain collective on contiguous
memory: the all-gather output IS the [n, D] feature matrix, a rank's anchors are ONE contiguous row range, and the
gradient with respect to the gathered features comes home through a reduce-scatter (each rank needs only the 1/W of
it that belongs to its own samples).
"""
from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.distributed as dist


def _is_gloo(pg) -> bool:
    return dist.get_backend(pg) == "gloo"


def _all_gather_flat(send: torch.Tensor, pg) -> torch.Tensor:
    """[m] per rank -> [W, m]"""
    world = dist.get_world_size(pg)
    out = torch.empty(world, send.numel(), dtype=send.dtype, device=send.device)
    dist.all_gather_into_tensor(out.view(-1), send.contiguous(), group=pg)
    return out


def _reduce_scatter_flat(buf: torch.Tensor, pg) -> torch.Tensor:
    """[W, m] per rank -> sum over ranks of row `rank`: [m]"""
    world, rank = dist.get_world_size(pg), dist.get_rank(pg)
    if _is_gloo(pg):                       # gloo has no reduce_scatter: all-reduce and keep the own row
        dist.all_reduce(buf, group=pg)
        return buf[rank].clone()
    out = torch.empty(buf.shape[1], dtype=buf.dtype, device=buf.device)
    dist.reduce_scatter_tensor(out, buf.view(-1), group=pg)
    return out


def global_views(rnc_local: torch.Tensor, y_local: torch.Tensor, pg) -> Tuple[torch.Tensor, torch.Tensor]:
    """rnc_local [2,B,D] (view, sample), y_local [B] -> feats_g [W*2B, D], y_g [W*2B] in rank-major row order.
    One collective: a rank sends its 2B feature rows followed by its 2B labels (each sample's label once per view)."""
    _, B, D = rnc_local.shape
    n_loc = 2 * B
    send = torch.empty(n_loc * D + n_loc, dtype=rnc_local.dtype, device=rnc_local.device)
    send[:n_loc * D] = rnc_local.reshape(-1)
    send[n_loc * D:n_loc * D + B] = y_local.to(rnc_local.dtype)
    send[n_loc * D + B:] = y_local.to(rnc_local.dtype)
    recv = _all_gather_flat(send, pg)                                    # [W, 2B*D + 2B]
    W = recv.shape[0]
    feats_g = recv[:, :n_loc * D].reshape(W * n_loc, D)                  # (copies: the label columns sit in between)
    y_g = recv[:, n_loc * D:].reshape(W * n_loc).to(y_local.dtype)
    return feats_g, y_g


def anchor_range(B: int, world: int, rank: int) -> Tuple[int, int]:
    """Rows of the global problem owned by `rank` (its samples in both views): one contiguous range."""
    return rank * 2 * B, (rank + 1) * 2 * B


def rnc_global(rnc_local: torch.Tensor, y_local: torch.Tensor, pg,
               rnc_fn: Callable[[torch.Tensor, torch.Tensor, int, int, torch.Tensor, torch.Tensor], None],
               extra: torch.Tensor = None):
    """Global Rank-N-Contrast over the sharded batch.  rnc_fn(feats_g, y_g, row_begin, row_end, loss, dfeats)
    adds the share of the loss of anchors [row_begin,row_end) to loss[0] and its gradient w.r.t. ALL rows to
    dfeats.  Returns (global loss [1], d loss / d rnc_local [2,B,D]).
    `extra` (optional 1-D tensor of the same dtype, e.g. the sums of squares of the MSE / RMSE terms) is summed over
    the ranks in place by the same reduce-scatter (every rank's chunk carries a copy of the local values): two
    collectives per step in total."""
    world, rank = dist.get_world_size(pg), dist.get_rank(pg)
    B, D = rnc_local.shape[1], rnc_local.shape[2]
    n_loc = 2 * B
    feats_g, y_g = global_views(rnc_local, y_local, pg)
    n = feats_g.shape[0]
    k = 0 if extra is None else extra.numel()
    dfeats = torch.zeros(n, D, dtype=feats_g.dtype, device=feats_g.device)
    loss = torch.zeros(1, dtype=feats_g.dtype, device=feats_g.device)
    lo, hi = anchor_range(B, world, rank)
    rnc_fn(feats_g, y_g, lo, hi, loss, dfeats)
    buf = torch.empty(world, n_loc * D + 1 + k, dtype=feats_g.dtype, device=feats_g.device)   # per destination rank:
    buf[:, :n_loc * D] = dfeats.view(world, n_loc * D)                                         # [its rows | loss | extra]
    buf[:, n_loc * D] = loss
    if k:
        assert extra.dtype == buf.dtype and extra.dim() == 1
        buf[:, n_loc * D + 1:] = extra
    out = _reduce_scatter_flat(buf, pg)
    if k:
        extra.copy_(out[n_loc * D + 1:])
    return out[n_loc * D:n_loc * D + 1], out[:n_loc * D].view(2, B, D)


def reduce_sums(sums: torch.Tensor, pg) -> torch.Tensor:
    """Sums of squares of the MSE / RMSE terms over the global batch (RMSE = sqrt of the GLOBAL mean)."""
    dist.all_reduce(sums, group=pg)
    return sums
