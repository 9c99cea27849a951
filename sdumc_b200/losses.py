"""Drop-in loss modules: same constructors and call signatures as toolkit/utils/loss.py of the
reference (MSELoss :19-33, RMSELoss :37-51, RnCLoss :271-315), computed by the sm_100a kernels
(csrc/loss.cu) and connected to autograd.  CUDA only — there is no CPU fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib, ops


def _flatten_pair(pred, target):
    # reshape rules of the reference (loss.py:26-31, :44-49)
    if pred.dim() == 1 or target.dim() == 1:
        return pred.reshape(-1, 1), target.reshape(-1, 1)
    if pred.dim() == 3 and target.dim() == 3:
        return pred.reshape(pred.shape[0], -1), target.reshape(target.shape[0], -1)
    return pred, target


def _require_cuda(t):
    if not t.is_cuda:
        raise _lib.SdumcError("sdumc_b200 losses run on CUDA (sm_100a) only; there is no CPU fallback")


class _SqDiff(torch.autograd.Function):
    """mode 0: sum((a-b)^2)/rows (MSELoss);  mode 1: sqrt(mean((a-b)^2)) (RMSELoss)."""

    @staticmethod
    def forward(ctx, a, b, mode):
        _require_cuda(a)
        a32, b32 = a.contiguous().float(), b.contiguous().float()
        if a32.shape != b32.shape:
            raise RuntimeError(f"shape mismatch {tuple(a32.shape)} vs {tuple(b32.shape)}")
        s = torch.zeros(1, dtype=torch.float32, device=a.device)
        ops.sqdiff_sum(a32, b32, s)
        rows, n = a32.shape[0], a32.numel()
        out = s / rows if mode == 0 else torch.sqrt(s / n)
        ctx.save_for_backward(a32, b32, out)
        ctx.mode, ctx.rows, ctx.n = mode, rows, n
        ctx.shapes = (a.shape, b.shape)
        return out.reshape(())

    @staticmethod
    def backward(ctx, go):
        a32, b32, out = ctx.saved_tensors
        if ctx.mode == 0:
            coef = go.reshape(1).float() * (2.0 / ctx.rows)
        else:
            coef = go.reshape(1).float() / (ctx.n * out)     # inf at out == 0, like torch.sqrt's backward
        da = torch.empty_like(a32)
        db = torch.empty_like(b32) if ctx.needs_input_grad[1] else None
        ops.sqdiff_grad(a32, b32, coef.contiguous(), da, db)
        return (da.view(ctx.shapes[0]) if ctx.needs_input_grad[0] else None,
                db.view(ctx.shapes[1]) if db is not None else None, None)


class MSELoss(nn.Module):
    def forward(self, pred, target):
        p, t = _flatten_pair(pred, target)
        return _SqDiff.apply(p, t, 0)


class RMSELoss(nn.Module):
    def forward(self, pred, target):
        p, t = _flatten_pair(pred, target)
        return _SqDiff.apply(p, t, 1)


class _RnC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, labels, temperature):
        _require_cuda(features)
        B, V, D = features.shape
        if V != 2 or labels.reshape(B, -1).shape[1] != 1:
            raise NotImplementedError("RnCLoss: features [B,2,D] and scalar labels [B,1] (the SDUMC use) are supported")
        feats = torch.cat([features[:, 0], features[:, 1]], dim=0).contiguous().float()      # loss.py:282
        y = labels.reshape(B).float().repeat(2).contiguous()                                  # loss.py:283
        loss = torch.zeros(1, dtype=torch.float32, device=features.device)
        need = ctx.needs_input_grad[0]
        dfeats = torch.zeros_like(feats) if need else None
        ops.rnc(feats, y, loss=loss, dfeats=dfeats, temperature=float(temperature))
        ctx.dfeats, ctx.B, ctx.D = dfeats, B, D
        return loss.reshape(())

    @staticmethod
    def backward(ctx, go):
        if ctx.dfeats is None:
            return None, None, None
        d = ctx.dfeats.view(2, ctx.B, ctx.D).transpose(0, 1) * go
        return d, None, None


class RnCLoss(nn.Module):
    def __init__(self, temperature=2, label_diff='l1', feature_sim='l2'):
        super().__init__()
        if label_diff != 'l1' or feature_sim != 'l2':
            raise ValueError("only label_diff='l1', feature_sim='l2' exist in the reference (loss.py:252,:266)")
        self.t = temperature

    def forward(self, features, labels):
        return _RnC.apply(features, labels, self.t)
