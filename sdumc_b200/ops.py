"""Torch-tensor front ends of the C-ABI operators (include/sdumc_b200.h).

PyTorch is used for device memory and streams only; every function launches hand-written sm_100a
kernels from libsdumc_b200.so on the current CUDA stream and raises SdumcError on failure.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import (ACT_NONE, ACT_RELU, ACT_TANH, EPI_GENERIC, EPI_INPROJ, EPI_KEYPROJ, OUT_ADD, OUT_ATOMIC,
                   OUT_STORE, STRUCTS, GemmDesc, call, check, current_stream, dropkey, ptr)



def _ld(t: torch.Tensor) -> int:
    assert t.dim() == 2 and t.stride(1) == 1, "row-major 2-D tensor expected"
    return t.stride(0)


def gemm(A: torch.Tensor, B: torch.Tensor, *, M: int, N: int, K: int, a_mn=False, b_mn=False, k_splits=1,
         block_n=0, max_ctas=0, bias=None, act=ACT_NONE, gate=None, gate_scale=1.0, drop_p=0.0, drop_site=0,
         fmask_site=0, out_f32=None, f32_mode=OUT_STORE, out_bf16=None, bf16_mode=OUT_STORE, epi_kind=EPI_GENERIC,
         targets=(), target_sites=(), qv=None, q_stride=0, nq=0, L=1, scores=None, seed=0, step=0, step_dev=None,
         fmask_site2=0, fmask_split=0) -> None:
    """C[M,N] = epilogue(op(A) op(B)) — tcgen05/TMA GEMM.

    A is stored [M,K] (a_mn=False) or [K,M]; B is stored [N,K] (b_mn=False, the nn.Linear weight
    layout) or [K,N].  bf16 tensors run kind::f16 MMAs, fp32 tensors kind::tf32 (K-major only).
    """
    assert A.is_cuda and B.is_cuda and A.dtype == B.dtype and A.dtype in (torch.bfloat16, torch.float32)
    d = GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.a_mn, d.b_mn = int(a_mn), int(b_mn)
    d.tf32 = int(A.dtype == torch.float32)
    d.k_splits, d.block_n, d.max_ctas = k_splits, block_n, max_ctas
    d.A, d.lda, d.B, d.ldb = ptr(A), _ld(A), ptr(B), _ld(B)
    d.epi_kind, d.act = epi_kind, act
    d.bias = ptr(bias)
    d.gate, d.ld_gate, d.gate_scale = ptr(gate), (_ld(gate) if gate is not None else 0), gate_scale
    d.drop_p, d.drop_site, d.fmask_site = drop_p, drop_site, fmask_site
    d.fmask_site2, d.fmask_split = fmask_site2, fmask_split
    d.out_f32, d.ld_f32, d.f32_mode = ptr(out_f32), (_ld(out_f32) if out_f32 is not None else 0), f32_mode
    d.out_bf16, d.ld_bf16, d.bf16_mode = ptr(out_bf16), (_ld(out_bf16) if out_bf16 is not None else 0), bf16_mode
    d.n_tgt = len(targets)
    for i, (t, s) in enumerate(zip(targets, target_sites)):
        d.tgt[i] = ptr(t)
        d.tgt_site[i] = s
        if d.ld_bf16 == 0:
            d.ld_bf16 = _ld(t)
    d.qv, d.q_stride, d.nq, d.L, d.scores = ptr(qv), q_stride, nq, L, ptr(scores)
    d.seed, d.step, d.step_dev = seed, step, ptr(step_dev)
    check(_lib.lib().sdumc_gemm(C.byref(d), current_stream()), "sdumc_gemm")


def frame_mask(seed: int, step: int, site: int, rows: int, cols: int, device="cuda") -> torch.Tensor:
    out = torch.empty(rows, cols, dtype=torch.float32, device=device)
    check(_lib.lib().sdumc_frame_mask(seed, step, site, rows, cols, ptr(out), current_stream()), "sdumc_frame_mask")
    return out


def elem_mask(seed: int, step: int, site: int, n: int, p: float, device="cuda") -> torch.Tensor:
    out = torch.empty(n, dtype=torch.float32, device=device)
    check(_lib.lib().sdumc_elem_mask(seed, step, site, n, p, ptr(out), current_stream()), "sdumc_elem_mask")
    return out


def cast_bf16(src: torch.Tensor, dst: torch.Tensor) -> None:
    assert src.dtype == torch.float32 and dst.dtype == torch.bfloat16 and src.numel() == dst.numel()
    assert src.is_contiguous() and dst.is_contiguous()
    check(_lib.lib().sdumc_cast_bf16(ptr(src), ptr(dst), src.numel(), current_stream()), "sdumc_cast_bf16")


def collate_pad(packed: torch.Tensor, row_offset: torch.Tensor, idx: torch.Tensor, Lpad: int, out: torch.Tensor,
                out_off: torch.Tensor = None) -> None:
    """out[b, Lpad, D] = right-zero-padded utterances idx of the packed device store (read_data.py:223-248); with
    out_off ([b+1] int32, device) the output is PACKED: only the valid rows, utterance i at rows out_off[i].. of out."""
    assert packed.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and packed.is_contiguous() and out.is_contiguous()
    assert row_offset.dtype == torch.int64 and idx.dtype == torch.int32
    b, D = idx.numel(), packed.shape[1]
    assert out_off is not None or out.numel() == b * Lpad * D
    assert out_off is None or (out_off.dtype == torch.int32 and out_off.numel() == b + 1)
    check(_lib.lib().sdumc_collate_pad(ptr(packed), ptr(row_offset), ptr(idx), b, Lpad, D, ptr(out), ptr(out_off),
                                       current_stream()), "sdumc_collate_pad")


def colsum_bf16(X: torch.Tensor, out: torch.Tensor) -> None:
    """out[cols] += column sums of the bf16 matrix X [rows,cols]."""
    assert X.dtype == torch.bfloat16 and out.numel() == X.shape[1] and out.dtype == torch.float32
    check(_lib.lib().sdumc_colsum_bf16(ptr(X), _ld(X), X.shape[0], X.shape[1], ptr(out), current_stream()),
          "sdumc_colsum_bf16")


def pool_fwd(X, S, *, B, L, nq, O_pre, out, out_stride_b, out_bf16=None, drop_p=0.0, site=0, seed=0, step=0,
             step_dev=None, alpha=0.3, Kt=None, Qp=None, qp_stride_b=0, row_off=None, Hpad=None, Kpad=None) -> None:
    """softmax over frames + pooling; with Kt/Qp the scores are computed in the kernel (S is output only)."""
    a = STRUCTS["sdumc_pool_fwd_args"]()
    a.X, a.ldx, a.S = ptr(X), _ld(X), ptr(S)
    a.Kt, a.ldk = ptr(Kt), (_ld(Kt) if Kt is not None else 0)
    a.Qp, a.qp_stride_b = ptr(Qp), qp_stride_b
    a.B, a.L, a.nq, a.alpha = B, L, nq, alpha
    a.O_pre, a.out, a.out_stride_b, a.out_bf16 = ptr(O_pre), ptr(out), out_stride_b, ptr(out_bf16)
    a.drop_p, a.site, a.key = drop_p, site, dropkey(seed, step, step_dev)
    a.G = X.shape[1]
    a.row_off, a.Hpad, a.Kpad = ptr(row_off), ptr(Hpad), ptr(Kpad)
    call("sdumc_pool_fwd", a)


def attn_bwd(X, Kt, P, dOut, *, dout_stride_b, O_pre, Qp, qp_stride_b, B, L, nq, out_drop_p, out_site, dZ, dH,
             dh_mode, fmask_site, dQp, dqp_stride_b, db, seed=0, step=0, step_dev=None, alpha=0.3, max_ctas=0,
             split_b=0, out_site2=0, fmask_site2=0) -> None:
    a = STRUCTS["sdumc_attn_bwd_args"]()
    a.X, a.ldx, a.Kt, a.ldk, a.P = ptr(X), _ld(X), ptr(Kt), _ld(Kt), ptr(P)
    a.dOut, a.dout_stride_b, a.O_pre = ptr(dOut), dout_stride_b, ptr(O_pre)
    a.Qp, a.qp_stride_b = ptr(Qp), qp_stride_b
    a.B, a.L, a.nq, a.alpha = B, L, nq, alpha
    a.out_drop_p, a.out_site = out_drop_p, out_site
    a.dZ, a.lddz, a.dH, a.lddh, a.dh_mode, a.fmask_site = ptr(dZ), _ld(dZ), ptr(dH), _ld(dH), dh_mode, fmask_site
    a.split_b, a.out_site2, a.fmask_site2 = split_b, out_site2, fmask_site2
    a.dQp, a.dqp_stride_b, a.db = ptr(dQp), dqp_stride_b, ptr(db)
    a.key = dropkey(seed, step, step_dev)
    a.G = X.shape[1]
    a.max_ctas = max_ctas
    call("sdumc_attn_bwd", a)


def act_bwd(dY, dZ, *, rows, cols, Y=None, scale=1.0, db=None, dY2=None) -> None:
    a = STRUCTS["sdumc_act_bwd_args"]()
    a.dY, a.ld_dy = ptr(dY), _ld(dY)
    a.dY2, a.ld_dy2 = ptr(dY2), (_ld(dY2) if dY2 is not None else 0)
    a.Y, a.ld_y = ptr(Y), (_ld(Y) if Y is not None else 0)
    a.scale, a.rows, a.cols = scale, rows, cols
    a.dZ, a.ld_dz, a.db = ptr(dZ), _ld(dZ), ptr(db)
    call("sdumc_act_bwd", a)


def gate_fwd(a2, Wg, bg, h, *, R, g, qin) -> None:
    a = STRUCTS["sdumc_gate_fwd_args"]()
    a.a2, a.ld_a2, a.Wg, a.bg = ptr(a2), _ld(a2), ptr(Wg), ptr(bg)
    a.h, a.ld_h, a.R, a.G = ptr(h), _ld(h), R, a2.shape[1]
    a.g, a.qin, a.qin_stride = ptr(g), ptr(qin), qin.stride(0)
    call("sdumc_gate_fwd", a)


def gate_bwd(dqin, dg_extra, g, h, a2, Wg, *, R, dh, da2, dWg, dbg) -> None:
    a = STRUCTS["sdumc_gate_bwd_args"]()
    a.dqin, a.dqin_stride, a.dg_extra, a.g = ptr(dqin), dqin.stride(0), ptr(dg_extra), ptr(g)
    a.h, a.ld_h, a.a2, a.ld_a2, a.Wg, a.R, a.G = ptr(h), _ld(h), ptr(a2), _ld(a2), ptr(Wg), R, a2.shape[1]
    a.dh, a.ld_dh, a.da2, a.ld_da2 = ptr(dh), _ld(dh), ptr(da2), _ld(da2)
    a.dWg, a.dbg = ptr(dWg), ptr(dbg)
    call("sdumc_gate_bwd", a)


def weight_fwd(c, g, *, R, W) -> None:
    a = STRUCTS["sdumc_weight_fwd_args"]()
    for m in range(3):
        a.c[m] = ptr(c[m])
    a.g, a.R, a.W = ptr(g), R, ptr(W)
    call("sdumc_weight_fwd", a)


def weight_bwd(dW, c, g, *, R, dc, dg, dc_extra=(None, None, None)) -> None:
    a = STRUCTS["sdumc_weight_bwd_args"]()
    a.dW = ptr(dW)
    for m in range(3):
        a.c[m] = ptr(c[m])
        a.dc[m] = ptr(dc[m])
        a.dc_extra[m] = ptr(dc_extra[m])
    a.g, a.R, a.dg = ptr(g), R, ptr(dg)
    call("sdumc_weight_bwd", a)


def final_fwd(x2, Wr, br, W, Wv, bv, *, R, r, f, vals) -> None:
    a = STRUCTS["sdumc_final_fwd_args"]()
    a.x2, a.ld_x2, a.Wr, a.br, a.W, a.Wv, a.bv = ptr(x2), _ld(x2), ptr(Wr), ptr(br), ptr(W), ptr(Wv), ptr(bv)
    a.R, a.r, a.f, a.vals = R, ptr(r), ptr(f), ptr(vals)
    call("sdumc_final_fwd", a)


def final_bwd(dvals, df_ext, x2, Wr, W, r, f, Wv, *, R, dWc, dx2, dWr, dbr, dWv, dbv) -> None:
    a = STRUCTS["sdumc_final_bwd_args"]()
    a.dvals, a.df_ext, a.x2, a.ld_x2 = ptr(dvals), ptr(df_ext), ptr(x2), _ld(x2)
    a.Wr, a.W, a.r, a.f, a.Wv, a.R = ptr(Wr), ptr(W), ptr(r), ptr(f), ptr(Wv), R
    a.dWc, a.dx2, a.ld_dx2 = ptr(dWc), ptr(dx2), _ld(dx2)
    a.dWr, a.dbr, a.dWv, a.dbv = ptr(dWr), ptr(dbr), ptr(dWv), ptr(dbv)
    call("sdumc_final_bwd", a)


def _loss_in(a, v0, v1, y, th0, th1, ct0, ct1, f0, f1, B):
    a.v0, a.v1, a.y = ptr(v0), ptr(v1), ptr(y)
    a.th0, a.th1, a.ct0, a.ct1, a.f0, a.f1, a.B = ptr(th0), ptr(th1), ptr(ct0), ptr(ct1), ptr(f0), ptr(f1), B
    a.G = th0.shape[-1]


def loss_sums(v0, v1, y, th0, th1, ct0, ct1, f0, f1, *, B, sums) -> None:
    a = STRUCTS["sdumc_loss_sums_args"]()
    _loss_in(a, v0, v1, y, th0, th1, ct0, ct1, f0, f1, B)
    a.sums = ptr(sums)
    call("sdumc_loss_sums", a)


def loss_finish(v0, v1, y, th0, th1, ct0, ct1, f0, f1, *, B, sums, rnc, B_global, w, terms, d_v0, d_v1, d_th1,
                d_ct1, d_f0, d_f1) -> None:
    a = STRUCTS["sdumc_loss_finish_args"]()
    _loss_in(getattr(a, "in"), v0, v1, y, th0, th1, ct0, ct1, f0, f1, B)
    a.sums, a.rnc, a.B_global = ptr(sums), ptr(rnc), B_global
    for i in range(6):
        a.w[i] = float(w[i])
    a.terms, a.d_v0, a.d_v1, a.d_th1, a.d_ct1, a.d_f0, a.d_f1 = (ptr(terms), ptr(d_v0), ptr(d_v1), ptr(d_th1),
                                                                 ptr(d_ct1), ptr(d_f0), ptr(d_f1))
    call("sdumc_loss_finish", a)


def sqdiff_sum(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor) -> None:
    check(_lib.lib().sdumc_sqdiff_sum(ptr(a), ptr(b), a.numel(), ptr(out), current_stream()), "sdumc_sqdiff_sum")


def sqdiff_grad(a, b, coef, da, db=None) -> None:
    check(_lib.lib().sdumc_sqdiff_grad(ptr(a), ptr(b), a.numel(), ptr(coef), ptr(da), ptr(db), current_stream()),
          "sdumc_sqdiff_grad")


def rnc_workspace_bytes(n: int, D: int, rows: int = 0) -> int:
    """Workspace of sdumc_rnc for calls with at most `rows` anchor rows (0: all n)."""
    if rows:
        return int(_lib.lib().sdumc_rnc_workspace_bytes_rows(n, D, rows))
    return int(_lib.lib().sdumc_rnc_workspace_bytes(n, D))


RNC_ALL, RNC_LABELS, RNC_FEATURES, RNC_SORT = 0, 1, 2, 3


def rnc(feats, labels, *, loss=None, dfeats=None, row_begin=0, row_end=None, temperature=2.0, grad_scale=1.0,
        workspace=None, reuse_sort=False, phase=RNC_ALL, D=None) -> None:
    """Rank-N-Contrast over feats [n,D] (rows = view-0 samples then view-1 samples), labels [n].
    phase = RNC_LABELS runs only what depends on the labels (sort, bucket index, boundaries; feats may be None),
    RNC_FEATURES the rest on the same workspace - a trainer hides the first under its forward pass."""
    n = labels.numel()
    D = feats.shape[1] if feats is not None else (D or 64)
    a = STRUCTS["sdumc_rnc_args"]()
    a.feats, a.labels, a.n, a.D = ptr(feats), ptr(labels), n, D
    a.row_begin, a.row_end = row_begin, (n if row_end is None else row_end)
    a.temperature, a.loss, a.dfeats, a.grad_scale = temperature, ptr(loss), ptr(dfeats), grad_scale
    if workspace is None:
        workspace = torch.empty(rnc_workspace_bytes(n, D), dtype=torch.uint8, device=labels.device)
    a.workspace, a.workspace_bytes = ptr(workspace), workspace.numel()
    a.reuse_sort = 1 if reuse_sort else 0
    a.phase = phase
    n_lab = (0 if reuse_sort else 2) + 1
    n_feat = 4 if dfeats is not None else 2
    call("sdumc_rnc", a, launches={RNC_ALL: n_lab + n_feat, RNC_LABELS: n_lab, RNC_FEATURES: n_feat, RNC_SORT: 2}[phase])


def adam(p, g, m, v, *, lr, step, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, grad_scale=1.0,
         p_bf16=None, n=None, step_dev=None, lr_dev=None) -> None:
    a = STRUCTS["sdumc_adam_args"]()
    a.p, a.g, a.m, a.v, a.p_bf16 = ptr(p), ptr(g), ptr(m), ptr(v), ptr(p_bf16)
    a.n = p.numel() if n is None else n
    a.lr, a.beta1, a.beta2, a.eps, a.weight_decay, a.grad_scale, a.step = lr, beta1, beta2, eps, weight_decay, \
        grad_scale, step
    a.step_dev, a.lr_dev = ptr(step_dev), ptr(lr_dev)
    call("sdumc_adam", a)
