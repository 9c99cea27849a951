"""Thin torch-tensor front ends of the C-ABI operators (device memory and streams only)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import (ACT_NONE, ACT_RELU, ACT_TANH, EPI_GENERIC, EPI_INPROJ, EPI_KEYPROJ, OUT_ADD, OUT_ATOMIC,
                   OUT_STORE, GemmDesc, check, current_stream, ptr)


def _ld(t: torch.Tensor) -> int:
    assert t.dim() == 2 and t.stride(1) == 1, "row-major 2-D tensor expected"
    return t.stride(0)


def gemm(A: torch.Tensor, B: torch.Tensor, *, M: int, N: int, K: int, a_mn=False, b_mn=False, k_splits=1,
         block_n=0, max_ctas=0, bias=None, act=ACT_NONE, gate=None, gate_scale=1.0, drop_p=0.0, drop_site=0,
         fmask_site=0, out_f32=None, f32_mode=OUT_STORE, out_bf16=None, bf16_mode=OUT_STORE, epi_kind=EPI_GENERIC,
         targets=(), target_sites=(), qv=None, q_stride=0, nq=0, L=1, scores=None, seed=0, step=0, dbg_lbo=0,
         dbg_sbo=0) -> None:
    """C[M,N] = epilogue(op(A) op(B)) on the current stream.

    A is stored [M,K] (a_mn=False) or [K,M]; B is stored [N,K] (b_mn=False, the nn.Linear weight
    layout) or [K,N].  bf16 tensors run kind::f16 MMAs, fp32 tensors kind::tf32.
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
    d.out_f32, d.ld_f32, d.f32_mode = ptr(out_f32), (_ld(out_f32) if out_f32 is not None else 0), f32_mode
    d.out_bf16, d.ld_bf16, d.bf16_mode = ptr(out_bf16), (_ld(out_bf16) if out_bf16 is not None else 0), bf16_mode
    d.n_tgt = len(targets)
    for i, (t, s) in enumerate(zip(targets, target_sites)):
        d.tgt[i] = ptr(t)
        d.tgt_site[i] = s
        if d.ld_bf16 == 0:
            d.ld_bf16 = _ld(t)
    d.qv, d.q_stride, d.nq, d.L, d.scores = ptr(qv), q_stride, nq, L, ptr(scores)
    d.seed, d.step = seed, step
    d.dbg_lbo, d.dbg_sbo = dbg_lbo, dbg_sbo
    check(_lib.lib().sdumc_gemm(C.byref(d), current_stream()), "sdumc_gemm")


def frame_mask(seed: int, step: int, site: int, rows: int, cols: int, device="cuda") -> torch.Tensor:
    out = torch.empty(rows, cols, dtype=torch.float32, device=device)
    check(_lib.lib().sdumc_frame_mask(seed, step, site, rows, cols, ptr(out), current_stream()), "sdumc_frame_mask")
    return out


def elem_mask(seed: int, step: int, site: int, n: int, p: float, device="cuda") -> torch.Tensor:
    out = torch.empty(n, dtype=torch.float32, device=device)
    check(_lib.lib().sdumc_elem_mask(seed, step, site, n, p, ptr(out), current_stream()), "sdumc_elem_mask")
    return out
