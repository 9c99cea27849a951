"""Drop-in for the reference's EncoderProjectorConcat (feature_extraction/llm4wav/extract_wavlm_vicuna.py:162-185):
the dense front end of the feat4 ("text decoded from audio") producer - k consecutive WavLM frames are concatenated
and sent through Linear(k*encoder_dim, 2048) -> ReLU -> Linear(2048, llm_dim) before they enter the LLM.  Same
constructor, forward() and state_dict keys (linear1.*, linear2.*).  SURVEY.md §8f N4.

Both layers run on the library's tcgen05 GEMM (bf16 operands, fp32 accumulation, bias / ReLU in the epilogue); the
hidden activation stays bf16.  Inference only, like its use in the reference (parameters frozen, :197-198).  CUDA
(sm_100a) only - there is no CPU fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib, ops


class EncoderProjectorConcat(nn.Module):
    def __init__(self, encoder_projector_ds_rate, encoder_dim, llm_dim):
        super().__init__()
        self.k = encoder_projector_ds_rate
        self.encoder_dim = encoder_dim
        self.llm_dim = llm_dim
        self.linear1 = nn.Linear(self.encoder_dim * self.k, 2048)
        self.relu = nn.ReLU()
        self.linear2 = nn.Linear(2048, llm_dim)
        self._shadow = None          # (version key, bf16 weights)

    def _weights(self):
        key = (self.linear1.weight._version, self.linear2.weight._version, self.linear1.weight.data_ptr())
        if self._shadow is None or self._shadow[0] != key:
            w1 = torch.empty(self.linear1.weight.shape, dtype=torch.bfloat16, device=self.linear1.weight.device)
            w2 = torch.empty(self.linear2.weight.shape, dtype=torch.bfloat16, device=self.linear2.weight.device)
            ops.cast_bf16(self.linear1.weight.detach().float().contiguous().view(-1), w1.view(-1))
            ops.cast_bf16(self.linear2.weight.detach().float().contiguous().view(-1), w2.view(-1))
            self._shadow = (key, w1, w2)
        return self._shadow[1], self._shadow[2]

    @torch.no_grad()
    def forward(self, x):
        if not x.is_cuda:
            raise _lib.SdumcError("sdumc_b200.EncoderProjectorConcat runs on CUDA (sm_100a) only; there is no CPU fallback")
        _lib.lib()
        batch_size, seq_len, dim = x.size()
        if dim != self.encoder_dim or (dim * self.k) % 8 != 0:
            raise ValueError(f"expected [B, T, {self.encoder_dim}] with k*dim a multiple of 8, got {tuple(x.shape)}")
        seq_len -= seq_len % self.k                       # trailing frames that do not fill a group are discarded (:177-179)
        m = batch_size * (seq_len // self.k)
        if m == 0:
            return x.new_zeros(batch_size, 0, self.llm_dim, dtype=torch.float32)
        xg = x[:, :seq_len, :].contiguous().view(m, dim * self.k)
        if xg.dtype == torch.bfloat16:
            xb = xg
        else:
            xb = torch.empty(m, dim * self.k, dtype=torch.bfloat16, device=x.device)
            ops.cast_bf16(xg.float().view(-1), xb.view(-1))
        w1, w2 = self._weights()
        h = torch.empty(m, 2048, dtype=torch.bfloat16, device=x.device)
        ops.gemm(xb, w1, M=m, N=2048, K=dim * self.k, bias=self.linear1.bias.detach().float().contiguous(),
                 act=ops.ACT_RELU, out_bf16=h)
        y = torch.empty(m, self.llm_dim, dtype=torch.float32, device=x.device)
        ops.gemm(h, w2, M=m, N=self.llm_dim, K=2048, bias=self.linear2.bias.detach().float().contiguous(), out_f32=y)
        return y.view(batch_size, seq_len // self.k, self.llm_dim)
