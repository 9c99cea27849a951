"""Drop-in nn.Module for the reference model class.

    WengnetMOSEIMultViewsTextMissing(args, output_dim1=1, output_dim2=1, layers='256,128', dropout=0.3)
    forward([audio[B,La,Da], text[B,Lt,Dt], video[B,Lv,Dv], missing_flag]) ->
        (vals_out[B,1], [fused[B,128], feat4rnc[B,64], text_hidden[B,256], cross_text[B,7,128]])

mirrors toolkit/models/wengnet_mosei_mult_views_text_missing.py:186-370 of the reference: same
constructor and forward signature, same state_dict keys (including the parameters the reference
constructs but never uses), train()/eval() toggling every dropout, outputs connected to autograd.
The computation runs on the hand-written sm_100a kernels (sdumc_b200.engine); there is no CPU path.
"""
from __future__ import annotations

import math
from typing import List

import torch
import torch.nn as nn

from . import _lib
from .engine import Cfg, Engine, Weights
from .params import ParamLayout, is_live


class _Node(nn.Module):
    """Plain container so parameter names nest exactly like the reference's sub-modules."""


class _UMCFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, need_grad, audio, text, video, *params):
        W = module._weights()
        B = audio.shape[0]
        cfg = Cfg(B=B, n_pass=1, frames={"a": audio.shape[1], "t0": text.shape[1], "v": video.shape[1]},
                  dropout=module.training, need_grad=need_grad, seed=module.dropout_seed, step=module._next_step())
        st = module._engine.forward(W, {"a": audio, "t0": text, "v": video}, cfg)
        vals, fused, rnc, th, ct = Engine.outputs(st)
        ctx.module, ctx.st = module, (st if need_grad else None)
        module._last_state = st if module.keep_last_state else None   # test hook (tests/parity_common.py)
        ctx.set_materialize_grads(False)
        return vals[0].clone(), fused[0].clone(), rnc[0].clone(), th[0].contiguous(), ct[0].clone()

    @staticmethod
    def backward(ctx, d_vals, d_fused, d_rnc, d_th, d_ct):
        module, st = ctx.module, ctx.st
        if st is None:
            raise RuntimeError("sdumc_b200: backward called on a forward pass that did not record gradients")
        layout = module.layout
        W = module._weights()
        W.grads = torch.zeros(layout.n_total, dtype=torch.float32, device=W.master.device)
        c = lambda g: None if g is None else g.contiguous().float()  # noqa: E731
        module._engine.backward(W, st, d_vals=c(d_vals), d_fused=c(d_fused), d_rnc=c(d_rnc), d_th=c(d_th), d_ct=c(d_ct))
        grads = [layout.view(W.grads, n) if is_live(n) else None for n in layout.names]
        ctx.st = None
        return (None, None, None, None, None, *grads)


class WengnetMOSEIMultViewsTextMissing(nn.Module):
    def __init__(self, args, output_dim1=1, output_dim2=1, layers='256,128', dropout=0.3):
        super().__init__()
        if layers != '256,128' or output_dim1 != 1 or output_dim2 != 1 or abs(dropout - 0.3) > 1e-12:
            # the reference CLI never reaches these arguments (models/__init__.py:67 passes args only)
            raise NotImplementedError("sdumc_b200 implements the configuration the reference trains: "
                                      "layers='256,128', output dims 1, dropout 0.3")
        # `general_dim` is hard-coded to 256 in the reference (:191); args.general_dim is an additive knob
        self.layout = ParamLayout(args.input_dims, int(getattr(args, "general_dim", 256)))
        self.dropout_seed = int(getattr(args, "seed", 100))
        self._step = 0
        self.keep_last_state = False
        self._last_state = None
        self._engine = None
        self._flat = None
        self._shadow = None
        for name, shape in self.layout.spec:
            parent = self
            *mods, leaf = name.split(".")
            for mname in mods:
                if not hasattr(parent, mname):
                    parent.add_module(mname, _Node())
                parent = getattr(parent, mname)
            parent.register_parameter(leaf, nn.Parameter(self._init_tensor(name, shape)))
        self._rebuild_flat()

    # ---- parameters -----------------------------------------------------------------------
    @staticmethod
    def _init_tensor(name: str, shape) -> torch.Tensor:
        """torch default initialisers of the reference's layers (nn.Linear: kaiming_uniform(a=sqrt 5) =
        U(+-1/sqrt(fan_in)); context vectors xavier_normal, reference :52; PReLU 0.25; LayerNorm 1/0)."""
        t = torch.empty(shape)
        if name == "prelu.weight":
            return t.fill_(0.25)
        if name == "layer_normali.weight":
            return t.fill_(1.0)
        if name == "layer_normali.bias":
            return t.zero_()
        if name.endswith("attention_context_vector"):
            return nn.init.xavier_normal_(t)
        if name.endswith(".weight"):
            bound = 1.0 / math.sqrt(shape[1])
            return t.uniform_(-bound, bound)
        return t  # bias: filled right after its weight (needs fan_in) in _rebuild_flat's first call

    def _named_param_dict(self):
        return dict(self.named_parameters())

    def _rebuild_flat(self):
        """(Re)creates the flat master buffer on the parameters' device and re-points every parameter at
        its slice, so optimizers that update parameters in place keep the kernels' view in sync."""
        params = self._named_param_dict()
        first = self._flat is None
        dev = next(iter(params.values())).device
        flat = torch.zeros(self.layout.n_total, dtype=torch.float32, device=dev)
        for name in self.layout.names:
            p = params[name]
            if first and name.endswith(".bias") and not name.startswith("layer_normali"):
                fan_in = params[name[:-5] + ".weight"].shape[1]
                bound = 1.0 / math.sqrt(fan_in)
                p.data.uniform_(-bound, bound)
            v = self.layout.view(flat, name)
            v.copy_(p.data.float())
            p.data = v
        self._flat = flat
        self._shadow = torch.empty(self.layout.n_total, dtype=torch.bfloat16, device=dev) if dev.type == "cuda" else None
        self._engine = Engine(self.layout, dev) if dev.type == "cuda" else None

    def _apply(self, fn, *a, **kw):
        super()._apply(fn, *a, **kw)
        self._rebuild_flat()
        return self

    def _flat_is_current(self) -> bool:
        base = self._flat.data_ptr()
        for name, p in self.named_parameters():
            if p.data_ptr() != base + 4 * self.layout.entries[name].offset or p.dtype != torch.float32:
                return False
        return True

    def _weights(self) -> Weights:
        if not self._flat_is_current():
            self._rebuild_flat()
        return Weights(self.layout, self._flat, self._shadow)

    def _next_step(self) -> int:
        self._step += 1
        return self._step

    # ---- forward --------------------------------------------------------------------------
    def forward(self, batch):
        audio, text, video = batch[0], batch[1], batch[2]   # batch[-1] = missing_flag: read, never used (:278)
        if not audio.is_cuda:
            raise _lib.SdumcError("sdumc_b200 runs on CUDA (sm_100a) only; there is no CPU fallback. "
                                  "Move the model and the batch to a B200.")
        _lib.lib()  # raises if the extension has not been built
        W = self._weights()
        if W.master.device != audio.device:
            raise RuntimeError(f"model is on {W.master.device}, batch on {audio.device}")
        W.refresh_shadow()
        params: List[torch.Tensor] = [self._named_param_dict()[n] for n in self.layout.names]
        # (inside autograd.Function.forward grad mode is always off, so decide here)
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        vals, fused, rnc, th, ct = _UMCFunction.apply(self, need_grad, audio, text, video, *params)
        return vals, [fused, rnc, th, ct]


class get_models(nn.Module):
    """toolkit.models.get_models (reference toolkit/models/__init__.py:29-70): wrapper with `.model`."""

    def __init__(self, args):
        super().__init__()
        args.dim = 1024  # reference :34
        name = getattr(args, "model", "wengnet_mosei_mult_views_text_missing")
        if name != "wengnet_mosei_mult_views_text_missing":
            raise KeyError(f"sdumc_b200 implements 'wengnet_mosei_mult_views_text_missing' only, got {name!r}")
        self.model = WengnetMOSEIMultViewsTextMissing(args)

    def forward(self, batch):
        return self.model(batch)
