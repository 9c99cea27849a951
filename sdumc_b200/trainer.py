"""Fused self-distillation train step and two-pass scoring on the sm_100a kernels.

Replaces the hot loop of train_or_eval_model() (reference main_frame_val_text_missing.py:89-158;
inference variant main_frame_val_text_missing_inference.py:158-175): both passes (full / text-missing)
run as ONE batch of 2B utterance rows, the audio and video in-projections are shared by the two passes,
the 6-term loss (:148), its backward, and Adam (:317, :150) follow on the same stream, and for a
single GPU the whole step is captured into one CUDA graph (dropout step counter, Adam step and learning
rate live in device memory so replays differ).

Data parallel (one process per GPU, torch.distributed / NCCL): the batch is sharded; the only
exchanges are the ones the algorithm needs to stay identical to a single-process batch —
  all_reduce of the 5 sums of squares (RMSE is the sqrt of the GLOBAL mean, loss.py:50),
  all_gather of the RnC features / labels (RnC couples all 2B x 2B pairs, loss.py:282-313) and an
  all_reduce of their gradient, and the gradient all_reduce over the contiguous live-parameter range.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib, ops
from .engine import Cfg, Engine, Weights
from .params import ParamLayout

DEFAULT_LOSS_W = (0.5, 0.5, 0.1, 0.7, 0.1, 0.8)   # main_frame_val_text_missing.py:234-239
LOSS_TERMS = ("mse_full", "mse_missing", "rmse_text_hidden", "rmse_cross_text", "rmse_fused", "rnc")


def lr_lambda(epoch: int, warm_up_epochs: int = 5, gamma: float = 0.9, stepsize: int = 10) -> float:
    """LambdaLR schedule of the reference (main…:318-320), stepped once per epoch."""
    return (epoch + 1) / warm_up_epochs if epoch < warm_up_epochs else gamma ** ((epoch + 1 - warm_up_epochs) // stepsize)


def default_state_dict(layout: ParamLayout) -> Dict[str, torch.Tensor]:
    """Fresh parameters drawn from torch's global CPU generator with the initialisers of the reference's layers
    (model.WengnetMOSEIMultViewsTextMissing._init_tensor): what `get_models(args)` would hold after construction."""
    from .model import WengnetMOSEIMultViewsTextMissing
    sd = {}
    for name, shape in layout.spec:
        sd[name] = WengnetMOSEIMultViewsTextMissing._init_tensor(name, shape)
    for name, shape in layout.spec:   # biases need fan_in of their weight
        if name.endswith(".bias") and not name.startswith("layer_normali"):
            bound = 1.0 / (sd[name[:-5] + ".weight"].shape[1] ** 0.5)
            sd[name] = torch.empty(shape).uniform_(-bound, bound)
    return sd


class Trainer:
    """Owns the flat parameter / gradient / Adam buffers and static batch buffers of one rank."""

    def __init__(self, input_dims: Sequence[int], B: int, frames: Sequence[int], device, *, lr=1e-4,
                 weight_decay=1e-5, loss_w: Sequence[float] = DEFAULT_LOSS_W, seed: int = 100,
                 state_dict: Optional[Dict[str, torch.Tensor]] = None, process_group=None, use_graph: bool = True,
                 general_dim: int = 256):
        _lib.lib()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.SdumcError("sdumc_b200.Trainer needs a CUDA (sm_100a) device; there is no CPU fallback")
        self.layout = ParamLayout(input_dims, general_dim)   # general_dim: additive knob, 256 = the reference (:191)
        self.dims = tuple(int(d) for d in input_dims)          # (Da, Dt, Dv[, D4 == Dt])
        self.B = int(B)
        self.frames = dict(zip(("a", "t0", "v", "t1"), (int(f) for f in frames)))   # (La, Lt, Lv, L4)
        self.lr, self.weight_decay, self.loss_w, self.seed = float(lr), float(weight_decay), tuple(loss_w), int(seed)
        self.pg = process_group
        if process_group is not None:
            import torch.distributed as dist
            self.world, self.rank = dist.get_world_size(process_group), dist.get_rank(process_group)
        else:
            self.world, self.rank = 1, 0
        # dropout counters index LOCAL rows, so data-parallel ranks fold their rank into the Philox key: the masks of
        # the W shards are independent draws, like the W*B rows of a single-process batch (same seed on every rank
        # would repeat the same mask pattern in every shard)
        self.drop_seed = self.seed if self.world == 1 else (self.seed + 0x9E3779B97F4A7C15 * self.rank) & (2 ** 64 - 1)
        # the data-parallel step (NCCL all_gather / all_reduce inside) is captured too: NCCL collectives are graph
        # nodes like any kernel; SDUMC_DP_GRAPH=0 falls back to eager launches for debugging
        import os
        self.use_graph = use_graph and (self.world == 1 or os.environ.get("SDUMC_DP_GRAPH", "1") != "0")
        dev, L = self.device, self.layout
        z = lambda n, dt=torch.float32: torch.zeros(n, dtype=dt, device=dev)  # noqa: E731
        self.master, self.grads, self.m, self.v = z(L.n_total), z(L.n_total), z(L.n_live), z(L.n_live)
        self.shadow = z(L.n_total, torch.bfloat16)
        self.W = Weights(L, self.master, self.shadow, self.grads)
        self.engine = Engine(L, dev)
        if state_dict is None:
            state_dict = default_state_dict(L)
        self.load_state_dict(state_dict)
        Da, Dt, Dv = self.dims[:3]
        fr = self.frames
        # static input buffers sized for the largest batch (B utterances x `frames`); a batch padded to fewer
        # frames (the reference pads to the batch maximum, read_data.py:223-248) uses a prefix of each buffer
        self.in_dims = {"a": Da, "t0": Dt, "v": Dv, "t1": Dt}
        self.in_flat = {k: torch.zeros(B * fr[k] * self.in_dims[k], dtype=torch.bfloat16, device=dev) for k in fr}
        self.cur_frames = dict(fr)
        self.inputs = {k: self.in_flat[k].view(B, fr[k], self.in_dims[k]) for k in fr}
        self.labels = z(B)
        # device-resident scalars (CUDA-graph replay)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)      # optimisation step, 1-based after the bump
        self.lr_dev = torch.full((1,), self.lr, dtype=torch.float32, device=dev)
        self.terms = z(8)
        R = 2 * B
        Gd = self.layout.G
        # accumulators that every step starts from zero live side by side: ONE fill per step instead of five
        self._zero_block = z(16 + R * 64 + R * Gd + R * 896)
        zb, o = self._zero_block, 16
        self.sums, self.rnc_val = zb[:8], zb[8:9]
        self.d_rnc = zb[o:o + R * 64].view(R, 64)
        o += R * 64
        self.d_th = zb[o:o + R * Gd].view(R, Gd)
        o += R * Gd
        self.d_ct = zb[o:o + R * 896].view(R, 896)
        self.d_vals, self.d_f = z(R), z(R * 128).view(R, 128)
        self.y2 = z(R)
        n_g = 2 * B * self.world
        self.rnc_ws = torch.empty(ops.rnc_workspace_bytes(n_g, 64, rows=2 * B), dtype=torch.uint8, device=dev)
        self.y_glob = z(n_g)                  # labels of the global Rank-N-Contrast problem (rank-major rows)
        self._label_stream = None
        self._labels_ready = False
        self.cur_B = self.B
        self.train_dropout = True      # tests switch dropout off to compare against a deterministic reference
        self._comm_stream = None
        # first flat-buffer offset after the in-projection and FRA2UTT_new parameters (the reference constructor
        # registers them first, :193-226): gradients from there on are final before the last third of the backward
        late = [e for e in L.entries.values() if e.live and e.name.startswith(("frame_dim_reshape_", "fra2utt_"))]
        self._chain_begin = max(e.offset + (e.numel + 63) // 64 * 64 for e in late)
        assert all(e.offset >= self._chain_begin for e in L.entries.values()
                   if e.live and not e.name.startswith(("frame_dim_reshape_", "fra2utt_")))
        self.n_steps = 0
        self.n_replays = 0             # steps that were CUDA-graph replays (the CLI tests assert the fast path is taken)
        # captured steps, keyed by the batch shape (utterances, frames per stream): a graph replays a fixed shape, so
        # every shape that recurs (the full-capacity batch; the fixed trailing batch of an epoch; length buckets) gets
        # its own graph the second time it is seen.  Least-recently-used graphs are dropped beyond `max_graphs`
        # (each holds its own activation pool).
        self.max_graphs = 4
        self._graphs: Dict[tuple, dict] = {}
        self._score_graphs: Dict[tuple, dict] = {}
        self._seen: Dict[tuple, int] = {}
        self._outputs = None
        self._score_calls = 0

    # ---- parameters ------------------------------------------------------------------------
    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = False):
        """Loads parameters by name.  Keys may carry the `model.` prefix of get_models() (toolkit/models/__init__.py)
        and / or DataParallel's `module.` (stripped like main_frame_val_text_missing_inference.py:341).  Returns
        (missing, unexpected) like nn.Module.load_state_dict; raises if NOTHING matched (a silently random model is
        never what the caller wants) or, with strict=True, if anything is missing / unexpected."""
        norm = {}
        for k, v in sd.items():
            k2 = k.replace("module.", "")
            if k2.startswith("model."):
                k2 = k2[len("model."):]
            norm[k2] = v
        missing = [n for n in self.layout.names if n not in norm]
        unexpected = [k for k in norm if k not in self.layout.entries]
        if norm and len(missing) == len(self.layout.names):
            raise _lib.SdumcError("load_state_dict: no key of the checkpoint matches a model parameter "
                                  f"(first keys: {list(sd)[:3]})")
        if strict and (missing or unexpected):
            raise _lib.SdumcError(f"load_state_dict(strict): missing {missing[:5]}, unexpected {unexpected[:5]}")
        for name in self.layout.names:
            if name in norm:
                src = norm[name]
                if tuple(src.shape) != tuple(self.layout.entries[name].shape):
                    raise _lib.SdumcError(f"load_state_dict: {name} has shape {tuple(src.shape)}, expected "
                                          f"{tuple(self.layout.entries[name].shape)}")
                self.layout.view(self.master, name).copy_(src.to(self.device, torch.float32))
        self.W.refresh_shadow()
        return missing, unexpected

    def state_dict(self, prefix: str = "") -> Dict[str, torch.Tensor]:
        """Parameters by reference name; prefix='model.' gives the keys of get_models(args).state_dict()."""
        return {prefix + name: self.layout.view(self.master, name).clone() for name in self.layout.names}

    def optimizer_state_dict(self) -> dict:
        """torch.optim.Adam.state_dict() layout over model.parameters() in registration order (the reference's
        optimizer, main_frame_val_text_missing.py:317): parameters that never receive a gradient have no state."""
        from .params import is_live
        step = float(int(self.step_dev.item()))
        state = {}
        for i, name in enumerate(self.layout.names):
            if is_live(name) and step > 0:
                state[i] = {"step": torch.tensor(step), "exp_avg": self.layout.view(self.m, name).clone().cpu(),
                            "exp_avg_sq": self.layout.view(self.v, name).clone().cpu()}
        group = {"lr": self.lr, "betas": (0.9, 0.999), "eps": 1e-8, "weight_decay": self.weight_decay,
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                 "fused": None, "params": list(range(len(self.layout.names)))}
        return {"state": state, "param_groups": [group]}

    def load_optimizer_state_dict(self, osd: dict):
        """Restores Adam's moments and step count (so a resumed run continues the bias correction and the dropout
        stream where the saved one stopped)."""
        from .params import is_live
        step = 0
        self.m.zero_()
        self.v.zero_()
        for i, name in enumerate(self.layout.names):
            st_ = osd["state"].get(i)
            if st_ is None or not is_live(name):
                continue
            self.layout.view(self.m, name).copy_(st_["exp_avg"].to(self.device, torch.float32))
            self.layout.view(self.v, name).copy_(st_["exp_avg_sq"].to(self.device, torch.float32))
            step = max(step, int(float(st_["step"])))
        self.step_dev.fill_(step)
        if osd.get("param_groups"):
            g = osd["param_groups"][0]
            self.weight_decay = float(g.get("weight_decay", self.weight_decay))
            self.set_lr(float(g.get("lr", self.lr)))

    def checkpoint(self, epoch: int) -> dict:
        """The dict the reference saves (commented out at main_frame_val_text_missing.py:375):
        {'epoch', 'state_dict' (keys of get_models().state_dict(), i.e. `model.`-prefixed), 'optimizer'}."""
        return {"epoch": int(epoch), "state_dict": {k: v.cpu() for k, v in self.state_dict("model.").items()},
                "optimizer": self.optimizer_state_dict()}

    def load_checkpoint(self, ck: dict, strict: bool = False):
        out = self.load_state_dict(ck["state_dict"], strict=strict)
        if "optimizer" in ck:
            self.load_optimizer_state_dict(ck["optimizer"])
        return out

    def set_lr(self, lr: float):
        self.lr = float(lr)
        self.lr_dev.fill_(self.lr)

    # ---- batch -----------------------------------------------------------------------------
    def load_batch(self, audio, text, video, feat4, vals):
        """Copies a batch into the static device buffers (H2D when given pinned host tensors).  bf16 tensors
        are copied as they are; fp32 tensors are converted on the device."""
        b = audio.shape[0]
        if b > self.B:
            raise ValueError(f"batch of {b} exceeds the trainer's capacity {self.B}")
        self.cur_B = b                      # a batch smaller than B (end of an epoch) runs eagerly
        for key, src in (("a", audio), ("t0", text), ("v", video), ("t1", feat4)):
            L, D = int(src.shape[1]), int(src.shape[2])
            if src.shape[0] != b or D != self.in_dims[key] or L > self.frames[key] or L < 1:
                raise ValueError(f"stream {key}: got {tuple(src.shape)}, capacity [{self.B},{self.frames[key]},"
                                 f"{self.in_dims[key]}]")
            self.cur_frames[key] = L
            dst = self.in_flat[key][:b * L * D].view(b, L, D)
            self.inputs[key] = dst
            if src.dtype == torch.bfloat16:
                dst.copy_(src, non_blocking=True)
            else:
                tmp = src.to(self.device, non_blocking=True).float().contiguous()
                ops.cast_bf16(tmp.view(-1), dst.view(-1))
        self.labels[:b].copy_(vals.reshape(-1), non_blocking=True)

    def load_from_store(self, store, idx, labels: bool = True):
        """Builds the batch `idx` of a DeviceStore4F directly in the static input buffers with the collate kernel
        (gather + right-zero-pad to the batch maximum, read_data.py:223-248): no host->device copy of features.
        labels=False skips the label gather (scoring does not read them)."""
        b = len(idx)
        if b > self.B:
            raise ValueError(f"batch of {b} exceeds the trainer's capacity {self.B}")
        frames = store.batch_frames(idx)
        idx_dev = torch.from_numpy(np.asarray(idx, dtype=np.int32)).to(self.device, non_blocking=True)
        self.cur_B = b
        for key, s, L in zip(("a", "t0", "v", "t1"), ("audio", "text", "video", "feat4"), frames):
            D = self.in_dims[key]
            if store.packed[s].shape[1] != D or L > self.frames[key]:
                raise ValueError(f"stream {key}: store has D={store.packed[s].shape[1]}, batch needs {L} frames; "
                                 f"capacity [{self.B},{self.frames[key]},{D}]")
            self.cur_frames[key] = L
            dst = self.in_flat[key][:b * L * D]
            ops.collate_pad(store.packed[s], store.offsets[s], idx_dev, L, dst)
            self.inputs[key] = dst.view(b, L, D)
        if labels:
            self.labels[:b].copy_(store.vals[idx_dev.long()])

    def _check_stream(self, key, src, b):
        L, D = int(src.shape[1]), int(src.shape[2])
        if src.shape[0] != b or D != self.in_dims[key] or L > self.frames[key] or L < 1:
            raise ValueError(f"stream {key}: got {tuple(src.shape)}, capacity [{self.B},{self.frames[key]},"
                             f"{self.in_dims[key]}]")
        return L, D

    def stage_batch(self, audio, text, video, feat4, vals):
        """Double buffering for host-resident data: starts the H2D copy of the NEXT batch (pinned bf16 host
        tensors) into a staging set on a copy stream and returns immediately; commit_staged() makes it the
        current batch.  The copy overlaps the train step of the current batch - the role of the reference's
        DataLoader workers + pin_memory (main_frame_val_text_missing.py:42-60).  The host tensors must stay
        untouched until commit_staged()."""
        b = audio.shape[0]
        if b > self.B:
            raise ValueError(f"batch of {b} exceeds the trainer's capacity {self.B}")
        if getattr(self, "_stage", None) is None:
            self._stage = {k: torch.empty_like(v) for k, v in self.in_flat.items()}
            self._stage_labels = torch.empty_like(self.labels)
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._stage_ready = torch.cuda.Event()
            self._stage_consumed = None
        meta = {}
        with torch.cuda.stream(self._copy_stream):
            if self._stage_consumed is not None:          # the previous commit's device copies read the staging set
                self._copy_stream.wait_event(self._stage_consumed)
            for key, src in (("a", audio), ("t0", text), ("v", video), ("t1", feat4)):
                L, D = self._check_stream(key, src, b)
                if src.dtype != torch.bfloat16:
                    raise ValueError("stage_batch expects bf16 host tensors (use load_batch for fp32 sources)")
                meta[key] = L
                self._stage[key][:b * L * D].view(b, L, D).copy_(src, non_blocking=True)
            self._stage_labels[:b].copy_(vals.reshape(-1), non_blocking=True)
            self._stage_ready.record(self._copy_stream)
        self._staged = (b, meta)

    def commit_staged(self):
        """Makes the staged batch current: the compute stream waits for its H2D copy, then moves it into the
        static input buffers the (captured) step reads - a device copy, ~0.4 ms for the 1.2 GB S0 batch."""
        b, meta = self._staged
        cur = torch.cuda.current_stream()
        cur.wait_event(self._stage_ready)
        self.cur_B = b
        for key, L in meta.items():
            D = self.in_dims[key]
            self.cur_frames[key] = L
            dst = self.in_flat[key][:b * L * D]
            dst.copy_(self._stage[key][:b * L * D], non_blocking=True)
            self.inputs[key] = dst.view(b, L, D)
        self.labels[:b].copy_(self._stage_labels[:b], non_blocking=True)
        self._stage_consumed = torch.cuda.Event()
        self._stage_consumed.record(cur)
        self._staged = None

    # ---- the step --------------------------------------------------------------------------
    def _forward(self, dropout: bool, need_grad: bool, on_pools_done=None, on_frames_done=None):
        b = self.cur_B
        cfg = Cfg(B=b, n_pass=2, frames=dict(self.cur_frames), dropout=dropout, need_grad=need_grad,
                  seed=self.drop_seed, step=0, step_dev=self.step_dev)
        return self.engine.forward(self.W, self.inputs, cfg, on_pools_done=on_pools_done, on_frames_done=on_frames_done)

    def _loss_and_seeds(self, st):
        """6-term loss (:134-148) -> self.terms, gradient seeds -> self.d_*."""
        B, W_, rank = self.cur_B, self.world, self.rank
        vals, f, rnc, th, ct = Engine.outputs(st)
        th = th.contiguous()                                      # [2,B,256] out of the strided Q buffer
        y = self.labels[:B]
        R = 2 * B
        d_vals, d_f, d_rnc, d_th, d_ct, y2 = (self.d_vals[:R], self.d_f[:R], self.d_rnc[:R], self.d_th[:R],
                                              self.d_ct[:R], self.y2[:R])
        self._zero_block.zero_()                                  # sums, rnc_val, d_rnc, d_th, d_ct
        ops.loss_sums(vals[0], vals[1], y, th[0], th[1], ct[0], ct[1], f[0], f[1], B=B, sums=self.sums)
        w6 = self.loss_w[5]
        if not self._labels_ready:                                        # (a caller that skipped _step_body)
            self._rnc_labels_early()
        self._labels_ready = False
        torch.cuda.current_stream().wait_stream(self._label_stream)      # the label-only part (_rnc_labels_early)
        if W_ == 1:
            ops.rnc(st.t["rnc"], y2, loss=self.rnc_val, dfeats=d_rnc, grad_scale=w6, workspace=self.rnc_ws,
                    phase=ops.RNC_FEATURES)
        else:
            from . import dp

            def rnc_fn(feats_g, y_g, lo, hi, loss, dfeats):   # this rank's anchors: one contiguous row range
                ops.rnc(feats_g, self.y_glob[:feats_g.shape[0]], loss=loss, dfeats=dfeats, row_begin=lo, row_end=hi,
                        grad_scale=w6, workspace=self.rnc_ws, phase=ops.RNC_FEATURES)
            # one all_gather (features + labels) and one reduce_scatter (RnC gradient + loss + the sums of squares)
            loss_g, d_local = dp.rnc_global(rnc.contiguous(), y.contiguous(), self.pg, rnc_fn, extra=self.sums)
            self.rnc_val.copy_(loss_g)
            d_rnc.view(2, B, 64).copy_(d_local)
        ops.loss_finish(vals[0], vals[1], y, th[0], th[1], ct[0], ct[1], f[0], f[1], B=B, sums=self.sums,
                        rnc=self.rnc_val, B_global=B * W_, w=self.loss_w, terms=self.terms,
                        d_v0=d_vals[:B], d_v1=d_vals[B:], d_th1=d_th[B:], d_ct1=d_ct[B:], d_f0=d_f[:B], d_f1=d_f[B:])
        return d_vals, d_f, d_rnc, d_th, d_ct

    def _rnc_labels_early(self, part: int = 0):
        """Everything of the Rank-N-Contrast term that depends on the labels only (the global labels, their sort, the
        bucket index and the four boundaries of every (anchor, element) pair: most of the term's instructions) runs on a
        side stream under the latency-bound utterance chains of the forward pass, which leave most SMs idle (forked at
        the start of the step it only delays the in-projections: the big kernels want whole SMs):
        part 1 - label all-gather (data-parallel), sort, bucket index - is forked after the FRA2UTT_new blocks (chain A),
        part 2 - the boundaries - after the Cross_Attention blocks (chain B); part 0 = both.  _loss_and_seeds joins."""
        B, W_ = self.cur_B, self.world
        y = self.labels[:B]
        main = torch.cuda.current_stream()
        if self._label_stream is None:
            self._label_stream = torch.cuda.Stream(device=self.device)
        ls = self._label_stream
        ls.wait_stream(main)
        with torch.cuda.stream(ls):
            y2 = self.y2[:2 * B]
            y_g = y2 if W_ == 1 else self.y_glob[:W_ * 2 * B]
            lo, hi = (0, 2 * B)
            if W_ > 1:
                from . import dp
                lo, hi = dp.anchor_range(B, W_, self.rank)
            if part in (0, 1):
                y2[:B].copy_(y)
                y2[B:].copy_(y)
                if W_ > 1:
                    import torch.distributed as dist
                    dist.all_gather_into_tensor(y_g, y2, group=self.pg)       # rank-major rows, like dp.global_views
                ops.rnc(None, y_g, row_begin=lo, row_end=hi, workspace=self.rnc_ws, phase=ops.RNC_SORT)
            if part in (0, 2):
                ops.rnc(None, y_g, row_begin=lo, row_end=hi, workspace=self.rnc_ws, phase=ops.RNC_LABELS, reuse_sort=True)
                self._labels_ready = True

    def _step_body(self):
        self.step_dev.add_(1)
        st = self._forward(dropout=self.train_dropout, need_grad=True, on_pools_done=lambda: self._rnc_labels_early(1),
                           on_frames_done=lambda: self._rnc_labels_early(2))
        d_vals, d_f, d_rnc, d_th, d_ct = self._loss_and_seeds(st)
        self.grads.zero_()
        if self.world == 1:
            self.engine.backward(self.W, st, d_vals=d_vals, d_fused=d_f, d_rnc=d_rnc, d_th=d_th, d_ct=d_ct)
        else:
            # gradient all-reduce overlapped with the backward pass, in two buckets that are contiguous ranges of
            # the flat gradient buffer (params.ParamLayout orders the live tensors like the reference constructor):
            #   early  [chain_begin, n_live): utterance chain, heads and Cross_Attention blocks - final once the modality
            #          MLPs are done; reduced on a side stream while the FRA2UTT_new blocks and the in-projection weight
            #          gradients (the last ~1/3 of the backward pass) still run;
            #   late   [0, chain_begin): in-projections + FRA2UTT_new blocks, reduced when the backward pass ends.
            import torch.distributed as dist
            main = torch.cuda.current_stream()
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream(device=self.device)
            comm, split = self._comm_stream, self._chain_begin

            def early_bucket():
                comm.wait_stream(main)
                with torch.cuda.stream(comm):
                    dist.all_reduce(self.grads[split:self.layout.n_live], group=self.pg)
            self.engine.backward(self.W, st, d_vals=d_vals, d_fused=d_f, d_rnc=d_rnc, d_th=d_th, d_ct=d_ct,
                                 on_chain_grads_final=early_bucket)
            dist.all_reduce(self.grads[:split], group=self.pg)
            main.wait_stream(comm)
        n = self.layout.n_live
        ops.adam(self.master, self.grads, self.m, self.v, lr=self.lr, step=1, weight_decay=self.weight_decay,
                 p_bf16=self.shadow, n=n, step_dev=self.step_dev, lr_dev=self.lr_dev)
        self._outputs = Engine.outputs(st)

    def _shape_key(self):
        return (self.cur_B, tuple(self.cur_frames[k] for k in ("a", "t0", "v", "t1")), self.train_dropout)

    def _full_key(self):
        return (self.B, tuple(self.frames[k] for k in ("a", "t0", "v", "t1")), self.train_dropout)

    def _cache_put(self, cache, key, entry):
        cache[key] = entry
        while len(cache) > self.max_graphs:            # dicts iterate in insertion order: the first key is the LRU one
            cache.pop(next(iter(cache)))

    @staticmethod
    def _cache_touch(cache, key):
        cache[key] = cache.pop(key)

    def train_step(self):
        """One optimisation step on the batch currently in the static buffers.  Returns nothing; read
        `terms` (device tensor: 6 loss terms + total) and `predictions()` when needed.

        The first step of a trainer runs eagerly (one-time kernel attribute setup).  After that a batch shape is
        captured into a CUDA graph the second time it is seen (the full-capacity shape: immediately) and replayed
        from then on; shapes seen once (ragged real-data batches) run the same kernels eagerly."""
        key = self._shape_key()
        ent = self._graphs.get(key)
        if ent is not None:
            self._cache_touch(self._graphs, key)
            ent["graph"].replay()
            self.n_replays += 1
            self._outputs = ent["outputs"]               # the replay refreshed THESE tensors (not an eager step's)
        else:
            seen = self._seen.get(key, 0)
            self._seen[key] = seen + 1
            if self.world > 1 and seen == 0:
                self._check_equal_shards()
            capture = self.use_graph and self.n_steps > 0 and (key == self._full_key() or seen >= 1)
            if not capture:
                self._step_body()
            else:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):   # capture does not execute ...
                    self._step_body()
                self._cache_put(self._graphs, key, {"graph": g, "outputs": self._outputs})
                g.replay()                  # ... so this replay is the step
        self.n_steps += 1

    def _check_equal_shards(self):
        """Data-parallel steps need the same shard size on every rank (all-gather counts, the B*world normalisation,
        the anchor ranges of the global RnC).  Checked once per new batch shape; a mismatch would otherwise hang or
        corrupt the collectives."""
        import torch.distributed as dist
        t = torch.tensor([self.cur_B, -self.cur_B], dtype=torch.int64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.pg)
        hi, lo = int(t[0].item()), -int(t[1].item())
        if hi != lo:
            raise _lib.SdumcError(f"data-parallel step with unequal shards: this rank has {self.cur_B} utterances, the "
                                  f"ranks range over [{lo}, {hi}]; use dataset.batch_chunks(lockstep=True)")

    @property
    def _graph(self):
        """The captured full-capacity step (None until it exists)."""
        for tr_drop in (True, False):
            ent = self._graphs.get(self._full_key()[:2] + (tr_drop,))
            if ent is not None:
                return ent["graph"]
        return None

    def close(self):
        """Drops the captured graph (it holds NCCL kernels of the process group in the data-parallel case) and
        drains the device; call before torch.distributed.destroy_process_group()."""
        torch.cuda.synchronize(self.device)
        self._graphs.clear()
        self._score_graphs.clear()
        self._outputs = None
        torch.cuda.synchronize(self.device)

    def predictions(self):
        """(vals_full [B,1], vals_missing [B,1]) of the last step (device tensors)."""
        vals = self._outputs[0]
        return vals[0], vals[1]

    # ---- scoring (main_frame_val_text_missing_inference.py:158-175) --------------------------
    @staticmethod
    def _score_dict(st):
        vals, f, rnc, th, ct = Engine.outputs(st)
        return {"val_preds_full": vals[0], "val_preds_missing": vals[1], "full_rep": f[0], "missing_rep": f[1],
                "full_rnc": rnc[0], "missing_rnc": rnc[1], "text_rep_query_full": th[0].contiguous(),
                "text_rep_query_missing": th[1].contiguous(), "text_rep_full": ct[0], "text_rep_missing": ct[1]}

    SCORE_KEYS = ("val_preds_full", "val_preds_missing", "full_rep", "missing_rep", "full_rnc", "missing_rnc",
                  "text_rep_query_full", "text_rep_query_missing", "text_rep_full", "text_rep_missing")

    def _score_body(self):
        d = self._score_dict(self._forward(dropout=False, need_grad=False))
        self._pack_scores(d)
        return d

    def _pack_scores(self, d):
        """All outputs of a scored batch side by side in one [b, sum of widths] fp32 tensor (part of the captured graph):
        a caller that collects every batch on the host needs ONE device->host copy per batch instead of ten."""
        b = d["val_preds_full"].shape[0]
        self.score_layout = [(k, tuple(d[k].shape[1:])) for k in self.SCORE_KEYS]
        self.last_packed = torch.cat([d[k].reshape(b, -1).float() for k in self.SCORE_KEYS], dim=1)

    def unpack_scores(self, packed):
        """Splits a (host or device, torch or numpy) [n, width] array of packed score rows into the dict of score()."""
        out, c = {}, 0
        for k, shp in self.score_layout:
            w = 1
            for x in shp:
                w *= x
            out[k] = packed[:, c:c + w].reshape((packed.shape[0],) + shp)
            c += w
        return out

    @torch.no_grad()
    def score_varlen(self, store, idx):
        """Scores the ragged batch `idx` of a DeviceStore4F WITHOUT touching padded frames, with the
        reference's semantics: the reference pads every modality to the batch maximum and has no mask, so padded frames
        (in-projection = bias) take part in both softmaxes (read_data.py:223-248; …text_missing.py:63, :90).  Here the
        batch is gathered PACKED (valid frames only), every frame-level GEMM runs on sum_T rows instead of B * L_max,
        and the pooling kernels add the padded frames' closed-form share.  Equal to score() on the padded batch up to
        rounding (tests/test_trainer_gpu.py); eval mode only - train-mode input dropout makes padded rows differ."""
        b = len(idx)
        if b > self.B:
            raise ValueError(f"batch of {b} exceeds the trainer's capacity {self.B}")
        frames = store.batch_frames(idx)
        idx_dev = torch.tensor(list(idx), dtype=torch.int32).to(self.device, non_blocking=True)
        inputs, row_off = {}, {}
        for key, sname, Lpad in zip(("a", "t0", "v", "t1"), ("audio", "text", "video", "feat4"), frames):
            D = self.in_dims[key]
            if store.packed[sname].shape[1] != D or Lpad > self.frames[key]:
                raise ValueError(f"stream {key}: store has D={store.packed[sname].shape[1]}, batch needs {Lpad} frames; "
                                 f"capacity [{self.B},{self.frames[key]},{D}]")
            lens = [store.lengths[sname][i] for i in idx]
            cum = [0]
            for n in lens:
                cum.append(cum[-1] + n)
            off = torch.tensor(cum, dtype=torch.int32).to(self.device, non_blocking=True)
            dst = self.in_flat[key][:cum[-1] * D].view(cum[-1], D)
            ops.collate_pad(store.packed[sname], store.offsets[sname], idx_dev, Lpad, dst, out_off=off)
            inputs[key], row_off[key] = dst, off
        cfg = Cfg(B=b, n_pass=2, frames=dict(zip(("a", "t0", "v", "t1"), frames)), dropout=False, need_grad=False,
                  seed=self.drop_seed, step=0, step_dev=self.step_dev, row_off=row_off)
        d = self._score_dict(self.engine.forward(self.W, inputs, cfg))
        self._pack_scores(d)
        return d

    @torch.no_grad()
    def score(self):
        """Both passes in eval mode on the batch in the static buffers.  Returns the dict of device tensors
        the inference CLI collects: predictions + the 4 embeddings of each pass.  Full-size batches replay a
        captured CUDA graph (~100 launches per batch otherwise dominate at the reference's inference batch of
        128): the returned tensors are then overwritten by the next score() call on the same batch shape - copy what
        must survive.  Other recurring shapes get their own graph like train_step().  `last_packed` holds the same
        outputs as one [b, width] tensor (see _pack_scores / unpack_scores)."""
        key = self._shape_key()[:2]
        ent = self._score_graphs.get(key)
        if ent is not None:
            self._cache_touch(self._score_graphs, key)
            ent["graph"].replay()
            self.last_packed = ent["packed"]
            return ent["outputs"]
        seen = self._seen.get(("score",) + key, 0)
        self._seen[("score",) + key] = seen + 1
        self._score_calls += 1
        capture = self.use_graph and self._score_calls > 1 and (key == self._full_key()[:2] or seen >= 1)
        if not capture:
            return self._score_body()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = self._score_body()
        self._cache_put(self._score_graphs, key, {"graph": g, "outputs": out, "packed": self.last_packed})
        g.replay()
        return out
