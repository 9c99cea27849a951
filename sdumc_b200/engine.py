"""Forward / backward orchestration of the UMC model on the sm_100a kernels.

Host-side mirror of WengnetMOSEIMultViewsTextMissing.forward
(reference toolkit/models/wengnet_mosei_mult_views_text_missing.py:275-370) and of its autograd
backward.  Everything that touches data is a kernel launch through the C ABI (sdumc_b200.ops);
torch supplies device buffers and the stream only.  One call processes NP passes (1 for the drop-in
nn.Module path, 2 = full + text-missing for the fused train / scoring step) as one batch of
R = NP*B utterance rows, so every utterance-level GEMM of both passes is a single launch and the
audio / video in-projections are computed once for both passes.

Data layout in HBM (all row-major):
  X_s   bf16 [B*L_s, D_s]   input stream s in {a, v, t0, t1}
  Xf/Xc bf16 [B*L_s, 256]   dropped copies of H_s = X_s W^T + b for the FRA2UTT_new / Cross_Attention
                            blocks of each (pass, modality) unit; in eval mode both alias H_s
  K     bf16 [B*L_s, 256]   tanh key projections (kept only when a backward pass will follow)
  S     fp32 [B*L_s, nq]    attention scores, overwritten by the softmax probabilities
  utterance-level activations fp32 [R, .] with bf16 copies where a weight-gradient GEMM reads them
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from .params import CROSS_MLPS, MODALITY_MLPS, QUERY_MLPS, ParamLayout

NQ = 7
FRAME_P = 0.5   # nn.Dropout(0.5) inside FRA2UTT_new / Cross_Attention (reference :54, :77)
MLP_P = 0.3     # MLP() dropout (reference :187, :270)
NUM_SMS = 148


def dropout_site_names() -> List[str]:
    """Dropout sites in the order the reference forward() visits its nn.Dropout modules."""
    s: List[str] = []
    for i in range(3):
        s += [f"fra2utt_{i}.in", f"fra2utt_{i}.out"]
    for m in (*MODALITY_MLPS, "attention_mlp"):
        s += [f"{m}.0", f"{m}.1"]
    for q in QUERY_MLPS:
        s.append(f"{q}.0")
    for i in range(3):
        s += [f"cross_att_fra2utt_{i}.in", f"cross_att_fra2utt_{i}.out"]
    for m in (*CROSS_MLPS, "cross_attention_mlp"):
        s += [f"{m}.0", f"{m}.1"]
    return s


_SITE_INDEX = {n: i for i, n in enumerate(dropout_site_names())}


def site_id(name: str, pass_idx: int = 0) -> int:
    """RNG site of a dropout layer.  Frame-level and pooled-output sites are per pass (both passes index
    the same rows); utterance-level MLP sites are shared, their rows r = pass*B + b already differ."""
    return 1 + 64 * pass_idx + _SITE_INDEX[name]


class Weights:
    """Flat fp32 master, bf16 shadow and fp32 gradient buffers sharing one ParamLayout."""

    def __init__(self, layout: ParamLayout, master: torch.Tensor, shadow: torch.Tensor,
                 grads: Optional[torch.Tensor] = None):
        self.layout, self.master, self.shadow, self.grads = layout, master, shadow, grads

    def f32(self, name):
        return self.layout.view(self.master, name)

    def bf16(self, name):
        return self.layout.view(self.shadow, name)

    def grad(self, name):
        return self.layout.view(self.grads, name)

    def refresh_shadow(self):
        ops.cast_bf16(self.master, self.shadow)


@dataclass
class Cfg:
    B: int
    n_pass: int                       # 1 or 2
    frames: Dict[str, int]            # stream -> L   (streams: a, v, t0[, t1])
    dropout: bool                     # train-mode dropout
    need_grad: bool                   # keep what the backward pass needs
    seed: int = 0
    step: int = 0
    step_dev: Optional[torch.Tensor] = None   # device uint32 counter added to step (CUDA-graph replay)
    # varlen (eval only): per stream an int32 device tensor [B+1] of cumulative valid-frame counts.
    # The inputs then hold only the valid frames, packed ([sum_T, D]); `frames` stays the padded length the
    # reference's collater would produce (the padded frames enter the softmaxes in closed form, ops.pool_fwd).
    row_off: Optional[Dict[str, torch.Tensor]] = None


@dataclass
class State:
    cfg: Cfg
    t: Dict[str, torch.Tensor] = field(default_factory=dict)   # named buffers


def _unit_stream(p: int, m: int) -> str:
    return ("a", f"t{p}", "v")[m]


INPROJ = ("frame_dim_reshape_0", "frame_dim_reshape_1", "frame_dim_reshape_2")


def _stream_mod(s: str) -> int:
    return {"a": 0, "t": 1, "v": 2}[s[0]]


def _ksplits(rows_k: int, m_gemm: int, n_gemm: int, ctas: int = NUM_SMS) -> int:
    tiles = ((m_gemm + 127) // 128) * ((n_gemm + 255) // 256)
    nkb = max(1, (rows_k + 63) // 64)
    return max(1, min(nkb, (ctas + tiles - 1) // tiles))


def _shares(weights: Dict, total: int = NUM_SMS, even: bool = True) -> Dict:
    """Splits the SMs among kernels that run side by side, proportionally to their work; even counts (CTA pairs)."""
    tot = float(sum(weights.values()))
    unit = 2 if even else 1
    out = {k: max(unit, int(total * w / tot) // unit * unit) for k, w in weights.items()}
    # hand the remainder to the largest shares
    left = total - sum(out.values())
    for k in sorted(weights, key=lambda k: -weights[k]):
        if left < unit:
            break
        out[k] += unit
        left -= unit
    return out


class Engine:
    def __init__(self, layout: ParamLayout, device, multi_stream: bool = True):
        import os
        # SDUMC_SM_SHARES=1 (experiment, off): the frame-level kernels of the units that run side by side on parallel
        # streams each get a share of the SMs proportional to their rows instead of a grid sized for the whole GPU.
        # Measured 5.17 vs 3.99 ms per step (profiles/r2_experiments.md): the non-persistent kernels of a unit only find
        # the SMs its own predecessor frees, so the short (text) pipelines become the long pole.
        self.sm_shares = os.environ.get("SDUMC_SM_SHARES", "0") == "1" and multi_stream
        # Frame-level work that nothing on the critical path waits for (the Cross_Attention key projections during
        # utterance chain A, the Cross_Attention weight gradients during the backward chain B7-B9) runs on an auxiliary
        # stream UNDER the latency-bound utterance chain, on at most `overlap_ctas` SMs so that the chain's small
        # kernels always find free SMs.  SDUMC_OVERLAP=0 restores the serial schedule.
        self.overlap = os.environ.get("SDUMC_OVERLAP", "1") == "1" and multi_stream
        # the forward half (key projections under chain A) measured slower than the serial schedule (335 vs 311 us: the
        # chain's kernels are memory-latency-bound and stretch 3x next to a streaming GEMM), the backward half faster
        # (363 vs 471 us): profiles/r2_experiments.md
        self.overlap_fwd = os.environ.get("SDUMC_OVERLAP_FWD", "0") == "1" and self.overlap
        self.overlap_ctas = int(os.environ.get("SDUMC_OVERLAP_CTAS", "96"))
        self.layout = layout
        self.G = layout.G
        self.device = device
        self.multi_stream = multi_stream
        self._streams: List[torch.cuda.Stream] = []

    # ------------------------------------------------------------------ helpers
    def _parallel(self, n: int, fn):
        """Run fn(0..n-1) on n side streams forked from the current stream, then join.  The branches of the
        model that do not depend on each other (three modalities, seven query MLPs, ...) are small kernels
        that each occupy a fraction of the SMs; running them side by side hides their launch / pipeline-fill
        latency (and becomes parallel branches of the captured CUDA graph).  Tensors shared with the main
        stream must be allocated before the fork."""
        if not self.multi_stream or n <= 1:
            for i in range(n):
                fn(i)
            return
        main = torch.cuda.current_stream()
        while len(self._streams) < n:
            self._streams.append(torch.cuda.Stream(device=self.device))
        for i in range(n):
            side = self._streams[i]
            side.wait_stream(main)
            with torch.cuda.stream(side):
                fn(i)
        for i in range(n):
            main.wait_stream(self._streams[i])

    def _side(self, fn, *keep):
        """Run fn() on an auxiliary stream forked from the current one; joined by _join_side().  For work that is off the
        critical path of the backward pass (weight / bias gradients: nothing downstream reads them before the
        optimizer).  `keep` tensors are held until the join so that their memory is not reused while fn still reads it."""
        if not self.multi_stream:
            fn()
            return
        cur = torch.cuda.current_stream()
        if not hasattr(self, "_aux"):
            self._aux, self._aux_rr, self._aux_pending, self._aux_keep = [], 0, [], []
        if len(self._aux) < 4:
            self._aux.append(torch.cuda.Stream(device=self.device))
        aux = self._aux[self._aux_rr % len(self._aux)]
        self._aux_rr += 1
        aux.wait_stream(cur)
        with torch.cuda.stream(aux):
            fn()
        self._aux_pending.append(aux)
        self._aux_keep.extend(keep)

    def _join_side(self):
        if not getattr(self, "_aux_pending", None):
            return
        cur = torch.cuda.current_stream()
        for aux in dict.fromkeys(self._aux_pending):
            cur.wait_stream(aux)
        self._aux_pending.clear()
        self._aux_keep.clear()

    def _new(self, st: State, name: str, shape, dtype=torch.float32, zero=False):
        fn = torch.zeros if zero else torch.empty
        t = fn(shape, dtype=dtype, device=self.device)
        st.t[name] = t
        return t

    def _linear_fwd(self, W: Weights, st: State, lname: str, x: torch.Tensor, y: torch.Tensor, *, relu: bool,
                    drop_site: int = 0, y_bf16: Optional[torch.Tensor] = None):
        """y = [dropout(relu(]x W^T + b[))]  as one tf32 tcgen05 GEMM (fp32 activations and master weights)."""
        cfg = st.cfg
        w = W.f32(lname + ".weight")
        N, K = w.shape
        p = MLP_P if (cfg.dropout and drop_site) else 0.0
        ops.gemm(x, w, M=x.shape[0], N=N, K=K, bias=W.f32(lname + ".bias"),
                 act=ops.ACT_RELU if relu else ops.ACT_NONE, drop_p=p, drop_site=drop_site, out_f32=y,
                 out_bf16=y_bf16, seed=cfg.seed, step=cfg.step, step_dev=cfg.step_dev)

    def _linear_bwd(self, W: Weights, st: State, lname: str, dY: Optional[torch.Tensor], x_bf16: torch.Tensor, *,
                    Y: Optional[torch.Tensor], dropped: bool, dX: Optional[torch.Tensor], dX_mode=ops.OUT_STORE,
                    dY2: Optional[torch.Tensor] = None, dZ: Optional[torch.Tensor] = None,
                    below: Optional[Tuple[torch.Tensor, bool]] = None) -> Optional[torch.Tensor]:
        """Backward of y = [drop(relu(]x W^T + b[))]: bias/weight gradients accumulate into W.grads.

        Critical path: act_bwd (dY -> bf16 dZ, bias gradient) -> dX = dZ W.  The weight gradient dW += dZ^T x runs on a
        side stream (nothing reads it before the optimizer).  With `dZ` given (the layer above already produced it, see
        `below`) act_bwd is skipped and the bias gradient is a column sum on the side stream.  With
        below = (Y_below, dropped_below) the input x of this layer is the output of another Linear+ReLU(+dropout):
        the dX GEMM applies that layer's ReLU / dropout gate in its epilogue and returns its dZ directly (bf16 for the
        GEMMs + the unrounded fp32 values for the bias gradient) - one kernel on the critical path per layer
        instead of three."""
        cfg = st.cfg
        N, K = W.f32(lname + ".weight").shape
        if dZ is None:
            rows = dY.shape[0]
            dZ = torch.empty(rows, N, dtype=torch.bfloat16, device=self.device)
            scale = 1.0 / (1.0 - MLP_P) if (dropped and cfg.dropout) else 1.0
            ops.act_bwd(dY, dZ, rows=rows, cols=N, Y=Y, scale=scale, db=W.grad(lname + ".bias"), dY2=dY2)
            self._side(lambda: self._dw(W, lname, dZ, x_bf16), dZ)
        else:
            dZ, dZ32 = dZ                             # bf16 for the GEMMs, fp32 (same values before rounding) for the bias

            def grads():
                scratch = torch.empty_like(dZ)
                ops.act_bwd(dZ32, scratch, rows=dZ.shape[0], cols=N, db=W.grad(lname + ".bias"))
                self._dw(W, lname, dZ, x_bf16)
            self._side(grads, dZ, dZ32)
        rows = dZ.shape[0]
        if below is not None:
            Yb, dropped_b = below
            dZb = torch.empty(rows, K, dtype=torch.bfloat16, device=self.device)
            dZb32 = torch.empty(rows, K, dtype=torch.float32, device=self.device)
            sc = 1.0 / (1.0 - MLP_P) if (dropped_b and cfg.dropout) else 1.0
            ops.gemm(dZ, W.bf16(lname + ".weight"), M=rows, N=K, K=N, b_mn=True, gate=Yb, gate_scale=sc, out_bf16=dZb,
                     out_f32=dZb32)
            return dZb, dZb32
        if dX is not None:
            # dX[rows,K] = dZ W : A K-major, B = W[N,K] read MN-major
            ops.gemm(dZ, W.bf16(lname + ".weight"), M=rows, N=K, K=N, b_mn=True, out_f32=dX, f32_mode=dX_mode)
        return None

    def _dw(self, W: Weights, lname: str, dZ: torch.Tensor, x_bf16: torch.Tensor):
        # dW[N,K] += dZ^T x : both operands MN-major (reduction over the rows), split-K with fp32 atomics
        rows, N = dZ.shape
        K = x_bf16.shape[1]
        ops.gemm(dZ, x_bf16, M=N, N=K, K=rows, a_mn=True, b_mn=True, k_splits=_ksplits(rows, N, K),
                 out_f32=W.grad(lname + ".weight"), f32_mode=ops.OUT_ATOMIC)

    def _mlp2_bwd(self, W: Weights, st: State, name: str, dY: torch.Tensor, *, Y_hi, x_hi_bf16, Y_lo, x_lo_bf16, dX,
                  dX_mode=ops.OUT_STORE):
        """Backward of the two-layer MLP() blocks (reference :264-273): name.3 (above) then name.0 (below)."""
        dZ_lo = self._linear_bwd(W, st, name + ".3", dY, x_hi_bf16, Y=Y_hi, dropped=True, dX=None, below=(Y_lo, True))
        self._linear_bwd(W, st, name + ".0", None, x_lo_bf16, Y=None, dropped=True, dX=dX, dX_mode=dX_mode, dZ=dZ_lo)

    # ------------------------------------------------------------------ forward
    def forward(self, W: Weights, inputs: Dict[str, torch.Tensor], cfg: Cfg, on_pools_done=None,
                on_frames_done=None) -> State:
        """on_pools_done() / on_frames_done() are called (on the launching stream) once the FRA2UTT_new blocks / the last
        frame-level kernel of the forward pass have been issued: what follows is the latency-bound utterance chain A / B,
        which leaves most SMs idle - the place for side work of the caller (the trainer forks the label-only parts of
        the Rank-N-Contrast term there)."""
        st = State(cfg)
        G = self.G
        B, NP = cfg.B, cfg.n_pass
        R = NP * B
        dev = self.device
        drop, keep = cfg.dropout, cfg.need_grad
        seed, step = cfg.seed, cfg.step
        streams = ["a", "v"] + [f"t{p}" for p in range(NP)]
        units = [(p, m) for p in range(NP) for m in range(3)]

        # 1. inputs -> bf16, in-projection (+ the dropped copies each attention block consumes); the streams run side by
        #    side, each on a share of the SMs proportional to its FLOPs (outputs are allocated before the fork)
        plan = []
        varlen = cfg.row_off is not None
        # modality m can run its passes as one GEMM when the passes' streams have the same shape (audio / video: the same
        # stream; text: the text and the text-substitute streams, equal in the fixed-length configurations)
        mrg = {m: drop and not varlen and len({cfg.frames[_unit_stream(p, m)] for p in range(NP)}) == 1 for m in range(3)}
        # without input dropout the passes of a modality fed by ONE stream (audio, video) see identical frames: its
        # FRA2UTT block and its Cross_Attention key projection are computed once and shared by the passes (scoring, the
        # validation passes, --no_dropout training)
        same = {m: (not drop) and NP > 1 and len({_unit_stream(p, m) for p in range(NP)}) == 1 for m in range(3)}
        if varlen:
            assert not drop and not keep, "the varlen layout is defined for eval mode (no dropout, no backward)"
        nrows: Dict[str, int] = {}
        for s in streams:
            x = inputs[s]
            L, D = cfg.frames[s], x.shape[-1]
            if varlen:
                assert x.dim() == 2, f"stream {s}: packed [sum_T, D] expected"
                x2 = x
            else:
                assert x.shape[0] == B and x.shape[1] == L, f"stream {s}: expected [{B},{L},D], got {tuple(x.shape)}"
                x2 = x.reshape(B * L, D)
            nrows[s] = x2.shape[0]
            if x2.dtype == torch.float32:
                xb = torch.empty(B * L, D, dtype=torch.bfloat16, device=dev)
                ops.cast_bf16(x2.contiguous(), xb)
            else:
                assert x2.dtype == torch.bfloat16
                xb = x2.contiguous()
            st.t[f"X.{s}"] = xb
            mod = _stream_mod(s)
            users = [(p, m) for (p, m) in units if _unit_stream(p, m) == s]
            if drop:
                # the dropped copies of one (block, modality) are contiguous over the passes ([NP, B*L, G]): the passes
                # share the block's weights, so its key projection and its weight gradient run as ONE GEMM over both
                tg, sites = [], []
                for (p, m) in users:
                    for blk in ("fra2utt", "cross_att_fra2utt"):
                        if mrg[m]:
                            key_ = f"X{blk[0]}.all.{m}"
                            if key_ not in st.t:
                                self._new(st, key_, (NP, B * L, G), torch.bfloat16)
                            st.t[f"X{blk[0]}.{p}.{m}"] = st.t[key_][p]
                        else:
                            self._new(st, f"X{blk[0]}.{p}.{m}", (B * L, G), torch.bfloat16)
                        tg.append(st.t[f"X{blk[0]}.{p}.{m}"])
                        sites.append(site_id(f"{blk}_{m}.in", p))
                plan.append((s, xb, L, D, INPROJ[mod], tg, sites, None))
            else:
                H = self._new(st, f"H.{s}", (nrows[s], G), torch.bfloat16)
                for (p, m) in users:
                    st.t[f"Xf.{p}.{m}"] = H
                    st.t[f"Xc.{p}.{m}"] = H
                plan.append((s, xb, L, D, INPROJ[mod], None, None, H))
        in_share = _shares({i: pl[2] * pl[3] for i, pl in enumerate(plan)}) if self.sm_shares else {}

        def inproj(i):
            s, xb, L, D, wname, tg, sites, H = plan[i]
            if tg is not None:
                ops.gemm(xb, W.bf16(wname + ".weight"), M=B * L, N=G, K=D, bias=W.f32(wname + ".bias"),
                         epi_kind=ops.EPI_INPROJ, targets=tg, target_sites=sites, seed=seed, step=step,
                         step_dev=cfg.step_dev, max_ctas=in_share.get(i, 0))
            else:
                ops.gemm(xb, W.bf16(wname + ".weight"), M=xb.shape[0], N=G, K=D, bias=W.f32(wname + ".bias"),
                         epi_kind=ops.EPI_INPROJ, out_bf16=H, max_ctas=in_share.get(i, 0))
        self._parallel(len(plan), inproj)

        # varlen: the constants of a padded frame per modality / block: h_pad = in-projection bias (as stored: bf16),
        # k_pad = tanh(W_in h_pad + b_in)
        pad_c: Dict[Tuple[str, int], Tuple[torch.Tensor, torch.Tensor]] = {}
        if varlen:
            for m in range(3):
                hp = W.bf16(INPROJ[m] + ".bias").reshape(1, G)
                for blk in ("fra2utt", "cross_att_fra2utt"):
                    kp = torch.empty(1, G, dtype=torch.bfloat16, device=dev)
                    ops.gemm(hp, W.bf16(f"{blk}_{m}.input_proj.weight"), M=1, N=G, K=G,
                             bias=W.f32(f"{blk}_{m}.input_proj.bias"), act=ops.ACT_TANH, out_bf16=kp)
                    pad_c[(blk, m)] = (hp, kp)

        # 2. FRA2UTT_new per unit: key projection + scores (GEMM epilogue), softmax + pooling
        u_pool = [self._new(st, f"u.{m}", (R, G)) for m in range(3)]
        u_pool_b = [self._new(st, f"u_bf16.{m}", (R, G), torch.bfloat16) for m in range(3)]
        for m in range(3):
            merged = mrg[m]                         # pass-contiguous X' exist: one key-projection GEMM per modality
            nr = nrows[_unit_stream(0, m)]
            if merged:
                Sall = self._new(st, f"Sf.all.{m}", (NP, nr, 1))
                Kall = self._new(st, f"Kf.all.{m}", (NP, nr, G), torch.bfloat16) if keep else None
            for p in range(NP):
                nr_p = nrows[_unit_stream(p, m)]
                if same[m] and p > 0:                  # shared with pass 0 (read-only from here on)
                    for nm in ("Sf", "Kf", "Of_pre"):
                        if f"{nm}.0.{m}" in st.t:
                            st.t[f"{nm}.{p}.{m}"] = st.t[f"{nm}.0.{m}"]
                    continue
                if merged:
                    st.t[f"Sf.{p}.{m}"] = Sall[p]
                    if keep:
                        st.t[f"Kf.{p}.{m}"] = Kall[p]
                else:
                    self._new(st, f"Sf.{p}.{m}", (nr_p, 1))
                    if keep:
                        self._new(st, f"Kf.{p}.{m}", (nr_p, G), torch.bfloat16)
                if merged:
                    if p == 0:
                        self._new(st, f"Of_pre.all.{m}", (NP, B, 1, G))
                    st.t[f"Of_pre.{p}.{m}"] = st.t[f"Of_pre.all.{m}"][p]
                else:
                    self._new(st, f"Of_pre.{p}.{m}", (B, 1, G))

        unit_share = _shares({i: cfg.frames[_unit_stream(p, m)] for i, (p, m) in enumerate(units)}) if self.sm_shares else {}

        def fra2utt_mod(m):                         # merged: ONE key projection (+ scores) over both passes, then the pools
            pre = f"fra2utt_{m}"
            L = cfg.frames[_unit_stream(0, m)]
            nr = nrows[_unit_stream(0, m)]
            Kall = st.t[f"Kf.all.{m}"].view(NP * nr, G) if keep else None
            ops.gemm(st.t[f"Xf.all.{m}"].view(NP * nr, G), W.bf16(pre + ".input_proj.weight"), M=NP * nr, N=G, K=G,
                     bias=W.f32(pre + ".input_proj.bias"), act=ops.ACT_TANH, epi_kind=ops.EPI_KEYPROJ, out_bf16=Kall,
                     qv=W.f32(pre + ".attention_context_vector"), q_stride=0, nq=1, L=L,
                     scores=st.t[f"Sf.all.{m}"].view(NP * nr, 1))
            for p in range(NP):
                ops.pool_fwd(st.t[f"Xf.{p}.{m}"], st.t[f"Sf.{p}.{m}"], B=B, L=L, nq=1, O_pre=st.t[f"Of_pre.{p}.{m}"],
                             out=u_pool[m][p * B:(p + 1) * B], out_stride_b=G, out_bf16=u_pool_b[m][p * B:(p + 1) * B],
                             drop_p=FRAME_P, site=site_id(pre + ".out", p), seed=seed, step=step, step_dev=cfg.step_dev,
                             Kt=None, Qp=None, qp_stride_b=0)

        def fra2utt_unit(i):
            p, m = units[i]
            L = cfg.frames[_unit_stream(p, m)]
            nr = nrows[_unit_stream(p, m)]
            vl = dict(row_off=cfg.row_off[_unit_stream(p, m)], Hpad=pad_c[("fra2utt", m)][0],
                      Kpad=pad_c[("fra2utt", m)][1]) if varlen else {}
            mc = unit_share.get(i, 0)
            X = st.t[f"Xf.{p}.{m}"]
            S = st.t[f"Sf.{p}.{m}"]
            Kt = st.t[f"Kf.{p}.{m}"] if keep else None
            pre = f"fra2utt_{m}"
            ctx = W.f32(pre + ".attention_context_vector")
            if G == 256:
                # one query: the score is a single dot product per row, free in the GEMM epilogue (K is stored only
                # when the backward pass needs it), and the pooling kernel reads X' alone
                ops.gemm(X, W.bf16(pre + ".input_proj.weight"), M=nr, N=G, K=G, bias=W.f32(pre + ".input_proj.bias"),
                         act=ops.ACT_TANH, epi_kind=ops.EPI_KEYPROJ, out_bf16=Kt, qv=ctx, q_stride=0, nq=1, L=L, scores=S,
                         max_ctas=mc)
                Kp, Qc = None, (ctx if varlen else None)      # varlen: the pooling kernel scores the padded frame itself
            else:
                # wider models: a row spans several N tiles of the key projection, so the scores come from the pooling
                # kernel's tensor-core product over the stored K (like the 7-query blocks)
                Kp = Kt if keep else torch.empty(nr, G, dtype=torch.bfloat16, device=dev)
                ops.gemm(X, W.bf16(pre + ".input_proj.weight"), M=nr, N=G, K=G, bias=W.f32(pre + ".input_proj.bias"),
                         act=ops.ACT_TANH, out_bf16=Kp, max_ctas=mc)
                Qc = ctx
            ops.pool_fwd(X, S, B=B, L=L, nq=1, O_pre=st.t[f"Of_pre.{p}.{m}"], out=u_pool[m][p * B:(p + 1) * B],
                         out_stride_b=G, out_bf16=u_pool_b[m][p * B:(p + 1) * B], drop_p=FRAME_P if drop else 0.0,
                         site=site_id(pre + ".out", p), seed=seed, step=step, step_dev=cfg.step_dev, Kt=Kp, Qp=Qc,
                         qp_stride_b=0, **vl)
        fu_jobs = []                                # per modality when merged, else per (pass, modality) unit
        for m in range(3):
            if mrg[m] and G == 256:
                fu_jobs.append((fra2utt_mod, m))
            else:
                fu_jobs += [(fra2utt_unit, i) for i, (p_, m_) in enumerate(units) if m_ == m and not (same[m] and p_ > 0)]
        self._parallel(len(fu_jobs), lambda j: fu_jobs[j][0](fu_jobs[j][1]))
        for m in range(3):
            if same[m]:
                for p in range(1, NP):                  # the pooled output (and its dropout-free copy) of pass 0
                    u_pool[m][p * B:(p + 1) * B].copy_(u_pool[m][:B])
                    u_pool_b[m][p * B:(p + 1) * B].copy_(u_pool_b[m][:B])
        if on_pools_done is not None:
            on_pools_done()

        # 3'. the Cross_Attention key projections depend on the in-projections only: issued here on an auxiliary stream,
        #     they stream under utterance chain A (joined before the pooling kernels of step 4)
        for m in range(3):
            if mrg[m]:
                Kall = self._new(st, f"Kc.all.{m}", (NP, nrows[_unit_stream(0, m)], G), torch.bfloat16)
            for p in range(NP):
                if same[m] and p > 0:
                    st.t[f"Kc.{p}.{m}"] = st.t[f"Kc.0.{m}"]
                    continue
                st.t[f"Kc.{p}.{m}"] = Kall[p] if mrg[m] else torch.empty(nrows[_unit_stream(p, m)], G, dtype=torch.bfloat16,
                                                                         device=dev)

        def cross_keyproj_mod(m, mc):               # merged: one GEMM over both passes of the modality
            pre = f"cross_att_fra2utt_{m}"
            nr = nrows[_unit_stream(0, m)]
            ops.gemm(st.t[f"Xc.all.{m}"].view(NP * nr, G), W.bf16(pre + ".input_proj.weight"), M=NP * nr, N=G, K=G,
                     bias=W.f32(pre + ".input_proj.bias"), act=ops.ACT_TANH, out_bf16=st.t[f"Kc.all.{m}"].view(NP * nr, G),
                     max_ctas=mc)

        def cross_keyproj(i, mc):
            p, m = units[i]
            pre = f"cross_att_fra2utt_{m}"
            nr = nrows[_unit_stream(p, m)]
            # K is materialised (kept for the backward pass when training, a temporary when scoring): the 7 scores
            # per row come from the pooling kernel's tensor-core product instead of 7 x 256 SIMT FMAs in the GEMM
            # epilogue - writing and re-reading K costs less than those FMAs
            ops.gemm(st.t[f"Xc.{p}.{m}"], W.bf16(pre + ".input_proj.weight"), M=nr, N=G, K=G,
                     bias=W.f32(pre + ".input_proj.bias"), act=ops.ACT_TANH, out_bf16=st.t[f"Kc.{p}.{m}"], max_ctas=mc)
        early_k = self.overlap_fwd and not varlen
        kp_jobs = []
        for m in range(3):
            if mrg[m]:
                kp_jobs.append((cross_keyproj_mod, m))
            else:
                kp_jobs += [(cross_keyproj, i) for i, (p_, m_) in enumerate(units) if m_ == m and not (same[m] and p_ > 0)]
        if early_k:
            self._side(lambda: [fn(x, self.overlap_ctas) for fn, x in kp_jobs])

        # 3. utterance chain A: modality MLPs, raw gate, partial fusions, 7 query MLPs, query projections
        cat = self._new(st, "cat", (R, 3 * G))
        cat_b = self._new(st, "cat_bf16", (R, 3 * G), torch.bfloat16)
        for m in range(3):
            self._new(st, f"h1.{m}", (R, G))
            self._new(st, f"h1_bf16.{m}", (R, G), torch.bfloat16)

        def modality_mlp(m):
            name = MODALITY_MLPS[m]
            h1, h1b = st.t[f"h1.{m}"], st.t[f"h1_bf16.{m}"]
            self._linear_fwd(W, st, name + ".0", u_pool[m], h1, relu=True, drop_site=site_id(name + ".0"), y_bf16=h1b)
            self._linear_fwd(W, st, name + ".3", h1, cat[:, m * G:(m + 1) * G], relu=True,
                             drop_site=site_id(name + ".1"), y_bf16=cat_b[:, m * G:(m + 1) * G])
        self._parallel(3, modality_mlp)
        a1 = self._new(st, "a1", (R, G))
        a1b = self._new(st, "a1_bf16", (R, G), torch.bfloat16)
        a2 = self._new(st, "a2", (R, G))
        self._linear_fwd(W, st, "attention_mlp.0", cat, a1, relu=True, drop_site=site_id("attention_mlp.0"), y_bf16=a1b)
        self._linear_fwd(W, st, "attention_mlp.3", a1, a2, relu=True, drop_site=site_id("attention_mlp.1"))
        g = self._new(st, "g", (R, 4))
        qin = self._new(st, "qin", (4, R, G))
        ops.gate_fwd(a2, W.f32("fc_att.weight"), W.f32("fc_att.bias"), cat, R=R, g=g, qin=qin)
        qin_b = self._new(st, "qin_bf16", (4, R, G), torch.bfloat16)
        if keep:
            ops.cast_bf16(qin, qin_b)
        Q = self._new(st, "Q", (R, NQ * G))
        Q_b = self._new(st, "Q_bf16", (R, NQ * G), torch.bfloat16)
        def query_mlp(i):
            name = QUERY_MLPS[i]
            x = qin[i] if i < 4 else cat[:, (i - 4) * G:(i - 3) * G]
            self._linear_fwd(W, st, name + ".0", x, Q[:, i * G:(i + 1) * G], relu=True, drop_site=site_id(name + ".0"),
                             y_bf16=Q_b[:, i * G:(i + 1) * G])
        self._parallel(NQ, query_mlp)
        Qp = [self._new(st, f"Qp.{m}", (R * NQ, G)) for m in range(3)]
        self._parallel(3, lambda m: self._linear_fwd(W, st, f"cross_att_fra2utt_{m}.query_proj", Q.view(R * NQ, G),
                                                     Qp[m], relu=False))

        # 4. Cross_Attention per unit
        C = [self._new(st, f"C.{m}", (R * NQ, G)) for m in range(3)]
        C_b = [self._new(st, f"C_bf16.{m}", (R * NQ, G), torch.bfloat16) for m in range(3)]
        for (p, m) in units:
            L = cfg.frames[_unit_stream(p, m)]
            nr = nrows[_unit_stream(p, m)]
            if mrg[m]:                                   # pass-contiguous (the stacked backward of a modality, see backward)
                if p == 0:
                    self._new(st, f"Sc.all.{m}", (NP, nr, NQ))
                    self._new(st, f"Oc_pre.all.{m}", (NP, B, NQ, G))
                st.t[f"Sc.{p}.{m}"] = st.t[f"Sc.all.{m}"][p]
                st.t[f"Oc_pre.{p}.{m}"] = st.t[f"Oc_pre.all.{m}"][p]
            else:
                self._new(st, f"Sc.{p}.{m}", (nr, NQ))
                self._new(st, f"Oc_pre.{p}.{m}", (B, NQ, G))
        if early_k:
            self._join_side()
        else:
            self._parallel(len(kp_jobs), lambda j: kp_jobs[j][0](kp_jobs[j][1], 0))

        def cross_unit(i):
            p, m = units[i]
            L = cfg.frames[_unit_stream(p, m)]
            X = st.t[f"Xc.{p}.{m}"]
            S = st.t[f"Sc.{p}.{m}"]
            pre = f"cross_att_fra2utt_{m}"
            qp = Qp[m][p * B * NQ:(p + 1) * B * NQ]
            vl = dict(row_off=cfg.row_off[_unit_stream(p, m)], Hpad=pad_c[("cross_att_fra2utt", m)][0],
                      Kpad=pad_c[("cross_att_fra2utt", m)][1]) if varlen else {}
            Kt = st.t[f"Kc.{p}.{m}"]
            ops.pool_fwd(X, S, B=B, L=L, nq=NQ, O_pre=st.t[f"Oc_pre.{p}.{m}"], out=C[m][p * B * NQ:(p + 1) * B * NQ],
                         out_stride_b=NQ * G, out_bf16=C_b[m][p * B * NQ:(p + 1) * B * NQ],
                         drop_p=FRAME_P if drop else 0.0, site=site_id(pre + ".out", p), seed=seed, step=step,
                         step_dev=cfg.step_dev, Kt=Kt, Qp=qp, qp_stride_b=NQ * G, **vl)
        self._parallel(len(units), cross_unit)
        if on_frames_done is not None:
            on_frames_done()

        # 5. utterance chain B
        c = []
        for m in range(3):
            self._new(st, f"c1.{m}", (R * NQ, 256))
            self._new(st, f"c1_bf16.{m}", (R * NQ, 256), torch.bfloat16)
            c.append(self._new(st, f"c.{m}", (R * NQ, 128)))

        def cross_mlp(m):
            name = CROSS_MLPS[m]
            c1, c1b = st.t[f"c1.{m}"], st.t[f"c1_bf16.{m}"]
            self._linear_fwd(W, st, name + ".0", C[m], c1, relu=True, drop_site=site_id(name + ".0"), y_bf16=c1b)
            self._linear_fwd(W, st, name + ".3", c1, c[m], relu=True, drop_site=site_id(name + ".1"))
        self._parallel(3, cross_mlp)
        Wc = self._new(st, "Wc", (R, NQ * 128))
        ops.weight_fwd(c, g, R=R, W=Wc)
        Wc_b = self._new(st, "Wc_bf16", (R, NQ * 128), torch.bfloat16)
        if keep:
            ops.cast_bf16(Wc, Wc_b)
        x1 = self._new(st, "x1", (R, 256))
        x1b = self._new(st, "x1_bf16", (R, 256), torch.bfloat16)
        x2 = self._new(st, "x2", (R, 128))
        self._linear_fwd(W, st, "cross_attention_mlp.0", Wc, x1, relu=True, drop_site=site_id("cross_attention_mlp.0"),
                         y_bf16=x1b)
        self._linear_fwd(W, st, "cross_attention_mlp.3", x1, x2, relu=True, drop_site=site_id("cross_attention_mlp.1"))
        r = self._new(st, "r", (R, 8))
        f = self._new(st, "f", (R, 128))
        vals = self._new(st, "vals", (R,))
        ops.final_fwd(x2, W.f32("cross_fc_att.weight"), W.f32("cross_fc_att.bias"), Wc, W.f32("fc_out_v.weight"),
                      W.f32("fc_out_v.bias"), R=R, r=r, f=f, vals=vals)
        f_b = self._new(st, "f_bf16", (R, 128), torch.bfloat16)
        if keep:
            ops.cast_bf16(f, f_b)
        o1 = self._new(st, "o1", (R, 64))
        o1b = self._new(st, "o1_bf16", (R, 64), torch.bfloat16)
        rnc = self._new(st, "rnc", (R, 64))
        self._linear_fwd(W, st, "orgin_linear_change.0", f, o1, relu=True, y_bf16=o1b)
        self._linear_fwd(W, st, "orgin_linear_change.2", o1, rnc, relu=False)
        return st

    @staticmethod
    def outputs(st: State):
        """(vals [NP,B,1], fused [NP,B,128], rnc [NP,B,64], text_hidden [NP,B,256] (strided), cross_text [NP,B,7,128])"""
        cfg = st.cfg
        NP, B = cfg.n_pass, cfg.B
        t = st.t
        G = t["Q"].shape[1] // NQ
        return (t["vals"].view(NP, B, 1), t["f"].view(NP, B, 128), t["rnc"].view(NP, B, 64),
                t["Q"].view(NP, B, NQ, G)[:, :, 5, :], t["c.1"].view(NP, B, NQ, 128))

    # ------------------------------------------------------------------ backward
    def backward(self, W: Weights, st: State, d_vals=None, d_fused=None, d_rnc=None, d_th=None, d_ct=None,
                 on_chain_grads_final=None):
        """Accumulates parameter gradients into W.grads.  d_* are fp32, contiguous, shaped like outputs()
        flattened over passes ([R,...]); None = zero.  on_chain_grads_final() is called (on the launching stream) once
        every gradient except those of the in-projections and the FRA2UTT_new blocks has been issued - the
        data-parallel trainer starts their all-reduce there, under the rest of the backward pass."""
        cfg = st.cfg
        G = self.G
        assert cfg.need_grad, "forward was run without need_grad"
        B, NP = cfg.B, cfg.n_pass
        R = NP * B
        t = st.t
        dev = self.device
        drop = cfg.dropout
        seed, step = cfg.seed, cfg.step
        units = [(p, m) for p in range(NP) for m in range(3)]
        z = lambda *shape: torch.zeros(*shape, dtype=torch.float32, device=dev)  # noqa: E731
        e = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)  # noqa: E731

        # B1. RnC head  rnc = Linear(64,64)(relu(Linear(128,64)(f)))
        df = d_fused.reshape(R, 128).clone() if d_fused is not None else z(R, 128)
        if d_rnc is not None:
            dZ_o1 = self._linear_bwd(W, st, "orgin_linear_change.2", d_rnc.reshape(R, 64), t["o1_bf16"], Y=None,
                                     dropped=False, dX=None, below=(t["o1"], False))
            self._linear_bwd(W, st, "orgin_linear_change.0", None, t["f_bf16"], Y=None, dropped=False, dX=df,
                             dX_mode=ops.OUT_ADD, dZ=dZ_o1)
        # B2. head + query gate
        dWc = e(R, NQ * 128)
        dx2 = e(R, 128)
        dv = d_vals.reshape(R).contiguous() if d_vals is not None else None
        ops.final_bwd(dv, df, t["x2"], W.f32("cross_fc_att.weight"), t["Wc"], t["r"], t["f"], W.f32("fc_out_v.weight"),
                      R=R, dWc=dWc, dx2=dx2, dWr=W.grad("cross_fc_att.weight"), dbr=W.grad("cross_fc_att.bias"),
                      dWv=W.grad("fc_out_v.weight"), dbv=W.grad("fc_out_v.bias"))
        # B3. cross_attention_mlp
        self._mlp2_bwd(W, st, "cross_attention_mlp", dx2, Y_hi=t["x2"], x_hi_bf16=t["x1_bf16"], Y_lo=t["x1"],
                       x_lo_bf16=t["Wc_bf16"], dX=dWc, dX_mode=ops.OUT_ADD)
        # B4. gate-weighted sum
        dc = [e(R * NQ, 128) for _ in range(3)]
        dg_extra = e(R, 4)
        extra = (None, d_ct.reshape(R * NQ, 128).contiguous() if d_ct is not None else None, None)
        ops.weight_bwd(dWc, [t[f"c.{m}"] for m in range(3)], t["g"], R=R, dc=dc, dg=dg_extra, dc_extra=extra)
        # B5. cross MLPs -> gradient of the (dropped) pooled cross-attention outputs
        dC = [e(R * NQ, G) for _ in range(3)]

        def cross_mlp_bwd(m):
            self._mlp2_bwd(W, st, CROSS_MLPS[m], dc[m], Y_hi=t[f"c.{m}"], x_hi_bf16=t[f"c1_bf16.{m}"], Y_lo=t[f"c1.{m}"],
                           x_lo_bf16=t[f"C_bf16.{m}"], dX=dC[m])
        self._parallel(3, cross_mlp_bwd)
        # B6. Cross_Attention blocks
        dH: Dict[str, torch.Tensor] = {}
        # a modality whose passes read DIFFERENT streams (text / text substitute) has one dH per pass: with pass-contiguous
        # tensors (forward) both passes of a block run as ONE attn_bwd + ONE dH GEMM over 2B samples, each half with its
        # own dropout sites (split_b / fmask_split) - the short text launches are dominated by their fixed cost
        stacked = {m: NP == 2 and f"Xc.all.{m}" in t and f"Kf.all.{m}" in t and
                   len({_unit_stream(p, m) for p in range(NP)}) == NP for m in range(3)}
        dH_all: Dict[int, torch.Tensor] = {}
        for (p, m) in units:                       # one dH per input stream, written first by the cross block
            s_ = _unit_stream(p, m)
            if s_ not in dH:
                if stacked[m]:
                    if m not in dH_all:
                        dH_all[m] = torch.empty(NP, B * cfg.frames[s_], G, dtype=torch.bfloat16, device=dev)
                    dH[s_] = dH_all[m][p]
                else:
                    dH[s_] = torch.empty(B * cfg.frames[s_], G, dtype=torch.bfloat16, device=dev)
        started: Dict[str, bool] = {}
        zq = z(4, R * NQ, G)                       # one fill for the four accumulators below
        dQp = [zq[m] for m in range(3)]            # attn_bwd accumulates (a sample may be split over CTAs)
        # the three modalities' blocks run side by side, each on a share of the SMs proportional to its frames
        mod_share = _shares({m: cfg.frames[_unit_stream(0, m)] for m in range(3)}) if self.sm_shares else {}

        deferred = [] if self.overlap else None     # the blocks' weight-gradient GEMMs: nothing reads them before Adam
        mrg = {m: f"Xc.all.{m}" in t for m in range(3)}   # pass-contiguous X' (see forward): one weight-gradient GEMM

        def block_dw(blk, m, dZall, mc=0):          # dW_in += dZ^T X' over both passes
            pre = f"{blk}_{m}"
            Xall = t[f"X{blk[0]}.all.{m}"]
            rows = Xall.shape[0] * Xall.shape[1]
            ops.gemm(dZall.view(rows, G), Xall.view(rows, G), M=G, N=G, K=rows, a_mn=True, b_mn=True,
                     k_splits=_ksplits(rows, G, G, mc or NUM_SMS), out_f32=W.grad(pre + ".input_proj.weight"),
                     f32_mode=ops.OUT_ATOMIC, max_ctas=mc)

        def new_dz(m):
            return torch.empty(NP, B * cfg.frames[_unit_stream(0, m)], G, dtype=torch.bfloat16, device=dev) if mrg[m] else None
        dZc = [new_dz(m) for m in range(3)]

        def cross_attn_bwd(m):                     # passes of one modality accumulate into the same dH: in order
            if stacked[m]:
                self._attn_block_bwd_stacked(W, st, m, "cross_att_fra2utt", NQ, dOut=dC[m], Qp=t[f"Qp.{m}"],
                                             qp_stride=NQ * G, dQp=dQp[m], dH_all=dH_all[m], started=started, dZ=dZc[m])
            for p in range(NP if not stacked[m] else 0):
                self._attn_block_bwd(W, st, p, m, "cross_att_fra2utt", NQ, dOut=dC[m][p * B * NQ:(p + 1) * B * NQ],
                                     Qp=t[f"Qp.{m}"][p * B * NQ:(p + 1) * B * NQ], qp_stride=NQ * G,
                                     dQp=dQp[m][p * B * NQ:(p + 1) * B * NQ], dH=dH, started=started,
                                     max_ctas=mod_share.get(m, 0), defer_dw=deferred,
                                     dZ=dZc[m][p] if mrg[m] else None)
            if mrg[m]:
                if deferred is not None:
                    deferred.append((lambda mc, m=m: block_dw("cross_att_fra2utt", m, dZc[m], mc), (dZc[m],)))
                else:
                    block_dw("cross_att_fra2utt", m, dZc[m])
        self._parallel(3, cross_attn_bwd)
        if deferred:
            # ... so they stream under the latency-bound backward chain B7-B9 (joined before the early gradient bucket)
            self._side(lambda: [fn(self.overlap_ctas) for fn, _ in deferred], *[k for _, ks in deferred for k in ks])
        # B7. query projections -> dQ
        dQ = zq[3]                                   # the three blocks add their share side by side (fp32 reds)
        self._parallel(3, lambda m: self._linear_bwd(W, st, f"cross_att_fra2utt_{m}.query_proj", dQp[m],
                                                     t["Q_bf16"].view(R * NQ, G), Y=None, dropped=False, dX=dQ,
                                                     dX_mode=ops.OUT_ATOMIC))
        dQ = dQ.view(R, NQ * G)
        # B8. 7 query MLPs
        dqin = e(4, R, G)
        dcat = e(R, 3 * G)
        Qv = t["Q"]
        dth = d_th.reshape(R, G).contiguous() if d_th is not None else None
        def query_mlp_bwd(i):
            name = QUERY_MLPS[i]
            xb = t["qin_bf16"][i] if i < 4 else t["cat_bf16"][:, (i - 4) * G:(i - 3) * G]
            dX = dqin[i] if i < 4 else dcat[:, (i - 4) * G:(i - 3) * G]
            self._linear_bwd(W, st, name + ".0", dQ[:, i * G:(i + 1) * G], xb, Y=Qv[:, i * G:(i + 1) * G],
                             dropped=True, dX=dX, dY2=dth if i == 5 else None)
        self._parallel(NQ, query_mlp_bwd)
        # gate + partial fusions
        da2 = e(R, G)
        ops.gate_bwd(dqin, dg_extra, t["g"], t["cat"], t["a2"], W.f32("fc_att.weight"), R=R, dh=dcat, da2=da2,
                     dWg=W.grad("fc_att.weight"), dbg=W.grad("fc_att.bias"))
        self._mlp2_bwd(W, st, "attention_mlp", da2, Y_hi=t["a2"], x_hi_bf16=t["a1_bf16"], Y_lo=t["a1"],
                       x_lo_bf16=t["cat_bf16"], dX=dcat, dX_mode=ops.OUT_ADD)
        # B9. modality MLPs -> gradient of the (dropped) FRA2UTT outputs
        du = [e(R, G) for _ in range(3)]

        def modality_mlp_bwd(m):
            self._mlp2_bwd(W, st, MODALITY_MLPS[m], dcat[:, m * G:(m + 1) * G], Y_hi=t["cat"][:, m * G:(m + 1) * G],
                           x_hi_bf16=t[f"h1_bf16.{m}"], Y_lo=t[f"h1.{m}"], x_lo_bf16=t[f"u_bf16.{m}"], dX=du[m])
        self._parallel(3, modality_mlp_bwd)
        if on_chain_grads_final is not None:
            self._join_side()                         # the chain's weight gradients ran on side streams
            on_chain_grads_final()
        # B10. FRA2UTT_new blocks
        dZf = [new_dz(m) for m in range(3)]

        def fra2utt_bwd(m):
            pre = f"fra2utt_{m}"
            if stacked[m]:
                self._attn_block_bwd_stacked(W, st, m, "fra2utt", 1, dOut=du[m],
                                             Qp=W.f32(pre + ".attention_context_vector"), qp_stride=0,
                                             dQp=W.grad(pre + ".attention_context_vector"), dH_all=dH_all[m],
                                             started=started, dZ=dZf[m])
            for p in range(NP if not stacked[m] else 0):
                self._attn_block_bwd(W, st, p, m, "fra2utt", 1, dOut=du[m][p * B:(p + 1) * B],
                                     Qp=W.f32(pre + ".attention_context_vector"), qp_stride=0,
                                     dQp=W.grad(pre + ".attention_context_vector"), dH=dH, started=started,
                                     max_ctas=mod_share.get(m, 0), dZ=dZf[m][p] if mrg[m] else None)
            if mrg[m]:
                block_dw("fra2utt", m, dZf[m])
        self._parallel(3, fra2utt_bwd)
        # B11. in-projection weight / bias gradients (inputs carry no gradient)
        items = list(dH.items())
        dw_share = _shares({i: t[f"X.{s_}"].numel() for i, (s_, _) in enumerate(items)}) if self.sm_shares else {}

        def inproj_bwd(i):
            s, dHs = items[i]
            wname = INPROJ[_stream_mod(s)]
            X = t[f"X.{s}"]
            rows, D = X.shape
            mc = dw_share.get(i, 0)
            ops.gemm(dHs, X, M=G, N=D, K=rows, a_mn=True, b_mn=True, k_splits=_ksplits(rows, G, D, mc or NUM_SMS),
                     out_f32=W.grad(wname + ".weight"), f32_mode=ops.OUT_ATOMIC, max_ctas=mc)
            ops.colsum_bf16(dHs, W.grad(wname + ".bias"))
        self._parallel(len(items), inproj_bwd)
        self._join_side()

    def _attn_block_bwd_stacked(self, W: Weights, st: State, m: int, blk: str, nq: int, *, dOut, Qp, qp_stride, dQp,
                                dH_all: torch.Tensor, started: Dict[str, bool], dZ: torch.Tensor):
        """Both passes of one attention block of a modality whose passes have their own streams (and their own dH):
        ONE attn_bwd over 2B samples and ONE dH GEMM over 2*B*L rows; samples / rows of pass 1 use their own dropout
        sites with their own sample / row index (split_b, fmask_split).  The weight gradient is the caller's."""
        cfg = st.cfg
        G, t, B, NP = self.G, st.t, cfg.B, cfg.n_pass
        L = cfg.frames[_unit_stream(0, m)]
        tag, pre = blk[0], f"{blk}_{m}"
        rows = NP * B * L
        first = not started.get(_unit_stream(0, m), False)
        for p in range(NP):
            started[_unit_stream(p, m)] = True
        fm = [site_id(pre + ".in", p) for p in range(NP)]
        om = [site_id(pre + ".out", p) for p in range(NP)]
        ops.attn_bwd(t[f"X{tag}.all.{m}"].view(rows, G), t[f"K{tag}.all.{m}"].view(rows, G),
                     t[f"S{tag}.all.{m}"].view(rows, nq), dOut, dout_stride_b=nq * G,
                     O_pre=t[f"O{tag}_pre.all.{m}"].view(NP * B, nq, G), Qp=Qp, qp_stride_b=qp_stride, B=NP * B, L=L, nq=nq,
                     out_drop_p=FRAME_P, out_site=om[0], dZ=dZ.view(rows, G), dH=dH_all.view(rows, G),
                     dh_mode=0 if first else 1, fmask_site=fm[0], dQp=dQp, dqp_stride_b=nq * G,
                     db=W.grad(pre + ".input_proj.bias"), seed=cfg.seed, step=cfg.step, step_dev=cfg.step_dev,
                     split_b=B, out_site2=om[1], fmask_site2=fm[1])
        ops.gemm(dZ.view(rows, G), W.bf16(pre + ".input_proj.weight"), M=rows, N=G, K=G, b_mn=True, fmask_site=fm[0],
                 fmask_site2=fm[1], fmask_split=B * L, out_bf16=dH_all.view(rows, G), bf16_mode=ops.OUT_ADD, seed=cfg.seed,
                 step=cfg.step, step_dev=cfg.step_dev)

    def _attn_block_bwd(self, W: Weights, st: State, p: int, m: int, blk: str, nq: int, *, dOut, Qp, qp_stride, dQp,
                        dH: Dict[str, torch.Tensor], started: Dict[str, bool], max_ctas: int = 0, defer_dw=None,
                        dZ: Optional[torch.Tensor] = None):
        """One (pass, modality) attention block: attn_bwd (row-wise part), dH += (dZ W_in) * M_in and - unless the caller
        passed its own dZ buffer and runs the weight gradient over both passes at once - dW_in += dZ^T X'."""
        cfg = st.cfg
        G = self.G
        t = st.t
        B = cfg.B
        s = _unit_stream(p, m)
        L = cfg.frames[s]
        tag = blk[0]  # 'f' or 'c'
        pre = f"{blk}_{m}"
        X = t[f"X{tag}.{p}.{m}"]
        Kt = t[f"K{tag}.{p}.{m}"]
        P = t[f"S{tag}.{p}.{m}"]
        Opre = t[f"O{tag}_pre.{p}.{m}"]
        own_dw = dZ is None
        if own_dw:
            dZ = torch.empty(B * L, G, dtype=torch.bfloat16, device=self.device)
        first = not started.get(s, False)          # the first block of a stream stores, later ones accumulate
        started[s] = True
        fmask = site_id(pre + ".in", p) if cfg.dropout else 0
        ops.attn_bwd(X, Kt, P, dOut, dout_stride_b=nq * G, O_pre=Opre, Qp=Qp, qp_stride_b=qp_stride, B=B, L=L, nq=nq,
                     out_drop_p=FRAME_P if cfg.dropout else 0.0, out_site=site_id(pre + ".out", p), dZ=dZ, dH=dH[s],
                     dh_mode=0 if first else 1, fmask_site=fmask, dQp=dQp, dqp_stride_b=nq * G,
                     db=W.grad(pre + ".input_proj.bias"), seed=cfg.seed, step=cfg.step, step_dev=cfg.step_dev,
                     max_ctas=max_ctas)
        # dH += (dZ W_in) * M_in
        ops.gemm(dZ, W.bf16(pre + ".input_proj.weight"), M=B * L, N=G, K=G, b_mn=True, fmask_site=fmask,
                 out_bf16=dH[s], bf16_mode=ops.OUT_ADD, seed=cfg.seed, step=cfg.step, step_dev=cfg.step_dev,
                 max_ctas=max_ctas)
        # dW_in += dZ^T X'
        def dw(mc):
            ops.gemm(dZ, X, M=G, N=G, K=B * L, a_mn=True, b_mn=True, k_splits=_ksplits(B * L, G, G, mc or NUM_SMS),
                     out_f32=W.grad(pre + ".input_proj.weight"), f32_mode=ops.OUT_ATOMIC, max_ctas=mc)
        if not own_dw:
            return
        if defer_dw is not None:
            defer_dw.append((dw, (dZ, X)))
        else:
            dw(max_ctas)
