"""Shared helpers of the GPU parity tests: run the CUDA path and the oracle on the same seeded
inputs / parameters / dropout masks and report normalised errors per tensor."""
from __future__ import annotations

import types
from typing import Dict

import torch

from oracle import sdumc_oracle as O


def nerr(got: torch.Tensor, ref: torch.Tensor, floor: float = 1e-12) -> float:
    """max |got - ref| / max |ref|"""
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(floor))


def _r16(x):
    return x.float().bfloat16().to(x.dtype)


def _tf32(x):
    """fp32 value as the tensor core reads it for kind::tf32: low 13 mantissa bits ignored."""
    y = x.float().contiguous()
    return (y.view(torch.int32) & -8192).view(torch.float32).to(x.dtype)


def _st(x, fn):
    """straight-through rounding: value fn(x), gradient of x"""
    return x + (fn(x.detach()) - x.detach())


class _ChainLinear(torch.autograd.Function):
    """Utterance-level nn.Linear with the rounding points of the CUDA path: forward = tf32 operands, exact
    accumulation; backward = dZ, saved input and weight as bf16 into the two GEMMs, bias gradient unrounded."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        return _tf32(x) @ _tf32(w).t() + b

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dz = _r16(dy)
        dz2, x2 = dz.reshape(-1, dz.shape[-1]), _r16(x).reshape(-1, x.shape[-1])
        return (dz @ _r16(w)), dz2.t() @ x2, dy.reshape(-1, dy.shape[-1]).sum(0)


_EXACT = ("fc_att", "cross_fc_att", "fc_out_v")       # SIMT fp32 dot products


def emu_linear(P, name, x):
    """Every nn.Linear of the model with the CUDA path's storage / operand precisions."""
    w, b = P[f"{name}.weight"], P[f"{name}.bias"]
    if name.startswith("frame_dim_reshape"):          # bf16 operands, fp32 accumulate, H stored as bf16
        return _st(x @ _st(w, _r16).t() + b, _r16)
    if name.endswith("input_proj"):                   # bf16 operands (x is already bf16-valued), fp32 result
        return x @ _st(w, _r16).t() + b
    if name in _EXACT:
        return x @ w.t() + b
    return _ChainLinear.apply(x, w, b)


def emu_round(tag, x):
    return _st(x, _r16)                               # K = tanh(...) is stored (and scored) as bf16


def l2err(got: torch.Tensor, ref: torch.Tensor, floor: float = 1e-12) -> float:
    """||got - ref||_2 / ||ref||_2"""
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return float((got - ref).norm() / ref.norm().clamp_min(floor))


def cuda_relu_gates(net):
    """ReLU(+dropout) on/off pattern of every MLP layer in the forward that just ran, keyed like the oracle."""
    t = net._last_state.t
    G = net.layout.G
    B = t["a2"].shape[0]
    g = {}
    for m, name in enumerate(("audio_mlp", "text_mlp", "video_mlp")):
        g[f"{name}.0"] = t[f"h1.{m}"] > 0
        g[f"{name}.3"] = t["cat"][:, m * G:(m + 1) * G] > 0
    g["attention_mlp.0"], g["attention_mlp.3"] = t["a1"] > 0, t["a2"] > 0
    for i, name in enumerate(O.QUERY_MLPS):
        g[f"{name}.0"] = t["Q"][:, i * G:(i + 1) * G] > 0
    for m, name in enumerate(("cross_audio_mlp", "cross_text_mlp", "cross_video_mlp")):
        g[f"{name}.0"] = (t[f"c1.{m}"] > 0).view(B, 7, 256)
        g[f"{name}.3"] = (t[f"c.{m}"] > 0).view(B, 7, 128)
    g["cross_attention_mlp.0"], g["cross_attention_mlp.3"] = t["x1"] > 0, t["x2"] > 0
    g["orgin_linear_change.0"] = t["o1"] > 0
    return {k: v.cpu() for k, v in g.items()}


def state_relu_gates(st, pass_idx: int):
    """ReLU on/off pattern of pass `pass_idx` of a fused two-pass forward (Trainer: rows [p*B, (p+1)*B) of every
    utterance-level buffer belong to pass p), keyed like the oracle."""
    t = st.t
    G = t["a2"].shape[1]
    B = st.cfg.B
    rows = slice(pass_idx * B, (pass_idx + 1) * B)
    rows7 = slice(pass_idx * B * 7, (pass_idx + 1) * B * 7)
    g = {}
    for m, name in enumerate(("audio_mlp", "text_mlp", "video_mlp")):
        g[f"{name}.0"] = t[f"h1.{m}"][rows] > 0
        g[f"{name}.3"] = t["cat"][rows, m * G:(m + 1) * G] > 0
    g["attention_mlp.0"], g["attention_mlp.3"] = t["a1"][rows] > 0, t["a2"][rows] > 0
    for i, name in enumerate(O.QUERY_MLPS):
        g[f"{name}.0"] = t["Q"][rows, i * G:(i + 1) * G] > 0
    for m, name in enumerate(("cross_audio_mlp", "cross_text_mlp", "cross_video_mlp")):
        g[f"{name}.0"] = (t[f"c1.{m}"][rows7] > 0).view(B, 7, 256)
        g[f"{name}.3"] = (t[f"c.{m}"][rows7] > 0).view(B, 7, 128)
    g["cross_attention_mlp.0"], g["cross_attention_mlp.3"] = t["x1"][rows] > 0, t["x2"][rows] > 0
    g["orgin_linear_change.0"] = t["o1"][rows] > 0
    return {k: v.cpu() for k, v in g.items()}


def make_relu_from_gates(gates, masks=None, stats=None, pin=True):
    """relu(name, z) = z * gate.  Where the oracle's own sign disagrees with the gate AND the unit is not
    dropped, the unit sits within forward rounding noise of 0 (counted in `stats`; `sample_flips` counts the
    samples that own at least one such unit).  pin=False only counts: the oracle keeps its own ReLU."""
    def relu(name, z):
        gate = gates[name].reshape(z.shape)
        if stats is not None:
            flip = ((z > 0) & ~gate) | ((z <= 0) & gate)
            if masks is not None:                      # train mode: a dropped unit is 0 on both sides, not a flip
                base, idx = name.rsplit(".", 1)
                site = f"{base}.{int(idx) // 3}"
                if site in masks:
                    flip = flip & (masks[site].reshape(z.shape) > 0)
            stats["units"] = stats.get("units", 0) + z.numel()
            stats["flipped"] = stats.get("flipped", 0) + int(flip.sum())
            per = flip.reshape(flip.shape[0], -1).any(dim=1)
            stats["sample_flips"] = per if "sample_flips" not in stats else (stats["sample_flips"] | per)
        return z * gate.to(z.dtype) if pin else torch.relu(z)
    return relu


def build_model(dims, P: Dict[str, torch.Tensor], seed=100, device="cuda", general_dim=256):
    from sdumc_b200.model import WengnetMOSEIMultViewsTextMissing
    net = WengnetMOSEIMultViewsTextMissing(types.SimpleNamespace(input_dims=dims, seed=seed, general_dim=general_dim))
    net.load_state_dict({k: v.float() for k, v in P.items()}, strict=True)
    return net.to(device)


def kernel_masks(net, B, frames_amv, pass_idx: int):
    """The dropout masks the kernels applied in the forward that just ran (net._step), keyed by the
    oracle's site names, shaped like the tensors the oracle drops."""
    from sdumc_b200 import ops
    from sdumc_b200.engine import FRAME_P, MLP_P, dropout_site_names, site_id
    seed, step = net.dropout_seed, net._step
    G = net.layout.G
    La, Lt, Lv = frames_amv
    Ls = (La, Lt, Lv)
    masks = {}
    for name in dropout_site_names():
        # the nn.Module path runs one pass per call: every site uses pass index 0 of that call
        sid = site_id(name, 0)
        if name.endswith(".in"):
            m = int(name.split(".")[0][-1])
            masks[name] = ops.frame_mask(seed, step, sid, B * Ls[m], G).view(B, Ls[m], G)
        elif name.endswith(".out"):
            nq = 1 if name.startswith("fra2utt") else 7
            mk = ops.elem_mask(seed, step, sid, B * nq * G, FRAME_P).view(B, nq, G)
            masks[name] = mk[:, 0] if nq == 1 else mk
        else:
            base, idx = name.rsplit(".", 1)
            seven = base.startswith("cross_") and base.endswith("_mlp") and base not in ("cross_attention_mlp",) \
                and "query" not in base
            width = {"cross_audio_mlp": (256, 128), "cross_text_mlp": (256, 128), "cross_video_mlp": (256, 128),
                     "cross_attention_mlp": (256, 128)}.get(base, (G, G))[int(idx)]
            rows = B * 7 if seven else B
            mk = ops.elem_mask(seed, step, sid, rows * width, MLP_P).view(rows, width)
            masks[name] = mk.view(B, 7, width) if seven else mk
    return {k: v.double().cpu() for k, v in masks.items()}


def run_parity(dims, frames, B, gain, train: bool, data_seed=4321, device="cuda", loss_w=None, emulate=False,
               cotangent=False, count_flips=False, pin=None, general_dim=256):
    """Returns dict of normalised errors: outputs of both passes, the 6 loss terms, every live gradient.

    cotangent=True replaces the distillation loss by sum_p <output_p, C_p> with fixed random cotangents C:
    the vector-Jacobian product of the model alone (the RMSE / RnC terms are direction-like functions of
    differences of nearly equal features at initialisation and amplify forward rounding noise)."""
    from sdumc_b200.losses import MSELoss, RMSELoss, RnCLoss
    P = O.init_params(dims, seed=100, gain=gain, dtype=torch.float64, general_dim=general_dim)
    batch = O.synth_batch(B, dims, frames, seed=data_seed)
    # the CUDA path stores its inputs as bf16: hand the oracle the same rounded values
    b64 = {k: (v.bfloat16().double() if k != "vals" else v.double()) for k, v in batch.items()}
    P_bf = {k: v for k, v in P.items()}
    net = build_model(dims, P, device=device, general_dim=general_dim)
    net.keep_last_state = True
    net.train(train)
    dev = {k: v.bfloat16().float().to(device) if k != "vals" else v.to(device) for k, v in batch.items()}

    La, Lt, Lv, L4 = frames
    v0, e0 = net([dev["audio"], dev["text"], dev["video"], False])
    masks0 = kernel_masks(net, B, (La, Lt, Lv), 0) if train else None
    gates0 = cuda_relu_gates(net) if (emulate or count_flips or pin) else None
    v1, e1 = net([dev["audio"], dev["feat4"], dev["video"], True])
    masks1 = kernel_masks(net, B, (La, L4, Lv), 1) if train else None
    gates1 = cuda_relu_gates(net) if (emulate or count_flips or pin) else None
    stats = {}
    pin = emulate if pin is None else pin     # pin the oracle's ReLU pattern to the CUDA forward's (default: with emulation)
    relu0 = make_relu_from_gates(gates0, masks0, stats=stats, pin=pin) if (emulate or count_flips or pin) else None
    relu1 = make_relu_from_gates(gates1, masks1, stats=stats, pin=pin) if (emulate or count_flips or pin) else None

    w = {**O.DEFAULT_LOSS_W, **(loss_w or {})}
    d0 = O.make_drop_from_masks(masks0) if train else None
    d1 = O.make_drop_from_masks(masks1) if train else None
    if cotangent:
        g = torch.Generator().manual_seed(data_seed + 17)
        outs_dev = [v0, *e0, v1, *e1]
        cts = [torch.randn(tuple(t.shape), generator=g, dtype=torch.float64) for t in outs_dev]
        loss = sum((t * c.float().to(device)).sum() for t, c in zip(outs_dev, cts))
        loss.backward()
        torch.cuda.synchronize()
        leaves = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
        lin, rnd = (emu_linear, emu_round) if emulate else (None, None)
        o0 = O.forward(leaves, b64["audio"], b64["text"], b64["video"], d0, lin, rnd, relu0)
        o1 = O.forward(leaves, b64["audio"], b64["feat4"], b64["video"], d1, lin, rnd, relu1)
        outs_ref = [o0[0], *o0[1], o1[0], *o1[1]]
        oloss = sum((t * c).sum() for t, c in zip(outs_ref, cts))
        names = list(leaves)
        gs = torch.autograd.grad(oloss, [leaves[k] for k in names], allow_unused=True)
        ograds = dict(zip(names, gs))
        # <outputs, C> is a random-sign sum: normalise its error by the sum of |contributions|
        scale = float(sum((t.detach() * c).abs().sum() for t, c in zip(outs_ref, cts)))
        res = {"loss": abs(float(loss) - float(oloss)) / scale}
        if stats:
            res["stat/relu_flip_frac"] = stats["flipped"] / max(1, stats["units"])
            res["stat/sample_flip_frac"] = float(stats["sample_flips"].double().mean())
        for tag, (v, e), (ov, oe) in (("p0", (v0, e0), o0), ("p1", (v1, e1), o1)):
            res[f"{tag}/vals"] = nerr(v, ov)
            for nm, a, b in zip(("fused", "rnc", "text_hidden", "cross_text"), e, oe):
                res[f"{tag}/{nm}"] = nerr(a, b)
        params = dict(net.named_parameters())
        gmax = max(float(x.abs().max()) for x in ograds.values() if x is not None)
        for name, og in ograds.items():
            if og is None:
                continue
            res[f"grad/{name}"] = nerr(params[name].grad, og, floor=1e-6 * gmax)
            res[f"gradl2/{name}"] = l2err(params[name].grad, og, floor=1e-4 * gmax * og.numel() ** 0.5)
        return res
    mse, rmse, rnc = MSELoss(), RMSELoss(), RnCLoss()
    f0, r0, th0, ct0 = e0
    f1, r1, th1, ct1 = e1
    terms = [mse(v0, dev["vals"]), mse(v1, dev["vals"]), rmse(th1, th0.detach()), rmse(ct1, ct0.detach()),
             rmse(f1, f0), rnc(torch.stack((r0, r1), dim=1), dev["vals"].unsqueeze(1))]
    loss = (w["full_mse_loss_w"] * terms[0] + w["missing_mse_loss_w"] * terms[1] + w["text_feat_loss_w"] * terms[2]
            + w["text_query_feat_loss_w"] * terms[3] + w["features_loss_w"] * terms[4] + w["rnc_loss_w"] * terms[5])
    loss.backward()
    torch.cuda.synchronize()

    oloss, oterms, ograds, (o0, o1) = O.loss_and_grads(P_bf, b64["audio"], b64["text"], b64["feat4"], b64["video"],
                                                       b64["vals"], w, d0, d1, emu_linear if emulate else None,
                                                       emu_round if emulate else None, relu0, relu1)
    res = {}
    for tag, (v, e), (ov, oe) in (("p0", (v0, e0), o0), ("p1", (v1, e1), o1)):
        res[f"{tag}/vals"] = nerr(v, ov)
        for nm, a, b in zip(("fused", "rnc", "text_hidden", "cross_text"), e, oe):
            res[f"{tag}/{nm}"] = nerr(a, b)
    for nm, a, key in zip(("mse_full", "mse_missing", "rmse_text_hidden", "rmse_cross_text", "rmse_fused", "rnc"),
                          terms, ("mse_full", "mse_missing", "rmse_text_hidden", "rmse_cross_text", "rmse_fused",
                                  "rnc")):
        res[f"term/{nm}"] = nerr(a.reshape(()), oterms[key].reshape(()))
    res["loss"] = nerr(loss.reshape(()), oloss.reshape(()))
    params = dict(net.named_parameters())
    gmax = max(float(g.abs().max()) for g in ograds.values() if g is not None)
    for name, og in ograds.items():
        p = params[name]
        if og is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, f"dead parameter {name} received a gradient"
            continue
        assert p.grad is not None, f"{name}: no gradient"
        # tensors whose true gradient is (numerically) zero are compared against the global scale
        res[f"grad/{name}"] = nerr(p.grad, og, floor=1e-6 * gmax)
        res[f"gradl2/{name}"] = l2err(p.grad, og, floor=1e-4 * gmax * og.numel() ** 0.5)
    return res
