"""GPU bring-up probe: each non-GEMM operator (and the GEMM-composed attention / linear backward)
against torch autograd on the same inputs."""
import sys
import traceback
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from sdumc_b200 import ops  # noqa: E402

dev = "cuda"
G = 256


def nerr(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))


RESULTS = {}


def report(tag, d):
    RESULTS[tag] = d
    print(f"== {tag}: " + "  ".join(f"{k}={v:.2e}" for k, v in d.items()), flush=True)


def attention_block(nq, train, B=5, L=70, fused_scores=True):
    """fused_scores: training path (K from a plain GEMM, scores inside pool_fwd); otherwise the scoring path
    (scores from the key-projection GEMM epilogue)."""
    torch.manual_seed(1)
    seed, step, site_in, site_out = 77, 3, 11, 12
    H = (torch.randn(B * L, G, device=dev) * 1.5).bfloat16()
    W = (torch.randn(G, G, device=dev) / 16).bfloat16()
    b = torch.randn(G, device=dev) * 0.1
    if nq == 1:
        Qp = torch.randn(1, 1, G, device=dev) * 0.5
        qstride = 0
    else:
        Qp = torch.randn(B, nq, G, device=dev) * 0.5
        qstride = nq * G
    dOut = torch.randn(B, nq, G, device=dev)
    # forward (kernels)
    if train:
        Min = ops.frame_mask(seed, step, site_in, B * L, G)
        X = (H.float() * Min).bfloat16()
    else:
        Min = torch.ones(B * L, G, device=dev)
        X = H
    S = torch.zeros(B * L, nq, device=dev)
    Kt = torch.zeros(B * L, G, device=dev, dtype=torch.bfloat16)
    Opre = torch.zeros(B, nq, G, device=dev)
    out = torch.zeros(B, nq, G, device=dev)
    if fused_scores:
        ops.gemm(X, W, M=B * L, N=G, K=G, bias=b, act=ops.ACT_TANH, out_bf16=Kt)
        ops.pool_fwd(X, S, B=B, L=L, nq=nq, O_pre=Opre, out=out, out_stride_b=nq * G, drop_p=0.5 if train else 0.0,
                     site=site_out, seed=seed, step=step, Kt=Kt, Qp=Qp, qp_stride_b=qstride)
    else:
        ops.gemm(X, W, M=B * L, N=G, K=G, bias=b, act=ops.ACT_TANH, epi_kind=ops.EPI_KEYPROJ, out_bf16=Kt, qv=Qp,
                 q_stride=qstride, nq=nq, L=L, scores=S)
        ops.pool_fwd(X, S, B=B, L=L, nq=nq, O_pre=Opre, out=out, out_stride_b=nq * G, drop_p=0.5 if train else 0.0,
                     site=site_out, seed=seed, step=step)
    Mout = ops.elem_mask(seed, step, site_out, B * nq * G, 0.5).view(B, nq, G) if train else torch.ones_like(out)
    # backward (kernels)
    dZ = torch.zeros(B * L, G, device=dev, dtype=torch.bfloat16)
    dH = torch.zeros(B * L, G, device=dev, dtype=torch.bfloat16)
    dQp = torch.zeros(B, nq, G, device=dev) if nq > 1 else torch.zeros(1, 1, G, device=dev)
    db = torch.zeros(G, device=dev)
    dW = torch.zeros(G, G, device=dev)
    fm = site_in if train else 0
    ops.attn_bwd(X, Kt, S, dOut, dout_stride_b=nq * G, O_pre=Opre, Qp=Qp, qp_stride_b=qstride, B=B, L=L, nq=nq,
                 out_drop_p=0.5 if train else 0.0, out_site=site_out, dZ=dZ, dH=dH, dh_mode=0, fmask_site=fm, dQp=dQp,
                 dqp_stride_b=nq * G, db=db, seed=seed, step=step)
    ops.gemm(dZ, W, M=B * L, N=G, K=G, b_mn=True, fmask_site=fm, out_bf16=dH, bf16_mode=ops.OUT_ADD, seed=seed, step=step)
    ops.gemm(dZ, X, M=G, N=G, K=B * L, a_mn=True, b_mn=True, k_splits=2, out_f32=dW, f32_mode=ops.OUT_ATOMIC)
    torch.cuda.synchronize()
    # torch reference (fp32, same rounded operands)
    Hr = H.float().view(B, L, G).clone().requires_grad_(True)
    Wr = W.float().clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    Qr = Qp.clone().requires_grad_(True)
    Xr = Hr * Min.view(B, L, G)
    Kr = torch.tanh(Xr @ Wr.t() + br)
    Sr = Kr @ (Qr.expand(B, nq, G)).transpose(1, 2)
    Pr = torch.softmax(0.3 * Sr, dim=1)
    Or = Pr.transpose(1, 2) @ Xr
    outr = Or * Mout
    (outr * dOut).sum().backward()
    report(f"attn nq={nq} train={train} fused={fused_scores}", dict(
        K=nerr(Kt.view(B, L, G), Kr), P=nerr(S.view(B, L, nq), Pr), Opre=nerr(Opre, Or), out=nerr(out, outr),
        dH=nerr(dH.view(B, L, G), Hr.grad), dW=nerr(dW, Wr.grad), db=nerr(db, br.grad),
        dQp=nerr(dQp, Qr.grad)))


def linear_bwd(rows, N, K, dropped):
    from sdumc_b200.engine import Cfg, Engine, State, Weights
    from sdumc_b200.params import ParamLayout
    torch.manual_seed(2)
    x = torch.randn(rows, K, device=dev)
    w = torch.randn(N, K, device=dev) / K ** 0.5
    bias = torch.randn(N, device=dev) * 0.1
    mask = (torch.rand(rows, N, device=dev) > 0.3).float() / 0.7 if dropped else torch.ones(rows, N, device=dev)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    yr = torch.relu(xr @ wr.t() + br) * mask
    dY = torch.randn(rows, N, device=dev)
    (yr * dY).sum().backward()
    Y = yr.detach().contiguous()
    dZ = torch.zeros(rows, N, device=dev, dtype=torch.bfloat16)
    db = torch.zeros(N, device=dev)
    ops.act_bwd(dY, dZ, rows=rows, cols=N, Y=Y, scale=1 / 0.7 if dropped else 1.0, db=db)
    dW = torch.zeros(N, K, device=dev)
    ops.gemm(dZ, x.bfloat16(), M=N, N=K, K=rows, a_mn=True, b_mn=True, k_splits=3, out_f32=dW, f32_mode=ops.OUT_ATOMIC)
    dX = torch.zeros(rows, K, device=dev)
    ops.gemm(dZ, w.bfloat16(), M=rows, N=K, K=N, b_mn=True, out_f32=dX)
    torch.cuda.synchronize()
    dZr = dY * (Y > 0).float() * (1 / 0.7 if dropped else 1.0)
    report(f"linear_bwd rows={rows} N={N} K={K} drop={dropped}", dict(
        dZ=nerr(dZ, dZr), db=nerr(db, br.grad), dW=nerr(dW, wr.grad), dX=nerr(dX, xr.grad)))


def linear_fwd(rows, N, K):
    torch.manual_seed(3)
    x = torch.randn(rows, K, device=dev)
    w = torch.randn(N, K, device=dev) / K ** 0.5
    bias = torch.randn(N, device=dev) * 0.1
    y = torch.zeros(rows, N, device=dev)
    yb = torch.zeros(rows, N, device=dev, dtype=torch.bfloat16)
    ops.gemm(x, w, M=rows, N=N, K=K, bias=bias, act=ops.ACT_RELU, out_f32=y, out_bf16=yb)
    torch.cuda.synchronize()
    ref = torch.relu(x.double() @ w.double().t() + bias.double())
    report(f"linear_fwd tf32 rows={rows} N={N} K={K}", dict(y=nerr(y, ref), y_bf16=nerr(yb, ref)))


def glue(R=37):
    torch.manual_seed(4)
    a2 = torch.randn(R, G, device=dev)
    Wg = torch.randn(3, G, device=dev) / 16
    bg = torch.randn(3, device=dev) * 0.1
    h = torch.randn(R, 3 * G, device=dev)
    g = torch.zeros(R, 4, device=dev)
    qin = torch.zeros(4, R, G, device=dev)
    ops.gate_fwd(a2, Wg, bg, h, R=R, g=g, qin=qin)
    a2r, Wgr, bgr, hr = (t.clone().requires_grad_(True) for t in (a2, Wg, bg, h))
    gr = a2r @ Wgr.t() + bgr
    ha, ht, hv = hr[:, :G], hr[:, G:2 * G], hr[:, 2 * G:]
    ga, gt, gv = gr[:, 0:1], gr[:, 1:2], gr[:, 2:3]
    qr = torch.stack([ga * ha + gt * ht + gv * hv, ga * ha + gt * ht, gt * ht + gv * hv, ga * ha + gv * hv])
    dqin = torch.randn(4, R, G, device=dev)
    dgx = torch.randn(R, 4, device=dev)
    ((qr * dqin).sum() + (gr * dgx[:, :3]).sum()).backward()
    dh = torch.zeros(R, 3 * G, device=dev)
    da2 = torch.zeros(R, G, device=dev)
    dWg = torch.zeros(3, G, device=dev)
    dbg = torch.zeros(3, device=dev)
    ops.gate_bwd(dqin, dgx, g, h, a2, Wg, R=R, dh=dh, da2=da2, dWg=dWg, dbg=dbg)
    torch.cuda.synchronize()
    report("gate", dict(g=nerr(g[:, :3], gr), qin=nerr(qin, qr), dh=nerr(dh, hr.grad), da2=nerr(da2, a2r.grad),
                        dWg=nerr(dWg, Wgr.grad), dbg=nerr(dbg, bgr.grad)))

    c = [torch.randn(R, 7, 128, device=dev) for _ in range(3)]
    gg = torch.randn(R, 4, device=dev)
    Wc = torch.zeros(R, 7, 128, device=dev)
    ops.weight_fwd(c, gg, R=R, W=Wc)
    cr = [t.clone().requires_grad_(True) for t in c]
    ggr = gg.clone().requires_grad_(True)
    Wr = sum(ggr[:, m].view(R, 1, 1) * cr[m] for m in range(3))
    dW = torch.randn(R, 7, 128, device=dev)
    ex = torch.randn(R, 7, 128, device=dev)
    ((Wr * dW).sum() + (cr[1] * ex).sum()).backward()
    dc = [torch.zeros(R, 7, 128, device=dev) for _ in range(3)]
    dg = torch.zeros(R, 4, device=dev)
    ops.weight_bwd(dW, c, gg, R=R, dc=dc, dg=dg, dc_extra=(None, ex, None))
    torch.cuda.synchronize()
    report("weight", dict(W=nerr(Wc, Wr), dc0=nerr(dc[0], cr[0].grad), dc1=nerr(dc[1], cr[1].grad),
                          dc2=nerr(dc[2], cr[2].grad), dg=nerr(dg[:, :3], ggr.grad[:, :3])))

    x2 = torch.randn(R, 128, device=dev)
    Wr_ = torch.randn(7, 128, device=dev) / 11
    br_ = torch.randn(7, device=dev) * 0.1
    Wv = torch.randn(1, 128, device=dev) / 11
    bv = torch.randn(1, device=dev)
    Wt = torch.randn(R, 7, 128, device=dev)
    r = torch.zeros(R, 8, device=dev)
    f = torch.zeros(R, 128, device=dev)
    vals = torch.zeros(R, device=dev)
    ops.final_fwd(x2, Wr_, br_, Wt, Wv, bv, R=R, r=r, f=f, vals=vals)
    x2r, Wrr, brr, Wvr, bvr, Wtr = (t.clone().requires_grad_(True) for t in (x2, Wr_, br_, Wv, bv, Wt))
    rr = x2r @ Wrr.t() + brr
    fr = (Wtr * rr.unsqueeze(2)).sum(1)
    vr = fr @ Wvr.t() + bvr
    dv = torch.randn(R, device=dev)
    dfx = torch.randn(R, 128, device=dev)
    ((vr[:, 0] * dv).sum() + (fr * dfx).sum()).backward()
    dWc = torch.zeros(R, 7, 128, device=dev)
    dx2 = torch.zeros(R, 128, device=dev)
    dWr, dbr, dWv, dbv = (torch.zeros_like(t) for t in (Wr_, br_, Wv, bv))
    ops.final_bwd(dv, dfx, x2, Wr_, Wt, r, f, Wv, R=R, dWc=dWc, dx2=dx2, dWr=dWr, dbr=dbr, dWv=dWv, dbv=dbv)
    torch.cuda.synchronize()
    report("final", dict(r=nerr(r[:, :7], rr), f=nerr(f, fr), vals=nerr(vals, vr[:, 0]), dWc=nerr(dWc, Wtr.grad),
                         dx2=nerr(dx2, x2r.grad), dWr=nerr(dWr, Wrr.grad), dbr=nerr(dbr, brr.grad),
                         dWv=nerr(dWv, Wvr.grad), dbv=nerr(dbv, bvr.grad)))


def losses(B=19):
    from oracle import sdumc_oracle as O
    torch.manual_seed(5)
    mk = lambda *s: torch.randn(*s, device=dev)  # noqa: E731
    v0, v1, y = mk(B), mk(B), mk(B)
    th0, th1, ct0, ct1, f0, f1 = mk(B, 256), mk(B, 256), mk(B, 896), mk(B, 896), mk(B, 128), mk(B, 128)
    sums = torch.zeros(8, device=dev)
    ops.loss_sums(v0, v1, y, th0, th1, ct0, ct1, f0, f1, B=B, sums=sums)
    w = [0.5, 0.5, 0.1, 0.7, 0.1, 0.8]
    terms = torch.zeros(8, device=dev)
    outs = [torch.zeros_like(t) for t in (v0, v1, th1, ct1, f0, f1)]
    rncv = torch.full((1,), 1.25, device=dev)
    ops.loss_finish(v0, v1, y, th0, th1, ct0, ct1, f0, f1, B=B, sums=sums, rnc=rncv, B_global=B, w=w, terms=terms,
                    d_v0=outs[0], d_v1=outs[1], d_th1=outs[2], d_ct1=outs[3], d_f0=outs[4], d_f1=outs[5])
    torch.cuda.synchronize()
    rq = [t.clone().requires_grad_(True) for t in (v0, v1, th1, ct1, f0, f1)]
    l = (w[0] * O.mse_loss(rq[0].view(B, 1), y) + w[1] * O.mse_loss(rq[1].view(B, 1), y) + w[2] * O.rmse_loss(rq[2], th0)
         + w[3] * O.rmse_loss(rq[3].view(B, 7, 128), ct0.view(B, 7, 128)) + w[4] * O.rmse_loss(rq[5], rq[4]) + w[5] * 1.25)
    l.backward()
    report("loss", dict(total=nerr(terms[6], l), **{f"d{i}": nerr(outs[i], rq[i].grad) for i in range(6)}))

    for n_b, tied in ((16, False), (16, True), (150, True), (512, False)):
        feats = torch.randn(n_b, 2, 64, device=dev) * 0.7
        labels = (torch.randint(-9, 10, (n_b, 1), device=dev).float() / 3.0) if tied else torch.randn(n_b, 1, device=dev)
        fr = feats.double().cpu().requires_grad_(True)
        lr = O.rnc_loss(fr, labels.double().cpu())
        lr.backward()
        ff = torch.cat([feats[:, 0], feats[:, 1]], 0).contiguous()
        yy = labels.view(-1).repeat(2).contiguous()
        loss = torch.zeros(1, device=dev)
        dfe = torch.zeros_like(ff)
        ops.rnc(ff, yy, loss=loss, dfeats=dfe)
        torch.cuda.synchronize()
        dref = torch.cat([fr.grad[:, 0], fr.grad[:, 1]], 0)
        report(f"rnc B={n_b} tied={tied}", dict(loss=nerr(loss[0], lr), dfeats=nerr(dfe, dref)))
        # sliced anchors (data-parallel use)
        loss2 = torch.zeros(1, device=dev)
        dfe2 = torch.zeros_like(ff)
        n = 2 * n_b
        for lo, hi in ((0, n // 3), (n // 3, n)):
            ops.rnc(ff, yy, loss=loss2, dfeats=dfe2, row_begin=lo, row_end=hi)
        torch.cuda.synchronize()
        report(f"rnc sliced B={n_b}", dict(loss=nerr(loss2[0], lr), dfeats=nerr(dfe2, dref)))


def adam():
    torch.manual_seed(6)
    n = 100003
    p = torch.randn(n, device=dev)
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=1e-3, weight_decay=1e-5)
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    pb = torch.zeros(n, device=dev, dtype=torch.bfloat16)
    for t in range(1, 4):
        g = torch.randn(n, device=dev)
        pr.grad = g.clone()
        opt.step()
        ops.adam(p, g, m, v, lr=1e-3, step=t, weight_decay=1e-5, p_bf16=pb)
    torch.cuda.synchronize()
    report("adam", dict(p=nerr(p, pr), shadow=nerr(pb, pr)))


if __name__ == "__main__":
    tests = [lambda: attention_block(1, False), lambda: attention_block(7, False), lambda: attention_block(1, True),
             lambda: attention_block(7, True), lambda: attention_block(7, True, B=3, L=300),
             lambda: linear_fwd(8, 256, 256), lambda: linear_fwd(56, 128, 896), lambda: linear_fwd(1000, 64, 64),
             lambda: linear_bwd(8, 256, 256, True), lambda: linear_bwd(56, 128, 256, True),
             lambda: linear_bwd(1000, 256, 768, False), lambda: linear_bwd(16, 64, 128, False),
             glue, losses, adam]
    for t in tests:
        try:
            t()
        except Exception:  # noqa: BLE001
            traceback.print_exc()
            try:
                torch.cuda.synchronize()
            except Exception as e:  # noqa: BLE001
                print("cuda error, aborting:", e)
                break
