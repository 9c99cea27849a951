"""GPU parity of every C-ABI operator against torch autograd / the oracle on the same inputs
(tools/probe_ops.py holds the cases), plus the device Philox against a numpy re-implementation."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BF16_TOL = 1e-2     # bf16 operands, fp32 accumulation
F32_TOL = 2e-5      # fp32 SIMT kernels


def _run(fn, *a, **kw):
    from tools import probe_ops
    probe_ops.RESULTS.clear()
    fn(probe_ops, *a, **kw)
    return dict(probe_ops.RESULTS)


@pytest.mark.parametrize("nq,train,B,L", [(1, False, 5, 70), (7, False, 5, 70), (1, True, 5, 70), (7, True, 5, 70),
                                          (7, True, 3, 300), (7, False, 130, 2), (1, True, 2, 1000)])
def test_pooling_attention_block_forward_backward(nq, train, B, L):
    """training path: K from the tanh GEMM, scores + softmax + pooling in pool_fwd, tensor-core backward"""
    res = _run(lambda p: p.attention_block(nq, train, B=B, L=L))
    (d,) = res.values()
    assert all(v < BF16_TOL for v in d.values()), d


@pytest.mark.parametrize("nq,B,L", [(1, 5, 70), (7, 5, 70), (7, 3, 300), (7, 130, 2), (1, 2, 1000)])
def test_pooling_attention_block_scoring_path(nq, B, L):
    """scoring path: the scores come from the key-projection GEMM epilogue (K is never re-read)"""
    res = _run(lambda p: p.attention_block(nq, False, B=B, L=L, fused_scores=False))
    (d,) = res.values()
    assert all(v < BF16_TOL for v in d.values()), d


@pytest.mark.parametrize("rows,N,K", [(8, 256, 256), (56, 128, 896), (1000, 64, 64), (1, 256, 768)])
def test_linear_forward_tf32(rows, N, K):
    (d,) = _run(lambda p: p.linear_fwd(rows, N, K)).values()
    assert d["y"] < 2e-3 and d["y_bf16"] < BF16_TOL, d


@pytest.mark.parametrize("rows,N,K,drop", [(8, 256, 256, True), (56, 128, 256, True), (1000, 256, 768, False),
                                           (16, 64, 128, False)])
def test_linear_backward_bf16(rows, N, K, drop):
    (d,) = _run(lambda p: p.linear_bwd(rows, N, K, drop)).values()
    assert all(v < BF16_TOL for v in d.values()), d


def test_gate_weight_final_kernels():
    res = _run(lambda p: p.glue(R=37))
    for tag, d in res.items():
        assert all(v < F32_TOL for v in d.values()), (tag, d)


def test_loss_and_rnc_kernels():
    res = _run(lambda p: p.losses(B=19))
    for tag, d in res.items():
        assert all(v < 2e-5 for v in d.values()), (tag, d)


def test_adam_matches_torch():
    (d,) = _run(lambda p: p.adam()).values()
    assert d["p"] < 1e-6 and d["shadow"] < BF16_TOL, d


def test_device_philox_matches_numpy():
    from sdumc_b200 import ops
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85

    def philox(c, k):
        c, k = list(c), list(k)
        for _ in range(10):
            p0, p1 = M0 * c[0], M1 * c[2]
            c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
            k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
        return c
    seed, step, site = 0x123456789ABC, 9, 42
    key = [seed & 0xFFFFFFFF, seed >> 32]
    n, p = 1000, 0.3
    got = ops.elem_mask(seed, step, site, n, p).cpu().numpy()
    thr = int(p * 2 ** 32)
    want = np.array([(1 / (1 - p)) if philox([e >> 2, 0x5D0C, site, step], key)[e & 3] >= thr else 0.0 for e in range(n)],
                    dtype=np.float32)
    assert np.allclose(got, want)
    rows, cols = 5, 256
    fm = ops.frame_mask(seed, step, site, rows, cols).cpu().numpy()
    for r in range(rows):
        for c0 in range(0, cols, 32):
            w = philox([r, c0 >> 7, site, step], key)[(c0 >> 5) & 3]
            want_row = np.array([2.0 if (w >> j) & 1 else 0.0 for j in range(32)], dtype=np.float32)
            assert np.array_equal(fm[r, c0:c0 + 32], want_row)


def test_rnc_phases_equal_the_single_call():
    """sdumc_rnc split into its label-only phase(s) (sort + bucket index, boundaries) and its feature phase - what the
    trainer runs on a side stream under the forward pass - gives the same loss and gradient as the single call,
    for a full problem and for an anchor slice (a data-parallel rank), with tied labels; a row-sized workspace
    (sdumc_rnc_workspace_bytes_rows) is enough for the slice."""
    from sdumc_b200 import ops
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(11)
    for n, D, (lo, hi) in ((600, 64, (0, 600)), (2048, 64, (512, 1024))):
        feats = (torch.randn(n, D, device=dev, generator=g) * 0.4).contiguous()
        y = ((torch.randn(n // 2, device=dev, generator=g).clamp(-3, 3) * 4).round() / 4).repeat(2).contiguous()

        def run(phases, ws):
            loss = torch.zeros(1, device=dev)
            df = torch.zeros(n, D, device=dev)
            for ph, kw in phases:
                ops.rnc(feats if ph in (ops.RNC_ALL, ops.RNC_FEATURES) else None, y, loss=loss, dfeats=df, row_begin=lo,
                        row_end=hi, grad_scale=0.8, workspace=ws, phase=ph, D=D, **kw)
            return loss.clone(), df.clone()
        ws_full = torch.empty(ops.rnc_workspace_bytes(n, D), dtype=torch.uint8, device=dev)
        ws_rows = torch.empty(ops.rnc_workspace_bytes(n, D, rows=hi - lo), dtype=torch.uint8, device=dev)
        assert ws_rows.numel() <= ws_full.numel()
        l0, d0 = run([(ops.RNC_ALL, {})], ws_full)
        l1, d1 = run([(ops.RNC_LABELS, {}), (ops.RNC_FEATURES, {})], ws_rows)
        l2, d2 = run([(ops.RNC_SORT, {}), (ops.RNC_LABELS, {"reuse_sort": True}), (ops.RNC_FEATURES, {})], ws_rows)
        assert float(l0) != 0.0 and torch.isfinite(d0).all()
        for l, d in ((l1, d1), (l2, d2)):
            # loss and gradient are accumulated with fp32 atomics (the order varies between launches): equal to rounding
            assert abs(float(l) - float(l0)) <= 1e-5 * abs(float(l0)), (float(l), float(l0))
            assert float((d - d0).abs().max()) <= 5e-6 * float(d0.abs().max())


def test_rnc_kernels_at_data_parallel_size():
    """n = 4096 rows (the global Rank-N-Contrast problem of a 4-GPU job): the large-n launch configuration
    (1024-thread anchor kernel, tiled row-gradient kernel, split column kernel, shared label sort).
    Loss and gradient against fp64 (autograd) evaluations of loss.py:278-315 on the device for subsets of the
    anchors, and the sum over 8 anchor slices (what 8 calls of a data-parallel rank compute) against one full call."""
    from sdumc_b200 import ops
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(5)
    Bq, D = 2048, 64
    n = 2 * Bq
    base = torch.randn(Bq, D, device=dev, generator=g) * 0.4
    feats = torch.cat([base, base + 0.05 * torch.randn(Bq, D, device=dev, generator=g)], 0).contiguous()
    y = (torch.randn(Bq, device=dev, generator=g).clamp(-3, 3) * 4).round() / 4      # many tied labels
    y2 = y.repeat(2).contiguous()
    ws = torch.empty(ops.rnc_workspace_bytes(n, D), dtype=torch.uint8, device=dev)

    def run(f, grads=True, slices=1):
        loss = torch.zeros(1, device=dev)
        df = torch.zeros(n, D, device=dev) if grads else None
        step = n // slices
        for k in range(slices):
            ops.rnc(f, y2, loss=loss, dfeats=df, row_begin=k * step, row_end=(k + 1) * step, workspace=ws,
                    reuse_sort=k > 0)
        return loss, df

    loss, df = run(feats)
    # fp64 reference of the loss (anchor loop of the reference, vectorised over positives and negatives)
    f64, y64 = feats.double(), y2.double()
    logit = -torch.cdist(f64, f64) / 2.0
    dlab = (y64[:, None] - y64[None, :]).abs()
    off = ~torch.eye(n, dtype=torch.bool, device=dev)
    total = torch.zeros((), dtype=torch.float64, device=dev)
    for i in range(0, n, 7):                   # every 7th anchor: ~600 rows of the 4096
        li, di = logit[i][off[i]], dlab[i][off[i]]
        li = li - li.max()
        member = di[None, :] >= (di[:, None] - 0.0001)
        denom = (member.double() * li.exp()[None, :]).sum(1)
        total = total - (li - denom.log()).sum() / (n * (n - 1))
    sub = torch.zeros(1, device=dev)
    for i in range(0, n, 7):
        ops.rnc(feats, y2, loss=sub, dfeats=None, row_begin=i, row_end=i + 1, workspace=ws, reuse_sort=i > 0)
    assert abs(float(sub) - float(total)) <= 1e-4 * abs(float(total)), (float(sub), float(total))
    # gradient: fp64 autograd of the same formula restricted to the first 64 anchors vs the kernels on that range
    A = 64
    fr = feats.double().clone().requires_grad_(True)
    tot = torch.zeros((), dtype=torch.float64, device=dev)
    ar = torch.arange(n, device=dev)
    for i in range(A):
        m = ar != i
        li = -(fr[i][None, :] - fr[m]).norm(dim=1) / 2.0
        di = dlab[i][m]
        li = li - li.max().detach()
        member = di[None, :] >= (di[:, None] - 0.0001)
        denom = (member.double() * li.exp()[None, :]).sum(1)
        tot = tot - (li - denom.log()).sum() / (n * (n - 1))
    tot.backward()
    lossA, dfA = torch.zeros(1, device=dev), torch.zeros(n, D, device=dev)
    ops.rnc(feats, y2, loss=lossA, dfeats=dfA, row_begin=0, row_end=A, workspace=ws)
    assert abs(float(lossA) - float(tot)) <= 1e-4 * abs(float(tot)), (float(lossA), float(tot))
    gerr = float((dfA.double() - fr.grad).abs().max() / fr.grad.abs().max())
    assert gerr <= 1e-3, gerr
    # 8 anchor slices sharing one sort == one call
    loss8, df8 = run(feats, slices=8)
    assert abs(float(loss8) - float(loss)) <= 1e-5 * abs(float(loss))
    assert float((df8 - df).abs().max()) <= 1e-5 * float(df.abs().max()) + 1e-9


def test_encoder_projector_concat_matches_oracle():
    """N4 drop-in (feature_extraction/llm4wav/extract_wavlm_vicuna.py:162-185) at the reference's real shape
    EncoderProjectorConcat(5, 1024, 4096): bf16 tcgen05 GEMMs vs the fp64 oracle (itself pinned to the reference class by
    tests/golden/projector_small.npz), max|err| / max|ref| <= 1e-2; trailing frames that do not fill a group of 5 are
    discarded; the state_dict keys are the reference's."""
    from oracle import sdumc_oracle as O
    from sdumc_b200.projector import EncoderProjectorConcat
    torch.manual_seed(3)
    dev = torch.device("cuda", 0)
    net = EncoderProjectorConcat(5, 1024, 4096).to(dev)
    assert list(net.state_dict()) == ["linear1.weight", "linear1.bias", "linear2.weight", "linear2.bias"]
    x = torch.randn(3, 203, 1024, device=dev)                       # 203 % 5 = 3 frames discarded
    y = net(x)
    assert y.shape == (3, 40, 4096)
    P = {k: v.double().cpu() for k, v in net.state_dict().items()}
    ref = O.encoder_projector_concat(P, x.double().cpu(), 5)
    err = float((y.double().cpu() - ref).abs().max() / ref.abs().max())
    assert err <= 1e-2, err
    with pytest.raises(Exception):
        net.cpu()(x.cpu())
