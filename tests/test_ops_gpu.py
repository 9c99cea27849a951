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
