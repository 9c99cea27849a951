"""BASELINE config 2 at full size (B = 512, S0 dims 1024/4096/1024/4096, frames 384/64/256/64) on the GPU:
  * the scoring forward of all 512 utterances against the oracle's fp32 forward on the host (eval mode has no
    cross-sample coupling, so the oracle runs the batch in chunks of 64) - the M = 196,608 tile schedule, the full-grid
    epilogues and the cross-sample persistent units are checked against an independent answer;
  * the six loss terms of a dropout-off train step against the oracle's distill_loss on the ORACLE's own outputs;
  * a 64-utterance batch through the same capacity-512 trainer: loss terms and in-projection weight gradients against
    the oracle's fp32 autograd;
  * size-independent properties (bitwise repeatability, permutation equivariance, pass equality)."""
import pytest
import torch

from oracle import sdumc_oracle as O

pytestmark = pytest.mark.gpu

B = 512
KEYS = ("val_preds_full", "val_preds_missing", "full_rep", "missing_rep", "full_rnc", "missing_rnc",
        "text_rep_query_full", "text_rep_query_missing", "text_rep_full", "text_rep_missing")


@pytest.fixture(scope="module")
def setup():
    from sdumc_b200.data import S0_DIMS, S0_FRAMES, synth_batch
    from sdumc_b200.trainer import Trainer
    dev = torch.device("cuda", 0)
    torch.manual_seed(100)
    tr = Trainer(S0_DIMS, B, S0_FRAMES, dev, seed=100)
    batch = synth_batch(B, S0_DIMS, S0_FRAMES, seed=1234, device=dev)
    return tr, batch


def _load(tr, batch, perm=None, feat4=None):
    sel = (lambda x: x) if perm is None else (lambda x: x[perm].contiguous())
    tr.load_batch(sel(batch["audio"]), sel(batch["text"]), sel(batch["video"]),
                  sel(batch["feat4"] if feat4 is None else feat4), sel(batch["vals"]))


def test_scoring_is_bitwise_reproducible_and_permutation_equivariant(setup):
    """Eval mode has no cross-sample coupling (no BatchNorm, no mask): permuting the utterances of the batch
    permutes every output, bit for bit - each row's dot products do not depend on its tile or CTA - and
    scoring the same batch twice gives identical bits (no atomics on the forward path)."""
    tr, batch = setup
    _load(tr, batch)
    ref = {k: v.clone() for k, v in tr.score().items()}
    again = tr.score()
    for k in KEYS:
        assert torch.equal(ref[k], again[k]), k
    perm = torch.randperm(B, device=tr.device, generator=torch.Generator(device=tr.device).manual_seed(3))
    _load(tr, batch, perm=perm)
    got = tr.score()
    for k in KEYS:
        assert torch.equal(ref[k][perm], got[k]), k
    assert all(bool(torch.isfinite(v).all()) for v in got.values())


def test_identical_text_streams_make_the_two_passes_identical(setup):
    """The text-missing pass sends feat4 through the text branch (…text_missing.py:275-283): with feat4 := text
    both passes compute the same function of the same bytes."""
    tr, batch = setup
    _load(tr, batch, feat4=batch["text"])
    out = tr.score()
    for a, b in (("val_preds_full", "val_preds_missing"), ("full_rep", "missing_rep"), ("full_rnc", "missing_rnc"),
                 ("text_rep_query_full", "text_rep_query_missing"), ("text_rep_full", "text_rep_missing")):
        assert torch.equal(out[a], out[b]), (a, b)


def test_loss_terms_match_the_oracle_on_the_same_forward(setup):
    """The six loss terms of a (dropout-off) train step against the oracle's distill_loss (loss.py:19-51, :278-315;
    RnC over all 1024 x 1023 pairs) evaluated in fp64 on the CUDA forward's outputs, and the first steps on one
    batch reduce the loss."""
    tr, batch = setup
    tr.train_dropout = False
    _load(tr, batch)
    out = {k: v.double().cpu() for k, v in tr.score().items()}
    o0 = (out["val_preds_full"], (out["full_rep"], out["full_rnc"], out["text_rep_query_full"], out["text_rep_full"]))
    o1 = (out["val_preds_missing"], (out["missing_rep"], out["missing_rnc"], out["text_rep_query_missing"],
                                     out["text_rep_missing"]))
    _, ref = O.distill_loss(o0, o1, batch["vals"].double().cpu())
    tr.train_step()                     # eval forward == dropout-off train forward; terms are pre-update
    torch.cuda.synchronize()
    got = tr.terms.tolist()
    names = ("mse_full", "mse_missing", "rmse_text_hidden", "rmse_cross_text", "rmse_fused", "rnc")
    for i, name in enumerate(names):
        r = float(ref[name])
        assert abs(got[i] - r) <= 2e-4 * max(1.0, abs(r)), (name, got[i], r)
    first = got[6]
    for _ in range(5):
        tr.train_step()
    torch.cuda.synchronize()
    assert tr.terms[6].item() < first


# ---- full-size numeric parity against the oracle (fp32 on the host cores) -----------------------------------
FULL_PRED_TOL = 1e-2     # max|err| / max|ref| on predictions (north_star)
FULL_EMB_TOL = 1e-2      # ... on the four embeddings of each pass
FULL_TERM_TOL = 1e-2     # |term - ref| <= tol * max(1, |ref|)
FULL_GRAD_TOL = 4e-2     # relative L2 of the in-projection weight gradients of the DISTILLATION loss vs the plain fp32
#                          oracle's autograd with its ReLU on/off pattern taken from the CUDA forward.  Measured (round 2):
#                          audio 1.1e-2, video 1.0e-2, text 3.0e-2 - the RMSE / RnC terms are direction-like functions of
#                          differences of nearly equal features and amplify forward rounding noise (the text stream, fed by
#                          both the text and the feat4 tensor, sees it twice); the model's plain vector-Jacobian product
#                          holds 2e-2 on every tensor (tests/test_unemulated_gpu.py)
FULL_GRAD_TOL_FREE = 0.2 # ... with the oracle's own ReLU pattern: units within bf16 rounding noise of 0 (~5e-4 of them)
#                          fall on the other side and each switches a whole unit's backward signal on or off - a
#                          discontinuity of the model (tests/test_unemulated_gpu.py), not a kernel error


def _oracle_forward_chunked(P, batch, chunk=64):
    outs = [[], []]
    with torch.no_grad():
        for s in range(0, B, chunk):
            sl = slice(s, s + chunk)
            a, t, v, f4 = (batch[k][sl].float().cpu() for k in ("audio", "text", "video", "feat4"))
            for p, txt in enumerate((t, f4)):
                vals, embs = O.forward(P, a, txt, v)
                outs[p].append((vals, *embs))
    cat = lambda p: tuple(torch.cat([o[i] for o in outs[p]], dim=0) for i in range(5))  # noqa: E731
    o0, o1 = cat(0), cat(1)
    return (o0[0], o0[1:]), (o1[0], o1[1:])


def _nerr(got, ref):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-12))


def test_fullsize_forward_and_loss_terms_match_oracle(setup):
    tr, batch = setup
    tr.train_dropout = False
    torch.set_num_threads(max(1, __import__("os").cpu_count() or 1))
    P = {k: v.float().cpu() for k, v in tr.state_dict().items()}
    _load(tr, batch)
    got = {k: v.clone() for k, v in tr.score().items()}
    o0, o1 = _oracle_forward_chunked(P, batch)
    pairs = (("val_preds_full", o0[0], FULL_PRED_TOL), ("val_preds_missing", o1[0], FULL_PRED_TOL),
             ("full_rep", o0[1][0], FULL_EMB_TOL), ("missing_rep", o1[1][0], FULL_EMB_TOL),
             ("full_rnc", o0[1][1], FULL_EMB_TOL), ("missing_rnc", o1[1][1], FULL_EMB_TOL),
             ("text_rep_query_full", o0[1][2], FULL_EMB_TOL), ("text_rep_query_missing", o1[1][2], FULL_EMB_TOL),
             ("text_rep_full", o0[1][3], FULL_EMB_TOL), ("text_rep_missing", o1[1][3], FULL_EMB_TOL))
    errs = {k: _nerr(got[k], ref) for k, ref, _ in pairs}
    bad = {k: errs[k] for k, _, tol in pairs if not errs[k] <= tol}
    assert not bad, f"full-size forward vs oracle: {bad} (all: {errs})"
    # loss terms of the (pre-update) dropout-off train forward against the oracle's loss on the oracle's outputs
    to64 = lambda o: (o[0].double(), tuple(e.double() for e in o[1]))  # noqa: E731
    _, ref = O.distill_loss(to64(o0), to64(o1), batch["vals"].double().cpu())
    tr.train_step()
    torch.cuda.synchronize()
    t = tr.terms.tolist()
    names = ("mse_full", "mse_missing", "rmse_text_hidden", "rmse_cross_text", "rmse_fused", "rnc")
    for i, name in enumerate(names):
        r = float(ref[name])
        assert abs(t[i] - r) <= FULL_TERM_TOL * max(1.0, abs(r)), (name, t[i], r)


def test_subbatch_step_through_the_capacity_512_trainer_matches_oracle_autograd():
    """64 utterances at the true S0 frame counts through a Trainer sized for 512 (the eager path of a smaller batch in
    the static buffers): the 6 loss terms and the in-projection weight gradients (70 % of the model's FLOPs; the
    deepest point of the backward pass) against the oracle's fp32 autograd of the same step."""
    from sdumc_b200.data import S0_DIMS, S0_FRAMES, synth_batch
    from sdumc_b200.trainer import Trainer
    dev = torch.device("cuda", 0)
    torch.manual_seed(100)
    tr = Trainer(S0_DIMS, B, S0_FRAMES, dev, seed=100, use_graph=False)
    tr.train_dropout = False
    n = 64
    batch = synth_batch(n, S0_DIMS, S0_FRAMES, seed=77, device=dev)
    P = {k: v.float().cpu() for k, v in tr.state_dict().items()}
    tr.load_batch(batch["audio"], batch["text"], batch["video"], batch["feat4"], batch["vals"])
    # keep the gradients: run the step body without Adam
    tr.step_dev.add_(1)
    st = tr._forward(dropout=False, need_grad=True)
    seeds = tr._loss_and_seeds(st)
    tr.grads.zero_()
    tr.engine.backward(tr.W, st, d_vals=seeds[0], d_fused=seeds[1], d_rnc=seeds[2], d_th=seeds[3], d_ct=seeds[4])
    torch.cuda.synchronize()
    torch.set_num_threads(max(1, __import__("os").cpu_count() or 1))
    cpu = {k: v.float().cpu() for k, v in batch.items()}
    from tests.parity_common import make_relu_from_gates, state_relu_gates
    relu0, relu1 = (make_relu_from_gates(state_relu_gates(st, p_)) for p_ in (0, 1))
    _, terms, grads, _ = O.loss_and_grads(P, cpu["audio"], cpu["text"], cpu["feat4"], cpu["video"], cpu["vals"],
                                          relu0=relu0, relu1=relu1)
    _, _, grads_free, _ = O.loss_and_grads(P, cpu["audio"], cpu["text"], cpu["feat4"], cpu["video"], cpu["vals"])
    got = tr.terms.tolist()
    for i, name in enumerate(("mse_full", "mse_missing", "rmse_text_hidden", "rmse_cross_text", "rmse_fused", "rnc")):
        r = float(terms[name])
        assert abs(got[i] - r) <= FULL_TERM_TOL * max(1.0, abs(r)), (name, got[i], r)
    errs, errs_free = {}, {}
    for i in range(3):
        name = f"frame_dim_reshape_{i}.weight"
        g, ref, ref_free = tr.W.grad(name).double().cpu(), grads[name].double(), grads_free[name].double()
        errs[name] = float((g - ref).norm() / ref.norm())
        errs_free[name] = float((g - ref_free).norm() / ref_free.norm())
    assert all(e <= FULL_GRAD_TOL for e in errs.values()), (errs, errs_free)
    assert all(e <= FULL_GRAD_TOL_FREE for e in errs_free.values()), errs_free
