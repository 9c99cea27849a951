"""Generate tests/golden/*.npz by running the REFERENCE modules themselves (imported by file
path from /root/reference — only possible in the build container; the GPU box has no reference).

    python oracle/make_golden.py

What is recorded (fp64, small shapes):
  * eval-mode forward of both passes (prediction + 4 embeddings), the 6 loss terms and total loss,
    fingerprints of every parameter gradient, parameters after 2 reference Adam steps;
  * the same with dropout active, the masks injected into the reference's nn.Dropout modules
    (oracle.seeded_mask) so the restatement can replay them;
  * RnC / MSE / RMSE on stand-alone inputs, including tied labels.
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))
from oracle import sdumc_oracle as O  # noqa: E402

DIMS = (96, 160, 64, 160)
FRAMES = (20, 7, 13, 9)
B = 6
SEED = 4321
GAIN = 1.3
GAIN_TRAIN = 1.0     # dropout's 2x / 1.43x rescaling inflates activations; keep the RnC term finite


def load_reference():
    cwd = os.getcwd()
    os.chdir(REF)
    sys.path.insert(0, str(REF))
    try:
        spec = importlib.util.spec_from_file_location(
            "ref_model", REF / "toolkit/models/wengnet_mosei_mult_views_text_missing.py")
        ref_model = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref_model)
        spec2 = importlib.util.spec_from_file_location("ref_loss", REF / "toolkit/utils/loss.py")
        ref_loss = importlib.util.module_from_spec(spec2)
        spec2.loader.exec_module(ref_loss)
    finally:
        os.chdir(cwd)
        sys.path.remove(str(REF))
    return ref_model, ref_loss


class DropInjector:
    """Replaces nn.Dropout.forward; checks the call order against oracle.dropout_sites()."""

    def __init__(self):
        self.sites = O.dropout_sites()
        self.pass_idx = 0
        self.pos = 0
        self.active = False

    def start(self, pass_idx):
        self.pass_idx, self.pos, self.active = pass_idx, 0, True

    def stop(self):
        assert self.pos == len(self.sites), (self.pos, len(self.sites))
        self.active = False

    def __call__(self, module, x):
        if not self.active or not module.training:
            return x
        name, p = self.sites[self.pos]
        assert abs(module.p - p) < 1e-12, (name, module.p, p)
        m = O.seeded_mask(SEED, self.pass_idx, self.pos, x.shape, p).to(x.dtype)
        self.pos += 1
        return x * m


def reference_step(ref_model, ref_loss, P, batch, inject=None, adam_steps=0):
    torch.set_default_dtype(torch.float64)
    try:
        net = ref_model.WengnetMOSEIMultViewsTextMissing(types.SimpleNamespace(input_dims=DIMS))
        missing, unexpected = net.load_state_dict({k: v.double() for k, v in P.items()}, strict=True)
        mse, rmse, rnc = ref_loss.MSELoss(), ref_loss.RMSELoss(), ref_loss.RnCLoss()
        w = O.DEFAULT_LOSS_W
        opt = torch.optim.Adam(net.parameters(), lr=1e-4, weight_decay=1e-5)
        rec = {}
        for it in range(max(1, adam_steps)):
            net.train(inject is not None)
            opt.zero_grad()
            if inject:
                inject.start(0)
            v0, (f0, r0, th0, ct0) = net([batch["audio"], batch["text"], batch["video"], False])
            if inject:
                inject.stop()
                inject.start(1)
            v1, (f1, r1, th1, ct1) = net([batch["audio"], batch["feat4"], batch["video"], True])
            if inject:
                inject.stop()
            vals = batch["vals"]
            terms = [mse(v0, vals), mse(v1, vals), rmse(th1, th0.detach()), rmse(ct1, ct0.detach()),
                     rmse(f1, f0), rnc(torch.stack((r0, r1), dim=1), vals.unsqueeze(1))]
            loss = (w["full_mse_loss_w"] * terms[0] + w["missing_mse_loss_w"] * terms[1]
                    + w["text_feat_loss_w"] * terms[2] + w["text_query_feat_loss_w"] * terms[3]
                    + w["features_loss_w"] * terms[4] + w["rnc_loss_w"] * terms[5])
            loss.backward()
            if it == 0:
                rec["vals0"], rec["vals1"] = v0.detach(), v1.detach()
                for nm, t in zip(("fused", "rnc", "text_hidden", "cross_text"), (f0, r0, th0, ct0)):
                    rec[f"emb0_{nm}"] = t.detach()
                for nm, t in zip(("fused", "rnc", "text_hidden", "cross_text"), (f1, r1, th1, ct1)):
                    rec[f"emb1_{nm}"] = t.detach()
                rec["terms"] = torch.stack([t.detach() for t in terms])
                rec["loss"] = loss.detach()
                dead = []
                for name, p in net.named_parameters():
                    if p.grad is None:
                        dead.append(name)
                    else:
                        rec[f"grad/{name}"] = O.grad_fingerprint(p.grad)
                rec["dead"] = dead
            if adam_steps:
                opt.step()
        if adam_steps:
            for name, p in net.named_parameters():
                rec[f"param_after/{name}"] = O.grad_fingerprint(p.detach())
        return rec
    finally:
        torch.set_default_dtype(torch.float32)


def main():
    ref_model, ref_loss = load_reference()
    out = {}
    P = O.init_params(DIMS, seed=100, gain=GAIN, dtype=torch.float64)
    batch = {k: v.double() for k, v in O.synth_batch(B, DIMS, FRAMES, seed=SEED).items()}

    # state_dict inventory of the reference itself (names/shapes/order) pins oracle.param_spec
    net = ref_model.WengnetMOSEIMultViewsTextMissing(types.SimpleNamespace(input_dims=O.S0_DIMS))
    out["spec_names"] = np.array(list(net.state_dict().keys()))
    out["spec_shapes"] = np.array([",".join(map(str, v.shape)) for v in net.state_dict().values()])

    inj = DropInjector()
    orig = torch.nn.Dropout.forward
    torch.nn.Dropout.forward = lambda self, x: inj(self, x)
    try:
        rec_eval = reference_step(ref_model, ref_loss, P, batch, inject=None, adam_steps=2)
        P_train = O.init_params(DIMS, seed=100, gain=GAIN_TRAIN, dtype=torch.float64)
        rec_train = reference_step(ref_model, ref_loss, P_train, batch, inject=inj, adam_steps=0)
    finally:
        torch.nn.Dropout.forward = orig

    for tag, rec in (("eval", rec_eval), ("train", rec_train)):
        for k, v in rec.items():
            out[f"{tag}/{k}"] = np.array(v) if k == "dead" else v.numpy()

    # stand-alone losses
    g = torch.Generator().manual_seed(77)
    feats = torch.randn(16, 2, 64, generator=g, dtype=torch.float64)
    y_cont = torch.randn(16, 1, generator=g, dtype=torch.float64)
    y_tied = torch.randint(-3, 4, (16, 1), generator=g).double() / 3.0 * 3.0     # MOSEI-like discrete labels
    torch.set_default_dtype(torch.float64)
    rnc = ref_loss.RnCLoss()
    for tag, y in (("cont", y_cont), ("tied", y_tied)):
        ff = feats.clone().requires_grad_(True)
        l = rnc(ff, y)
        l.backward()
        out[f"rnc/{tag}/loss"] = l.detach().numpy()
        out[f"rnc/{tag}/grad"] = ff.grad.numpy()
        out[f"rnc/{tag}/labels"] = y.numpy()
    out["rnc/feats"] = feats.numpy()
    a = torch.randn(5, 7, 12, generator=g, dtype=torch.float64)
    b = torch.randn(5, 7, 12, generator=g, dtype=torch.float64)
    out["loss/a"], out["loss/b"] = a.numpy(), b.numpy()
    out["loss/mse3d"] = ref_loss.MSELoss()(a, b).numpy()
    out["loss/rmse3d"] = ref_loss.RMSELoss()(a, b).numpy()
    out["loss/mse1d"] = ref_loss.MSELoss()(a[:, 0, :1], b[:, 0, 0]).numpy()
    torch.set_default_dtype(torch.float32)

    out["meta"] = np.array([f"dims={DIMS}", f"frames={FRAMES}", f"B={B}", f"seed={SEED}", f"gain={GAIN}", f"gain_train={GAIN_TRAIN}",
                            f"torch={torch.__version__}"])
    dst = ROOT / "tests" / "golden" / "sdumc_small.npz"
    dst.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(dst, **out)
    print(f"wrote {dst} ({dst.stat().st_size / 1024:.1f} KiB, {len(out)} arrays)")


if __name__ == "__main__":
    main()
