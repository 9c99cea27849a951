"""Synthetic MER2024 / CMU-MOSEI-shaped frame features ("S0") generated directly on the
device, and the pinned-host staging used by the end-to-end benchmark.  (No dataset or checkpoint can be
downloaded in the build environment; the reference's own loader reads per-utterance .npy files,
toolkit/utils/read_data.py:22-49.)"""
from __future__ import annotations

import math
from typing import Dict, Sequence

import torch

S0_DIMS = (1024, 4096, 1024, 4096)      # audio (WavLM-large), text (Vicuna-7B), video (MANet), feat4 (LLM-decoded)
S0_FRAMES = (384, 64, 256, 64)


def synth_batch(B: int, dims: Sequence[int] = S0_DIMS, frames: Sequence[int] = S0_FRAMES, seed: int = 1234,
                device="cuda", dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
    """x[b,l,:] = mu_b + eps with per-utterance means; text / feat4 carry 8 'massive' LLM channels;
    feat4 is a noisy re-ordering of text (never identical: RMSE has an infinite gradient at 0);
    labels come from the text mean, clipped to the MOSEI range [-3, 3]."""
    g = torch.Generator(device=device).manual_seed(seed)
    Da, Dt, Dv, D4 = dims
    La, Lt, Lv, L4 = frames
    assert D4 == Dt, "feat4 goes through frame_dim_reshape_1 and must have the text dimension"

    def stream(L, D, massive):
        mu = torch.randn(B, 1, D, generator=g, device=device) * 2.0
        x = mu + torch.randn(B, L, D, generator=g, device=device)
        if massive:
            idx = torch.randperm(D, generator=g, device=device)[:8]
            x[:, :, idx] *= 50.0
        return x, mu[:, 0]

    audio, _ = stream(La, Da, False)
    text, mu_t = stream(Lt, Dt, True)
    video, _ = stream(Lv, Dv, False)
    idx = torch.arange(L4, device=device) % Lt
    feat4 = text[:, idx.flip(0)] + 0.5 * torch.randn(B, L4, D4, generator=g, device=device)
    wv = torch.randn(Dt, generator=g, device=device)
    vals = ((mu_t @ wv) / math.sqrt(Dt) + 0.3 * torch.randn(B, generator=g, device=device)).clamp(-3, 3)
    return dict(audio=audio.to(dtype), text=text.to(dtype), video=video.to(dtype), feat4=feat4.to(dtype),
                vals=vals.float())


def pin(batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Pinned host copies of a batch (the form a data loader hands to the train step)."""
    return {k: v.detach().cpu().pin_memory() for k, v in batch.items()}
