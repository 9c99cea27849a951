"""Parameter inventory and flat-buffer layout of WengnetMOSEIMultViewsTextMissing.

Names, shapes and registration order follow the reference constructor
(toolkit/models/wengnet_mosei_mult_views_text_missing.py:187-262) so that `state_dict()` is
key-compatible with reference checkpoints.  All parameters live in ONE flat fp32 buffer (live
parameters first, each tensor 64-element aligned so every view satisfies the 16-byte TMA/vector
alignment), mirrored by a bf16 shadow with the same offsets; gradients and Adam moments use the
same layout, so the optimizer and the data-parallel all-reduce each touch a single contiguous range.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

GENERAL_DIM = 256
NUM_QUERIES = 7
ALIGN = 64

QUERY_MLPS = ("cross_fused_query_mlp", "cross_at_query_mlp", "cross_tv_query_mlp", "cross_av_query_mlp",
              "cross_audio_query_mlp", "cross_text_query_mlp", "cross_video_query_mlp")
MODALITY_MLPS = ("audio_mlp", "text_mlp", "video_mlp")
CROSS_MLPS = ("cross_audio_mlp", "cross_text_mlp", "cross_video_mlp")

# parameters that are constructed by the reference but never reached by forward() (grad stays None)
DEAD_PREFIXES = ("missing_text_imagination_mlp.", "missing_cross_text_query_imagination_mlp.", "fc_out_e.",
                 "fc_out_ev.", "prelu.", "layer_normali.")


def param_spec(input_dims: Sequence[int], general_dim: int = GENERAL_DIM) -> List[Tuple[str, Tuple[int, ...]]]:
    """general_dim = 256 is the reference (`general_dim = 256`, `fused_layer = '256,256'`, :191, :199).  Other values
    are the additive `args.general_dim` knob (BASELINE config 4, "hidden 1024"): every dimension the reference ties
    to general_dim OR to fused_layer scales together - the reference's own forward only type-checks when the two are
    equal (fc_att is Linear(general_dim, 3) on the fused_layer-wide attention_mlp output, :222-223, :303)."""
    G = general_dim
    spec: List[Tuple[str, Tuple[int, ...]]] = []

    def lin(name, out_f, in_f):
        spec.append((name + ".weight", (out_f, in_f)))
        spec.append((name + ".bias", (out_f,)))

    for i in range(3):
        lin(f"frame_dim_reshape_{i}", G, int(input_dims[i]))
    # ResidualAE(layers=[mid], n_blocks=1, input_dim=d) (:202-203): the bottleneck widths 128 / 64 are literals
    for pre, d, mid in (("missing_text_imagination_mlp", G, 128), ("missing_cross_text_query_imagination_mlp", 128, 64)):
        lin(f"{pre}.transition.0", d, 3 * d)
        lin(f"{pre}.transition.2", d, d)
        lin(f"{pre}.encoder_0.0", mid, d)
        lin(f"{pre}.decoder_0.0", d, mid)
    for i in range(3):
        spec.append((f"fra2utt_{i}.attention_context_vector", (1, G)))
        lin(f"fra2utt_{i}.input_proj", G, G)
    for m in MODALITY_MLPS:
        lin(f"{m}.0", G, G)
        lin(f"{m}.3", G, G)
    lin("attention_mlp.0", G, 3 * G)
    lin("attention_mlp.3", G, G)
    lin("fc_att", 3, G)
    for q in QUERY_MLPS:
        lin(f"{q}.0", G, G)
    for i in range(3):
        lin(f"cross_att_fra2utt_{i}.query_proj", G, G)
        lin(f"cross_att_fra2utt_{i}.input_proj", G, G)
    for m in CROSS_MLPS:
        lin(f"{m}.0", 256, G)
        lin(f"{m}.3", 128, 256)
    lin("cross_attention_mlp.0", 256, 128 * NUM_QUERIES)
    lin("cross_attention_mlp.3", 128, 256)
    lin("cross_fc_att", NUM_QUERIES, 128)
    lin("fc_out_e", 1, 128)
    lin("fc_out_v", 1, 128)
    lin("fc_out_ev", 1, 1)
    lin("orgin_linear_change.0", 64, 128)
    lin("orgin_linear_change.2", 64, 64)
    spec.append(("prelu.weight", (6,)))
    spec.append(("layer_normali.weight", (G,)))
    spec.append(("layer_normali.bias", (G,)))
    return spec


def is_live(name: str) -> bool:
    return not name.startswith(DEAD_PREFIXES)


@dataclass(frozen=True)
class Entry:
    name: str
    shape: Tuple[int, ...]
    offset: int
    numel: int
    live: bool


class ParamLayout:
    def __init__(self, input_dims: Sequence[int], general_dim: int = GENERAL_DIM):
        self.input_dims = tuple(int(d) for d in input_dims[:3])
        self.G = int(general_dim)
        if self.G not in (256, 1024):
            raise ValueError(f"general_dim must be 256 (reference) or 1024 (stress configuration), got {self.G}")
        for d in self.input_dims:
            if d % 8 != 0:
                raise ValueError(f"input feature dims must be multiples of 8 for the bf16 TMA path, got {self.input_dims}")
        self.spec = param_spec(self.input_dims, self.G)
        self.entries: Dict[str, Entry] = {}
        off = 0
        for live_pass in (True, False):
            for name, shape in self.spec:
                if is_live(name) != live_pass:
                    continue
                n = int(math.prod(shape))
                self.entries[name] = Entry(name, shape, off, n, live_pass)
                off += (n + ALIGN - 1) // ALIGN * ALIGN
            if live_pass:
                self.n_live = off          # padded extent of the live region: optimizer / all-reduce range
        self.n_total = off
        self.names = [n for n, _ in self.spec]

    def view(self, flat, name: str):
        e = self.entries[name]
        return flat[e.offset:e.offset + e.numel].view(e.shape)

    @property
    def n_live_params(self) -> int:
        return sum(e.numel for e in self.entries.values() if e.live)

    @property
    def n_params(self) -> int:
        return sum(e.numel for e in self.entries.values())
