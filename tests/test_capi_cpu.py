"""CPU-only checks of the drop-in boundary: the shared library builds, loads and exports every symbol
include/sdumc_b200.h declares (no compute call is made), the ctypes structs match the C layouts, the
parameter inventory matches the reference's state_dict, and the product fails loudly without a GPU."""
import ctypes as C
import re
import subprocess
import types
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from sdumc_b200 import build
    build.build()
    from sdumc_b200 import _lib
    return _lib.lib()


def test_header_symbols_are_exported(lib):
    from sdumc_b200 import _lib
    hdr = (ROOT / "include" / "sdumc_b200.h").read_text()
    declared = set(re.findall(r"\b(sdumc_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.exported_symbols())
    assert len(declared) >= 24
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (sdumc_[a-z0-9_]+)", out))
    assert declared <= exported, declared - exported
    assert lib.sdumc_version() == 1


def test_struct_layouts_match_library(lib):
    from sdumc_b200 import _lib
    for i, name in enumerate(_lib.STRUCT_ORDER):
        assert C.sizeof(_lib.STRUCTS[name]) == lib.sdumc_struct_size(i), name
    assert lib.sdumc_struct_size(99) == -1


def test_argument_errors_are_return_codes_not_crashes(lib):
    from sdumc_b200 import _lib
    d = _lib.GemmDesc()            # M = N = K = 0 and null pointers
    rc = lib.sdumc_gemm(C.byref(d), None)
    assert rc < 0 and b"gemm" in lib.sdumc_last_error()
    assert lib.sdumc_gemm(None, None) < 0
    a = _lib.STRUCTS["sdumc_rnc_args"]()
    assert lib.sdumc_rnc(C.byref(a), None) < 0


def test_rnc_workspace_sizes_and_phase_argument(lib):
    """sdumc_rnc_workspace_bytes_rows: a caller with at most `rows` anchors per call (a data-parallel rank) needs the
    label structures + rows * n coefficients and boundary quadruples, never more than the all-rows size; an unknown
    `phase` is an argument error (return code, no CUDA call)."""
    from sdumc_b200 import _lib
    lib.sdumc_rnc_workspace_bytes.restype = C.c_uint64
    lib.sdumc_rnc_workspace_bytes_rows.restype = C.c_uint64
    n, D = 8192, 64
    full = lib.sdumc_rnc_workspace_bytes(n, D)
    part = lib.sdumc_rnc_workspace_bytes_rows(n, D, 1024)
    assert lib.sdumc_rnc_workspace_bytes_rows(n, D, n) == full == lib.sdumc_rnc_workspace_bytes_rows(n, D, 0)
    assert 1024 * n * 12 <= part <= 1024 * n * 12 + 5 * (n + 64) * 4 + 2048 < full
    a = _lib.STRUCTS["sdumc_rnc_args"]()
    dummy = (C.c_float * 4)()
    a.feats = a.labels = a.loss = a.workspace = C.addressof(dummy)
    a.n, a.D, a.row_begin, a.row_end, a.workspace_bytes, a.phase = n, D, 0, 1024, part, 7
    assert lib.sdumc_rnc(C.byref(a), None) < 0 and b"phase" in lib.sdumc_last_error()


def test_kernels_are_blackwell_native():
    """SASS of the shipped library: tcgen05 MMAs (UTCHMMA, the CTA-pair form .2CTA), TMEM loads (LDTM), TMA loads
    (UTMALDG 2-D for the GEMMs, 3-D for the attention backward), tensor stores and the L2-side tensor reduce-add of the
    attention backward (UTMASTG / UTMAREDG), programmatic dependent launch (ACQBULK = griddepcontrol.wait, PREEXIT =
    griddepcontrol.launch_dependents)."""
    from sdumc_b200 import _lib
    sass = subprocess.run(["cuobjdump", "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG.2D", "UTMALDG.3D", "UTMASTG.3D", "UTMAREDG.3D.ADD",
                     "UTCBAR", "UTCBAR.2CTA.MULTICAST", "ACQBULK", "PREEXIT"):
        assert mnemonic in sass, mnemonic
    assert "HGMMA" not in sass


def test_param_layout_matches_reference_state_dict(golden):
    from sdumc_b200.params import ParamLayout, is_live
    L = ParamLayout((1024, 4096, 1024, 4096))
    assert L.names == list(golden["spec_names"])
    assert [",".join(map(str, s)) for _, s in L.spec] == list(golden["spec_shapes"])
    assert L.n_params == 4268884 and L.n_live_params == 3857291
    assert {n for n in L.names if not is_live(n)} == set(golden["eval/dead"])
    offs = sorted((e.offset, e.numel, e.live) for e in L.entries.values())
    for (o0, n0, _), (o1, _, _) in zip(offs, offs[1:]):
        assert o0 + n0 <= o1 and o1 % 64 == 0
    assert all(live for o, n, live in offs if o < L.n_live) and not any(live for o, n, live in offs if o >= L.n_live)


def test_module_state_dict_is_reference_compatible(golden):
    from sdumc_b200.model import WengnetMOSEIMultViewsTextMissing, get_models
    from oracle import sdumc_oracle as O
    args = types.SimpleNamespace(input_dims=(1024, 4096, 1024, 4096), model="wengnet_mosei_mult_views_text_missing")
    wrap = get_models(args)
    assert args.dim == 1024
    assert list(wrap.state_dict().keys()) == ["model." + n for n in golden["spec_names"]]
    net = wrap.model
    P = O.init_params(O.S0_DIMS)
    net.load_state_dict(P, strict=True)
    for k, v in net.state_dict().items():
        assert torch.equal(v, P[k].float())
    assert net._flat_is_current() and len(list(net.buffers())) == 0
    # optimizers update parameters in place: the flat buffer the kernels read must follow
    p = dict(net.named_parameters())["fc_att.bias"]
    with torch.no_grad():
        p.add_(1.0)
    assert torch.equal(net.layout.view(net._flat, "fc_att.bias"), p.detach())


def test_dropout_site_table_matches_reference_call_order():
    from sdumc_b200.engine import dropout_site_names, site_id
    from oracle import sdumc_oracle as O
    assert dropout_site_names() == [n for n, _ in O.dropout_sites()]
    ids = {site_id(n, p) for n in dropout_site_names() for p in (0, 1)}
    assert len(ids) == 70 and 0 not in ids


def test_no_cpu_fallback():
    from sdumc_b200 import _lib
    from sdumc_b200.losses import MSELoss
    from sdumc_b200.model import WengnetMOSEIMultViewsTextMissing
    from sdumc_b200.trainer import Trainer
    net = WengnetMOSEIMultViewsTextMissing(types.SimpleNamespace(input_dims=(64, 64, 64, 64)))
    x = torch.zeros(2, 3, 64)
    with pytest.raises(_lib.SdumcError):
        net([x, x, x, False])
    with pytest.raises(_lib.SdumcError):
        MSELoss()(torch.zeros(3, 1), torch.zeros(3))
    with pytest.raises(_lib.SdumcError):
        Trainer((64, 64, 64, 64), 2, (3, 3, 3, 3), "cpu")


def test_philox_host_reimplementation_is_self_consistent():
    """numpy Philox4x32-10 against the published Random123 known-answer vectors (the kernels' RNG)."""
    import numpy as np
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85

    def philox(c, k):
        c = [int(x) for x in c]
        k = [int(x) for x in k]
        for _ in range(10):
            p0, p1 = M0 * c[0], M1 * c[2]
            c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
            k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
        return c
    assert philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
