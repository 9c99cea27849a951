"""ctypes binding of libsdumc_b200.so (C ABI declared in include/sdumc_b200.h).

There is no CPU or PyTorch fallback: if the shared object is missing every product call raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libsdumc_b200.so"

ACT_NONE, ACT_RELU, ACT_TANH = 0, 1, 2
OUT_STORE, OUT_ADD, OUT_ATOMIC = 0, 1, 2
EPI_GENERIC, EPI_INPROJ, EPI_KEYPROJ = 0, 1, 2


class SdumcError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("a_mn", C.c_int32), ("b_mn", C.c_int32), ("tf32", C.c_int32),
        ("k_splits", C.c_int32), ("block_n", C.c_int32), ("max_ctas", C.c_int32),
        ("A", C.c_void_p), ("lda", C.c_int64), ("B", C.c_void_p), ("ldb", C.c_int64),
        ("epi_kind", C.c_int32), ("act", C.c_int32),
        ("bias", C.c_void_p),
        ("gate", C.c_void_p), ("ld_gate", C.c_int64), ("gate_scale", C.c_float),
        ("drop_p", C.c_float), ("drop_site", C.c_uint32), ("fmask_site", C.c_uint32),
        ("out_f32", C.c_void_p), ("ld_f32", C.c_int64), ("f32_mode", C.c_int32),
        ("out_bf16", C.c_void_p), ("ld_bf16", C.c_int64), ("bf16_mode", C.c_int32),
        ("n_tgt", C.c_int32), ("tgt", C.c_void_p * 4), ("tgt_site", C.c_uint32 * 4),
        ("qv", C.c_void_p), ("q_stride", C.c_int64), ("nq", C.c_int32), ("L", C.c_int32),
        ("scores", C.c_void_p),
        ("seed", C.c_uint64), ("step", C.c_uint32),
        ("dbg_lbo", C.c_uint32), ("dbg_sbo", C.c_uint32),
    ]


_lib = None


def lib() -> C.CDLL:
    """Load the shared object (once).  Raises SdumcError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise SdumcError(
            f"{LIB_PATH} is missing: build it with `python -m sdumc_b200.build` "
            "(sdumc_b200 has no CPU/PyTorch fallback path)")
    L = C.CDLL(str(LIB_PATH))
    L.sdumc_version.restype = C.c_int
    L.sdumc_last_error.restype = C.c_char_p
    _declare(L)
    _lib = L
    return L


# name -> argtypes; every function returns int (0 = ok)
_SIGNATURES = {
    "sdumc_gemm": [C.POINTER(GemmDesc), C.c_void_p],
    "sdumc_frame_mask": [C.c_uint64, C.c_uint32, C.c_uint32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p],
    "sdumc_elem_mask": [C.c_uint64, C.c_uint32, C.c_uint32, C.c_int64, C.c_float, C.c_void_p, C.c_void_p],
}


def _declare(L):
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(L, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int


def exported_symbols():
    return ["sdumc_version", "sdumc_last_error", *_SIGNATURES.keys()]


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().sdumc_last_error()
        raise SdumcError(f"{what or 'sdumc call'} failed (rc={rc}): {msg.decode() if msg else ''}")


def ptr(t) -> int:
    """Device pointer of a torch tensor (or 0 for None)."""
    return 0 if t is None else t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
