"""ctypes binding of libsdumc_b200.so.

The C header include/sdumc_b200.h is the single source of truth: argument structs and function
prototypes are parsed from it at import time (the library's sdumc_struct_size() cross-checks the
layouts).  There is no CPU or PyTorch fallback: if the shared object is missing every product call
raises SdumcError.
"""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libsdumc_b200.so"
HEADER_PATH = _PKG.parent / "include" / "sdumc_b200.h"

ACT_NONE, ACT_RELU, ACT_TANH = 0, 1, 2
OUT_STORE, OUT_ADD, OUT_ATOMIC = 0, 1, 2
EPI_GENERIC, EPI_INPROJ, EPI_KEYPROJ = 0, 1, 2


class SdumcError(RuntimeError):
    pass


_SCALARS = {
    "int32_t": C.c_int32, "uint32_t": C.c_uint32, "int64_t": C.c_int64, "uint64_t": C.c_uint64,
    "float": C.c_float, "int": C.c_int, "double": C.c_double,
}


def _strip_comments(text: str) -> str:
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", " ", text)


def _ctype_of(base: str, structs: dict):
    base = base.replace("const", "").strip()
    if "*" in base:
        return C.c_void_p
    if base in _SCALARS:
        return _SCALARS[base]
    if base in structs:
        return structs[base]
    raise SdumcError(f"sdumc header parser: unknown type {base!r}")


def parse_header(path: Path = HEADER_PATH):
    """-> (structs: {name: ctypes.Structure subclass}, functions: {name: (restype, [argtypes])})."""
    text = _strip_comments(path.read_text())
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    structs: dict = {}
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s*\{(.*?)\}\s*(\w+)\s*;", text, flags=re.S):
        name, body = m.group(3), m.group(2)
        fields = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            # "<type tokens> name1, name2[3], *name3"
            mm = re.match(r"^(.*?)([\w\[\]\s,\*]+)$", decl)
            # split type from declarators: the type is everything up to the last type-ish token
            toks = decl.replace("*", " * ").split()
            # find the split point: declarators start at the first identifier following the type
            # type = leading tokens among {const, known type, SDUMC_BF16, void, struct names, '*'}
            i = 0
            tparts = []
            while i < len(toks) and (toks[i] in ("const", "void", "SDUMC_BF16", "*") or toks[i] in _SCALARS
                                     or toks[i] in structs):
                tparts.append(toks[i])
                i += 1
            # a '*' collected into the type belongs to the first declarator only when several are declared;
            # the header never mixes pointer and non-pointer declarators in one statement.
            base = " ".join(tparts)
            rest = "".join(toks[i:])
            for d in rest.split(","):
                d = d.strip()
                if not d:
                    continue
                am = re.match(r"^(\w+)(?:\[(\d+)\])?$", d)
                if not am:
                    raise SdumcError(f"sdumc header parser: cannot parse declarator {d!r} in {name}")
                ct = _ctype_of(base, structs)
                if am.group(2):
                    ct = ct * int(am.group(2))
                fields.append((am.group(1), ct))
        structs[name] = type(name, (C.Structure,), {"_fields_": fields})
    text_nostruct = re.sub(r"typedef\s+struct\s+\w+\s*\{.*?\}\s*\w+\s*;", " ", text, flags=re.S)
    functions: dict = {}
    for m in re.finditer(r"(const\s+char\s*\*|uint64_t|int)\s+(sdumc_\w+)\s*\((.*?)\)\s*;", text_nostruct, flags=re.S):
        ret, fname, args = m.group(1), m.group(2), " ".join(m.group(3).split())
        restype = C.c_char_p if "char" in ret else (C.c_uint64 if ret == "uint64_t" else C.c_int)
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(C.c_void_p)
                else:
                    argtypes.append(_ctype_of(a.rsplit(" ", 1)[0], structs))
        functions[fname] = (restype, argtypes)
    return structs, functions


STRUCTS, FUNCTIONS = parse_header()
GemmDesc = STRUCTS["sdumc_gemm_desc"]
DropKey = STRUCTS["sdumc_dropkey"]

# order used by sdumc_struct_size()
STRUCT_ORDER = ["sdumc_gemm_desc", "sdumc_pool_fwd_args", "sdumc_attn_bwd_args", "sdumc_act_bwd_args",
                "sdumc_gate_fwd_args", "sdumc_gate_bwd_args", "sdumc_weight_fwd_args", "sdumc_weight_bwd_args",
                "sdumc_final_fwd_args", "sdumc_final_bwd_args", "sdumc_loss_sums_args", "sdumc_loss_finish_args",
                "sdumc_rnc_args", "sdumc_adam_args"]

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
    for name, (restype, argtypes) in FUNCTIONS.items():
        fn = getattr(L, name)  # AttributeError here = header/library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    for i, sname in enumerate(STRUCT_ORDER):
        got, want = C.sizeof(STRUCTS[sname]), L.sdumc_struct_size(i)
        if got != want:
            raise SdumcError(f"ABI mismatch: sizeof({sname}) python={got} library={want}")
    _lib = L
    return L


def exported_symbols():
    return list(FUNCTIONS.keys())


KERNEL_LAUNCHES = [0]   # kernels launched through this binding (bench.py reports it per step)


def check(rc: int, what: str = "", launches: int = 1) -> None:
    KERNEL_LAUNCHES[0] += launches
    if rc != 0:
        msg = lib().sdumc_last_error()
        raise SdumcError(f"{what or 'sdumc call'} failed (rc={rc}): {msg.decode() if msg else ''}")


def ptr(t) -> int:
    """Device pointer of a torch tensor (0 for None)."""
    return 0 if t is None else t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def dropkey(seed: int, step: int, step_dev=None):
    return DropKey(seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF, step & 0xFFFFFFFF, 0, ptr(step_dev))


def call(name: str, args_struct, what: str = "", launches: int = 1) -> None:
    """Invoke `int name(const struct*, void* stream)` on the current stream."""
    check(getattr(lib(), name)(C.byref(args_struct), current_stream()), what or name, launches)
