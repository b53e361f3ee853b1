"""ctypes binding of ``libsisr_b200.so`` (the C ABI declared in ``include/sisr_b200.h``).

The prototypes are parsed from the header itself so that the binding cannot drift from the ABI.
There is no fallback of any kind: if the library is missing or a call fails, an exception is
raised (the product path must never run on a CPU or library substitute).
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), "include", "sisr_b200.h")
LIB_PATH = os.path.join(HERE, "libsisr_b200.so")


class ConvDesc(ctypes.Structure):
    """struct sisr_conv_desc"""
    _fields_ = [(n, ctypes.c_int) for n in
                ("n", "h", "w", "cin", "oh", "ow", "cout", "k", "stride", "pad", "ps_r")]


class SisrError(RuntimeError):
    pass


def _ctype_of(decl: str):
    decl = decl.strip()
    if "*" in decl:
        return ctypes.c_void_p
    base = decl.rsplit(" ", 1)[0].replace("const", "").strip() if " " in decl else decl
    return {"int": ctypes.c_int, "float": ctypes.c_float, "long long": ctypes.c_longlong,
            "size_t": ctypes.c_size_t, "void": None}[base]


def parse_header(path: str = HEADER) -> Dict[str, Tuple[object, List[object]]]:
    """{function name: (restype, [argtypes])} for every prototype in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(int|size_t|const char\*)\s+(sisr_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        restype = {"int": ctypes.c_int, "size_t": ctypes.c_size_t, "const char*": ctypes.c_char_p}[ret]
        arglist = [a for a in (s.strip() for s in args.split(",")) if a and a != "void"]
        protos[name] = (restype, [_ctype_of(a) for a in arglist])
    return protos


_lib = None
_protos = None


def load():
    global _lib, _protos
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SisrError(
            f"{LIB_PATH} is missing: build it with `python single-image-super-resolution_b200/build.py` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path.")
    import torch  # noqa: F401  (loads the CUDA runtime the library shares with PyTorch)
    lib = ctypes.CDLL(LIB_PATH)
    _protos = parse_header()
    for name, (restype, argtypes) in _protos.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def _conv(a):
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    if isinstance(a, ctypes.Structure):
        return ctypes.addressof(a)
    return a


# kernels launched per ABI call (for bench.py's ``gpu_launches`` claim); default 1
_KERNELS_PER_CALL = {"sisr_sn_power_iteration": 3, "sisr_weight_grad_finish": 2, "sisr_dhead_forward": 2,
                     "sisr_dhead_backward": 3, "sisr_adam_multi": 2, "sisr_conv_wgrad": 3, "sisr_conv_wgrad_fused": 2}
LAUNCHES = [0]


def call(name: str, *args):
    """Invoke an ``int sisr_*`` entry; tensors are passed as device pointers; raises on error."""
    lib = load()
    LAUNCHES[0] += _KERNELS_PER_CALL.get(name, 1)
    rc = getattr(lib, name)(*[_conv(a) for a in args])
    if rc != 0:
        raise SisrError(f"{name}: {lib.sisr_last_error().decode()}")


def query(name: str, *args):
    """Invoke an entry that returns a value (size_t / int) instead of a status."""
    return getattr(load(), name)(*[_conv(a) for a in args])
