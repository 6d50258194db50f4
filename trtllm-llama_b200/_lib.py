"""ctypes binding of the C ABI declared in include/trtllm_b200.h.

The product path has NO fallback: if the shared library is missing or a symbol cannot be
resolved this raises, and every wrapper raises on a non-zero return code."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtrtllm_llama_b200.so")


class LibraryNotBuilt(RuntimeError):
    pass


class KernelError(RuntimeError):
    pass


vp, i32, i64, f32, sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes); mirrors include/trtllm_b200.h one to one
SIGNATURES = {
    "tb_version": (C.c_char_p, []),
    "tb_check_device": (i32, []),
    "tb_rmsnorm": (i32, [vp, vp, vp, vp, vp, f32, i32, i32, vp]),
    "tb_rmsnorm_quant": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, f32, i32, i32, i32, i32, vp]),
    "tb_quantize_per_token": (i32, [vp, vp, vp, i32, i32, i32, vp]),
    "tb_quantize_tensor": (i32, [vp, vp, i64, vp, i32, vp]),
    "tb_gemv": (i32, [i32, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, i32, i32, i32, i32, vp]),
    "tb_gemm_tc_workspace_bytes": (sz, [i32, i32, i32]),
    "tb_gemm_tc": (i32, [i32, vp, i32, vp, vp, vp, vp, vp, i32, i32, vp, i32, i32, i32, vp, sz, i32, i32, vp]),
    "tb_mmha_workspace_bytes": (sz, [i32, i32, i32]),
    "tb_mmha_num_splits": (i32, [i32, i32, i32, i32]),
    "tb_mmha_decode": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, f32, i32,
                             i32, vp]),
    "tb_context_attention": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, f32, i32, vp]),
    "tb_embedding": (i32, [vp, vp, vp, i32, i32, i32, vp]),
    "tb_swiglu": (i32, [vp, vp, vp, i32, i32, i32, vp]),
    "tb_add": (i32, [vp, vp, vp, i64, vp]),
    "tb_gather_last_token": (i32, [vp, vp, vp, i32, i32, i32, vp]),
    "tb_argmax": (i32, [vp, vp, i32, i32, i32, vp]),
    "tb_advance_step": (i32, [vp, vp, vp, vp, vp, i32, i32, vp]),
    "tb_half_to_float": (i32, [vp, vp, i64, vp]),
}

_lib = None


def load_library(path: str | None = None):
    """dlopen the in-tree library and type every declared entry point."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise LibraryNotBuilt(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or PyTorch fallback for this path)")
    h = C.CDLL(p, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(h, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = h
    return h


class _Lazy:
    def __getattr__(self, name):
        return getattr(load_library(), name)


lib = _Lazy()


def check(rc: int, what: str):
    if rc != 0:
        raise KernelError(f"{what} failed with code {rc}")
