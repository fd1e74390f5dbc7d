"""ctypes binding of the C-ABI shared library (include/univl_b200.h).

The library is the product's only compute backend: there is no CPU or PyTorch fallback.  `load()` raises if
the library is missing, and every wrapper raises `RuntimeError(univl_last_error_string())` on a non-zero
return code (SURVEY.md §8b error convention).
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libunivl_b200.so")

c_int = ctypes.c_int
c_ll = ctypes.c_longlong
c_ull = ctypes.c_ulonglong
c_f = ctypes.c_float
c_p = ctypes.c_void_p

# name -> argtypes; every entry returns int (0 = ok) unless listed in _SPECIAL.
SIGNATURES = {
    "univl_gemm_bf16": [c_p, c_ll, c_int, c_p, c_ll, c_int, c_int, c_int, c_int, c_p, c_ll, c_int, c_p, c_p, c_ll,
                        c_p, c_ll, c_f, c_int, c_int, c_p],
    "univl_layernorm_fwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_f, c_f, c_ull, c_ull, c_p],
    "univl_layernorm_bwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_f, c_ull,
                            c_ull, c_p],
    "univl_layernorm_f32_fwd": [c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_f, c_p],
    "univl_layernorm_f32_bwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_p],
    "univl_embed_text_fwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_p],
    "univl_embed_text_bwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_p],
    "univl_embed_cross_fwd": [c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_p],
    "univl_embed_cross_bwd": [c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_p],
    "univl_add_pos_fwd": [c_p, c_p, c_p, c_int, c_int, c_int, c_p],
    "univl_add_pos_bwd": [c_p, c_p, c_int, c_int, c_int, c_p],
    "univl_attention_fwd": [c_p, c_ll, c_p, c_ll, c_p, c_ll, c_p, c_p, c_ll, c_p, c_int, c_int, c_int, c_int, c_int,
                            c_int, c_f, c_f, c_ull, c_ull, c_p],
    "univl_attention_bwd": [c_p, c_ll, c_p, c_ll, c_p, c_ll, c_p, c_p, c_ll, c_p, c_ll, c_p, c_p, c_ll, c_p, c_ll,
                            c_p, c_ll, c_int, c_int, c_int, c_int, c_int, c_int, c_f, c_f, c_ull, c_ull, c_p],
    "univl_colsum_bf16": [c_p, c_ll, c_p, c_int, c_int, c_p],
    "univl_cast_f32_to_bf16": [c_p, c_p, c_ll, c_p],
    "univl_multi_cast_f32_to_bf16": [c_p, c_p, c_p, c_int, c_p],
    "univl_gather_rows_bf16": [c_p, c_ll, c_p, c_p, c_ll, c_int, c_int, c_p],
    "univl_scatter_add_rows_bf16": [c_p, c_ll, c_p, c_p, c_ll, c_int, c_int, c_p],
    "univl_meanpool_fwd": [c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_p],
    "univl_meanpool_bwd": [c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_p],
    "univl_sim_matmul_fwd": [c_p, c_p, c_p, c_int, c_int, c_int, c_p],
    "univl_sim_matmul_bwd": [c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_p],
    "univl_maxmargin_loss": [c_p, c_p, c_p, c_p, c_int, c_f, c_f, c_p],
    "univl_crossen_loss": [c_p, c_p, c_p, c_int, c_f, c_p],
    "univl_milnce_loss": [c_p, c_p, c_p, c_int, c_int, c_f, c_p],
    "univl_softmax_xent": [c_p, c_ll, c_p, c_p, c_p, c_p, c_ll, c_int, c_int, c_int, c_f, c_p],
    "univl_mfm_nce_loss": [c_p, c_ll, c_p, c_p, c_p, c_p, c_ll, c_int, c_f, c_p],
    "univl_pooler_tanh_fwd": [c_p, c_p, c_ll, c_p],
    "univl_pooler_tanh_bwd": [c_p, c_p, c_p, c_ll, c_p],
    "univl_bert_adam_step": [c_p, c_p, c_p, c_p, c_p, c_ll, c_f, c_f, c_f, c_f, c_f, c_f, c_p, c_p],
    "univl_grad_sqnorm": [c_p, c_ll, c_p, c_p],
    "univl_fill_f32": [c_p, c_f, c_ll, c_p],
}
_SPECIAL = {
    "univl_last_error_string": (ctypes.c_char_p, []),
    "univl_abi_version": (c_int, []),
}

_lib = None
_lock = threading.Lock()


def load(build_if_missing=False):
    """Load libunivl_b200.so and attach argtypes to every exported entry point that exists."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            if build_if_missing:
                from . import build as _build
                _build.build()
            else:
                raise RuntimeError(
                    "univl_b200: %s is missing — run `python -m univl_b200.build` (or __graft_entry__.build()); "
                    "there is no fallback compute path" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SPECIAL.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name, None)
            if fn is None:
                continue
            fn.restype = c_int
            fn.argtypes = args
        _lib = lib
        return lib


def call(name, *args):
    """Invoke a C-ABI entry point; raise on a non-zero status."""
    lib = load()
    fn = getattr(lib, name, None)
    if fn is None:
        raise RuntimeError("univl_b200: entry point %s is not exported by %s" % (name, LIB_PATH))
    rc = fn(*args)
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (name, rc, lib.univl_last_error_string().decode()))
    return rc
