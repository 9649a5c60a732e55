"""ctypes binding of include/revisionllm_b200.h.

There is no CPU fallback: if the shared library has not been built, or the process has no sm_100
GPU, the product path raises.  `python -m revisionllm_b200.build` (or `__graft_entry__.build()`)
produces the library in-tree.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "librevisionllm_b200.so")

RVL_OK = 0
GEMM_OUT_BF16, GEMM_OUT_F32, GEMM_ADD_F32 = 0, 1, 2
GEMM_FLAG_RELU, GEMM_FLAG_SWAP, GEMM_FLAG_STREAMK, GEMM_FLAG_W_CONST, GEMM_FLAG_SWIGLU = 1, 2, 4, 8, 16


class RvlError(RuntimeError):
    pass


class rvl_config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("hidden", "n_layers", "n_heads", "head_dim", "intermediate", "vocab",
                                          "adapter_dim", "max_pos", "kv_page_size", "device")] + \
               [("rms_eps", C.c_float), ("rope_theta", C.c_float)]


class rvl_layer_weights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("wqkv", "wo", "wgu", "wdown", "ln1", "ln2")]


class rvl_clip_layer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("in_proj_w", "in_proj_b", "out_proj_w", "out_proj_b", "linear1_w", "linear1_b", "linear2_w",
                                          "linear2_b", "norm1_w", "norm1_b", "norm2_w", "norm2_b")]


class rvl_clip_weights(C.Structure):
    _fields_ = [("t2v", rvl_clip_layer * 2), ("enc", rvl_clip_layer * 2), ("global_token", C.c_void_p), ("pos", C.c_void_p),
                ("pos_global", C.c_void_p), ("proj_w", C.c_void_p), ("proj_b", C.c_void_p), ("hidden", C.c_int32)]


class rvl_weights(C.Structure):
    _fields_ = [("embed_tokens", C.c_void_p), ("final_norm", C.c_void_p), ("lm_head", C.c_void_p),
                ("proj_w", C.c_void_p), ("proj_b", C.c_void_p), ("layers", C.POINTER(rvl_layer_weights)),
                ("wgu_layout", C.c_int32)]


_P, _I32, _I64, _SZ, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t, C.c_float

# name -> (restype, argtypes); mirrors the header one to one (tests check the exported symbols)
PROTOTYPES = {
    "rvl_abi_version": (C.c_int, []),
    "rvl_last_error": (C.c_char_p, [_P]),
    "rvl_create": (C.c_int, [C.POINTER(rvl_config), C.POINTER(_P)]),
    "rvl_destroy": (None, [_P]),
    "rvl_bind_weights": (C.c_int, [_P, C.POINTER(rvl_weights)]),
    "rvl_workspace_bytes": (_SZ, [_P, _I64, _I32]),
    "rvl_set_workspace": (C.c_int, [_P, _P, _SZ, _I64, _I32]),
    "rvl_kv_bytes": (_SZ, [_P, _I32]),
    "rvl_set_kv": (C.c_int, [_P, _P, _I32]),
    "rvl_project_splice": (C.c_int, [_P, _P, _P, _I32, _P, _P, _I32, _P, _I64, _P]),
    "rvl_gather_windows": (C.c_int, [_P, _P, _I32, _I32, _P, _I32, _P, _P]),
    "rvl_splice_rows": (C.c_int, [_P, _P, _P, _I32, _P, _P, _I32, _P, _I64, _P]),
    "rvl_prefill": (C.c_int, [_P, _P, _P, _I32, _I64, _I32, _P, _I32, _P, _I32, _P, _P, _P]),
    "rvl_decode_step": (C.c_int, [_P, _P, _P, _I32, _P, _I32, _I32, _P, _P]),
    "rvl_decode_n": (C.c_int, [_P, _I32, _P, _P, _P, _P, _I32, _I32, _P, _I32, _P, _I32, _I32, _P]),
    "rvl_sample_greedy": (C.c_int, [_P, _P, _I32, _I32, _P, _I32, _I32, _P, _P, _P]),
    "rvl_sample_multinomial": (C.c_int, [_P, _P, _I32, _I32, _F, C.c_uint64, C.c_uint32, _P, _I32, _I32, _P, _P, _P, _P]),
    "rvl_cosine_topk": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _P, _I32, _I32, _I32, _P, _P, _P]),
    "rvl_select_topk": (C.c_int, [_P, _P, _I32, _I32, _P, _P]),
    "rvl_merge_rank": (C.c_int, [_P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _P, _P, _P, _P]),
    "rvl_profile_enable": (C.c_int, [_P, _I32, _I32]),
    "rvl_profile_read": (C.c_int, [_P, _I32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "rvl_debug_gemm_timestamps": (None, [C.c_int, _P, C.c_int]),
    "rvl_reload_env": (None, []),
    "rvl_debug_sm_clock": (None, [_P, _P]),
    "rvl_debug_attn_timestamps": (None, [C.c_int, _P, C.c_int]),
    "rvl_gemm_bf16": (C.c_int, [_P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I32, _I32, _P, _I32, _P]),
    "rvl_rmsnorm": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _F, _P, _P]),
    "rvl_rope_kv": (C.c_int, [_P, _P, _I64, _P, _P, _P, _P, _I32, _I32, _P]),
    "rvl_swiglu": (C.c_int, [_P, _P, _P, _I64, _I32, _P]),
    "rvl_attn_prefill": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I64, _P]),
    "rvl_attn_decode": (C.c_int, [_P, _P, _P, _P, _I32, _P, _I32, _I32, _I32, _I32, _P]),
    "rvl_clip_encoder_workspace_bytes": (_SZ, [_I32, _I32, _I32, _I32]),
    "rvl_clip_encoder": (C.c_int, [_P, C.POINTER(rvl_clip_weights), _P, _P, _P, _P, _I32, _I32, _I32, _I32, _P, _SZ, _P, _P]),
    "rvl_layernorm": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _F, _P]),
    "rvl_mha96": (C.c_int, [_P, _P, _I64, _P, _I64, _P, _I64, _P, _I64, _I32, _I32, _I32, _I32, _P, _I32, _P, _P]),
}

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RvlError(
            f"{LIB_PATH} is missing: build it with `python -m revisionllm_b200.build` "
            "(there is no CPU or PyTorch fallback for the hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error(handle=None) -> str:
    msg = load().rvl_last_error(handle)
    return msg.decode() if msg else ""


def check(rc: int, handle=None, what: str = ""):
    if rc != RVL_OK:
        raise RvlError(f"{what} failed ({rc}): {last_error(handle)}")
