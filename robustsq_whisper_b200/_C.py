"""ctypes binding of libtsw_sm100.so (the C ABI declared in include/tsw.h).

The library is built in-tree by ``robustsq_whisper_b200/csrc/build.sh`` (``__graft_entry__.build()``).  There is no
CPU or PyTorch fallback: if the shared object is missing, or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtsw_sm100.so")

F32, BF16 = 0, 1
ABI_VERSION = 4
EPI_NONE, EPI_GELU, EPI_MUL_DGELU, EPI_GELU_SAVE_GRAD, EPI_MUL_AUX = 0, 1, 2, 3, 4
GEMM_AUTO, GEMM_SIMT, GEMM_TCGEN05, GEMM_SKINNY = 0, 1, 2, 3


class TswError(RuntimeError):
    pass


class GemmDesc(Structure):
    _fields_ = [
        ("M", c_int64), ("N", c_int64), ("K", c_int64),
        ("batch_outer", c_int32), ("batch_inner", c_int32),
        ("A", c_void_p), ("a_dtype", c_int32), ("a_mn_major", c_int32), ("lda", c_int64), ("a_stride_outer", c_int64), ("a_stride_inner", c_int64),
        ("B", c_void_p), ("b_dtype", c_int32), ("b_mn_major", c_int32), ("ldb", c_int64), ("b_stride_outer", c_int64), ("b_stride_inner", c_int64),
        ("D", c_void_p), ("d_dtype", c_int32), ("reserved0", c_int32), ("ldd", c_int64), ("d_stride_outer", c_int64), ("d_stride_inner", c_int64),
        ("bias", c_void_p),
        ("residual", c_void_p), ("res_dtype", c_int32), ("reserved1", c_int32), ("ldres", c_int64), ("res_stride_outer", c_int64), ("res_stride_inner", c_int64),
        ("res_row_mod", c_int64),
        ("aux_in", c_void_p), ("aux_out", c_void_p),
        ("epilogue", c_int32), ("impl", c_int32),
        ("alpha", c_float), ("beta", c_float),
        ("alpha_dev", c_void_p),
        ("A2", c_void_p), ("B2", c_void_p), ("K2", c_int64), ("lda2", c_int64), ("ldb2", c_int64),
        ("colsum_out", c_void_p),
        ("kgroups", c_int32), ("reserved2", c_int32),
        ("a_outer_step", c_int64), ("a_outer_off0", c_int64), ("a_outer_off_step", c_int64), ("a_outer_extent", c_int64), ("a_group_stride", c_int64),
        ("b_outer_step", c_int64), ("b_outer_off0", c_int64), ("b_outer_off_step", c_int64), ("b_outer_extent", c_int64), ("b_group_stride", c_int64),
    ]


# name -> (restype, argtypes); every symbol include/tsw.h declares
_P, _I64, _I, _F, _SZ = c_void_p, c_int64, c_int, c_float, c_size_t
SIGNATURES = {
    "tsw_abi_version": (c_int, []),
    "tsw_last_error": (c_char_p, []),
    "tsw_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "tsw_set_sm_reserve": (c_int, [_I]),
    "tsw_set_fmha_work_list": (c_int, [_I]),
    "tsw_logmel_init": (c_int, [_P, _I, _I]),
    "tsw_logmel_workspace_bytes": (_SZ, [_I64, _I64, _I]),
    "tsw_logmel_fwd": (c_int, [_P, _I64, _I64, _I64, _P, _I, _P, _SZ, _P]),
    "tsw_logmel_gather_fwd": (c_int, [_P, _P, _P, _I64, _I64, _P, _I, _P, _SZ, _P]),
    "tsw_dropout": (c_int, [_P, _P, _I, _I64, _F, ctypes.c_uint64, ctypes.c_uint64, _P]),
    "tsw_specaug_fwd": (c_int, [_P, _P, _I, _I64, _I64, _I64, _I64, _P, _P, _I, _P, _I, _I, _P]),
    "tsw_gemm_workspace_bytes": (_SZ, [POINTER(GemmDesc)]),
    "tsw_gemm": (c_int, [POINTER(GemmDesc), _P, _SZ, _P]),
    "tsw_layernorm_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _F, _I, _P]),
    "tsw_layernorm_bwd_workspace_bytes": (_SZ, [_I64, _I64]),
    "tsw_layernorm_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I, _P, _SZ, _P]),
    "tsw_cast": (c_int, [_P, _I, _P, _I, _I64, _P]),
    "tsw_colsum_workspace_bytes": (_SZ, [_I64, _I64]),
    "tsw_colsum": (c_int, [_P, _I, _I64, _I64, _I64, _P, _P, _SZ, _P]),
    "tsw_scale": (c_int, [_P, _P, _I, _I64, _F, _P, _P]),
    "tsw_add": (c_int, [_P, _P, _P, _I, _I64, _P]),
    "tsw_gelu_fwd": (c_int, [_P, _P, _I, _I64, _P]),
    "tsw_gelu_bwd": (c_int, [_P, _P, _P, _I, _I64, _P]),
    "tsw_im2col_k3": (c_int, [_P, _I, _I, _I64, _I64, _I64, _I, _P, _P]),
    "tsw_col2im_k3": (c_int, [_P, _I, _I64, _I64, _I64, _I, _P, _P]),
    "tsw_softmax_fwd": (c_int, [_P, _P, _I, _I64, _I64, _I64, _I64, _I64, _F, _P, _I, _P]),
    "tsw_softmax_bwd": (c_int, [_P, _P, _P, _I, _I64, _I64, _I64, _F, _P]),
    "tsw_fmha_fwd": (c_int, [_P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I64, _I64, _I64, _F, _P, _I, _P]),
    "tsw_fmha_bwd_workspace_bytes": (_SZ, [_I64, _I64, _I64]),
    "tsw_fmha_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I64, _I64, _I64, _I64, _F, _P, _I, _P, _P, _P, _SZ, _P]),
    "tsw_decode_attention": (c_int, [_P, _I64, _P, _P, _I64, _I64, _I64, _I64, _I64, _P, _F, _P, _I64, _I, _P, _P, _I64, _P]),
    "tsw_decoder_embed": (c_int, [_P, _P, _P, _I, _P, _I64, _I64, _I64, _I64, _I64, _P, _I, _P]),
    "tsw_decoder_embed_bwd": (c_int, [_P, _I, _P, _I64, _I64, _I64, _I64, _I64, _P, _P, _P, _I, _P]),
    "tsw_asp_pool_fwd": (c_int, [_P, _I, _I64, _I64, _I64, _F, _P, _P, _P, _P, _P]),
    "tsw_asp_pool_bwd": (c_int, [_P, _I, _I64, _I64, _I64, _F, _P, _P, _P, _P, _P, _P, _P]),
    "tsw_l2norm_fwd": (c_int, [_P, _P, _P, _I64, _I64, _F, _P]),
    "tsw_l2norm_bwd": (c_int, [_P, _P, _P, _P, _I64, _I64, _F, _P]),
    "tsw_aam_workspace_bytes": (_SZ, [_I64, _I64, _I64]),
    "tsw_aam_softmax_fwd_bwd": (c_int, [_P, _P, _P, _I64, _I64, _I64, _F, _F, _P, _P, _P, _P, _P, _SZ, _P]),
    "tsw_infonce_workspace_bytes": (_SZ, [_I64, _I64, _I64]),
    "tsw_arc_infonce_fwd_bwd": (c_int, [_P, _I, _I64, _I64, _I64, _P, _I64, _P, _P, _I64, _F, _F, _P, _P, _P, _P, _P, _SZ, _P]),
    "tsw_lsce_fwd_bwd": (c_int, [_P, _I, _I64, _I64, _I64, _P, _I64, _F, _F, _P, _P, _P, _I, _I64, _P]),
    "tsw_log_softmax": (c_int, [_P, _I, _I64, _I64, _I64, _P, _P]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library (once) and attach the argtypes. Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise TswError(
            f"{LIB_PATH} is missing: build it with robustsq_whisper_b200/csrc/build.sh "
            "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU / PyTorch fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.tsw_abi_version() != ABI_VERSION:
        raise TswError(f"ABI version mismatch: library reports {lib.tsw_abi_version()}")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().tsw_last_error().decode(errors="replace")
        raise TswError(f"{what} failed (code {rc}): {msg}")


def dtype_code(t: torch.dtype) -> int:
    if t == torch.float32:
        return F32
    if t == torch.bfloat16:
        return BF16
    raise TswError(f"unsupported dtype {t}: the sm_100a kernels take float32 or bfloat16")


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise TswError("tensor is not on a CUDA device: the B200 path has no CPU fallback")
