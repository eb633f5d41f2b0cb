"""CPU: the C-ABI shared library loads and exports every symbol include/tsw.h declares (no compute calls here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tsw.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tsw_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    from robustsq_whisper_b200 import _C
    if not os.path.isfile(_C.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _C.LIB_PATH


def test_header_declares_the_hot_path_ops():
    names = declared_symbols()
    for op in ("tsw_logmel_fwd", "tsw_gemm", "tsw_layernorm_fwd", "tsw_layernorm_bwd", "tsw_asp_pool_fwd", "tsw_asp_pool_bwd",
               "tsw_aam_softmax_fwd_bwd", "tsw_arc_infonce_fwd_bwd", "tsw_lsce_fwd_bwd", "tsw_last_error", "tsw_abi_version"):
        assert op in names


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing
    lib.tsw_abi_version.restype = ctypes.c_int
    assert lib.tsw_abi_version() == 4


def test_ctypes_signatures_cover_the_header(lib_path):
    from robustsq_whisper_b200 import _C
    assert sorted(_C.SIGNATURES) == declared_symbols()
    _C.load()


def test_library_is_sm100a_and_uses_tcgen05_tma(lib_path):
    out = subprocess.run(["cuobjdump", "-sass", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):  # tcgen05.mma / TMA load / tcgen05.ld (B200_PROFILING.md)
        assert mnemonic in out, mnemonic


def test_missing_gpu_fails_loudly(lib_path):
    import torch
    from robustsq_whisper_b200 import _C, kernels
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_C.TswError):
        kernels.logmel(torch.zeros(1, 16000))
