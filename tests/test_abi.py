"""CPU: the C-ABI library builds, loads and exports exactly what include/scat_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "scat_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(scat_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    from scat_b200 import build
    return build.build()


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("scat_head_forward", "scat_head_backward", "scat_head_train_step", "scat_proj_loss", "scat_gemm",
                 "scat_conv_pe_mask_fwd", "scat_conv_bwd", "scat_layernorm_fwd", "scat_layernorm_bwd",
                 "scat_attention_fwd", "scat_attention_bwd", "scat_regressor_fwd", "scat_lbs_fwd",
                 "scat_tokens_forward", "scat_head_workspace_bytes", "scat_last_error_string"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/scat_b200.h but not exported"


def test_ctypes_binding_covers_header(lib_path):
    from scat_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.scat_abi_version() == 1


def test_workspace_query_and_argument_errors_need_no_gpu(lib_path):
    from scat_b200 import _lib
    from scat_b200.functional import HeadConfig, workspace_bytes
    n96 = workspace_bytes(HeadConfig(n_masked=4, pl_reg=True), 96)
    n8 = workspace_bytes(HeadConfig(n_masked=4, pl_reg=True), 8)
    assert 0 < n8 < n96 < 400 * 2 ** 20
    lib = _lib.load()
    bad = HeadConfig(token_dim=783).desc(4)
    assert lib.scat_head_workspace_bytes(ctypes.byref(bad)) == 0
    assert b"token_dim" in lib.scat_last_error_string()


def test_kernels_are_blackwell_native(lib_path):
    """The shipped SASS is sm_100a; once the tensor-core GEMM is in, it must contain tcgen05 (UTC*MMA) + TMA."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out
