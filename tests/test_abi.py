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
    assert lib.scat_abi_version() == _lib.ABI_VERSION


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
    """The shipped SASS is sm_100a and the tensor-core kernels really are tcgen05 / TMEM / TMA kernels: per kernel, the
    SASS mnemonics of B200_PROFILING.md (UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA load / store)
    are counted from `cuobjdump -sass` (tools/sass_summary.py; profiles/sass_summary.txt is a committed run)."""
    import shutil
    import subprocess
    import sys
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import sass_summary
    rows = sass_summary.summarize(lib_path)
    names = sass_summary.demangle(list(rows))
    by = {}
    for nm, c in zip(names, rows.values()):
        by.setdefault(nm.split("<")[0], []).append(c)
    assert len(by["gemm_tc_kernel"]) == 32                              # bf16 / tf32 x BN 64 / 128 / 192 / 256 x operand majorness
    for fam, need in (("gemm_tc_kernel", ("UTC*MMA", "UTMALDG", "LDTM")), ("conv_fwd_tc_kernel", ("UTC*MMA", "UTMALDG", "LDTM")),
                      ("conv_wgrad_tc_kernel", ("UTC*MMA", "UTMALDG", "LDTM")),
                      ("conv_dgrad_tc_kernel", ("UTC*MMA", "UTMALDG", "UTMASTG", "LDTM")),
                      ("attention_fwd_tc128_kernel", ("UTC*MMA", "UTMALDG", "LDTM"))):
        assert fam in by, fam
        for c in by[fam]:
            for k in need:
                assert c.get(k, 0) > 0, (fam, k, dict(c))
    for fam in ("attention_fwd_mma_kernel", "attention_bwd_mma_kernel"):    # n = 21: warp-level mma.sync, by design
        assert by[fam][0].get("HMMA", 0) > 0 and by[fam][0].get("UTC*MMA", 0) == 0


def test_new_entry_points_validate_arguments_without_a_gpu(lib_path):
    """Optimiser, peer all-reduce and evaluation-metric entry points reject bad arguments before touching CUDA
    (negative SCAT_ERR_* code + message), so the checks run on a CPU-only box."""
    from scat_b200 import _lib
    lib = _lib.load()
    fake = ctypes.c_void_p(0x1000)                      # never dereferenced: validation fails first
    assert lib.scat_adam_step(None, fake, fake, fake, 16, 1e-4, 0.9, 0.999, 1e-8, 0.0, 1, None, None, None, None) == -1
    assert lib.scat_adam_step(fake, fake, fake, fake, 16, 1e-4, 0.9, 0.999, 1e-8, 0.0, 0, None, None, None, None) == -1
    assert b"step" in lib.scat_last_error_string()
    assert lib.scat_adam_step(fake, fake, fake, fake, 16, 1e-4, 1.0, 0.999, 1e-8, 0.0, 1, None, None, None, None) == -1
    misaligned = ctypes.c_void_p(0x1004)
    assert lib.scat_adam_step(misaligned, fake, fake, fake, 16, 1e-4, 0.9, 0.999, 1e-8, 0.0, 1, None, None, None, None) == -1
    assert b"aligned" in lib.scat_last_error_string()

    ptrs = (ctypes.c_void_p * 3)(0x1000, 0x2000, 0x3000)
    assert lib.scat_peer_allreduce(ptrs, ptrs, 0, 3, 0, 64, None) == -3          # 1, 2, 4 or 8 ranks
    assert lib.scat_peer_allreduce(ptrs, ptrs, 0, 2, 2, 64, None) == -1          # range not 16-byte aligned
    assert lib.scat_peer_allreduce(ptrs, ptrs, 2, 2, 0, 64, None) == -1          # rank out of range
    assert lib.scat_peer_allreduce(ptrs, ptrs, 0, 9, 0, 64, None) == -1
    assert lib.scat_peer_allreduce_part(ptrs, ptrs, 0, 2, 2, 64, 0, None) == -1  # the per-part form checks the same way
    assert lib.scat_peer_allreduce_part(ptrs, ptrs, 0, 3, 0, 64, 1, None) == -3
    assert lib.scat_head_train_step_hooked(None, None, None, None, None, None, None, None, 105, 1.0, 1.0, 1.0, None, None, None,
                                           None, None, None, None, None, 0, None, None, None) == -1        # desc is null
    assert lib.scat_peer_signal_bytes() >= 296 * 8 * 4

    assert lib.scat_eval_procrustes(fake, fake, 4, 2, fake, None, None) == -1     # fewer than 3 joints
    assert lib.scat_eval_procrustes(fake, None, 4, 21, fake, None, None) == -1
    thr = (ctypes.c_double * 65)()
    assert lib.scat_eval_joint_errors(fake, fake, 4, 21, 1000.0, thr, 65, fake, None, None) == -1
    assert b"thresholds" in lib.scat_last_error_string()
    assert lib.scat_eval_accel(fake, None, 2, 21, fake, None) == -1               # needs 3 frames
    assert lib.scat_head_train_step_phase.argtypes[-1] is ctypes.c_int32
