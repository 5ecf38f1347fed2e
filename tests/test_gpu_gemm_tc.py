"""GPU: the tcgen05/TMEM/TMA TF32 GEMM against float64, all operand layouts and epilogues.

TF32 keeps 10 mantissa bits (the tensor core truncates the fp32 operands), so a K-long dot product of O(1)
values carries ~1e-3 relative error w.r.t. the largest output: tolerance 3e-3 of max|ref|, and the result must
agree with the fp32 FFMA kernel of the same library to the same bound."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = 3e-3


def _mk(M, N, K, a, b, seed=1):
    g = np.random.Generator(np.random.PCG64(seed))
    A = torch.from_numpy(g.standard_normal((M, K)).astype(np.float32))
    B = torch.from_numpy(g.standard_normal((N, K)).astype(np.float32))
    Ad = A.cuda() if a == "k" else A.t().contiguous().cuda()
    Bd = B.cuda() if b == "k" else B.t().contiguous().cuda()
    return A, B, Ad, Bd, ((K, 1) if a == "k" else (1, M)), ((K, 1) if b == "k" else (1, N))


def _err(x, ref):
    return float((x.double().cpu() - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("M,N,K,a,b", [
    (128, 64, 32, "k", "k"), (2016, 1536, 784, "k", "k"), (2016, 392, 588, "k", "k"), (2016, 588, 784, "k", "k"),
    (2016, 784, 512, "k", "mn"), (2016, 392, 1536, "k", "mn"), (2016, 588, 392, "k", "mn"),
    (1536, 784, 2016, "mn", "mn"), (588, 784, 2016, "mn", "mn"), (196, 294, 2016, "mn", "mn"),
    (300, 200, 100, "mn", "k"), (1, 8, 8, "k", "k"), (129, 65, 33, "k", "k"), (4032, 1536, 196, "k", "k"),
    # single-wave launches on 192- / 256-wide tiles: the stacked dgrads of layer 0 (B operand MN-major) and a ragged N
    (4032, 784, 1536, "k", "mn"), (4032, 588, 392, "k", "mn"), (4032, 784, 588, "k", "mn"), (2016, 1000, 64, "k", "k"),
])
def test_tc_gemm_layouts(M, N, K, a, b):
    from scat_b200 import functional as SF
    A, B, Ad, Bd, sa, sb = _mk(M, N, K, a, b)
    if a == "mn" and M % 4 or b == "mn" and N % 4 or a == "k" and K % 4 or b == "k" and K % 4:
        with pytest.raises(RuntimeError, match="TMA"):          # not expressible as a tensor map: refused, no fallback
            SF.gemm(Ad, Bd, a_strides=sa, b_strides=sb, m=M, n=N, k=K, precision="tf32")
        return
    ref = A.double() @ B.double().t()
    out = SF.gemm(Ad, Bd, a_strides=sa, b_strides=sb, m=M, n=N, k=K, precision="tf32")
    assert _err(out, ref) < TOL
    simt = SF.gemm(Ad, Bd, a_strides=sa, b_strides=sb, m=M, n=N, k=K, precision="fp32")
    assert _err(out, simt.double().cpu()) < TOL
    if a == "mn" and b == "mn":                                           # weight-gradient shape: split-K path
        sk = SF.gemm(Ad, Bd, a_strides=sa, b_strides=sb, m=M, n=N, k=K, precision="tf32", split_k=True)
        assert _err(sk, ref) < TOL


def test_tc_gemm_epilogues_and_padded_ld():
    from scat_b200 import functional as SF
    M, N, K = 2016, 294, 392                       # fc1 of layer 1: hidden 294 lives in a ld=296 buffer
    A, B, Ad, Bd, sa, sb = _mk(M, N, K, "k", "k", seed=3)
    g = np.random.Generator(np.random.PCG64(9))
    bias = torch.from_numpy(g.standard_normal(N).astype(np.float32))
    res = torch.from_numpy(g.standard_normal((M, 296)).astype(np.float32))
    ref = A.double() @ B.double().t()
    out = torch.zeros(M, 296, device="cuda")
    y, z = SF.gemm(Ad, Bd, epilogue="bias_gelu", bias=bias.cuda(), precision="tf32", out=out[:, :N])
    assert _err(z[:, :N], ref + bias.double()) < TOL and _err(y[:, :N], F.gelu(ref + bias.double())) < TOL
    assert torch.all(out[:, N:] == 0)                                            # pad columns untouched
    r = SF.gemm(Ad, Bd, epilogue="bias_resid", bias=bias.cuda(), aux_in=res.cuda()[:, :N], precision="tf32")
    assert _err(r, ref + bias.double() + res[:, :N].double()) < TOL
    zz = res.double()[:, :N].requires_grad_(True)
    F.gelu(zz).sum().backward()
    d = SF.gemm(Ad, Bd, epilogue="dgelu", aux_in=res.cuda()[:, :N], precision="tf32")
    assert _err(d, ref * zz.grad) < TOL
    acc = torch.ones(M, N, device="cuda")
    SF.gemm(Ad, Bd, precision="tf32", out=acc, accumulate=True)
    assert _err(acc, ref + 1.0) < TOL


def test_tc_gemm_epilogues_on_wide_tiles():
    """The 192-wide single-wave launch of the stacked dZ GEMM of layer 0 (4032 x 588 x 392, B MN-major, dGELU with a saved
    pre-activation of 2016 rows read modulo) and the 192-wide qkv shape with a bias."""
    from scat_b200 import functional as SF
    M, N, K = 4032, 588, 392
    A, B, Ad, Bd, sa, sb = _mk(M, N, K, "k", "mn", seed=5)
    g = np.random.Generator(np.random.PCG64(11))
    z = torch.from_numpy(g.standard_normal((M, N)).astype(np.float32))
    ref = A.double() @ B.double().t()
    zz = z.double().requires_grad_(True)
    F.gelu(zz).sum().backward()
    d = SF.gemm(Ad, Bd, a_strides=sa, b_strides=sb, m=M, n=N, k=K, epilogue="dgelu", aux_in=z.cuda(), precision="tf32")
    assert _err(d, ref * zz.grad) < TOL
    M, N, K = 2016, 1536, 784
    A, B, Ad, Bd, sa, sb = _mk(M, N, K, "k", "k", seed=6)
    bias = torch.from_numpy(g.standard_normal(N).astype(np.float32))
    y = SF.gemm(Ad, Bd, epilogue="bias", bias=bias.cuda(), precision="tf32")
    assert _err(y, A.double() @ B.double().t() + bias.double()) < TOL


def test_tc_gemm_is_linear_and_exact_on_tf32_representable_inputs():
    """Size-independent property: with inputs that are exactly representable in TF32 (small integers) the
    tensor-core result is bit-exact against integer arithmetic."""
    from scat_b200 import functional as SF
    g = np.random.Generator(np.random.PCG64(4))
    M, N, K = 2016, 1536, 784
    A = torch.from_numpy(g.integers(-8, 9, (M, K)).astype(np.float32))
    B = torch.from_numpy(g.integers(-8, 9, (N, K)).astype(np.float32))
    ref = (A.double() @ B.double().t()).float()
    out = SF.gemm(A.cuda(), B.cuda(), precision="tf32")
    assert torch.equal(out.cpu(), ref)


# ---- bf16 operands (tcgen05 kind::f16) -------------------------------------------------------------------------
def _mk16(M, N, K, a, b, seed=1):
    """bf16 operands with 16-byte aligned leading dimensions (the head pads 588 -> 592, 196 -> 200, ...)."""
    g = np.random.Generator(np.random.PCG64(seed))
    A = torch.from_numpy(g.standard_normal((M, K)).astype(np.float32)).bfloat16()
    B = torch.from_numpy(g.standard_normal((N, K)).astype(np.float32)).bfloat16()
    pad = lambda v: (v + 7) // 8 * 8

    def place(X, major):                      # [rows, K] logical -> device buffer + (row stride, k stride)
        rows = X.shape[0]
        if major == "k":
            buf = torch.zeros(rows, pad(K), dtype=torch.bfloat16)
            buf[:, :K] = X
            return buf.cuda(), (pad(K), 1)
        buf = torch.zeros(K, pad(rows), dtype=torch.bfloat16)
        buf[:, :rows] = X.t()
        return buf.cuda(), (1, pad(rows))
    Ad, sa = place(A, a)
    Bd, sb = place(B, b)
    return A, B, Ad, Bd, sa, sb


@pytest.mark.parametrize("M,N,K,a,b", [
    (128, 64, 64, "k", "k"), (2016, 1536, 784, "k", "k"), (2016, 392, 588, "k", "k"), (2016, 196, 294, "k", "k"),
    (2016, 784, 512, "k", "mn"), (4032, 392, 1536, "k", "mn"), (2016, 588, 392, "k", "mn"), (2016, 294, 196, "k", "mn"),
    (1536, 784, 2016, "mn", "mn"), (588, 784, 2016, "mn", "mn"), (196, 294, 2016, "mn", "mn"), (392, 512, 2016, "mn", "mn"),
    (300, 200, 100, "mn", "k"), (1, 8, 8, "k", "k"), (129, 65, 33, "k", "k"),
    (4032, 784, 1536, "k", "mn"), (4032, 588, 392, "k", "mn"), (2016, 1000, 64, "k", "k"),      # 256- / 192-wide tiles
])
def test_bf16_gemm_layouts(M, N, K, a, b):
    """bf16 x bf16 products are exact in fp32, so the only error is fp32 accumulation order: tight tolerance."""
    from scat_b200 import functional as SF
    A, B, Ad, Bd, sa, sb = _mk16(M, N, K, a, b)
    ref = A.double() @ B.double().t()
    out, out16 = SF.gemm_bf16(Ad, Bd, a_strides=sa, b_strides=sb, m=M, n=N, k=K, want16=True)
    assert _err(out, ref) < 2e-5
    assert _err(out16[:, :N].float(), ref) < 6e-3                         # bf16 copy of the output: 8 mantissa bits
    if a == "mn" and b == "mn":                                           # weight-gradient shape: split-K path
        sk = SF.gemm_bf16(Ad, Bd, a_strides=sa, b_strides=sb, m=M, n=N, k=K, split_k=True)
        assert _err(sk, ref) < 2e-5


def test_bf16_gemm_epilogues():
    from scat_b200 import functional as SF
    M, N, K = 2016, 294, 392
    A, B, Ad, Bd, sa, sb = _mk16(M, N, K, "k", "k", seed=3)
    g = np.random.Generator(np.random.PCG64(9))
    bias = torch.from_numpy(g.standard_normal(N).astype(np.float32))
    res = torch.from_numpy(g.standard_normal((M, 296)).astype(np.float32))
    ref = A.double() @ B.double().t()
    out = torch.zeros(M, 296, device="cuda")
    y, z, y16 = SF.gemm_bf16(Ad, Bd, a_strides=sa, b_strides=sb, m=M, n=N, k=K, epilogue="bias_gelu", bias=bias.cuda(),
                             out=out[:, :N], want16=True)
    assert _err(z[:, :N], ref + bias.double()) < 2e-5 and _err(y[:, :N], F.gelu(ref + bias.double())) < 2e-5
    assert _err(y16[:, :N].float(), F.gelu(ref + bias.double())) < 6e-3
    assert torch.all(out[:, N:] == 0)
    r = SF.gemm_bf16(Ad, Bd, a_strides=sa, b_strides=sb, m=M, n=N, k=K, epilogue="bias_resid", bias=bias.cuda(),
                     aux_in=res.cuda()[:, :N])
    assert _err(r, ref + bias.double() + res[:, :N].double()) < 2e-5
    zz = res.double()[:, :N].requires_grad_(True)
    F.gelu(zz).sum().backward()
    d = SF.gemm_bf16(Ad, Bd, a_strides=sa, b_strides=sb, m=M, n=N, k=K, epilogue="dgelu", aux_in=res.cuda()[:, :N])
    assert _err(d, ref * zz.grad) < 2e-5


def test_bf16_gemm_exact_on_small_integers():
    from scat_b200 import functional as SF
    g = np.random.Generator(np.random.PCG64(4))
    M, N, K = 2016, 1536, 784
    A = torch.from_numpy(g.integers(-8, 9, (M, K)).astype(np.float32))
    B = torch.from_numpy(g.integers(-8, 9, (N, K)).astype(np.float32))
    ref = (A.double() @ B.double().t()).float()
    out = SF.gemm_bf16(A.bfloat16().cuda(), B.bfloat16().cuda())
    assert torch.equal(out.cpu(), ref)


# ---- 3xTF32 on mma.sync: the fp32-grade small-GEMM kernel ------------------------------------------------------
@pytest.mark.parametrize("M,N,K,a,b", [
    (2016, 147, 196, "k", "k"), (2016, 3, 147, "k", "k"), (4032, 147, 3, "k", "mn"), (4032, 196, 147, "k", "mn"),
    (147, 196, 2016, "mn", "mn"), (3, 147, 2016, "mn", "mn"), (66, 1024, 96, "mn", "mn"), (66, 66, 288, "mn", "mn"),
    (1, 1, 1, "k", "k"), (33, 65, 9, "k", "k"), (129, 31, 1000, "mn", "k"),
])
def test_tf32x3_gemm_is_fp32_grade(M, N, K, a, b):
    from scat_b200 import functional as SF
    A, B, Ad, Bd, sa, sb = _mk(M, N, K, a, b)
    ref = A.double() @ B.double().t()
    out = SF.gemm(Ad, Bd, a_strides=sa, b_strides=sb, m=M, n=N, k=K, precision="tf32x3")
    assert _err(out, ref) < 2e-6
    simt = SF.gemm(Ad, Bd, a_strides=sa, b_strides=sb, m=M, n=N, k=K, precision="fp32")
    assert _err(out, simt.double().cpu()) < 2e-6


def test_tf32x3_gemm_epilogues():
    from scat_b200 import functional as SF
    M, N, K = 2016, 147, 196
    A, B, Ad, Bd, sa, sb = _mk(M, N, K, "k", "k", seed=3)
    g = np.random.Generator(np.random.PCG64(9))
    bias = torch.from_numpy(g.standard_normal(N).astype(np.float32))
    res = torch.from_numpy(g.standard_normal((M, 148)).astype(np.float32))
    ref = A.double() @ B.double().t()
    out = torch.zeros(M, 148, device="cuda")
    y, z = SF.gemm(Ad, Bd, epilogue="bias_gelu", bias=bias.cuda(), precision="tf32x3", out=out[:, :N])
    assert _err(z[:, :N], ref + bias.double()) < 2e-6 and _err(y[:, :N], F.gelu(ref + bias.double())) < 2e-6
    assert torch.all(out[:, N:] == 0)
    r = SF.gemm(Ad, Bd, epilogue="bias_resid", bias=bias.cuda(), aux_in=res.cuda()[:, :N], precision="tf32x3")
    assert _err(r, ref + bias.double() + res[:, :N].double()) < 2e-6
    zz = res.double()[:, :N].requires_grad_(True)
    F.gelu(zz).sum().backward()
    d = SF.gemm(Ad, Bd, epilogue="dgelu", aux_in=res.cuda()[:, :N], precision="tf32x3")
    assert _err(d, ref * zz.grad) < 2e-6
    acc = torch.ones(M, N, device="cuda")
    SF.gemm(Ad, Bd, precision="tf32x3", out=acc, accumulate=True)
    assert _err(acc, ref + 1.0) < 2e-6
