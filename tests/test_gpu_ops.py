"""GPU parity of the single operators behind the C ABI against CPU restatements (oracle / torch fp64).

Tolerances: fp32 kernels are compared with float64 references at 2e-5 relative (summation-order noise of
K<=2016 fp32 dot products); index/masking work is bit-exact.
"""
import random

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import head_oracle
from scat_b200 import synth
from tests.util import rel_l2, rel_max

pytestmark = pytest.mark.gpu


def _rand(*shape, seed=0, scale=1.0):
    g = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy((scale * g.standard_normal(size=shape)).astype(np.float32))


def _gelu_grad(z):
    z = z.double().requires_grad_(True)
    F.gelu(z).sum().backward()
    return z.grad


@pytest.mark.parametrize("M,N,K", [(2016, 1536, 784), (2016, 392, 588), (42, 3, 147), (2016, 147, 196), (63, 200, 37),
                                   (1, 1, 1), (130, 70, 2016)])
def test_gemm_forward_layout_and_epilogues(M, N, K):
    from scat_b200 import functional as SF
    a, w, bias, res = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=0.1), _rand(N, seed=3), _rand(M, N, seed=4)
    ref = a.double() @ w.double().t()
    ad, wd, bd, rd = a.cuda(), w.cuda(), bias.cuda(), res.cuda()
    assert rel_max(SF.gemm(ad, wd), ref) < 2e-5
    assert rel_max(SF.gemm(ad, wd, epilogue="bias", bias=bd), ref + bias.double()) < 2e-5
    assert rel_max(SF.gemm(ad, wd, epilogue="bias_resid", bias=bd, aux_in=rd), ref + bias.double() + res.double()) < 2e-5
    assert rel_max(SF.gemm(ad, wd, epilogue="resid", aux_in=rd), ref + res.double()) < 2e-5
    y, z = SF.gemm(ad, wd, epilogue="bias_gelu", bias=bd)
    assert rel_max(z, ref + bias.double()) < 2e-5
    assert rel_max(y, F.gelu(ref + bias.double())) < 2e-5
    assert rel_max(SF.gemm(ad, wd, epilogue="dgelu", aux_in=rd), ref * _gelu_grad(res)) < 2e-5
    c0 = torch.ones(M, N, device="cuda")
    assert rel_max(SF.gemm(ad, wd, out=c0, accumulate=True), ref + 1.0) < 2e-5


@pytest.mark.parametrize("M,N,K", [(2016, 784, 512), (300, 294, 196), (21, 3, 147)])
def test_gemm_dgrad_and_wgrad_layouts(M, N, K):
    """dx = dy W (B operand N-major) and dW = dy^T x (both operands M-major) through the same kernel."""
    from scat_b200 import functional as SF
    dy, w, x = _rand(M, N, seed=5), _rand(N, K, seed=6, scale=0.1), _rand(M, K, seed=7)
    dyd, wd, xd = dy.cuda(), w.cuda(), x.cuda()
    dx = SF.gemm(dyd, wd, a_strides=(N, 1), b_strides=(1, K), m=M, n=K, k=N)
    assert rel_max(dx, dy.double() @ w.double()) < 2e-5
    dw = SF.gemm(dyd, xd, a_strides=(1, N), b_strides=(1, K), m=N, n=K, k=M)
    assert rel_max(dw, dy.double().t() @ x.double()) < 2e-5


@pytest.mark.parametrize("rows,dim", [(2016, 784), (2016, 392), (2016, 196), (7, 49), (1, 1000)])
def test_layernorm_fwd_bwd(rows, dim):
    from scat_b200 import functional as SF
    x, g, b, dy, res = (_rand(rows, dim, seed=1, scale=2.0) + 0.5, 1 + 0.1 * _rand(dim, seed=2), _rand(dim, seed=3),
                        _rand(rows, dim, seed=4), _rand(rows, dim, seed=5))
    xr = x.double().requires_grad_(True)
    gr, br = g.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = F.layer_norm(xr, (dim,), gr, br, 1e-5)
    yr.backward(dy.double())
    y, mean, rstd = SF.layernorm_fwd(x.cuda(), g.cuda(), b.cuda())
    assert rel_max(y, yr.detach()) < 5e-6
    dx, dg, db = SF.layernorm_bwd(dy.cuda(), x.cuda(), g.cuda(), mean, rstd, resid=res.cuda())
    assert rel_max(dx, xr.grad + res.double()) < 2e-5
    assert rel_max(dg, gr.grad) < 2e-5
    assert rel_max(db, br.grad) < 2e-5
    dx2, dg2, _ = SF.layernorm_bwd(dy.cuda(), x.cuda(), g.cuda(), mean, rstd, param_grads=False)
    assert dg2 is None and rel_max(dx2, xr.grad) < 2e-5


@pytest.mark.parametrize("B,n,heads", [(96, 21, 8), (3, 21, 4), (2, 128, 8), (5, 1, 2), (2, 64, 3)])
def test_attention_fwd_bwd(B, n, heads):
    from scat_b200 import functional as SF
    inner = 64 * heads
    qkv = _rand(B * n, 3 * inner, seed=11)
    d_o = _rand(B * n, inner, seed=12)
    q = qkv.double().requires_grad_(True)
    qq, kk, vv = q.view(B, n, 3 * inner).chunk(3, dim=-1)
    sp = lambda t: t.reshape(B, n, heads, 64).permute(0, 2, 1, 3)
    p_ref = (torch.matmul(sp(qq), sp(kk).transpose(-1, -2)) * 0.125).softmax(-1)
    o_ref = torch.matmul(p_ref, sp(vv)).permute(0, 2, 1, 3).reshape(B * n, inner)
    o, p = SF.attention_fwd(qkv.cuda(), B, n, heads)
    assert rel_max(p, p_ref.detach()) < 1e-5
    assert rel_max(o, o_ref.detach()) < 1e-5
    if n <= 64:
        o_ref.backward(d_o.double())
        dqkv = SF.attention_bwd(qkv.cuda(), p, d_o.cuda(), B, n, heads)
        assert rel_max(dqkv, q.grad) < 2e-5
    else:
        # n > 64 only exists on the inference-only token path (hand_net.py:193-203, the reference raises right after
        # it): there is no backward kernel for it and the entry point must say so instead of computing something
        with pytest.raises(RuntimeError, match="attention bwd: n=128 not in"):
            SF.attention_bwd(qkv.cuda(), p, d_o.cuda(), B, n, heads)


@pytest.mark.parametrize("B,heads", [(96, 8), (192, 8), (3, 4), (1, 1), (5, 3)])
def test_attention_tensor_core_fwd_bwd(B, heads):
    """n = 21 attention on mma.sync TF32 (one warp per (batch, head) problem): TF32-grade against float64."""
    from scat_b200 import functional as SF
    n, inner = 21, 64 * heads
    qkv = _rand(B * n, 3 * inner, seed=11)
    d_o = _rand(B * n, inner, seed=12)
    q = qkv.double().requires_grad_(True)
    qq, kk, vv = q.view(B, n, 3 * inner).chunk(3, dim=-1)
    sp = lambda t: t.reshape(B, n, heads, 64).permute(0, 2, 1, 3)
    p_ref = (torch.matmul(sp(qq), sp(kk).transpose(-1, -2)) * 0.125).softmax(-1)
    o_ref = torch.matmul(p_ref, sp(vv)).permute(0, 2, 1, 3).reshape(B * n, inner)
    o, p = SF.attention_fwd(qkv.cuda(), B, n, heads, tc=True)
    assert rel_max(p, p_ref.detach()) < 3e-3
    assert rel_max(o, o_ref.detach()) < 3e-3
    assert float((p.sum(-1) - 1).abs().max()) < 1e-5                  # rows of P are normalised in fp32
    o_ref.backward(d_o.double())
    dqkv = SF.attention_bwd(qkv.cuda(), p_ref.detach().float().cuda(), d_o.cuda(), B, n, heads, tc=True)
    assert rel_max(dqkv, q.grad) < 3e-3
    # against the fp32 FFMA kernels of the same library
    o32, p32 = SF.attention_fwd(qkv.cuda(), B, n, heads)
    d32 = SF.attention_bwd(qkv.cuda(), p32, d_o.cuda(), B, n, heads)
    assert rel_max(o, o32) < 3e-3 and rel_max(dqkv, d32) < 3e-3


@pytest.mark.parametrize("B,heads", [(256, 8), (3, 2), (1, 1)])
def test_attention_tcgen05_n128_fwd(B, heads):
    """n = 128 attention (config 4) as one tcgen05 tile per (batch, head) problem: TF32-grade against float64 and
    against the fp32 FFMA kernel of the same library."""
    from scat_b200 import functional as SF
    n, inner = 128, 64 * heads
    qkv = _rand(B * n, 3 * inner, seed=21)
    q = qkv.double()
    qq, kk, vv = q.view(B, n, 3 * inner).chunk(3, dim=-1)
    sp = lambda t: t.reshape(B, n, heads, 64).permute(0, 2, 1, 3)
    p_ref = (torch.matmul(sp(qq), sp(kk).transpose(-1, -2)) * 0.125).softmax(-1)
    o_ref = torch.matmul(p_ref, sp(vv)).permute(0, 2, 1, 3).reshape(B * n, inner)
    o, _ = SF.attention_fwd(qkv.cuda(), B, n, heads, tc=True)
    assert rel_max(o, o_ref) < 3e-3
    o32, _ = SF.attention_fwd(qkv.cuda(), B, n, heads)
    assert rel_max(o, o32) < 3e-3


@pytest.mark.parametrize("B,mask_rate,pos_embed", [(3, 0.2, True), (2, 0.9, True), (2, 0.0, True), (2, 0.2, False),
                                                   (5, 0.5, True)])
def test_conv_pe_mask_fwd_bwd(B, mask_rate, pos_embed):
    from scat_b200 import functional as SF
    x2, _, _ = synth.make_head_inputs(B, 3)
    W = synth.make_head_weights(8)
    cw, mt = torch.from_numpy(W["conv1x1_channel_reduction.weight"]), torch.from_numpy(W["mask_token"])
    pe = head_oracle.positional_encoding(21, 784)
    random.seed(5)
    idx = synth.mask_indices(mask_rate)
    xr = torch.from_numpy(x2).double().requires_grad_(True)
    cwr, mtr = cw.double().requires_grad_(True), mt.double().requires_grad_(True)
    fv_ref = F.conv2d(xr, cwr)
    tok_ref = fv_ref.view(B, 21, -1)
    if pos_embed:
        tok_ref = tok_ref + pe.double()
    else:
        tok_ref = tok_ref.clone()
    if idx:
        tok_ref[:, idx, :] = mtr
    idx_dev = torch.tensor(idx, dtype=torch.int32, device="cuda") if idx else None
    fv, tok = SF.conv_pe_mask_fwd(torch.from_numpy(x2).cuda(), cw.cuda().view(21, 512), pe[0].cuda(),
                                  mt.cuda().view(-1), idx_dev, pos_embed)
    keep = [t for t in range(21) if t not in idx]
    assert rel_max(fv.view(B, 21, -1)[:, keep], fv_ref.detach().view(B, 21, -1)[:, keep]) < 1e-5
    # masking / indexing is bit-exact: masked rows ARE the mask token, unmasked rows are fv (+ pe) exactly
    if idx:
        assert torch.equal(tok[:, idx, :].cpu(), mt.view(1, 1, -1).expand(B, len(idx), -1))
    if pos_embed:
        assert torch.equal(tok[:, keep].cpu(), (fv.view(B, 21, -1)[:, keep] + pe[0, keep].cuda()).cpu())
        assert rel_max(fv, fv_ref.detach()) < 1e-5
    else:
        assert tok.data_ptr() == fv.data_ptr()          # aliasing of hand_net.py:364,373
    d_tok = _rand(B, 21, 784, seed=9)
    tok_ref.backward(d_tok.double())
    x2g, wg, mg = SF.conv_bwd(d_tok.cuda(), torch.from_numpy(x2).cuda(), cw.cuda().view(21, 512), idx_dev)
    assert rel_max(x2g, xr.grad) < 2e-5
    assert rel_max(wg, cwr.grad.view(21, 512)) < 2e-5
    if idx:
        assert rel_max(mg, mtr.grad.view(-1)) < 2e-5
        assert torch.all(wg[idx] == 0)                   # masked tokens: exactly zero conv-weight gradient rows


def _seam_input(B, kind):
    """x2 >= 0 as the backbone's ReLU leaves it (resnet.py:151).  kinds: "gauss" = random low mantissa bits (what a
    truncating tensor core would shrink by 3.3e-4), "bf16" = values a bf16 / autocast backbone delivers (already
    TF32-representable: a compensation constant would BIAS them), "int" = small integers (same, post-quantisation data)."""
    x2, _, _ = synth.make_head_inputs(B, 3)
    x2 = torch.from_numpy(x2)
    if kind == "bf16":
        x2 = x2.bfloat16().float()
    elif kind == "int":
        x2 = torch.round(x2 * 4.0)
    return x2


@pytest.mark.parametrize("B,mask_rate,pos_embed,kind,seam", [
    (3, 0.2, True, "gauss", "fp32"), (2, 0.9, True, "gauss", "fp32"), (2, 0.0, True, "bf16", "fp32"),
    (2, 0.2, False, "gauss", "fp32"), (96, 0.2, True, "gauss", "fp32"), (96, 0.2, True, "bf16", "fp32"),
    (5, 0.2, True, "int", "fp32"),
    (3, 0.2, True, "bf16", "bf16"), (2, 0.2, False, "bf16", "bf16"), (96, 0.2, True, "bf16", "bf16"),
    (5, 0.5, True, "int", "bf16")])
def test_conv_tensor_core_fwd_bwd(B, mask_rate, pos_embed, kind, seam):
    """The conv front end on tcgen05 (csrc/conv_tc.cu, what the head runs in TF32 / BF16 mode) for both seam dtypes.
    fp32 seam: x2 is rounded to TF32-NEAREST in shared memory, so the error is zero-mean for every kind of input --
    Gaussian data, bf16-valued data and integers (where rounding is exact and only the TF32 weight rounds).
    bf16 seam: operands are exact on the tensor core and the weight enters as three bf16 terms: fp32-grade outputs,
    x2.grad with its final bf16 rounding.  Masking / indexing bit-exact in every case."""
    from scat_b200 import functional as SF
    x2 = _seam_input(B, kind)
    W = synth.make_head_weights(8)
    cw, mt = torch.from_numpy(W["conv1x1_channel_reduction.weight"]), torch.from_numpy(W["mask_token"])
    pe = head_oracle.positional_encoding(21, 784)
    random.seed(5)
    idx = synth.mask_indices(mask_rate)
    keep = [t for t in range(21) if t not in idx]
    xr = x2.double().requires_grad_(True)
    cwr, mtr = cw.double().requires_grad_(True), mt.double().requires_grad_(True)
    fv_ref = F.conv2d(xr, cwr)
    tok_ref = fv_ref.view(B, 21, -1) + pe.double() if pos_embed else fv_ref.view(B, 21, -1).clone()
    if idx:
        tok_ref[:, idx, :] = mtr
    idx_dev = torch.tensor(idx, dtype=torch.int32, device="cuda") if idx else None
    x2_dev = x2.cuda().bfloat16() if seam == "bf16" else x2.cuda()          # exact: kind is bf16-representable then
    fv, tok = SF.conv_pe_mask_fwd(x2_dev, cw.cuda().view(21, 512), pe[0].cuda(), mt.cuda().view(-1), idx_dev, pos_embed,
                                  tc=True)
    got, ref = fv.view(B, 21, -1)[:, keep].double().cpu(), fv_ref.detach().view(B, 21, -1)[:, keep]
    tol_out = 2e-5 if seam == "bf16" else 2e-3
    assert rel_max(got, ref) < tol_out
    big = ref.abs() > 0.5 * ref.abs().max()
    bias = float(((got - ref) / ref)[big].mean())
    assert abs(bias) < (2e-6 if seam == "bf16" else 5e-5), bias       # no systematic shrink / inflation for ANY input kind
    if idx:
        assert torch.equal(tok[:, idx, :].cpu(), mt.view(1, 1, -1).expand(B, len(idx), -1))
    if pos_embed:
        assert torch.equal(tok[:, keep].cpu(), (fv.view(B, 21, -1)[:, keep] + pe[0, keep].cuda()).cpu())
    else:
        assert tok.data_ptr() == fv.data_ptr()
    d_tok = _rand(B, 21, 784, seed=9)
    tok_ref.backward(d_tok.double())
    x2g, wg, mg = SF.conv_bwd(d_tok.cuda(), x2_dev, cw.cuda().view(21, 512), idx_dev, tc=True)
    assert x2g.dtype == x2_dev.dtype
    assert rel_max(x2g, xr.grad) < (6e-3 if seam == "bf16" else 2e-5)   # bf16: its own storage rounding (2^-9 of the value)
    assert rel_l2(x2g, xr.grad) < (3e-3 if seam == "bf16" else 1e-5)
    assert rel_max(wg, cwr.grad.view(21, 512)) < (2e-5 if seam == "bf16" else 1e-3)
    gbig = cwr.grad.view(21, 512).abs() > 0.5 * cwr.grad.abs().max()
    gbias = float(((wg.double().cpu() - cwr.grad.view(21, 512)) / cwr.grad.view(21, 512))[gbig].mean())
    assert abs(gbias) < 1e-4, gbias
    if idx:
        assert rel_max(mg, mtr.grad.view(-1)) < 2e-5
        assert torch.all(wg[idx] == 0)


@pytest.mark.parametrize("B,it", [(96, 3), (2, 1), (5, 0)])
def test_regressor_fwd(B, it):
    from scat_b200 import functional as SF
    W = synth.make_head_weights(8)
    _, mf, _ = synth.make_head_inputs(B, 1)
    feat_out = _rand(B, 63, seed=3, scale=0.05)
    mean = torch.from_numpy(synth.make_mean_params("hand"))
    w, b = torch.from_numpy(W["regressor.weight"]), torch.from_numpy(W["regressor.bias"])
    p0 = mean.double().repeat(B, 1)
    p0[:, 3:] += feat_out.double()
    ref = head_oracle.iterative_regressor(torch.from_numpy(mf).double(), p0, w.double(), b.double(), it)
    j = ref[:, 3:].view(B, 21, 3)
    ref = torch.cat([ref[:, :3], (j - j[:, 1:2]).reshape(B, 63)], dim=1)
    pred = SF.regressor_fwd(torch.from_numpy(mf).cuda(), feat_out.cuda(), mean.cuda().view(-1), w.cuda(), b.cuda(), it)
    assert rel_max(pred, ref) < 1e-5
    assert torch.all(pred[:, 6:9] == 0)


def test_h3dw_encoder_regressor():
    from scat_b200.hand_net import H3DWEncoder
    from tests.util import StubBackbone
    g = np.random.Generator(np.random.PCG64(5))
    mean = torch.from_numpy((0.1 * g.standard_normal(size=(1, 61))).astype(np.float32))
    enc = H3DWEncoder(None, mean, backbone=StubBackbone()).cuda()
    with torch.no_grad():
        for p in enc.parameters():
            p.copy_(torch.from_numpy((0.02 * g.standard_normal(size=tuple(p.shape))).astype(np.float32)))
    _, mf, _ = synth.make_head_inputs(7, 2)
    fc2, reg = enc.feat_encoder[1], enc.regressor[0]
    f_ref, p_ref = head_oracle.h3dw_regressor(torch.from_numpy(mf).double(), mean.double(), fc2.weight.detach().double().cpu(),
                                              fc2.bias.detach().double().cpu(), reg.weight.detach().double().cpu(),
                                              reg.bias.detach().double().cpu())
    with torch.no_grad():
        feat, pred = enc.forward_features(torch.from_numpy(mf).cuda())
    assert rel_max(feat, f_ref) < 1e-5
    assert rel_max(pred, p_ref) < 1e-5


@pytest.mark.parametrize("B,with_pl", [(96, True), (2, False), (1, True)])
def test_proj_loss_and_gradient(B, with_pl):
    from scat_b200 import functional as SF
    _, _, labels = synth.make_head_inputs(B, 4)
    pred = _rand(B, 66, seed=2, scale=0.05)
    pred[:, 0] += 5.0
    pl = _rand(B, 21, 28, 28, seed=3, scale=0.01) if with_pl else None
    pr = pred.double().requires_grad_(True)
    loss_ref, l3, l2, lpl = head_oracle.train_loss(pr, torch.from_numpy(labels).double(), None if pl is None else pl.double())
    loss_ref.backward()
    pd = pred.cuda().requires_grad_(True)
    loss, parts = SF.proj_loss(pd, torch.from_numpy(labels).cuda(), None if pl is None else pl.cuda())
    loss.backward()
    np.testing.assert_allclose(loss.item(), loss_ref.item(), rtol=2e-5)
    np.testing.assert_allclose(parts.cpu().numpy(), [loss_ref.item(), l3.item(), l2.item(), lpl.item()], rtol=5e-5, atol=1e-12)
    assert rel_max(pd.grad, pr.grad) < 2e-5


def test_proj_loss_wide_label_rows():
    """train.py:188-199 picks the ground truth by the row width: 105 = [63 3D | 42 2D], otherwise (FreiHAND / HO-3D)
    [61 pose | 63 3D | 42 2D].  166-wide rows must train against columns 61.. / 124.., any other width is refused."""
    from scat_b200 import functional as SF
    B = 6
    _, _, labels = synth.make_head_inputs(B, 4)
    labels = torch.from_numpy(labels)
    wide = torch.cat([_rand(B, 61, seed=8), labels], dim=1)            # pose / shape parameters in front
    pred = _rand(B, 66, seed=2, scale=0.05)
    pred[:, 0] += 5.0
    pr = pred.double().requires_grad_(True)
    loss_ref, l3, l2, _ = head_oracle.train_loss(pr, wide.double(), None)
    loss_ref.backward()
    pd = pred.cuda().requires_grad_(True)
    loss, parts = SF.proj_loss(pd, wide.cuda(), None)
    loss.backward()
    np.testing.assert_allclose(parts.cpu().numpy()[:3], [loss_ref.item(), l3.item(), l2.item()], rtol=5e-5)
    assert rel_max(pd.grad, pr.grad) < 2e-5
    loss105, _ = SF.proj_loss(pred.cuda(), labels.cuda(), None)
    np.testing.assert_allclose(loss.item(), loss105.item(), rtol=1e-6)  # same ground truth, same loss
    with pytest.raises(RuntimeError, match="105 wide .* or 166 wide"):
        SF.proj_loss(pred.cuda(), wide[:, :140].contiguous().cuda(), None)
