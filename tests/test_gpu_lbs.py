"""GPU parity of the fused MANO LBS kernel (models/mano.py:280-391) against the reference fixture, the numpy
oracle, and size-independent properties at sweep sizes (BASELINE config 5)."""
import numpy as np
import pytest
import torch

from oracle import mano_oracle
from scat_b200 import synth
from tests.util import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["tf32x3", "fp32"])
def layer(request):
    """Both forward paths: blend shapes on the tcgen05 GEMM at fp32 grade (csrc/lbs_tc.cu, the default) and the all-FFMA
    kernel (csrc/lbs.cu); the backward kernel is the same for both."""
    from scat_b200.mano import ManoLayer
    return ManoLayer(synth.make_mano_asset(), precision=request.param)


def _run(layer, rots, poses, betas):
    return layer(torch.from_numpy(rots).cuda(), torch.from_numpy(poses).cuda(), torch.from_numpy(betas).cuda()).cpu().numpy()


def test_lbs_matches_reference_fixture(layer):
    g = load_golden("mano_lbs")        # includes a zero global rotation and an all-zero local pose (Taylor branch)
    out = _run(layer, g["rots"], g["poses"], g["betas"])
    assert out.shape == g["out"].shape == (6, 799, 3)
    assert np.abs(out - g["out"]).max() < 2e-6            # metres; vertices are O(0.1 m) -> < 1e-4 relative
    assert np.abs(out - g["out_fp64_oracle"]).max() < 2e-6
    assert np.all(out[:, 1] == 0.0)                        # root joint exactly at the origin (mano.py:386-388)


@pytest.mark.parametrize("B", [1, 7, 8, 9, 1000])
def test_lbs_matches_oracle_ragged_batches(layer, B):
    rots, poses, betas = synth.make_mano_inputs(B, B)
    out = _run(layer, rots, poses, betas)
    ref = mano_oracle.rot_pose_beta_to_mesh(rots, poses, betas, synth.make_mano_asset())
    assert np.abs(out - ref).max() < 5e-6
    tips = list(synth.MANO_TIP_VERTS)
    assert np.array_equal(out[:, 16:21], out[:, [21 + t for t in tips]])   # fingertips ARE mesh vertices


def test_lbs_tensor_core_path_chunks_and_matches_ffma():
    """The tensor-core path cuts a call into L2-sized chunks (7936 samples); chunk boundaries and a deliberately small
    scratch must not change the result, and it must agree with the FFMA kernel to fp32 rounding."""
    from scat_b200 import _lib
    from scat_b200._lib import check, ptr, stream_ptr
    from scat_b200.mano import ManoLayer
    asset = synth.make_mano_asset()
    tc, ff = ManoLayer(asset, precision="tf32x3"), ManoLayer(asset, precision="fp32")
    B = 8192 + 8192 + 37
    r, p, b = (torch.from_numpy(a).cuda() for a in synth.make_mano_inputs(B, 12))
    a, c = tc(r, p, b), ff(r, p, b)
    assert float((a - c).abs().max()) < 5e-7
    lib = _lib.load()
    per = lib.scat_lbs_tc_scratch_floats(1)
    small = torch.empty(per * 100 + 5, device="cuda")                  # 100 samples per chunk
    out = torch.empty_like(a)
    check(lib.scat_lbs_fwd_tc(ptr(tc.derived), ptr(tc.table), ptr(tc.hands_mean), ptr(r), ptr(p), ptr(b), ptr(out), B, ptr(small),
                              small.numel(), stream_ptr()), "scat_lbs_fwd_tc")
    assert torch.equal(out, a)
    assert lib.scat_lbs_fwd_tc(ptr(tc.derived), ptr(tc.table), ptr(tc.hands_mean), ptr(r), ptr(p), ptr(b), ptr(out), B, ptr(small),
                               10, stream_ptr()) == -2                # scratch too small for a single sample


def test_lbs_properties_at_sweep_size(layer):
    """64k samples: finite, root at origin, and equivariance: a global rotation only rotates the output."""
    B = 65536
    rots, poses, betas = synth.make_mano_inputs(B, 3)
    out = _run(layer, rots, poses, betas)
    assert np.isfinite(out).all() and np.all(out[:, 1] == 0.0)
    out0 = _run(layer, np.zeros_like(rots), poses, betas)
    R = mano_oracle.rodrigues(rots[:64].astype(np.float64))
    rot = np.einsum("brc,bvc->bvr", R, out0[:64].astype(np.float64))
    assert np.abs(rot - out[:64]).max() < 5e-6


def _lbs_grads(layer, rots, poses, betas, cot):
    t = [torch.from_numpy(a).cuda().requires_grad_(True) for a in (rots, poses, betas)]
    out = layer(*t)
    (out * torch.from_numpy(cot).cuda()).sum().backward()
    return out.detach().cpu().numpy(), [x.grad.cpu().numpy() for x in t]


def test_lbs_backward_matches_reference_autograd_fixture(layer):
    """scat_lbs_bwd against the gradients the UNMODIFIED reference produced with autograd (oracle/make_golden.py mano):
    d/d rots, d/d poses, d/d betas of <cotangent, rot_pose_beta_to_mesh(...)>."""
    g = load_golden("mano_lbs_grad")
    cot = np.random.Generator(np.random.PCG64(int(g["cot_seed"]))).standard_normal((5, 799, 3)).astype(np.float32)
    out, (gr, gp, gb) = _lbs_grads(layer, g["rots"], g["poses"], g["betas"], cot)
    assert np.abs(out - g["out"]).max() < 2e-6
    for name, got in (("rots", gr), ("poses", gp), ("betas", gb)):
        ref32, ref64 = g["g_" + name], g["g_" + name + "_fp64"]
        assert np.abs(got - ref64).max() < 2e-5 * np.abs(ref64).max(), name      # fp32 kernel vs float64 truth
        assert np.abs(got - ref32).max() < 2e-5 * np.abs(ref32).max(), name      # ... and vs the reference's own fp32 autograd


@pytest.mark.parametrize("B", [1, 7, 8, 9, 33, 1000])
def test_lbs_backward_matches_oracle_autograd(layer, B):
    """Ragged batches (8 samples per CTA) against autograd through the differentiable float64 restatement."""
    rots, poses, betas = synth.make_mano_inputs(B, 50 + B)
    cot = np.random.Generator(np.random.PCG64(B)).standard_normal((B, 799, 3)).astype(np.float32)
    cot[:, 1] = 7.0                                      # joint 1 is identically the origin: its cotangent must not matter
    _, (gr, gp, gb) = _lbs_grads(layer, rots, poses, betas, cot)
    t = [torch.from_numpy(a.astype(np.float64)).requires_grad_(True) for a in (rots, poses, betas)]
    ref = mano_oracle.rot_pose_beta_to_mesh_torch(*t, synth.make_mano_asset())
    (ref * torch.from_numpy(cot).double()).sum().backward()
    for name, got, r in (("rots", gr, t[0].grad), ("poses", gp, t[1].grad), ("betas", gb, t[2].grad)):
        r = r.numpy()
        assert np.abs(got - r).max() < 3e-5 * np.abs(r).max(), (name, float(np.abs(got - r).max() / np.abs(r).max()))


def test_lbs_backward_taylor_rows_and_linearity(layer):
    """Zero global rotation (theta < 1e-30: the Taylor branch of mano.py:258-265) has a finite, correct gradient (the
    limit of the generic branch), and the backward is linear in the cotangent at sweep size."""
    B = 16
    rots, poses, betas = synth.make_mano_inputs(B, 5)
    rots[3] = 0.0
    cot = np.random.Generator(np.random.PCG64(9)).standard_normal((B, 799, 3)).astype(np.float32)
    _, (gr, gp, gb) = _lbs_grads(layer, rots, poses, betas, cot)
    assert np.isfinite(gr).all() and np.isfinite(gp).all() and np.isfinite(gb).all()
    near = rots.copy()
    near[3] = 1e-4                                       # generic branch right next to zero
    _, (gr2, _, _) = _lbs_grads(layer, near, poses, betas, cot)
    assert np.abs(gr[3] - gr2[3]).max() < 2e-3 * np.abs(gr2[3]).max()
    Bs = 4096
    rots, poses, betas = synth.make_mano_inputs(Bs, 6)
    c1 = np.random.Generator(np.random.PCG64(1)).standard_normal((Bs, 799, 3)).astype(np.float32)
    c2 = np.random.Generator(np.random.PCG64(2)).standard_normal((Bs, 799, 3)).astype(np.float32)
    _, ga = _lbs_grads(layer, rots, poses, betas, c1)
    _, gb_ = _lbs_grads(layer, rots, poses, betas, c2)
    _, gs = _lbs_grads(layer, rots, poses, betas, (c1 + 2.0 * c2).astype(np.float32))
    for a, b, s in zip(ga, gb_, gs):
        assert np.abs(s - (a + 2.0 * b)).max() < 1e-4 * np.abs(s).max()
