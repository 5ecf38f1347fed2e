"""GPU parity of the fused MANO LBS kernel (models/mano.py:280-391) against the reference fixture, the numpy
oracle, and size-independent properties at sweep sizes (BASELINE config 5)."""
import numpy as np
import pytest
import torch

from oracle import mano_oracle
from scat_b200 import synth
from tests.util import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def layer():
    from scat_b200.mano import ManoLayer
    return ManoLayer(synth.make_mano_asset())


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
