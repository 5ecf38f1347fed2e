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


@pytest.mark.skipif(__import__("os").environ.get("SCAT_EXPERIMENTAL") != "1",
                    reason="experimental LBS blockings (csrc/lbs.cu, SCAT_LBS_V2): enable with SCAT_EXPERIMENTAL=1")
@pytest.mark.parametrize("variant", ["8,1", "8,2", "16,1", "16,2", "32,1"])
def test_lbs_experimental_blockings(variant):
    """Each alternative blocking in its own process (the choice is read once per process) against the oracle, with its
    time for 64k samples printed; not part of the default suite until one of them is validated and promoted."""
    import os
    import subprocess
    import sys
    code = (
        "import numpy as np, torch\n"
        "from oracle import mano_oracle\n"
        "from scat_b200 import synth\n"
        "from scat_b200.mano import ManoLayer\n"
        "layer = ManoLayer(synth.make_mano_asset())\n"
        "for B in (1, 7, 33, 1000):\n"
        "    r, p, b = synth.make_mano_inputs(B, B)\n"
        "    out = layer(torch.from_numpy(r).cuda(), torch.from_numpy(p).cuda(), torch.from_numpy(b).cuda()).cpu().numpy()\n"
        "    ref = mano_oracle.rot_pose_beta_to_mesh(r, p, b, synth.make_mano_asset())\n"
        "    assert np.abs(out - ref).max() < 5e-6, (B, float(np.abs(out - ref).max()))\n"
        "r, p, b = [torch.from_numpy(a).cuda() for a in synth.make_mano_inputs(65536, 3)]\n"
        "layer(r, p, b); torch.cuda.synchronize()\n"
        "e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)\n"
        "e0.record()\n"
        "for _ in range(5): layer(r, p, b)\n"
        "e1.record(); torch.cuda.synchronize()\n"
        "print('LBS_V2_OK us_per_64k', e0.elapsed_time(e1) / 5 * 1e3)\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SCAT_LBS_V2=variant, PYTHONPATH=root)
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    print(variant, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:])
    assert r.returncode == 0 and "LBS_V2_OK" in r.stdout, r.stderr[-2000:]
