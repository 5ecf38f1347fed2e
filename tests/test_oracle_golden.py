"""CPU: the oracle restatement against the fixtures generated from the unmodified reference.

Fixtures: tests/golden/*.npz written by oracle/make_golden.py (reference imported from /root/reference
in the authoring container).  fp32 CPU results are reproducible to round-off across thread counts, so
tolerances are a few ulp-scale multiples rather than exact equality.
"""
import os
import random

import numpy as np
import pytest
import torch

from oracle import head_oracle, mano_oracle
from scat_b200 import synth

HEAD_CASES = ["head_kat_b2", "head_b3_mask50", "head_b2_nope_alias", "head_b2_h4_it1_nomask", "head_b2_mask90"]


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def test_analytic_anchors(golden_dir):
    # mask draws and positional-encoding values pinned from the reference
    random.seed(0)
    assert synth.mask_indices(0.2) == [10, 19, 17, 14]
    random.seed(1)
    assert synth.mask_indices(0.2) == [20, 17, 19, 11]
    for rate, n in ((0.05, 0), (0.1, 2), (0.2, 4), (0.5, 10), (0.9, 18), (0.95, 0)):
        assert len(synth.mask_indices(rate)) == n
    pe = head_oracle.positional_encoding(21, 784)
    assert np.array_equal(pe.numpy(), _load(golden_dir, "pos_encoding")["pe"])
    np.testing.assert_allclose(pe[0, 0, :4].numpy(), [0, 1, 0, 1], atol=0)
    np.testing.assert_allclose(pe[0, 1, :4].numpy(), [0.8415, 0.5403, 0.8287, 0.5597], atol=5e-5)


@pytest.mark.parametrize("name", HEAD_CASES)
def test_head_oracle_matches_reference_fixture(golden_dir, name):
    g = _load(golden_dir, name)
    B, heads, it = int(g["B"]), int(g["heads"]), int(g["iteration"])
    pos_embed, pl_reg = bool(g["pos_embed"]), bool(g["pl_reg"])
    W = synth.make_head_weights(heads, int(g["w_seed"]), str(g["regime"]))
    x2, mf, labels = synth.make_head_inputs(B, int(g["in_seed"]))
    # the regenerated inputs must be the bytes the fixture was made from
    assert abs(x2.astype(np.float64).sum() - g["in_x2_sum"][0]) < 1e-6 * abs(g["in_x2_sum"][0])
    assert abs(np.abs(mf.astype(np.float64)).sum() - g["in_main_feat_sum"][1]) < 1e-9 * g["in_main_feat_sum"][1]
    assert abs(sum(float(np.abs(v.astype(np.float64)).sum()) for v in W.values()) - g["w_sum"][0]) < 1e-9 * g["w_sum"][0]
    random.seed(int(g["mask_seed"]))
    mask_idx = synth.mask_indices(float(g["mask_rate"]))
    assert mask_idx == g["mask_idx"].tolist()          # bit-exact index draw

    P = {k: torch.from_numpy(v) for k, v in W.items()}
    mean = torch.from_numpy(synth.make_mean_params(str(g["mean_kind"])))
    step = head_oracle.train_step(P, torch.from_numpy(x2), torch.from_numpy(mf), torch.from_numpy(labels), mean,
                                  heads=heads, iteration=it, pos_embed=pos_embed, mask_idx=mask_idx, pl_reg=pl_reg)
    assert _rel(step["pred"], g["pred"]) < 2e-6
    assert _rel(step["feat_visual"], g["feat_visual"]) < 2e-6
    assert np.all(step["pred"].numpy()[:, 6:9] == 0.0)                       # joint 1 == 0 exactly
    if pl_reg:
        assert _rel(step["pl"], g["pl"]) < 2e-5
        if pos_embed and len(mask_idx):
            assert np.all(step["pl"].numpy().reshape(B, 21, -1)[:, mask_idx] == 0.0)   # masked tokens: zero VJP
    np.testing.assert_allclose(step["loss"].item(), g["loss"][0], rtol=2e-6)
    assert _rel(step["main_feat_grad"], g["main_feat_grad"]) < 2e-5
    xg = step["x2_grad"].double().reshape(-1).numpy()
    assert _rel(xg[g["x2_grad_idx"]], g["x2_grad_val"]) < 2e-5
    np.testing.assert_allclose(np.abs(xg).sum(), g["x2_grad_sum"][1], rtol=1e-5)
    for k in W:
        gg = step["grads"][k].double().reshape(-1).numpy()
        ref_abs = g["g_sum/" + k][1]
        np.testing.assert_allclose(np.abs(gg).sum(), ref_abs, rtol=2e-5, atol=1e-12, err_msg=k)
        scale = np.abs(gg).max() + 1e-30
        assert np.abs(gg[g["g_idx/" + k]] - g["g_val/" + k]).max() <= 2e-5 * scale, k


def test_fp64_oracle_close_to_fp32(golden_dir):
    g = _load(golden_dir, "head_kat_b2")
    W = synth.make_head_weights(8)
    x2, mf, _ = synth.make_head_inputs(2, 0)
    P = {k: torch.from_numpy(v).double() for k, v in W.items()}
    mean = torch.from_numpy(synth.make_mean_params("hand")).double()
    out = head_oracle.head_forward(P, torch.from_numpy(x2).double(), torch.from_numpy(mf).double(), mean,
                                   mask_idx=g["mask_idx"].tolist())
    assert _rel(out[0], g["pred"]) < 1e-5


def test_token_transformer_fixture(golden_dir):
    g = _load(golden_dir, "tokens_n128_d196")
    B, n, dim, heads = int(g["B"]), int(g["n"]), int(g["dim"]), int(g["heads"])
    W = synth.make_token_weights(dim, heads)
    P = {k: torch.from_numpy(v) for k, v in W.items()}
    tok = torch.from_numpy(synth.make_token_inputs(B, n, dim, int(g["in_seed"])))
    out, mean = head_oracle.token_transformer_forward(P, tok, heads=heads, mask_idx=g["mask_idx"].tolist())
    assert _rel(out, g["out"]) < 2e-6
    assert _rel(mean, g["mean"]) < 2e-6


def test_mano_oracle_fixture(golden_dir):
    g = _load(golden_dir, "mano_lbs")
    asset = synth.make_mano_asset()
    out = mano_oracle.rot_pose_beta_to_mesh(g["rots"], g["poses"], g["betas"], asset)
    assert out.shape == (g["rots"].shape[0], 799, 3)
    assert np.abs(out - g["out"]).max() < 2e-6
    assert np.all(out[:, 1] == 0.0)            # root joint (index 1) is the origin, mano.py:386-388
    a64 = {k: (np.asarray(v, dtype=np.float64) if v.dtype.kind == "f" else v) for k, v in asset.items()}
    o64 = mano_oracle.rot_pose_beta_to_mesh(g["rots"].astype(np.float64), g["poses"].astype(np.float64),
                                            g["betas"].astype(np.float64), a64)
    assert np.abs(o64 - g["out_fp64_oracle"]).max() < 1e-12


def test_mano_oracle_gradients_match_reference_autograd(golden_dir):
    """The differentiable restatement (oracle/mano_oracle.py: rot_pose_beta_to_mesh_torch) against the gradients the
    unmodified reference produced with autograd (tests/golden/mano_lbs_grad.npz)."""
    import torch
    g = _load(golden_dir, "mano_lbs_grad")
    asset = synth.make_mano_asset()
    cot = np.random.Generator(np.random.PCG64(int(g["cot_seed"]))).standard_normal((5, 799, 3)).astype(np.float32)
    t = [torch.from_numpy(g[k].astype(np.float64)).requires_grad_(True) for k in ("rots", "poses", "betas")]
    out = mano_oracle.rot_pose_beta_to_mesh_torch(*t, asset)
    assert np.abs(out.detach().numpy() - g["out"]).max() < 2e-6
    (out * torch.from_numpy(cot).double()).sum().backward()
    for name, x in zip(("rots", "poses", "betas"), t):
        ref = g["g_" + name]
        assert np.abs(x.grad.numpy() - ref).max() < 1e-5 * np.abs(ref).max(), name
        assert np.abs(x.grad.numpy() - g["g_" + name + "_fp64"]).max() < 1e-10 * np.abs(ref).max(), name
    # and the numpy / torch restatements are the same function
    o_np = mano_oracle.rot_pose_beta_to_mesh(g["rots"].astype(np.float64), g["poses"].astype(np.float64),
                                             g["betas"].astype(np.float64), asset)
    assert np.abs(o_np - out.detach().numpy()).max() < 1e-12


def test_adam_oracle_matches_torch_adam_fixture(golden_dir):
    """oracle/adam_oracle.py against tests/golden/adam.npz = torch.optim.Adam itself (train.py:60) run on the CPU:
    6 steps, warm-up style learning-rate changes, with and without weight decay."""
    from oracle import adam_oracle
    g = np.load(os.path.join(golden_dir, "adam.npz"))
    n = int(g["n_tensors"])
    for tag, wd in (("wd0", 0.0), ("wd1", 0.01)):
        for i in range(n):
            p = g[f"p0_{i}"].copy()
            m, v = np.zeros_like(p), np.zeros_like(p)
            for k in range(6):
                adam_oracle.adam_step(p, g[f"g{k}_{i}"], m, v, k + 1, float(g["lrs"][k]), weight_decay=wd)
            for got, key in ((p, "p"), (m, "m"), (v, "v")):
                ref = g[f"{tag}_{key}{i}"]
                assert np.abs(got - ref).max() <= 2e-7 * np.abs(ref).max(), (tag, key, i)


def test_warmup_schedule_restatement():
    """GradualWarmupScheduler(multiplier=1, total_epoch=15, StepLR(gamma=1)) stepped with epoch + 1
    (train.py:61-63,134): linear ramp base/15 .. base over 15 epochs, then constant.  The package itself is absent
    (parity unpinned, oracle/adam_oracle.py); product and oracle restatements must at least agree."""
    from oracle import adam_oracle
    from scat_b200 import optim
    base = 1e-4
    for e in range(40):
        want = base * min(e + 1, 15) / 15
        assert adam_oracle.gradual_warmup_lr(base, e) == pytest.approx(want, rel=1e-12)
        assert optim.gradual_warmup_lr(base, e) == adam_oracle.gradual_warmup_lr(base, e)

    class Opt:
        param_groups = [{"lr": base}]
    o = Opt()
    sched = optim.WarmupSchedule(o)
    assert o.param_groups[0]["lr"] == 0.0                      # _LRScheduler.__init__ leaves last_epoch = 0
    for e in range(20):
        sched.step(e + 1)
        assert o.param_groups[0]["lr"] == adam_oracle.gradual_warmup_lr(base, e)


@pytest.mark.parametrize("name", ["coarse_b3_mask20", "coarse_b2_nope_alias"])
def test_coarse_oracle_matches_reference_fixture(golden_dir, name):
    """oracle/head_oracle.coarse_forward against EncoderTransformerCoarse of the unmodified reference
    (hand_net.py:216-311 + vision_transformer_attn.py:88-113; oracle/make_golden.py coarse)."""
    import torch
    g = _load(golden_dir, name)
    W = synth.make_coarse_weights()
    P = {k: torch.from_numpy(v) for k, v in W.items()}
    x2, mf, _ = synth.make_head_inputs(int(g["B"]), int(g["in_seed"]))
    mean = torch.from_numpy(synth.make_mean_params("hand"))
    with torch.no_grad():
        pred, fv, attn = head_oracle.coarse_forward(P, torch.from_numpy(x2), torch.from_numpy(mf), mean,
                                                    pos_embed=bool(g["pos_embed"]), mask_idx=g["mask_idx"].tolist())
    assert _rel(pred.numpy(), g["pred"]) < 2e-6
    assert _rel(fv.numpy(), g["feat_visual"]) < 2e-6
    assert _rel(attn.numpy(), g["attn"]) < 2e-6
    assert np.all(pred.numpy()[:, 6:9] == 0.0)
    assert np.allclose(attn.numpy().sum(-1), 1.0, atol=1e-5)
