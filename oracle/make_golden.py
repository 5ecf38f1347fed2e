"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on the CPU.

Run in the authoring container only (the GPU box has no /root/reference):

    python oracle/make_golden.py

For every case it (1) runs the reference's own EncoderTransformer / Transformer / rot_pose_beta_to_mesh,
(2) runs the oracle restatement (oracle/head_oracle.py, oracle/mano_oracle.py) on the same inputs,
(3) asserts they agree, and (4) stores the reference outputs (full tensors where small, per-tensor
checksums + strided samples where large) together with input checksums.  Inputs and weights are
regenerated from seeds by scat_b200/synth.py (numpy PCG64), so fixtures stay small.

Shims needed to import the reference offline without a GPU: Tensor.cuda -> identity,
model_zoo.load_url -> {}; the backbone is replaced by a stub that returns the synthetic seam tensors.
"""
from __future__ import annotations

import os
import pickle
import random
import sys
import tempfile
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from scat_b200 import synth  # noqa: E402
from oracle import head_oracle, mano_oracle  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def _import_reference_head():
    sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self
    import torch.utils.model_zoo as mz
    mz.load_url = lambda *a, **k: {}
    from models import hand_net, vision_transformer
    return hand_net, vision_transformer


class _StubBackbone(torch.nn.Module):
    """Stands in for resnet50: returns (main_feat, x1-like, x2, None, None) (resnet.py:162)."""

    def __init__(self):
        super().__init__()
        self.main_feat = None
        self.x2 = None

    def forward(self, _img):
        return self.main_feat, self.x2[:, :1, :1, :1], self.x2, None, None


def summary(t: torch.Tensor, n_samples: int = 64):
    """Checksum record for a tensor too large to store: [sum, abs-sum, sum of squares] in float64
    plus a strided sample."""
    f = t.detach().double().reshape(-1)
    stride = max(1, f.numel() // n_samples)
    idx = torch.arange(0, f.numel(), stride)[:n_samples]
    return (np.array([f.sum().item(), f.abs().sum().item(), f.square().sum().item()]),
            idx.numpy().astype(np.int64), f[idx].numpy())


def run_head_case(hand_net, name, *, B, heads, iteration, pos_embed, mask_rate, pl_reg, mask_seed,
                  w_seed=20211011, in_seed=0, regime="unit", mean_kind="hand", backward=True):
    opt = SimpleNamespace(vit_heads=heads, pl_reg=pl_reg, iteration=iteration, pos_embed=pos_embed,
                          mask_rate=mask_rate)
    mean = torch.from_numpy(synth.make_mean_params(mean_kind))
    torch.manual_seed(0)
    net = hand_net.EncoderTransformer(opt, mean).train()
    W = synth.make_head_weights(heads, w_seed, regime)
    sd = net.state_dict()
    for k, v in W.items():
        assert tuple(sd[k].shape) == v.shape, (k, sd[k].shape, v.shape)
        sd[k].copy_(torch.from_numpy(v))
    stub = _StubBackbone()
    net.main_encoder = stub
    x2_np, mf_np, lab_np = synth.make_head_inputs(B, in_seed)
    x2 = torch.from_numpy(x2_np).requires_grad_(True)
    mf = torch.from_numpy(mf_np).requires_grad_(True)
    labels = torch.from_numpy(lab_np)
    stub.x2, stub.main_feat = x2, mf

    random.seed(mask_seed)
    outs = net(torch.zeros(B, 3, 8, 8))
    pred, fv = outs[0], outs[1]
    pl = outs[2] if pl_reg else None
    # what indices did the reference draw?  replay the host RNG (hand_net.py:370-372)
    random.seed(mask_seed)
    mask_idx = synth.mask_indices(mask_rate)

    rec = dict(B=B, heads=heads, iteration=iteration, pos_embed=int(pos_embed), mask_rate=mask_rate,
               pl_reg=int(pl_reg), mask_seed=mask_seed, w_seed=w_seed, in_seed=in_seed,
               regime=regime, mean_kind=mean_kind, mask_idx=np.array(mask_idx, dtype=np.int64),
               pred=pred.detach().numpy(), feat_visual=fv.detach().numpy().copy())
    for nm, arr in (("x2", x2_np), ("main_feat", mf_np), ("labels", lab_np)):
        rec["in_" + nm + "_sum"] = np.array([arr.astype(np.float64).sum(), np.abs(arr.astype(np.float64)).sum()])
    rec["w_sum"] = np.array([sum(float(np.abs(v.astype(np.float64)).sum()) for v in W.values())])
    if pl_reg:
        rec["pl"] = pl.detach().numpy()
        assert not pl.requires_grad

    # oracle vs reference, forward
    P = {k: torch.from_numpy(v) for k, v in W.items()}
    o = head_oracle.head_forward(P, torch.from_numpy(x2_np).requires_grad_(True), torch.from_numpy(mf_np),
                                 mean, heads=heads, iteration=iteration, pos_embed=pos_embed,
                                 mask_idx=mask_idx, pl_reg=pl_reg)
    d_pred = (o[0] - pred).abs().max().item()
    d_fv = (o[1] - fv).abs().max().item()
    assert d_pred <= 1e-5 * max(1.0, pred.abs().max().item()), (name, d_pred)
    assert d_fv <= 1e-5 * max(1.0, fv.abs().max().item()), (name, d_fv)
    diffs = dict(pred=d_pred, feat_visual=d_fv)
    if pl_reg:
        d_pl = (o[2] - pl).abs().max().item()
        assert d_pl <= 1e-5 * max(1e-3, pl.abs().max().item()), (name, d_pl)
        diffs["pl"] = d_pl

    if backward:
        # train-step body restated from train.py:165-206 (file itself unimportable offline)
        loss, l3, l2, lpl = head_oracle.train_loss(pred, labels, pl)
        loss.backward()
        rec["loss"] = np.array([loss.item(), l3.item(), l2.item(), lpl.item()])
        rec["main_feat_grad"] = mf.grad.numpy()
        s, i, v = summary(x2.grad)
        rec["x2_grad_sum"], rec["x2_grad_idx"], rec["x2_grad_val"] = s, i, v
        named = dict(net.named_parameters())
        ostep = head_oracle.train_step(P, torch.from_numpy(x2_np), torch.from_numpy(mf_np), labels, mean,
                                       heads=heads, iteration=iteration, pos_embed=pos_embed,
                                       mask_idx=mask_idx, pl_reg=pl_reg)
        assert abs(ostep["loss"].item() - loss.item()) <= 1e-5 * abs(loss.item()), (name, "loss")
        worst = 0.0
        for k in W:
            g = named[k].grad
            g = torch.zeros_like(named[k]) if g is None else g
            s, i, v = summary(g)
            rec["g_sum/" + k], rec["g_idx/" + k], rec["g_val/" + k] = s, i, v
            den = g.abs().max().item() + 1e-30
            worst = max(worst, (ostep["grads"][k] - g).abs().max().item() / den)
        assert worst < 2e-4, (name, "param grads", worst)
        dx = (ostep["x2_grad"] - x2.grad).abs().max().item() / (x2.grad.abs().max().item() + 1e-30)
        assert dx < 2e-4, (name, "x2 grad", dx)
        diffs["grads_rel"] = worst
        diffs["x2_grad_rel"] = dx
    rec["oracle_vs_reference"] = np.array([diffs.get(k, 0.0) for k in ("pred", "feat_visual", "pl", "grads_rel", "x2_grad_rel")])
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **rec)
    print(f"[golden] {name}: mask={mask_idx} oracle-vs-reference {diffs}")


def run_token_case(vision_transformer, hand_net, name, *, B, n, dim, heads, mask_rate, mask_seed, in_seed=3):
    """Config 4: the HRNet-variant token path up to feat.mean(dim=1) (hand_net.py:193-203)."""
    torch.manual_seed(0)
    tr = vision_transformer.Transformer(dim=dim, depth=3, heads=heads, dim_head=64, mlp_dim=2 * dim, dropout=0.0)
    pos = hand_net.PositionalEncoding(dim, max_len=n)
    W = synth.make_token_weights(dim, heads)
    for k, p in tr.named_parameters():
        p.data.copy_(torch.from_numpy(W["transformer." + k]))
    tok = torch.from_numpy(synth.make_token_inputs(B, n, dim, in_seed))
    random.seed(mask_seed)
    masked = list(range(n))
    random.shuffle(masked)
    masked = masked[: int(mask_rate * n)]
    feat = pos(tok)
    feat[:, masked, :] = torch.from_numpy(W["mask_token"])
    out = tr(feat, None)
    mean = out.mean(dim=1)
    P = {k: torch.from_numpy(v) for k, v in W.items()}
    o_out, o_mean = head_oracle.token_transformer_forward(P, tok, heads=heads, mask_idx=masked)
    d = (o_out - out).abs().max().item()
    assert d < 1e-5 * max(1.0, out.abs().max().item()), d
    rec = dict(B=B, n=n, dim=dim, heads=heads, mask_idx=np.array(masked, dtype=np.int64), in_seed=in_seed,
               out=out.detach().numpy(), mean=mean.detach().numpy(), pe_row1=pos.pe[0, 1, :8].numpy(),
               tok_sum=np.array([tok.double().sum().item()]))
    rec["w_sum"] = np.array([sum(float(np.abs(v.astype(np.float64)).sum()) for v in W.values())])
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **rec)
    print(f"[golden] {name}: oracle-vs-reference {d:.3e}")


def run_coarse_case(hand_net, name, *, B, pos_embed, mask_rate, mask_seed, in_seed=7):
    """EncoderTransformerCoarse (hand_net.py:216-311, `--net reg_transformer_coarse`) from the unmodified reference,
    pl_reg off (the inference path of eval.py:788-834), against oracle/head_oracle.coarse_forward."""
    opt = SimpleNamespace(vit_heads=8, pl_reg=False, iteration=3, pos_embed=pos_embed, mask_rate=mask_rate)
    mean = torch.from_numpy(synth.make_mean_params("hand"))
    torch.manual_seed(0)
    net = hand_net.EncoderTransformerCoarse(opt, mean).eval()
    W = synth.make_coarse_weights()
    sd = net.state_dict()
    head_keys = [k for k in sd if not k.startswith("main_encoder.") and k != "positionalEncoding.pe"]
    assert head_keys == list(W), (head_keys[:6], list(W)[:6])                 # same keys, same order as the reference
    for k, v in W.items():
        assert tuple(sd[k].shape) == v.shape, (k, sd[k].shape, v.shape)
        sd[k].copy_(torch.from_numpy(v))
    stub = _StubBackbone()
    net.main_encoder = stub
    x2_np, mf_np, _ = synth.make_head_inputs(B, in_seed)
    stub.x2, stub.main_feat = torch.from_numpy(x2_np), torch.from_numpy(mf_np)
    random.seed(mask_seed)
    with torch.no_grad():
        pred, fv, attn = net(torch.zeros(B, 3, 8, 8))
    random.seed(mask_seed)
    mask_idx = synth.mask_indices(mask_rate)
    P = {k: torch.from_numpy(v) for k, v in W.items()}
    with torch.no_grad():
        o = head_oracle.coarse_forward(P, torch.from_numpy(x2_np), torch.from_numpy(mf_np), mean, pos_embed=pos_embed,
                                       mask_idx=mask_idx)
    for nm, a, b in (("pred", o[0], pred), ("feat_visual", o[1], fv), ("attn", o[2], attn)):
        d = (a - b).abs().max().item()
        assert d <= 1e-6 * max(1.0, b.abs().max().item()), (name, nm, d)
        print(f"[golden] {name}: oracle-vs-reference {nm} {d:.3e}")
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), B=B, pos_embed=int(pos_embed), mask_rate=mask_rate,
                        mask_seed=mask_seed, in_seed=in_seed, mask_idx=np.array(mask_idx, dtype=np.int64),
                        pred=pred.numpy(), feat_visual=fv.numpy().copy(), attn=attn.numpy())


def run_mano_case(name, B=6):
    """MANO LBS: import models.mano from a temp cwd holding a synthetic MANO_RIGHT.pkl (mano.py:220)."""
    import scipy.sparse as sp
    asset = synth.make_mano_asset()
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "extra_data"))
    dd = dict(asset)
    dd["J_regressor"] = sp.csc_matrix(asset["J_regressor"])
    with open(os.path.join(tmp, "extra_data", "MANO_RIGHT.pkl"), "wb") as f:
        pickle.dump(dd, f)
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        sys.path.insert(0, REF)
        torch.Tensor.cuda = lambda self, *a, **k: self
        from models import mano as ref_mano
    finally:
        os.chdir(cwd)
    rots, poses, betas = synth.make_mano_inputs(B, 0)
    rots[1] = 0.0                       # exercises the theta -> 0 limit of the global rotation
    poses[2] = -asset["hands_mean"]     # all local rotations exactly zero -> NaN in the reference's n = r/theta
    out = ref_mano.rot_pose_beta_to_mesh(torch.from_numpy(rots), torch.from_numpy(poses), torch.from_numpy(betas))
    out = out.detach().numpy()
    o = mano_oracle.rot_pose_beta_to_mesh(rots, poses, betas, asset)
    finite = np.isfinite(out).all(axis=(1, 2))
    d = np.abs(o[finite] - out[finite]).max()
    assert d < 2e-5, d
    assert (np.isfinite(o).all(axis=(1, 2)) == finite).all(), "oracle must reproduce the reference's NaN rows"
    o64 = mano_oracle.rot_pose_beta_to_mesh(rots.astype(np.float64), poses.astype(np.float64),
                                            betas.astype(np.float64), {k: np.asarray(v, dtype=np.float64) if k != "kintree_table" and k != "f" else v for k, v in asset.items()})
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), rots=rots, poses=poses, betas=betas, out=out,
                        finite=finite, out_fp64_oracle=o64.astype(np.float64))
    print(f"[golden] {name}: oracle-vs-reference {d:.3e}; finite rows {finite.tolist()}")

    # ---- gradients: the reference function under the reference's own mechanism (autograd), generic inputs ----
    Bg = 5
    rots, poses, betas = synth.make_mano_inputs(Bg, 11)
    cot = np.random.Generator(np.random.PCG64(4242)).standard_normal((Bg, 799, 3)).astype(np.float32)
    tr, tp, tb = (torch.from_numpy(a.copy()).requires_grad_(True) for a in (rots, poses, betas))
    out_ref = ref_mano.rot_pose_beta_to_mesh(tr, tp, tb)
    (out_ref * torch.from_numpy(cot)).sum().backward()
    # the differentiable restatement (float64) must give the reference's gradients
    dr, dp_, db = (torch.from_numpy(a.astype(np.float64)).requires_grad_(True) for a in (rots, poses, betas))
    out_o = mano_oracle.rot_pose_beta_to_mesh_torch(dr, dp_, db, asset)
    (out_o * torch.from_numpy(cot).double()).sum().backward()
    for nm, a, b in (("rots", tr.grad, dr.grad), ("poses", tp.grad, dp_.grad), ("betas", tb.grad, db.grad)):
        e = float((a.double() - b).abs().max() / b.abs().max())
        assert e < 2e-4, (nm, e)                                           # fp32 autograd of the reference vs float64
        print(f"[golden] {name}_grad: d{nm} reference(fp32 autograd) vs oracle(fp64) rel {e:.2e}")
    np.savez_compressed(os.path.join(GOLD, name + "_grad.npz"), rots=rots, poses=poses, betas=betas, cot_seed=np.array(4242),
                        out=out_ref.detach().numpy(), g_rots=tr.grad.numpy(), g_poses=tp.grad.numpy(), g_betas=tb.grad.numpy(),
                        g_rots_fp64=dr.grad.numpy(), g_poses_fp64=dp_.grad.numpy(), g_betas_fp64=db.grad.numpy())


def run_adam_case(name="adam"):
    """torch.optim.Adam itself (the reference's optimiser, train.py:60) on a few small tensors: 6 steps, the learning
    rate changing between steps the way the warm-up schedule changes it, with and without weight decay."""
    from oracle import adam_oracle
    rng = np.random.Generator(np.random.PCG64(77))
    shapes = [(1, 1, 37), (5, 8, 1, 1), (13,), (6, 11)]
    p0 = [rng.standard_normal(s).astype(np.float32) * 0.05 for s in shapes]
    grads = [[(rng.standard_normal(s) * (10.0 ** rng.integers(-4, 1))).astype(np.float32) for s in shapes]
             for _ in range(6)]
    lrs = [1e-4 * (k + 1) / 15 for k in range(3)] * 2
    out = {"lrs": np.array(lrs), "n_tensors": np.array(len(shapes))}
    for tag, wd in (("wd0", 0.0), ("wd1", 0.01)):
        params = [torch.nn.Parameter(torch.from_numpy(a.copy())) for a in p0]
        opt = torch.optim.Adam(params, lr=lrs[0], weight_decay=wd)
        mine = [(a.copy(), np.zeros_like(a), np.zeros_like(a)) for a in p0]
        for k in range(6):
            for pg in opt.param_groups:
                pg["lr"] = lrs[k]
            for p, g in zip(params, grads[k]):
                p.grad = torch.from_numpy(g.copy())
            opt.step()
            for (a, m, v), g in zip(mine, grads[k]):
                adam_oracle.adam_step(a, g, m, v, k + 1, lrs[k], weight_decay=wd)
        for i, p in enumerate(params):
            ref = p.detach().numpy()
            err = np.abs(mine[i][0] - ref).max() / np.abs(ref).max()
            assert err < 2e-7, (tag, i, err)
            out[f"{tag}_p{i}"] = ref
            out[f"{tag}_m{i}"] = opt.state[p]["exp_avg"].numpy()
            out[f"{tag}_v{i}"] = opt.state[p]["exp_avg_sq"].numpy()
    for i, a in enumerate(p0):
        out[f"p0_{i}"] = a
    for k in range(6):
        for i, g in enumerate(grads[k]):
            out[f"g{k}_{i}"] = g
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print("wrote", name)


def _reference_functions(path, names, extra=None):
    """Execute only the named top-level function definitions of a reference source file (the file as a whole is not
    importable)."""
    import ast
    src = open(path, encoding="utf-8").read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert sorted(n.name for n in keep) == sorted(names), [n.name for n in keep]
    ns = {"np": np, "torch": torch}
    ns.update(extra or {})
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return ns


def run_eval_case(name="eval_metrics"):
    """The reference's own metric functions (eval.py:110-161,300-340; data_utils/eval_utils.py:6-47) on seeded joints."""
    from oracle import eval_oracle
    if not hasattr(np, "trapz"):
        np.trapz = np.trapezoid
    ev = _reference_functions(os.path.join(REF, "eval.py"),
                              ["batch_compute_similarity_transform_torch", "cal_PCK", "_area_under_curve"])
    eu = _reference_functions(os.path.join(REF, "data_utils", "eval_utils.py"), ["compute_accel", "compute_error_accel"])
    rng = np.random.Generator(np.random.PCG64(4242))
    B = 16
    gt = (rng.standard_normal((B, 21, 3)) * 0.04).astype(np.float32)
    gt -= gt[:, 1:2]
    # predictions = a similarity transform of the target + noise, so alignment matters
    pred = np.empty_like(gt)
    for b in range(B):
        q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
        if np.linalg.det(q) < 0:
            q[:, 0] *= -1
        pred[b] = (rng.uniform(0.7, 1.4) * gt[b] @ q.T + rng.standard_normal(3) * 0.05
                   + rng.standard_normal((21, 3)) * 0.012).astype(np.float32)
    pt, gtt = torch.from_numpy(pred), torch.from_numpy(gt)
    aligned = ev["batch_compute_similarity_transform_torch"](pt.clone(), gtt.clone())
    mine = eval_oracle.similarity_transform(pt, gtt)
    assert float((aligned - mine).abs().max()) < 2e-6, float((aligned - mine).abs().max())
    rnge = np.arange(20, 51, 5)
    out = {"pred": pred, "gt": gt, "aligned": aligned.numpy(), "rnge": rnge}
    for tag, p in (("raw", pt), ("pa", aligned)):
        pck = ev["cal_PCK"](p, gtt, rnge)
        assert np.array_equal(pck, eval_oracle.cal_pck(p, gtt, rnge))
        auc = ev["_area_under_curve"](rnge / rnge.max(), pck[:, -1])
        assert abs(auc - eval_oracle.area_under_curve(rnge / rnge.max(), pck[:, -1])) < 1e-12
        out[f"pck_{tag}"], out[f"auc_{tag}"] = pck, np.array(auc)
    vis = np.ones(B, dtype=bool)
    vis[[3, 9]] = False
    out["accel"] = eu["compute_accel"](pred)
    out["accel_err"] = eu["compute_error_accel"](joints_gt=gt, joints_pred=pred)
    out["accel_err_vis"] = eu["compute_error_accel"](joints_gt=gt, joints_pred=pred, vis=vis)
    out["vis"] = vis
    assert np.allclose(out["accel"], eval_oracle.compute_accel(pred), rtol=0, atol=0)
    assert np.array_equal(out["accel_err"], eval_oracle.compute_error_accel(gt, pred))
    assert np.array_equal(out["accel_err_vis"], eval_oracle.compute_error_accel(gt, pred, vis))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print("wrote", name)


def main():
    os.makedirs(GOLD, exist_ok=True)
    if sys.argv[1:] == ["eval"]:          # the evaluation-metric fixture alone
        run_eval_case("eval_metrics")
        return
    if sys.argv[1:] == ["adam"]:          # the optimiser fixture alone (needs torch only, not /root/reference)
        run_adam_case("adam")
        return
    if sys.argv[1:] == ["coarse"]:        # the reg_transformer_coarse fixtures alone
        hand_net, _ = _import_reference_head()
        run_coarse_case(hand_net, "coarse_b3_mask20", B=3, pos_embed=True, mask_rate=0.2, mask_seed=6)
        run_coarse_case(hand_net, "coarse_b2_nope_alias", B=2, pos_embed=False, mask_rate=0.5, mask_seed=8, in_seed=9)
        return
    if sys.argv[1:] == ["mano"]:          # the LBS fixtures alone (forward + autograd gradients of the reference)
        run_mano_case("mano_lbs")
        return
    hand_net, vt = _import_reference_head()
    # analytic anchors
    random.seed(0)
    assert synth.mask_indices(0.2) == [10, 19, 17, 14]
    random.seed(1)
    assert synth.mask_indices(0.2) == [20, 17, 19, 11]
    pe_ref = hand_net.PositionalEncoding(784, max_len=21).pe
    assert torch.equal(pe_ref, head_oracle.positional_encoding(21, 784))
    np.savez_compressed(os.path.join(GOLD, "pos_encoding.npz"), pe=pe_ref.numpy())

    run_head_case(hand_net, "head_kat_b2", B=2, heads=8, iteration=3, pos_embed=True, mask_rate=0.2,
                  pl_reg=True, mask_seed=0)
    run_head_case(hand_net, "head_b3_mask50", B=3, heads=8, iteration=3, pos_embed=True, mask_rate=0.5,
                  pl_reg=True, mask_seed=1, in_seed=1, regime="hand")
    run_head_case(hand_net, "head_b2_nope_alias", B=2, heads=8, iteration=2, pos_embed=False, mask_rate=0.2,
                  pl_reg=True, mask_seed=2, in_seed=2)
    run_head_case(hand_net, "head_b2_h4_it1_nomask", B=2, heads=4, iteration=1, pos_embed=True, mask_rate=0.0,
                  pl_reg=False, mask_seed=3, in_seed=3)
    run_head_case(hand_net, "head_b2_mask90", B=2, heads=8, iteration=3, pos_embed=True, mask_rate=0.9,
                  pl_reg=True, mask_seed=4, in_seed=4)
    run_token_case(vt, hand_net, "tokens_n128_d196", B=2, n=128, dim=196, heads=8, mask_rate=0.2, mask_seed=5)
    run_coarse_case(hand_net, "coarse_b3_mask20", B=3, pos_embed=True, mask_rate=0.2, mask_seed=6)
    run_coarse_case(hand_net, "coarse_b2_nope_alias", B=2, pos_embed=False, mask_rate=0.5, mask_seed=8, in_seed=9)
    run_mano_case("mano_lbs")
    run_adam_case("adam")
    run_eval_case("eval_metrics")


if __name__ == "__main__":
    main()
