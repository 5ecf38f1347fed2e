"""ORACLE (test infrastructure, not product code): numpy restatement of SCAT's MANO linear blend skinning.

Restates /root/reference/models/mano.py:236-391 (rodrigues, get_poseweights, rot_pose_beta_to_mesh)
sample by sample in float64/float32 numpy.  Pinned against the unmodified reference (imported with a
synthetic MANO_RIGHT.pkl) by oracle/make_golden.py -> tests/golden/mano_lbs.npz.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this file.
"""
from __future__ import annotations

import numpy as np

PARENTS = (-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14)      # mano.py:221-223
TIP_VERTS = (320, 443, 671, 554, 744)                                  # mano.py:373-377 (index, middle, pinky, ring, thumb)


def rodrigues(r: np.ndarray) -> np.ndarray:
    """r[N,3] -> R[N,3,3]; mano.py:236-267 (Taylor branch only where theta < 1e-30)."""
    r = np.asarray(r)
    theta = np.sqrt(np.sum(r * r, axis=1))

    def skew(n):
        z = np.zeros_like(n[:, 0])
        return np.stack([z, -n[:, 2], n[:, 1], n[:, 2], z, -n[:, 0], -n[:, 1], n[:, 0], z], axis=1).reshape(-1, 3, 3)

    eye = np.eye(3, dtype=r.dtype)[None]
    with np.errstate(divide="ignore", invalid="ignore"):
        n = r / theta[:, None]
    Sn = skew(n)
    R = eye + np.sin(theta)[:, None, None] * Sn + (1.0 - np.cos(theta))[:, None, None] * (Sn @ Sn)
    Sr = skew(r)
    t2 = theta ** 2
    R2 = eye + (1.0 - t2[:, None, None] / 6.0) * Sr + (0.5 - t2[:, None, None] / 24.0) * (Sr @ Sr)
    small = theta < 1e-30
    R[small] = R2[small]
    return R.astype(r.dtype)


def rot_pose_beta_to_mesh(rots, poses, betas, asset) -> np.ndarray:
    """rots[B,3], poses[B,45], betas[B,10] -> [B, 21+778, 3]; mano.py:280-391."""
    dt = np.asarray(rots).dtype
    v_template = asset["v_template"].astype(dt)          # [778,3]
    shapedirs = asset["shapedirs"].astype(dt)            # [778,3,10]
    posedirs = asset["posedirs"].astype(dt)              # [778,3,135]
    J_reg = np.asarray(asset["J_regressor"]).astype(dt)  # [16,778]
    weights = asset["weights"].astype(dt)                # [778,16]
    hands_mean = asset["hands_mean"].astype(dt)          # [45]
    B = rots.shape[0]

    theta = (hands_mean[None] + poses).reshape(B, 15, 3)                       # :284 (no PCA)
    theta = np.concatenate([np.zeros((B, 1, 3), dt), theta], axis=1)           # :286 root local rot = 0
    v_shaped = v_template[None] + np.einsum("vck,bk->bvc", shapedirs, betas)   # :288-292
    Rall = rodrigues(theta.reshape(-1, 3)).reshape(B, 16, 3, 3)                # :309-312
    pw = (Rall[:, 1:] - np.eye(3, dtype=dt)).reshape(B, 135)                   # :270-277
    v_posed = v_shaped + np.einsum("vck,bk->bvc", posedirs, pw)                # :296-300
    J = np.einsum("jv,bvc->bjc", J_reg, v_shaped)                              # :302-304 (from v_shaped)

    G = np.zeros((B, 16, 4, 4), dt)
    for i in range(16):                                                        # :318-327
        local = np.zeros((B, 4, 4), dt)
        local[:, :3, :3] = Rall[:, i]
        local[:, 3, 3] = 1.0
        if i == 0:
            local[:, :3, 3] = J[:, 0]
            G[:, 0] = local
        else:
            p = PARENTS[i]
            local[:, :3, 3] = J[:, i] - J[:, p]
            G[:, i] = G[:, p] @ local
    A = G.copy()                                                               # :331-337
    Jh = np.concatenate([J, np.zeros((B, 16, 1), dt)], axis=2)                 # (J,0)
    A[:, :, :, 3] -= np.einsum("bjrc,bjc->bjr", G, Jh)
    T = np.einsum("vj,bjrc->bvrc", weights, A)                                 # :339-340
    vh = np.concatenate([v_posed, np.ones((B, 778, 1), dt)], axis=2)
    v = np.einsum("bvrc,bvc->bvr", T, vh)[:, :, :3]                            # :341-348
    Jtr = np.concatenate([G[:, :, :3, 3], v[:, list(TIP_VERTS)]], axis=1)      # :355-380
    Rg = rodrigues(rots)                                                       # :351
    v = np.einsum("brc,bvc->bvr", Rg, v)                                       # :382
    Jtr = np.einsum("brc,bjc->bjr", Rg, Jtr)                                   # :383
    root = Jtr[:, 1:2].copy()                                                  # :386-388
    return np.concatenate([Jtr - root, v - root], axis=1).astype(dt)           # :391


def rot_pose_beta_to_mesh_torch(rots, poses, betas, asset):
    """The same function on torch tensors (any float dtype), differentiable: the checker for scat_lbs_bwd.

    mano.py:280-391 is differentiable end to end (it runs under autograd in the reference); this restatement keeps
    the reference's formulas, including rodrigues written in terms of n = r / theta (mano.py:246-253), so that its
    autograd gradient is the reference's.  Rows with theta < 1e-30 take the Taylor branch (mano.py:255-265)."""
    import torch
    dt, dev = rots.dtype, rots.device

    def cst(key):
        a = asset[key]
        if hasattr(a, "todense"):
            a = np.asarray(a.todense())
        return torch.as_tensor(np.asarray(a, dtype=np.float64)).to(dt).to(dev)

    v_template, shapedirs, posedirs = cst("v_template"), cst("shapedirs"), cst("posedirs")
    J_reg, weights, hands_mean = cst("J_regressor"), cst("weights"), cst("hands_mean")
    B = rots.shape[0]

    def skew(n):
        z = torch.zeros_like(n[:, 0])
        return torch.stack([z, -n[:, 2], n[:, 1], n[:, 2], z, -n[:, 0], -n[:, 1], n[:, 0], z], dim=1).view(-1, 3, 3)

    def rodrigues_t(r):
        theta = torch.sqrt(torch.sum(r * r, dim=1))
        eye = torch.eye(3, dtype=dt, device=dev)[None]
        small = theta < 1e-30
        safe = torch.where(small, torch.ones_like(theta), theta)      # keeps 0/0 out of the autograd graph of the other rows
        n = r / safe[:, None]
        Sn = skew(n)
        R = eye + torch.sin(theta)[:, None, None] * Sn + (1.0 - torch.cos(theta))[:, None, None] * (Sn @ Sn)
        Sr = skew(r)
        t2 = theta ** 2
        R2 = eye + (1.0 - t2[:, None, None] / 6.0) * Sr + (0.5 - t2[:, None, None] / 24.0) * (Sr @ Sr)
        return torch.where(small[:, None, None], R2, R)

    theta = (hands_mean[None] + poses).view(B, 15, 3)
    theta = torch.cat([torch.zeros(B, 1, 3, dtype=dt, device=dev), theta], dim=1)
    v_shaped = v_template[None] + torch.einsum("vck,bk->bvc", shapedirs, betas)
    Rall = rodrigues_t(theta.reshape(-1, 3)).view(B, 16, 3, 3)
    pw = (Rall[:, 1:] - torch.eye(3, dtype=dt, device=dev)).reshape(B, 135)
    v_posed = v_shaped + torch.einsum("vck,bk->bvc", posedirs, pw)
    J = torch.einsum("jv,bvc->bjc", J_reg, v_shaped)
    bottom = torch.tensor([0.0, 0.0, 0.0, 1.0], dtype=dt, device=dev).view(1, 1, 4).expand(B, 1, 4)
    G = [None] * 16
    for i in range(16):
        t = J[:, i] if i == 0 else J[:, i] - J[:, PARENTS[i]]
        local = torch.cat([torch.cat([Rall[:, i], t[:, :, None]], dim=2), bottom], dim=1)
        G[i] = local if i == 0 else G[PARENTS[i]] @ local
    G = torch.stack(G, dim=1)                                                     # [B,16,4,4]
    Jh = torch.cat([J, torch.zeros(B, 16, 1, dtype=dt, device=dev)], dim=2)
    corr = torch.einsum("bjrc,bjc->bjr", G, Jh)
    A = torch.cat([G[..., :3], (G[..., 3] - corr)[..., None]], dim=3)
    T = torch.einsum("vj,bjrc->bvrc", weights, A)
    vh = torch.cat([v_posed, torch.ones(B, 778, 1, dtype=dt, device=dev)], dim=2)
    v = torch.einsum("bvrc,bvc->bvr", T, vh)[:, :, :3]
    Jtr = torch.cat([G[:, :, :3, 3], v[:, list(TIP_VERTS)]], dim=1)
    Rg = rodrigues_t(rots)
    v = torch.einsum("brc,bvc->bvr", Rg, v)
    Jtr = torch.einsum("brc,bjc->bjr", Rg, Jtr)
    root = Jtr[:, 1:2]
    return torch.cat([Jtr - root, v - root], dim=1)
