"""Worker of test_gpu_peer.py: run under torchrun with N >= 2 GPUs.  Checks the one-kernel NVLink all-reduce
(dp.PeerMemory) against torch.distributed's NCCL all-reduce and the peer-mode train step against the NCCL-mode one."""
import os
import random
import sys
from types import SimpleNamespace

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scat_b200 import dp, synth  # noqa: E402
from scat_b200._lib import stream_ptr  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    # ---- raw all-reduce: odd length (padding), repeated launches, sub-range, integers (exact in any order) ----
    n = 1_000_003
    pm = dp.PeerMemory(n, dev)
    g = torch.Generator(device=dev).manual_seed(17 + rank)
    for it in range(5):
        x = torch.randint(-1000, 1000, (n,), generator=g, device=dev).float()
        pm.flat.copy_(x)
        ref = x.clone()
        dist.all_reduce(ref)
        pm.enqueue(stream_ptr())
        torch.cuda.synchronize()
        assert not pm.timed_out()
        assert torch.equal(pm.flat, ref), f"iteration {it}: peer all-reduce != NCCL all-reduce"
        assert float(pm.flat_padded[n:].abs().sum()) == 0.0
    x = torch.randn(n, generator=g, device=dev)
    pm.flat.copy_(x)
    lo, hi = 4096, 4096 + 40000
    ref = x.clone()
    part = ref[lo:hi].clone()
    dist.all_reduce(part)
    pm.enqueue(stream_ptr(), lo, hi)
    torch.cuda.synchronize()
    assert torch.equal(pm.flat[:lo], x[:lo]) and torch.equal(pm.flat[hi:], x[hi:])
    got = pm.flat[lo:hi]
    assert (got - part).abs().max() <= 1e-6 * part.abs().max()
    # every rank holds bit-identical sums (fixed rank order)
    mine = got.clone()
    theirs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(theirs, mine)
    assert all(torch.equal(t, mine) for t in theirs)
    pm.close()

    # ---- train step: comm="peer" (all-reduce inside the step's CUDA graph) == comm="nccl" ----
    from scat_b200.hand_net import EncoderTransformer
    from scat_b200.train_step import HeadTrainStep
    opt = SimpleNamespace(vit_heads=8, pl_reg=True, iteration=3, pos_embed=True, mask_rate=0.2)
    mean = torch.from_numpy(synth.make_mean_params("hand"))
    sd = {k: torch.from_numpy(v) for k, v in synth.make_head_weights(8).items()}
    B = 8
    x2, mf, lab = synth.make_head_inputs(B, 50 + rank)
    out = {}
    # "peer": the exchange in three parts hidden under the backward phases (default); "peer_serial": one exchange
    # kernel behind the step
    for mode in ("nccl", "peer", "peer_serial"):
        comm = "nccl" if mode == "nccl" else "peer"
        net = EncoderTransformer(opt, mean, precision="tf32", backbone=torch.nn.Identity())
        sd["positionalEncoding.pe"] = net.positionalEncoding.pe
        net.load_state_dict(sd, strict=True)
        net = net.to(dev)
        ts = HeadTrainStep(net, B, comm=comm, overlap_exchange=(mode == "peer"))
        assert ts.comm == comm and ts.overlap_exchange == (mode == "peer")
        ts.load_inputs(*[torch.from_numpy(a).to(dev) for a in (x2, mf, lab)])
        ts.set_mask(list(range(ts.n_masked)))
        for _ in range(3):
            ts.step()
        torch.cuda.synchronize()
        if ts.peer is not None:
            assert not ts.peer.timed_out()
        out[mode] = ts.bucket.flat.clone()
        if comm == "peer":
            # forward + backward + all-reduce + Adam in one graph: the replicas must stay bit-identical
            from scat_b200.optim import HeadAdam
            opt_ = HeadAdam(net.head_parameters(), lr=1e-4)
            ts.attach_optimizer(opt_)
            start = ts.flat_params.clone()
            for _ in range(3):
                ts.step(optimize=True)
            torch.cuda.synchronize()
            assert not ts.peer.timed_out()
            mine = ts.flat_params.clone()
            assert float((mine - start).abs().max()) > 0
            theirs = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(theirs, mine)
            assert all(torch.equal(t, mine) for t in theirs), "replicas diverged"
        ts.close()
    a = out["nccl"]
    err = max(float((a - out[m]).abs().max() / a.abs().max()) for m in ("peer", "peer_serial"))
    assert err < 1e-5, err            # split-K atomics reorder sums between runs; the exchange itself is exact
    dist.barrier()
    if rank == 0:
        print(f"PEER_OK world={world} train_step_rel_err={err:.2e}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
