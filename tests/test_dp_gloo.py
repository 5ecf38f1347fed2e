"""CPU, world_size 2, gloo: the data-parallel contract of scat_b200.dp.

Each rank runs the head's train-step arithmetic on its batch shard (here through the CPU oracle, as the
checker -- the CUDA kernels need a GPU), scales by 1/world, and sums the flat gradient bucket with the same
FlatGradBucket.all_reduce the GPU path uses over NCCL.  The result must equal one process on the full batch.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scat_b200 import dp, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle import head_oracle
    try:
        # every rank draws a DIFFERENT init; broadcast must make them identical to rank 0's
        W = synth.make_head_weights(8, seed=100 + rank)
        params = [torch.nn.Parameter(torch.from_numpy(v)) for v in W.values()]
        dp.broadcast_parameters(params, src=0)
        W0 = synth.make_head_weights(8, seed=100)
        assert all(torch.equal(p.data, torch.from_numpy(v)) for p, v in zip(params, W0.values()))
        mask = dp.broadcast_mask([3, 7, 11, 20] if rank == 0 else [0, 1, 2, 4], src=0)
        assert mask == [3, 7, 11, 20]

        GB = 4
        x2, mf, labels = synth.make_head_inputs(GB, 5)
        lo, hi = dp.shard_batch(GB, rank, world)
        P = {k: p.data for k, p in zip(W.keys(), params)}
        mean = torch.from_numpy(synth.make_mean_params("hand"))
        step = head_oracle.train_step(P, torch.from_numpy(x2[lo:hi]), torch.from_numpy(mf[lo:hi]),
                                      torch.from_numpy(labels[lo:hi]), mean, mask_idx=mask, pl_reg=True)
        bucket = dp.FlatGradBucket(params)
        for v, k in zip(bucket.views, W.keys()):
            v.copy_(step["grads"][k] / world)            # what grad_scale = 1/world does inside the loss kernel
        bucket.all_reduce()
        if rank == 0:
            full = head_oracle.train_step(P, torch.from_numpy(x2), torch.from_numpy(mf), torch.from_numpy(labels),
                                          mean, mask_idx=mask, pl_reg=True)
            ref = torch.cat([full["grads"][k].reshape(-1) for k in W.keys()])
            got = torch.cat([v.reshape(-1) for v in bucket.views])
            err = float((got - ref).norm() / ref.norm())
            np.save(os.path.join(out_dir, "err.npy"), np.array([err, bucket.payload_bytes(), bucket.nbytes()]))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    err, nbytes, padded = np.load(tmp_path / "err.npy")
    assert err < 1e-5
    assert int(nbytes) == 3795099 * 4          # one 15.18 MB fp32 bucket per step
    assert 0 <= int(padded) - int(nbytes) < 35 * 128           # + alignment gaps (each tensor on a 128-byte boundary)
