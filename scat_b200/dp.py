"""Data-parallel plumbing for the head: flat gradient bucket + one all-reduce per step.

The reference has no distributed code at all (DistributedDataParallel is imported at train.py:18 and never
used); batch sharding is new work.  The head's samples are independent in forward and backward, so the only
exchange is the sum of parameter gradients.  Device-agnostic on purpose: the world_size-2 ``gloo`` tests run
this exact code on CPU tensors, the GPU path runs it over NCCL / NVLink.
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


ALIGN = 32   # floats: every tensor of a flat buffer starts on a 128-byte boundary (TMA / 128-bit accesses need 16)


def flat_layout(tensors, align: int = ALIGN):
    """(offsets, total) of the tensors laid out one after the other, each start rounded up to ``align`` elements."""
    offsets, off = [], 0
    for t in tensors:
        offsets.append(off)
        off = (off + t.numel() + align - 1) // align * align
    return offsets, off


class FlatGradBucket:
    """One contiguous fp32 buffer whose slices are the ``.grad`` of the given parameters (``flat_layout`` order and
    alignment; the gaps stay zero).  ``flat`` lets the caller supply the storage (peer-mapped memory for the NVLink
    all-reduce, see ``PeerMemory``)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], flat: torch.Tensor = None):
        self.params: List[torch.nn.Parameter] = list(params)
        if not self.params:
            raise ValueError("FlatGradBucket needs at least one parameter")
        dev, dt = self.params[0].device, self.params[0].dtype
        self.offsets, total = flat_layout(self.params)
        if flat is None:
            flat = torch.zeros(total, device=dev, dtype=dt)
        elif flat.numel() != total or flat.dtype != dt or flat.device != dev or not flat.is_contiguous():
            raise ValueError("FlatGradBucket: supplied storage does not match the parameters")
        self.flat = flat
        self.views: List[torch.Tensor] = []
        for p, off in zip(self.params, self.offsets):
            v = self.flat[off: off + p.numel()].view_as(p)
            p.grad = v
            self.views.append(v)

    def payload_bytes(self) -> int:
        """Bytes of actual gradients (without the alignment gaps)."""
        return sum(p.numel() for p in self.params) * self.flat.element_size()

    def nbytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()

    def all_reduce(self, group=None, async_op: bool = False, lo: int = 0, hi: int = None):
        """Sum the bucket (or its slice [lo, hi)) over the process group (no-op for a single process)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        part = self.flat if (lo == 0 and hi is None) else self.flat[lo:hi]
        return dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def _flat_view(tensors: List[torch.Tensor]):
    """The 1-D tensor whose ``flat_layout`` slices the given tensors already are, or None."""
    first = tensors[0]
    offsets, total = flat_layout(tensors)
    base = first.storage_offset()
    for t, off in zip(tensors, offsets):
        if (not t.is_contiguous() or t.dtype != first.dtype or t.device != first.device
                or t.untyped_storage().data_ptr() != first.untyped_storage().data_ptr()
                or t.storage_offset() != base + off):
            return None
    if (first.untyped_storage().nbytes() // first.element_size()) - base < total:
        return None
    return torch.empty(0, dtype=first.dtype, device=first.device).set_(first.untyped_storage(), base, (total,))


def flatten_parameters(params: Iterable[torch.nn.Parameter]) -> torch.Tensor:
    """Make ``p.data`` of every parameter a view of ONE flat buffer (``flat_layout`` order and alignment, zero gaps)
    and return that buffer.  Idempotent: parameters that are already laid out that way are left alone.  Values are
    preserved; anything that captured the old ``data_ptr`` (a CUDA graph) must be built afterwards."""
    params = list(params)
    flat = _flat_view([p.data for p in params])
    if flat is not None:
        return flat
    offsets, total = flat_layout(params)
    flat = torch.zeros(total, dtype=params[0].dtype, device=params[0].device)
    for p, off in zip(params, offsets):
        n = p.numel()
        flat[off: off + n].copy_(p.data.reshape(-1))
        p.data = flat[off: off + n].view(p.shape)
    return flat


def flat_gradients(params: Iterable[torch.nn.Parameter]):
    """The flat buffer whose slices the ``.grad`` of the parameters are (a FlatGradBucket's), or None."""
    params = list(params)
    if any(p.grad is None for p in params):
        return None
    return _flat_view([p.grad for p in params])


class _DevicePtr:
    """A raw device allocation seen by torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, address: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (address, False), "version": 2}


class PeerMemory:
    """The gradient bucket of every rank of this node, mapped into this process, and the one-kernel all-reduce over
    it (csrc/dp_allreduce.cu, include/scat_b200.h scat_peer_*).

    Each rank allocates ``n_floats`` (rounded up to 4) of device memory plus a signal area through the library
    (plain cudaMalloc, so the CUDA IPC handle covers exactly the allocation), the 64-byte handles travel through
    one ``all_gather`` of the process group, and every rank opens its peers'.  ``flat`` is the local bucket as a
    torch tensor; ``enqueue(stream)`` launches the all-reduce of ``[lo, hi)`` (capturable into a CUDA graph).
    All ranks must sit on one node with peer access (NVLink / NVSwitch) and issue the same sequence of calls.
    """

    def __init__(self, n_floats: int, device, group=None):
        import ctypes as C

        from . import _lib
        self._C, self._check = C, _lib.check
        self.lib = _lib.load()
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world not in (1, 2, 4, 8):
            raise ValueError(f"PeerMemory: world size {self.world} (1, 2, 4 or 8 GPUs of one node)")
        self.device = torch.device(device)
        self.n = int(n_floats)
        self.n_pad = (self.n + 3) // 4 * 4
        self._opened: List[int] = []
        with torch.cuda.device(self.device):
            self._bucket = self._alloc(self.n_pad * 4)
            self._signal = self._alloc(int(self.lib.scat_peer_signal_bytes()))
            handles = torch.empty(128, dtype=torch.uint8)
            buf = (C.c_uint8 * 64)()
            for i, p in enumerate((self._bucket, self._signal)):
                self._check(self.lib.scat_peer_export(p, C.cast(buf, C.c_void_p)), "scat_peer_export")
                handles[64 * i: 64 * i + 64] = torch.frombuffer(bytearray(buf), dtype=torch.uint8)
            if self.world > 1:
                mine = handles.to(self.device)
                every = torch.empty(self.world * 128, dtype=torch.uint8, device=self.device)
                dist.all_gather_into_tensor(every, mine, group=group)
                every = every.cpu().view(self.world, 128)
            self.buckets, self.signals = [], []
            failure = None
            try:
                for r in range(self.world):
                    if r == self.rank:
                        self.buckets.append(self._bucket)
                        self.signals.append(self._signal)
                        continue
                    peer = self._peer_device(r)
                    if (torch.cuda.device_count() >= self.world and peer != self.device.index
                            and not torch.cuda.can_device_access_peer(self.device.index, peer)):
                        raise RuntimeError(f"no peer access from rank {self.rank} to rank {r}")
                    for i, dst in enumerate((self.buckets, self.signals)):
                        raw = (C.c_uint8 * 64)(*every[r, 64 * i: 64 * i + 64].tolist())
                        out = C.c_void_p()
                        self._check(self.lib.scat_peer_open(C.cast(raw, C.c_void_p), C.byref(out)), "scat_peer_open")
                        self._opened.append(out.value)
                        dst.append(out.value)
            except Exception as e:               # decide together: a rank that gave up must not leave the others waiting
                failure = e
            if self.world > 1:
                ok = torch.tensor([0 if failure else 1], dtype=torch.int32, device=self.device)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
                if int(ok.item()) == 0:
                    self._release()
                    raise RuntimeError(f"PeerMemory: peer mapping failed on at least one rank"
                                       f"{' (here: ' + str(failure) + ')' if failure else ''}")
            elif failure:
                raise failure
            self._bucket_arr = (C.c_void_p * self.world)(*self.buckets)
            self._signal_arr = (C.c_void_p * self.world)(*self.signals)
            self._holder = _DevicePtr(self._bucket, self.n_pad)
            self.flat_padded = torch.as_tensor(self._holder, device=self.device)
            self.flat = self.flat_padded[: self.n]
            if self.world > 1:
                dist.barrier(group=group)          # every peer has mapped everything before the first kernel
            torch.cuda.synchronize()

    def _peer_device(self, rank: int) -> int:
        """CUDA device index of ``rank`` on this node (one process per GPU, LOCAL_RANK == device index)."""
        return (self.device.index - self.rank + rank) % max(torch.cuda.device_count(), 1)

    def _release(self):
        for p in self._opened:
            self.lib.scat_peer_close(p)
        self._opened = []
        self.flat = self.flat_padded = self._holder = None
        for p in (self._bucket, self._signal):
            if p is not None:
                self.lib.scat_peer_free(p)
        self._bucket = self._signal = None

    def _alloc(self, nbytes: int) -> int:
        out = self._C.c_void_p()
        self._check(self.lib.scat_peer_alloc(nbytes, self._C.byref(out)), "scat_peer_alloc")
        return out.value

    def enqueue(self, stream_ptr: int, lo: int = 0, hi: int = None, last: bool = True):
        """Sum elements [lo, hi) (multiples of 4; hi=None: the whole bucket) over all ranks, in place, on ``stream``.
        ``last=False``: another exchange follows on the same stream and closes this one too (scat_peer_allreduce_part)."""
        hi = self.n_pad if hi is None else hi
        self._check(self.lib.scat_peer_allreduce_part(self._bucket_arr, self._signal_arr, self.rank, self.world, lo, hi,
                                                      1 if last else 0, stream_ptr), "scat_peer_allreduce_part")

    def error_word(self):
        """Device address (ctypes void pointer) of this rank's sticky time-out flag, for scat_adam_step's abort_flag."""
        return self._C.c_void_p(self.lib.scat_peer_error_word(self._signal))

    def timed_out(self) -> bool:
        """True when a kernel gave up waiting for a peer (synchronises the device)."""
        out = self._C.c_int32(0)
        self._check(self.lib.scat_peer_error(self._signal, self._C.byref(out)), "scat_peer_error")
        return bool(out.value)

    def close(self):
        """Unmap the peers and free the local memory (every rank, after a barrier: nobody may still be reading)."""
        if self._bucket is None:
            return
        torch.cuda.synchronize()
        if self.world > 1 and dist.is_initialized():
            dist.barrier(group=self.group)
        self._release()


def world_size(group=None) -> int:
    return dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1


def broadcast_parameters(params: Iterable[torch.Tensor], src: int = 0, group=None):
    """Every rank starts from rank ``src``'s weights (each process would otherwise draw its own init)."""
    if world_size(group) == 1:
        return
    for p in params:
        dist.broadcast(p.data if isinstance(p, torch.nn.Parameter) else p, src=src, group=group)


def broadcast_mask(mask, src: int = 0, group=None, device="cpu"):
    """Share rank ``src``'s host-drawn token mask so all shards see the same indices (parity with one big batch)."""
    if world_size(group) == 1:
        return list(mask)
    t = torch.tensor(list(mask), dtype=torch.int32, device=device)
    dist.broadcast(t, src=src, group=group)
    return t.cpu().tolist()


def shard_batch(global_batch: int, rank: int, world: int):
    """Contiguous batch shard [lo, hi) of rank ``rank``; remainders go to the lowest ranks."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
