"""NVLink peer memory and the fused gradient exchange + optimizer of the data-parallel step
(include/semcom_sm100.h: sc_peer_*, sc_dp_barrier, sc_dp_reduce_adam).

`peer_empty` allocates a flat device buffer through the library (plain cudaMalloc, exportable) and wraps it as a torch
tensor; `PeerReduceAdam` exchanges the handles of the flat gradient / parameter buckets between the ranks of a process
group once and then runs an optimizer step as   barrier -> reduce-scatter + sharded Adam + all-gather -> barrier
over peer loads / stores — three launches of the library's own kernels, CUDA-graph capturable, no NCCL on the data
path.  The NCCL all-reduce + sc_adam_flat schedule stays available (`PPOAgent(fused_dp=False)`, `SC_DP_FUSED=0`).
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, stream_ptr


class _PeerAllocation:
    """owns one sc_peer_alloc'd buffer; torch sees it through __cuda_array_interface__ as bytes"""

    def __init__(self, nbytes: int, device: torch.device):
        self.nbytes, self.device = int(nbytes), device
        p = C.c_void_p()
        with torch.cuda.device(device):
            check(_lib.load().sc_peer_alloc(C.byref(p), self.nbytes), "sc_peer_alloc")
        self.ptr = int(p.value)
        self.__cuda_array_interface__ = {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False),
                                         "version": 2, "strides": None}

    def __del__(self):
        ptr, self.ptr = getattr(self, "ptr", 0), 0
        if ptr:
            try:
                _lib.load().sc_peer_free(ptr)
            except Exception:                       # interpreter shutdown
                pass


def peer_empty(numel: int, dtype: torch.dtype, device) -> torch.Tensor:
    """zero-filled flat tensor in exportable device memory (the tensor keeps the allocation alive)"""
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.SemcomError("peer memory is device memory (no CPU fallback)")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    esize = torch.empty((), dtype=dtype).element_size()
    alloc = _PeerAllocation(max(numel * esize, 16), device)
    t = torch.as_tensor(alloc, device=device)[: numel * esize].view(dtype)
    assert t.data_ptr() == alloc.ptr and t.device == device, (t.data_ptr(), alloc.ptr, t.device, device)
    t._sc_peer_alloc = alloc                        # reachable for is_peer() / export
    return t


def is_peer(t: torch.Tensor) -> bool:
    return getattr(t, "_sc_peer_alloc", None) is not None


def export_handle(t: torch.Tensor) -> bytes:
    buf = C.create_string_buffer(_lib.SC_PEER_HANDLE_BYTES)
    check(_lib.load().sc_peer_export(t.data_ptr(), buf), "sc_peer_export")
    return buf.raw


def import_handle(handle: bytes, device) -> int:
    p = C.c_void_p()
    with torch.cuda.device(device):
        check(_lib.load().sc_peer_import(handle, C.byref(p)), "sc_peer_import")
    return int(p.value)


def _ptr_array(ptrs):
    return (C.c_void_p * len(ptrs))(*ptrs)


def preload_kernels(device) -> None:
    """Run the barrier and the reduce + Adam kernel once on a 1-rank dummy problem.  CUDA loads a kernel's code at its
    first launch, and that load can wait for running kernels: a rank whose barrier kernel is already spinning must not
    find out at the next launch that the code is not resident yet (with several ranks driven by ONE host thread — the
    single-device tests — that would be a deadlock; with one process per GPU it is only a stall)."""
    device = torch.device(device)
    lib = _lib.load()
    with torch.cuda.device(device):
        z = torch.zeros(5 * 16, dtype=torch.float32, device=device)
        g, p, m, v, sh = (z[16 * i:16 * (i + 1)] for i in range(5))
        fl = torch.zeros(_lib.SC_DP_MAX_WORLD, dtype=torch.int64, device=device)
        ep = torch.zeros(1, dtype=torch.int64, device=device)
        st = torch.cuda.current_stream(device).cuda_stream
        check(lib.sc_dp_barrier(_ptr_array([fl.data_ptr()]), ep.data_ptr(), 1, 0, st), "sc_dp_barrier")
        check(lib.sc_dp_reduce_adam(_ptr_array([g.data_ptr()]), _ptr_array([p.data_ptr()]), _ptr_array([sh.data_ptr()]),
                                    m.data_ptr(), v.data_ptr(), 1, 0, 16, 1e-3, 0.9, 0.999, 1e-8, 1, None, 1, st),
              "sc_dp_reduce_adam")
        torch.cuda.synchronize(device)


class PeerReduceAdam:
    """Fused data-parallel optimizer step over peer memory.

    flat: FlatParams whose g32 / p32 / p16 were allocated with peer_empty (FlatParams(peer=True)).  Construction is a
    collective over `pg` (handles travel by all_gather_object).  step() enqueues the three launches on the current
    stream.  Optimizer state is sharded: flat.m / flat.v are only maintained for this rank's slice [lo, hi);
    gather_moments() completes them on every rank (checkpoints)."""

    def __init__(self, flat, pg=None, *, max_ctas: int = 0):
        self.flat, self.pg = flat, pg
        self.world, self.rank = dist.get_world_size(pg), dist.get_rank(pg)
        if self.world > _lib.SC_DP_MAX_WORLD:
            raise _lib.SemcomError(f"PeerReduceAdam supports up to {_lib.SC_DP_MAX_WORLD} ranks")
        for name in ("g32", "p32", "p16"):
            if not is_peer(getattr(flat, name)):
                raise _lib.SemcomError(f"flat.{name} is not in peer memory: build the policy with peer memory "
                                       "(SemComPolicy(..., peer=True)) — there is no staging fallback")
        self.device = flat.g32.device
        self.max_ctas = int(max_ctas)
        self.flags = peer_empty(_lib.SC_DP_MAX_WORLD, torch.int64, self.device)
        self.epoch = torch.zeros(1, dtype=torch.int64, device=self.device)
        local = {"g32": flat.g32, "p32": flat.p32, "p16": flat.p16, "flags": self.flags}
        # Every rank takes part in every collective below even when its own export / import fails: the failure is agreed on
        # afterwards (one all-reduce) and raised on ALL ranks, so a caller may fall back to the NCCL schedule collectively.
        err = None
        try:
            mine = {k: export_handle(t) for k, t in local.items()}
        except _lib.SemcomError as e:
            err, mine = e, {}
        mine["numel"] = int(flat.numel)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=pg)
        self._imported = []
        ptrs = {k: [] for k in local}
        for r, hs in enumerate(everyone):
            if err is not None:
                break
            try:
                if hs["numel"] != flat.numel:
                    raise _lib.SemcomError(f"rank {r} holds {hs['numel']} flat parameters, this rank {flat.numel}")
                for k, t in local.items():
                    if r == self.rank:
                        ptrs[k].append(t.data_ptr())
                    else:
                        if k not in hs:
                            raise _lib.SemcomError(f"rank {r} could not export its buffers")
                        if os.environ.get("SC_PEER_TEST_FAIL_IMPORT") == str(self.rank):   # test hook: this rank cannot map
                            raise _lib.SemcomError("import refused (SC_PEER_TEST_FAIL_IMPORT)")
                        p = import_handle(hs[k], self.device)
                        self._imported.append(p)
                        ptrs[k].append(p)
            except _lib.SemcomError as e:
                err = e
        failed = torch.tensor([0 if err is None else 1], dtype=torch.int32, device=self.device)
        dist.all_reduce(failed, group=pg)
        if int(failed.item()) > 0:
            self.close()
            raise _lib.SemcomError(f"peer-memory mapping failed on {int(failed.item())} of {self.world} ranks"
                                   + (f" (this rank: {err})" if err is not None else ""))
        self._g, self._p32, self._p16, self._fl = (_ptr_array(ptrs[k]) for k in ("g32", "p32", "p16", "flags"))
        n4 = flat.numel // 4
        base, rem = divmod(n4, self.world)
        lo4 = self.rank * base + min(self.rank, rem)
        self.lo, self.hi = 4 * lo4, 4 * (lo4 + base + (1 if self.rank < rem else 0))
        self.moments_complete = self.world == 1
        preload_kernels(self.device)
        dist.barrier(group=pg)                       # every rank has mapped every buffer before the first step

    def barrier(self):
        check(_lib.load().sc_dp_barrier(self._fl, self.epoch.data_ptr(), self.world, self.rank,
                                        stream_ptr(self.flat.g32)), "sc_dp_barrier")

    def step(self, lr, beta1, beta2, eps, step: int, step_dev=None):
        f = self.flat
        self.barrier()
        check(_lib.load().sc_dp_reduce_adam(self._g, self._p32, self._p16, f.m.data_ptr(), f.v.data_ptr(), self.world,
                                            self.rank, f.numel, lr, beta1, beta2, eps, step,
                                            None if step_dev is None else step_dev.data_ptr(), self.max_ctas,
                                            stream_ptr(f.g32)), "sc_dp_reduce_adam")
        self.barrier()
        self.moments_complete = self.world == 1

    def gather_moments(self):
        """complete flat.m / flat.v on every rank from the owners' slices (before a checkpoint is written)"""
        if self.moments_complete:
            return
        f = self.flat
        n4 = f.numel // 4
        for buf in (f.m, f.v):
            for r in range(self.world):
                base, rem = divmod(n4, self.world)
                lo4 = r * base + min(r, rem)
                hi4 = lo4 + base + (1 if r < rem else 0)
                dist.broadcast(buf[4 * lo4:4 * hi4], src=dist.get_global_rank(self.pg, r) if self.pg is not None else r,
                               group=self.pg)
        self.moments_complete = True

    def close(self):
        torch.cuda.synchronize(self.device)
        for p in self._imported:
            _lib.load().sc_peer_close(p)
        self._imported = []
