"""GPU tests of the peer-memory data-parallel optimizer step (sc_peer_*, sc_dp_barrier, sc_dp_reduce_adam;
SURVEY.md §8e / §8f rank 2) on ONE device:

* two "ranks" live in one process (two sets of buffers, two streams, grids capped so that both kernels are resident):
  the full barrier -> reduce-scatter + sharded Adam + all-gather -> barrier protocol must reproduce
  sc_adam_flat(g0 + g1) bit for bit on both ranks, step after step, eagerly and as replayed CUDA graphs;
* a peer allocation exported by this process is mapped by a CHILD process (cudaIpc handle), read and written there.

The real two-GPU run (NVLink, one process per GPU, against the NCCL schedule) is scripts/gpu_dp_fused_check.py ->
profiles/r2_dp_fused_check_n2.json.
"""
import ctypes as C
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _arr(ptrs):
    return (C.c_void_p * len(ptrs))(*ptrs)


class _Rank:
    def __init__(self, n, seed):
        from multimodal_sc_b200.peer import peer_empty
        g = torch.Generator().manual_seed(seed)
        self.g32 = peer_empty(n, torch.float32, "cuda")
        self.p32 = peer_empty(n, torch.float32, "cuda")
        self.p16 = peer_empty(n, torch.bfloat16, "cuda")
        self.flags = peer_empty(16, torch.int64, "cuda")
        self.m = torch.zeros(n, device="cuda")
        self.v = torch.zeros(n, device="cuda")
        self.epoch = torch.zeros(1, dtype=torch.int64, device="cuda")
        self.stream = torch.cuda.Stream()
        self.gen = g


def _launch_step(lib, ranks, r, n, step, step_dev=None):
    from multimodal_sc_b200._lib import check
    me = ranks[r]
    world = len(ranks)
    fl = _arr([x.flags.data_ptr() for x in ranks])
    st = me.stream.cuda_stream
    check(lib.sc_dp_barrier(fl, me.epoch.data_ptr(), world, r, st), "sc_dp_barrier")
    check(lib.sc_dp_reduce_adam(_arr([x.g32.data_ptr() for x in ranks]), _arr([x.p32.data_ptr() for x in ranks]),
                                _arr([x.p16.data_ptr() for x in ranks]), me.m.data_ptr(), me.v.data_ptr(), world, r, n,
                                3e-4, 0.9, 0.999, 1e-8, step, None if step_dev is None else step_dev.data_ptr(), 16, st),
          "sc_dp_reduce_adam")
    check(lib.sc_dp_barrier(fl, me.epoch.data_ptr(), world, r, st), "sc_dp_barrier")


@pytest.mark.parametrize("world,n", [(2, 4096 * 4 + 8), (3, 100_004), (1, 1024)])
def test_reduce_adam_over_peer_buffers_equals_adam_of_the_summed_gradients(cuda_lib, world, n):
    from multimodal_sc_b200 import ops
    from multimodal_sc_b200.peer import preload_kernels
    lib = cuda_lib
    preload_kernels("cuda")                          # one host thread drives every rank here: no code loading mid-protocol
    ranks = [_Rank(n, 10 + r) for r in range(world)]
    p0 = torch.randn(n, generator=torch.Generator().manual_seed(1)).cuda()
    for x in ranks:
        x.p32.copy_(p0)
        x.p16.copy_(p0.bfloat16())
    # reference: one bucket, Adam on the rank-ordered sum
    rp, rm, rv, rb = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda"), p0.bfloat16()
    torch.cuda.synchronize()
    for step in range(1, 4):
        gs = [torch.randn(n, generator=x.gen).cuda() * (0.1 * step) for x in ranks]
        for x, g in zip(ranks, gs):
            x.g32.copy_(g)
        torch.cuda.synchronize()
        for r in range(world):                       # both ranks' launches are in flight together (own streams)
            _launch_step(lib, ranks, r, n, step)
        torch.cuda.synchronize()
        gsum = gs[0].clone()
        for g in gs[1:]:
            gsum += g                                # rank order, like the kernel
        ops.adam_flat(rp, gsum, rm, rv, rb, 3e-4, 0.9, 0.999, 1e-8, step)
        torch.cuda.synchronize()
        n4 = n // 4
        for r, x in enumerate(ranks):
            assert torch.equal(x.p32, rp), f"step {step} rank {r}: fp32 parameters differ"
            assert torch.equal(x.p16.view(torch.int16), rb.view(torch.int16)), f"step {step} rank {r}: bf16 shadow differs"
            base, rem = divmod(n4, world)
            lo = 4 * (r * base + min(r, rem))
            hi = lo + 4 * (base + (1 if r < rem else 0))
            assert torch.equal(x.m[lo:hi], rm[lo:hi]) and torch.equal(x.v[lo:hi], rv[lo:hi]), "moments of the own slice"
            assert int(x.epoch.item()) == 2 * step
            assert float(x.m[:lo].abs().sum()) == 0.0 and float(x.m[hi:].abs().sum()) == 0.0, "foreign slices untouched"


def test_reduce_adam_replays_as_cuda_graphs_with_device_step_counter(cuda_lib):
    from multimodal_sc_b200 import ops
    from multimodal_sc_b200.peer import preload_kernels
    lib = cuda_lib
    preload_kernels("cuda")
    world, n = 2, 65536
    ranks = [_Rank(n, 20 + r) for r in range(world)]
    p0 = torch.randn(n, generator=torch.Generator().manual_seed(2)).cuda()
    for x in ranks:
        x.p32.copy_(p0)
        x.p16.copy_(p0.bfloat16())
        x.g32.copy_(torch.randn(n, generator=x.gen).cuda())
    dev_step = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in ranks]
    rp, rm, rv, rb = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda"), p0.bfloat16()
    gsum = ranks[0].g32.clone() + ranks[1].g32
    torch.cuda.synchronize()
    graphs = []
    for r, x in enumerate(ranks):                    # capture only records: no rank waits for the other here
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=x.stream):
            dev_step[r].add_(1)
            _launch_step(lib, ranks, r, n, 0, step_dev=dev_step[r])
        graphs.append(g)
    torch.cuda.synchronize()
    for step in range(1, 4):
        for r, x in enumerate(ranks):
            with torch.cuda.stream(x.stream):
                graphs[r].replay()
        torch.cuda.synchronize()
        ops.adam_flat(rp, gsum, rm, rv, rb, 3e-4, 0.9, 0.999, 1e-8, step)
        torch.cuda.synchronize()
        for r, x in enumerate(ranks):
            assert torch.equal(x.p32, rp), f"replay {step} rank {r}"
            assert int(x.epoch.item()) == 2 * step


_CHILD = r"""
import ctypes as C, sys, torch
sys.path.insert(0, {root!r})
from multimodal_sc_b200 import _lib, peer
h = bytes.fromhex(sys.argv[1]); n = int(sys.argv[2])
torch.cuda.init()
p = peer.import_handle(h, torch.device("cuda", 0))
class V:
    __cuda_array_interface__ = {{"shape": (n,), "typestr": "<f4", "data": (p, False), "version": 2, "strides": None}}
t = torch.as_tensor(V(), device="cuda:0")
assert torch.equal(t.cpu(), torch.arange(n, dtype=torch.float32)), "child does not see the parent's data"
t.mul_(2.0)
torch.cuda.synchronize()
del t
_lib.check(_lib.load().sc_peer_close(p), "sc_peer_close")
print("child ok")
"""


def test_peer_allocation_is_mapped_by_another_process(cuda_lib):
    from multimodal_sc_b200 import peer
    n = 1 << 16
    t = peer.peer_empty(n, torch.float32, "cuda")
    assert peer.is_peer(t) and float(t.abs().sum()) == 0.0          # zero-filled
    t.copy_(torch.arange(n, dtype=torch.float32))
    torch.cuda.synchronize()
    h = peer.export_handle(t)
    assert len(h) == 64
    r = subprocess.run([sys.executable, "-c", _CHILD.format(root=ROOT), h.hex(), str(n)], capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0 and "child ok" in r.stdout, r.stdout + r.stderr
    torch.cuda.synchronize()
    assert torch.equal(t.cpu(), 2.0 * torch.arange(n, dtype=torch.float32)), "the child's writes are not visible"


def test_policy_in_peer_memory_trains_like_the_default_one(cuda_lib):
    """FlatParams(peer=True): same forward / step results as ordinary torch allocations (one rank)"""
    from multimodal_sc_b200 import synthetic
    from multimodal_sc_b200.model import SemComPolicy
    from multimodal_sc_b200.ppo import PPOAgent
    from multimodal_sc_b200.spec import SpecV0
    from multimodal_sc_b200.peer import is_peer
    sp = SpecV0()
    mb = {k: v.cuda() for k, v in synthetic.minibatch(sp, 8, seed=5).items()}
    args = (mb["cam"], mb["bev"], mb["actions"], mb["logp_old"], mb["adv"], None, mb["ret"])
    out = []
    for peer_mem in (False, True):
        pol = SemComPolicy(sp, seed=3, peer=peer_mem)
        assert is_peer(pol.flat.p32) == peer_mem and pol.flat.peer == peer_mem
        pol.engine.deterministic = True
        ag = PPOAgent(pol, sp, minibatch=8, seed=77)
        stats = [ag.minibatch_step(*args) for _ in range(2)]
        out.append((pol.flat.p32.clone(), stats))
    # not bit-equal: the column-sum / statistics atomics of the memory-bound kernels are order-dependent
    # (Adam's first steps move every weight by ~lr * sign(g): a gradient that is zero up to summation noise may flip)
    d = (out[0][0] - out[1][0]).abs()
    assert float((d > 0.02 * sp.lr).float().mean()) < 0.01 and float(d.max()) <= 2.5 * sp.lr
    for a, b in zip(out[0][1], out[1][1]):
        for k in a:
            assert abs(a[k] - b[k]) <= 1e-4 * max(1.0, abs(a[k])), (k, a[k], b[k])


def test_peer_reduce_adam_class_with_a_one_rank_group(cuda_lib, tmp_path):
    """peer.PeerReduceAdam end to end in this process: handle export, the collective construction (one-rank gloo group),
    step() == sc_adam_flat bit for bit over the model's real flat bucket, gather_moments(), close()"""
    import torch.distributed as dist
    from multimodal_sc_b200 import ops
    from multimodal_sc_b200.model import FlatParams
    from multimodal_sc_b200.peer import PeerReduceAdam
    from multimodal_sc_b200.spec import SpecV0
    sp = SpecV0()
    assert not dist.is_initialized()
    dist.init_process_group("gloo", init_method=f"file://{tmp_path}/pg", rank=0, world_size=1)
    try:
        flat = FlatParams(sp, torch.device("cuda", 0), peer=True)
        g = torch.Generator(device="cuda").manual_seed(9)
        flat.p32.copy_(torch.randn(flat.numel, generator=g, device="cuda") * 0.05)
        flat.sync_shadow()
        red = PeerReduceAdam(flat)
        assert (red.lo, red.hi) == (0, flat.numel) and red.world == 1
        rp, rm, rv, rb = flat.p32.clone(), torch.zeros_like(flat.m), torch.zeros_like(flat.v), flat.p16.clone()
        for step in (1, 2, 3):
            flat.g32.copy_(torch.randn(flat.numel, generator=g, device="cuda") * 1e-3)
            red.step(sp.lr, sp.beta1, sp.beta2, sp.adam_eps, step)
            ops.adam_flat(rp, flat.g32, rm, rv, rb, sp.lr, sp.beta1, sp.beta2, sp.adam_eps, step)
            torch.cuda.synchronize()
            assert torch.equal(flat.p32, rp) and torch.equal(flat.p16.view(torch.int16), rb.view(torch.int16))
            assert torch.equal(flat.m, rm) and torch.equal(flat.v, rv)
        assert int(red.epoch) == 6
        red.gather_moments()
        red.close()
    finally:
        dist.destroy_process_group()

