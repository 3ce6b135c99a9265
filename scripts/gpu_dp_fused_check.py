"""torchrun --nproc-per-node N scripts/gpu_dp_fused_check.py [batch] [steps]

The data-parallel optimizer step over NVLink peer memory (peer.PeerReduceAdam: sc_dp_barrier -> sc_dp_reduce_adam ->
sc_dp_barrier) against the NCCL schedule (all-reduce of the flat fp32 bucket -> sc_adam_flat), same seeds, same per-rank
minibatches, real GPUs, one process per GPU:

* parameters after K optimizer steps: fused vs NCCL (differences = fp32 summation order of the gradients only), and
  bit-identical across the ranks of the fused run (each element is computed by exactly one rank);
* optimizer state: the owner's slice of m / v equals the NCCL run's, gather_moments() completes it everywhere;
* time: the exchange + optimizer part of an eager step (CUDA events), and whole steps replayed as CUDA graphs, A/B in
  the same process.
Writes gpurun_out/dp_fused_check_n<world>.json (rank 0)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_sc_b200 import synthetic  # noqa: E402
from multimodal_sc_b200.model import SemComPolicy  # noqa: E402
from multimodal_sc_b200.ppo import PPOAgent  # noqa: E402
from multimodal_sc_b200.spec import SpecV0  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 20
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
sp = SpecV0()
mbs = [{k: v.to(dev) for k, v in synthetic.minibatch(sp, B, seed=100 + 2 * rank + i, obs_dtype=torch.bfloat16).items()}
       for i in range(2)]


def args_of(mb):
    return (mb["cam"], mb["bev"], mb["actions"], mb["logp_old"], mb["adv"], None, mb["ret"])


def make(fused):
    pol = SemComPolicy(sp, device=dev, seed=0, peer=fused)
    ag = PPOAgent(pol, sp, minibatch=B, snr_db=10.0, channel="awgn", seed=1234, fused_dp=fused)
    assert (ag._fused is not None) == fused
    return pol, ag


def timed_graph_steps(ag, n):
    for i in range(4):                                            # capture (one graph per input set) + warm-up
        ag.minibatch_step(*args_of(mbs[i % 2]), sample_offset=rank * B, sync_stats=False, graph="inplace")
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        ag.minibatch_step(*args_of(mbs[i % 2]), sample_offset=rank * B, sync_stats=False, graph="inplace")
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


out = {"world": world, "batch_per_rank": B, "flat_params": None}
pol_f, ag_f = make(True)
pol_n, ag_n = make(False)
out["flat_params"] = pol_f.flat.numel
out["slice"] = [ag_f._fused.lo, ag_f._fused.hi]

# ---- parity: K eager steps on identical inputs -------------------------------------------------------------------
K = 4
ex_ms = {"fused": [], "nccl": []}
for i in range(K):
    for name, ag in (("fused", ag_f), ("nccl", ag_n)):
        ag.minibatch_step(*args_of(mbs[i % 2]), sample_offset=rank * B, sync_stats=False)
        ex_ms[name].append(ag.allreduce_ms())
torch.cuda.synchronize()
d = (pol_f.flat.p32 - pol_n.flat.p32).abs()
out["params_fused_vs_nccl"] = {"max_abs": float(d.max()), "in_units_of_lr": float(d.max()) / sp.lr,
                               "frac_above_0.02_lr": float((d > 0.02 * sp.lr).float().mean())}
d16 = (pol_f.flat.p16.float() - pol_n.flat.p16.float()).abs()
out["shadow_fused_vs_nccl_max_abs"] = float(d16.max())
# the same bits on every rank (fused run): compare a checksum of the raw words
chk = pol_f.flat.p32.view(torch.int32).long().sum().reshape(1)
chk16 = pol_f.flat.p16.view(torch.int16).long().sum().reshape(1)
alls = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(alls, chk)
alls16 = [torch.zeros_like(chk16) for _ in range(world)]
dist.all_gather(alls16, chk16)
out["fused_ranks_bit_identical"] = bool(all(int(a) == int(alls[0]) for a in alls) and
                                        all(int(a) == int(alls16[0]) for a in alls16))
lo, hi = ag_f._fused.lo, ag_f._fused.hi
out["own_slice_moments_vs_nccl"] = {
    "m_max_abs": float((pol_f.flat.m[lo:hi] - pol_n.flat.m[lo:hi]).abs().max()),
    "m_scale": float(pol_n.flat.m[lo:hi].abs().max()),
    "v_max_abs": float((pol_f.flat.v[lo:hi] - pol_n.flat.v[lo:hi]).abs().max()),
    "v_scale": float(pol_n.flat.v[lo:hi].abs().max())}
ag_f._fused.gather_moments()
out["gathered_moments_vs_nccl"] = {"m_max_abs": float((pol_f.flat.m - pol_n.flat.m).abs().max()),
                                   "v_max_abs": float((pol_f.flat.v - pol_n.flat.v).abs().max())}
out["exchange_plus_optimizer_ms_eager"] = {
    "fused (barrier + reduce_adam + barrier)": ex_ms["fused"][1:],
    "nccl (all-reduce only; sc_adam_flat follows)": ex_ms["nccl"][1:]}

# ---- time: whole steps as CUDA graphs, A/B/A/B ---------------------------------------------------------------------
ab = {"fused": [], "nccl": []}
for rep in range(2):
    for name, ag in (("fused", ag_f), ("nccl", ag_n)):
        ab[name].append(timed_graph_steps(ag, STEPS))
out["graph_step_ms"] = ab

# ---- the exchange alone (no backward in front, ranks aligned by a host barrier): launch-to-completion device time ----
def exchange_alone(ag, pol, fused, n=20):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if fused:
            ag._fused.step(sp.lr, sp.beta1, sp.beta2, sp.adam_eps, 5)
        else:
            dist.all_reduce(pol.flat.g32)
            from multimodal_sc_b200 import ops
            ops.adam_flat(pol.flat.p32, pol.flat.g32, pol.flat.m, pol.flat.v, pol.flat.p16, sp.lr, sp.beta1, sp.beta2,
                          sp.adam_eps, 5)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[2:])
    t = torch.tensor([ts[len(ts) // 2]], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


pol_f.flat.g32.zero_()
pol_n.flat.g32.zero_()
out["exchange_and_adam_alone_ms_median"] = {"fused": exchange_alone(ag_f, pol_f, True),
                                            "nccl_allreduce_plus_adam": exchange_alone(ag_n, pol_n, False)}
if rank == 0:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", f"dp_fused_check_n{world}.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))
# tear-down as in bench.py: the captured graphs hold the NCCL communicator, so they go first; the process then leaves
# without the collective destroy (the driver unmaps the peer buffers)
ag_f._graphs.clear()
ag_n._graphs.clear()
torch.cuda.synchronize()
sys.stdout.flush()
sys.stderr.flush()
os._exit(0)
