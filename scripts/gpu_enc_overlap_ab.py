"""A/B of the second encoder stream for small inference batches (Engine.enc_overlap): cfg2 forward at B = 256 (and a few
other batch sizes), eagerly and as a CUDA-graph replay, outputs compared bit for bit.  Writes gpurun_out/enc_overlap_ab.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_sc_b200 import synthetic
from multimodal_sc_b200.model import SemComPolicy
from multimodal_sc_b200.spec import SpecV0

sp = SpecV0()
pol = SemComPolicy(sp)
eng = pol.engine
out = {}


def timeit(fn, n=50, warm=10):
    for _ in range(warm): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for B in (32, 256, 1024):
    mb = synthetic.minibatch(sp, B, seed=5, device="cuda", obs_dtype=torch.bfloat16)
    res, outs = {}, {}
    for ov in (0, 1, 0, 1):
        eng.enc_overlap = bool(ov)
        fwd = lambda: eng.forward(mb["cam"], mb["bev"], 10.0, "awgn", 77, None, None, 0, train=False)
        ho, _ = fwd()
        torch.cuda.synchronize()
        outs[ov] = ho.clone()
        ms_e = timeit(fwd)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fwd()
        ms_g = timeit(g.replay)
        g.replay(); torch.cuda.synchronize()
        same_graph = bool(torch.equal(eng.arena(B)["ac.ho"], outs[ov]))
        res.setdefault(f"overlap={ov}", []).append({"eager_ms": ms_e, "graph_ms": ms_g, "graph_samples_per_s": B / ms_g * 1e3,
                                                    "graph_equals_eager": same_graph})
        del g
    res["outputs_bit_equal"] = bool(torch.equal(outs[0], outs[1]))
    out[f"B={B}"] = res
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/enc_overlap_ab.json", "w"), indent=1)
print(json.dumps(out, indent=1))

# ---- training: forward + L_sem + backward (cfg4: Rayleigh, B = 1024; and B = 256), gradients compared ----
train = {}
for B in (256, 1024):
    mb = synthetic.minibatch(sp, B, seed=6, device="cuda", obs_dtype=torch.bfloat16)
    loss = torch.zeros(1, device="cuda")
    dho = torch.randn(B, sp.n_actions + 1, device="cuda") / B
    res, grads = {}, {}

    def step():
        pol.flat.g32.zero_()
        loss.zero_()
        _, A = eng.forward(mb["cam"], mb["bev"], 5.0, "rayleigh", 77, None, None, 0, train=True)
        eng.l_sem(A, B, loss, grad_coef=1.0)
        eng.backward(A, B, dho, None, "rayleigh", 5.0, 77, None, 0)

    for ov in (0, 1, 0, 1):
        eng.enc_overlap = bool(ov)
        step()
        torch.cuda.synchronize()
        grads[ov] = pol.flat.g32.clone()
        ms_e = timeit(step, n=20, warm=3)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step()
        ms_g = timeit(g.replay, n=20, warm=3)
        res.setdefault(f"overlap={ov}", []).append({"eager_ms": ms_e, "graph_ms": ms_g, "graph_samples_per_s": B / ms_g * 1e3})
        del g
    d = (grads[0] - grads[1]).abs().max().item()
    res["grad_max_abs_diff"] = d
    res["grad_scale"] = grads[0].abs().max().item()
    train[f"B={B}"] = res
out["train_fwd_lsem_bwd_rayleigh"] = train
json.dump(out, open("gpurun_out/enc_overlap_ab.json", "w"), indent=1)
print(json.dumps(train, indent=1))
