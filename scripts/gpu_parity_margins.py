"""How far the real-size parity tests are from their tolerances (run under gpurun): worst per-tensor gradient error / scale for
cfg4 (B = 1024) and cfg3 (B = 4096), three repetitions each (the split-K atomics make runs differ at fp32 rounding level)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_gpu_parity_full as T
from multimodal_sc_b200 import ops, synthetic
from multimodal_sc_b200.ppo import PPOAgent
from oracle.ppo import PPOAgent as OracleAgent

torch.backends.cuda.matmul.allow_tf32 = False
for rep in range(3):
    o, n = T._pair()
    B = 1024
    cam, bev = T._inputs(B, seed=1)
    snr = torch.tensor([-5.0, 0.0, 5.0, 10.0, 15.0, 20.0]).repeat(B // 6 + 1)[:B].cuda()
    g = torch.Generator().manual_seed(9)
    gl, gv = torch.randn(B, 9, generator=g).cuda() / B, torch.randn(B, generator=g).cuda() / B
    for m in (o, n):
        l, v, aux = m(cam, bev, snr, channel="rayleigh", seed=77)
        ((l * gl).sum() + (v * gv).sum() + aux["l_sem"]).backward()
    tab = T._grad_table([(k, p.grad) for k, p in o.named_parameters()], [(k, p.grad) for k, p in n.named_parameters()])
    print(f"cfg4 rep {rep}: worst gradient tensors (err/scale, name): {[(round(r[0], 4), r[1]) for r in tab[:3]]}", flush=True)
    del o, n
    torch.cuda.empty_cache()
for rep in range(3):
    o, n = T._pair()
    B = 4096
    mb = synthetic.minibatch(T.SPEC, B, seed=2, device="cuda")
    ao, an = OracleAgent(o, T.SPEC, minibatch=B, seed=1234), PPOAgent(n, T.SPEC, minibatch=B, seed=1234)
    so = ao.minibatch_step(mb["cam"], mb["bev"], mb["actions"], mb["logp_old"], mb["adv"], None, mb["ret"])
    grads = {}
    orig = ops.adam_flat
    def spy(p32, g32, *a, **k):
        grads["g"] = g32.clone()
        return orig(p32, g32, *a, **k)
    ops.adam_flat = spy
    sn = an.minibatch_step(mb["cam"], mb["bev"], mb["actions"], mb["logp_old"], mb["adv"], None, mb["ret"])
    ops.adam_flat = orig
    flat = n.flat
    gn = [(name, grads["g"][off:off + cnt].view(flat.shapes[name])) for name, (off, cnt) in flat.offsets.items()]
    tab = T._grad_table([(k, p.grad) for k, p in o.named_parameters()], gn)
    st = {k: round(abs(sn[k] - so[k]) / max(abs(so[k]), 0.05), 5) for k in ("pi", "v", "ent", "kl", "l_sem", "total")}
    print(f"cfg3 rep {rep}: worst gradient tensors: {[(round(r[0], 4), r[1]) for r in tab[:3]]}; stat rel diffs {st}", flush=True)
    del o, n, ao, an
    torch.cuda.empty_cache()
