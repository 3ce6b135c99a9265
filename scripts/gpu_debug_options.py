"""debug helper: forward / gradient error tables of the round-2 spec options against the oracle (run under gpurun)"""
import dataclasses, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import test_gpu_parity_full as T

torch.backends.cuda.matmul.allow_tf32 = False
for kw in ({}, {"snr_cond": True}, {"gelu": "erf"}, {"quant_bits": 4}, {"quant_bits": 8}):
    spec = dataclasses.replace(T.SPEC, **kw)
    o, n = T._pair(spec)
    B = 64
    cam, bev = T._inputs(B, seed=8)
    snr = torch.linspace(-5, 20, B).cuda()
    g = torch.Generator().manual_seed(2)
    gl, gv = torch.randn(B, spec.n_actions, generator=g).cuda() / B, torch.randn(B, generator=g).cuda() / B

    def run(model):
        l, v, aux = model(cam, bev, snr, channel="rayleigh", seed=21)
        ((l * gl).sum() + (v * gv).sum() + aux["l_sem"]).backward()
        return l.detach(), v.detach(), {k: a.detach() for k, a in aux.items()}

    lo, vo, ao = run(o)
    ln, vn, an = run(n)
    print(kw, "logits", T._rel(ln, lo), "value", T._rel(vn, vo), "lsem", float(an["l_sem"]), float(ao["l_sem"]),
          "y", T._rel(an["y"], ao["y"]), "z", T._rel(an["z"], ao["z"]))
    tab = T._grad_table([(k, p.grad) for k, p in o.named_parameters()], [(k, p.grad) for k, p in n.named_parameters()])
    for r in tab[:6]:
        print("   ", r)
