"""sc_mlp_bwd_dgrad: the single-CTA kernel (default) vs the CTA-pair kernel (mlp_dbg bit 32), M = 262144 / 524288 / 16384 tokens, L2 flushed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_sc_b200 import ops

BF16 = torch.bfloat16
D, R = 256, 1024
g = lambda *s: (torch.randn(*s, device="cuda") * 0.1).to(BF16)
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
w1, w2 = g(R, D), g(D, R)
for M in (262144, 524288, 16384):
    dy, hpre = g(M, D), g(M, R)
    dhpre, dx = torch.empty(M, R, device="cuda", dtype=BF16), torch.empty(M, D, device="cuda", dtype=BF16)
    out = {}
    res = {}
    for mode, dbg in (("pair", 32), ("single", 0), ("pair", 32), ("single", 0)):
        ops.debug_set("mlp_dbg", dbg)
        for _ in range(3):
            ops.mlp_bwd_dgrad(dy, w1, w2, hpre, dhpre=dhpre, dx=dx)
        tot = 0.0
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.mlp_bwd_dgrad(dy, w1, w2, hpre, dhpre=dhpre, dx=dx); e1.record(); torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        res.setdefault(mode, []).append(round(tot / 10 * 1e3, 1))
        out[mode] = (dhpre.clone(), dx.clone())
    ops.debug_set("mlp_dbg", 0)
    same = torch.equal(out["pair"][0], out["single"][0]), float((out["pair"][1].float() - out["single"][1].float()).abs().max())
    fl = 4.0 * M * D * R
    by = 2.0 * M * D * 2 + 2.0 * M * R * 2
    print(f"M {M}: pair {res['pair']} us ({fl / res['pair'][-1] / 1e6:.0f} TF, {by / res['pair'][-1] / 1e3:.0f} GB/s), single {res['single']} us; "
          f"dhpre identical {same[0]}, dx max diff {same[1]:.3g}", flush=True)
