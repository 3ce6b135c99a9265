"""A/B of sc_debug_set switches on the step's main GEMM shapes: python scripts/gpu_gemm_ab.py key=value [key=value ...] -- shape ...
Prints us per launch (M = 262144 tokens, L2 flushed between launches) with the switches off / on."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from gpu_gemm_bottleneck import make
from multimodal_sc_b200 import ops

args = sys.argv[1:]
sw = dict(a.split("=") for a in args[:args.index("--")]) if "--" in args else {}
names = args[args.index("--") + 1:] if "--" in args else args
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
for name in names or ["proj_ln", "fc2_ln", "proj", "qkv", "fc1"]:
    spec = make(name)
    A, kw = spec["A"](), spec["kw"]()
    out = torch.empty(spec["M"], spec["N"], device="cuda", dtype=torch.bfloat16)
    run = lambda: ops.gemm(A, spec["B"], b_mn=spec.get("b_mn", False), out=out, **kw)
    res, ref, same = {}, None, True
    for mode in ("off", "on", "off", "on"):
        for k, v in sw.items():
            ops.debug_set(k, int(v) if mode == "on" else 0)
        for _ in range(3):
            run()
        tot = 0.0
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        res.setdefault(mode, []).append(round(tot / 10 * 1e3, 1))
        snap = [out.clone()] + [v.clone() for k_, v in kw.items() if k_ == "aux_out"] + ([kw["ln"]["y"].clone()] if "ln" in kw else [])
        same = same and all(torch.equal(a_, b_) for a_, b_ in zip(ref, snap)) if mode == "on" else same
        if mode == "off":
            ref = snap
    for k in sw:
        ops.debug_set(k, 0)
    print(f"{name:10s} M {spec['M']} N {spec['N']} K {A.shape[1]}: off {res['off']} us, on {res['on']} us  ({sw}); outputs identical: {same}", flush=True)
