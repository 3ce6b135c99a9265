"""max |n_kernel - n_float64| of the channel kernel's SFU Box-Muller and of the accurate variant over 2^26 normals (run under gpurun)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multimodal_sc_b200 import ops
from oracle import philox as oph
n_quads = 1 << 22
worst = {False: 0.0, True: 0.0}
near = 0.0
for chunk in range(4):
    off = chunk * n_quads + (7 << 30) * chunk
    w = oph.philox_words(99, n_quads, 0, off)
    u = oph.u32_to_uniform(w)
    n0, n1 = oph.box_muller(u[:, 0], u[:, 1]); n2, n3 = oph.box_muller(u[:, 2], u[:, 3])
    ref = torch.from_numpy(np.stack([n0, n1, n2, n3], axis=-1))
    for exact in (False, True):
        got = ops.philox_normals(n_quads, 99, off, exact=exact).cpu().double()
        worst[exact] = max(worst[exact], float((got - ref).abs().max()))
        if not exact:
            m = torch.from_numpy(u[:, 0] > 1 - 2.0 ** -6)
            near = max(near, float((got[m, :2] - ref[m, :2]).abs().max()))
print("max |dn| fast", worst[False], "exact", worst[True], "near u->1 (fast)", near)
