"""micro-benchmarks of the memory-bound kernels against the measured HBM peak (profiles/r1_microbench.json)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_sc_b200 import ops
from multimodal_sc_b200.spec import SpecV0

sp = SpecV0()
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)

def t(fn, n=20):
    for _ in range(3): fn()
    tot = 0.0
    for _ in range(n):
        flush.zero_()                                   # evict L2 between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e-3

out = {}
def rec(name, secs, nbytes, units=None, unit_name=None):
    d = {"us": secs * 1e6, "algorithmic_bytes": nbytes, "gbs": nbytes / secs / 1e9, "frac_of_measured_hbm": nbytes / secs / 1e9 / peak}
    if units: d[unit_name] = units / secs
    out[name] = d
    print(name, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items()})

# pure-write and pure-read streams (what a write-heavy / read-heavy kernel can expect, next to the 50/50 copy peak)
big = torch.empty(1 << 29, device="cuda")                       # 2 GiB fp32
st8 = torch.zeros(8, device="cuda")
s = t(lambda: ops.step_begin(big, st8, None))
rec("pure_write_2GiB_sc_step_begin", s, big.numel() * 4)
s = t(lambda: big.zero_())
rec("pure_write_2GiB_torch_memset", s, big.numel() * 4)
b16 = big.view(torch.bfloat16).view(-1, 256)
cs = torch.zeros(256, device="cuda")
s = t(lambda: ops.colsum(b16, cs))
rec("pure_read_2GiB_sc_colsum", s, big.numel() * 4)
dst = torch.empty(1 << 28, device="cuda")
s = t(lambda: dst.copy_(big[:1 << 28]))
rec("copy_1GiB_torch", s, dst.numel() * 8)
del big, b16, dst
# patchify (both observation shapes, bf16 and fp32 sources)
for name, shape, P in (("cam", (4096, 3, 128, 128), 16), ("bev", (4096, 1, 64, 64), 8)):
    for dt in (torch.bfloat16, torch.float32):
        img = torch.rand(*shape, device="cuda").to(dt)
        outp = torch.empty(shape[0] * (shape[2] // P) ** 2, shape[1] * P * P, device="cuda", dtype=torch.bfloat16)
        s = t(lambda: ops.patchify(img, P, out=outp))
        rec(f"patchify_{name}_{str(dt).split('.')[-1]}_B4096", s, img.numel() * img.element_size() + outp.numel() * 2)
        del img, outp

k = sp.k_sym
for B in (4096, 32768):
    z = torch.randn(2, B, k, device="cuda")
    y16 = torch.empty(B, 2, k, device="cuda", dtype=torch.bfloat16)
    y32 = torch.empty(B, 2, k, device="cuda")
    scale = torch.empty(B, 2, device="cuda")
    zs = (k, B * k, B, k)
    for kind in ("awgn", "rayleigh"):
        s = t(lambda: ops.pnorm_channel_fwd(z, kind, 10.0, seed=1, y_bf16=y16, scale=scale, z_strides=zs))
        rec(f"pnorm_channel_fwd_{kind}_bf16out_B{B}", s, B * 2 * k * 6, B, "samples_per_s")
    s = t(lambda: ops.pnorm_channel_fwd(z, "awgn", 10.0, seed=1, y_f32=y32, scale=scale, z_strides=zs))
    rec(f"pnorm_channel_fwd_awgn_f32out_B{B}", s, B * 2 * k * 8, B, "samples_per_s")
    dz = torch.empty(2, B, k, device="cuda", dtype=torch.bfloat16)
    s = t(lambda: ops.pnorm_channel_bwd(y32, z, scale, "awgn", 10.0, seed=1, dz_bf16=dz, z_strides=zs))
    rec(f"pnorm_channel_bwd_awgn_B{B}", s, B * 2 * k * 10, B, "samples_per_s")
    del z, y16, y32, dz

for T, N in ((128, 32), (128, 256), (128, 65536)):
    r, v, d = (torch.randn(T, N, device="cuda") for _ in range(3))
    lv = torch.randn(N, device="cuda")
    adv, ret = torch.empty_like(r), torch.empty_like(r)
    sums = torch.zeros(2, device="cuda", dtype=torch.float64)
    s = t(lambda: ops.gae(r, v, d, lv, sp.gamma, sp.lam, adv=adv, ret=ret, sums=sums))
    rec(f"gae_T{T}_N{N}", s, T * N * 20 + 4 * N, T * N, "steps_per_s")

M = 262144
x = torch.randn(M, 256, device="cuda").bfloat16()
g, b = torch.ones(256, device="cuda"), torch.zeros(256, device="cuda")
y, mean, rstd = torch.empty_like(x), torch.empty(M, device="cuda"), torch.empty(M, device="cuda")
s = t(lambda: ops.layernorm_fwd(x, g, b, 1e-5, y=y, mean=mean, rstd=rstd))
rec("layernorm_fwd_M262144", s, M * (1024 + 8))
dg, db, cs = (torch.zeros(256, device="cuda") for _ in range(3))
dx = torch.empty_like(x)
s = t(lambda: ops.layernorm_bwd(y, x, mean, rstd, g, dg, db, dres=y, dx=dx, colsum=cs))
rec("layernorm_bwd_M262144", s, M * (2048 + 8))
for Mm in (4096, 1 << 20):
    ho = torch.randn(Mm, 10, device="cuda"); act = torch.randint(0, 9, (Mm,), device="cuda")
    lp, ad, rt = (torch.randn(Mm, device="cuda") for _ in range(3))
    dho, st = torch.empty_like(ho), torch.zeros(8, device="cuda")
    s = t(lambda: ops.ppo_loss(ho, act, lp, ad, rt, 9, 0.2, 0.5, 0.01, 1.0 / Mm, dho=dho, stats=st))
    rec(f"ppo_loss_M{Mm}", s, Mm * (40 + 8 + 12 + 40), Mm, "samples_per_s")
n = 5165866 // 4 * 4
p_, g_, m_, v_ = (torch.randn(n, device="cuda") for _ in range(4)); pb = torch.empty(n, device="cuda", dtype=torch.bfloat16)
s = t(lambda: ops.adam_flat(p_, g_, m_.abs_(), v_.abs_(), pb, 3e-4, 0.9, 0.999, 1e-8, 3))
rec("adam_flat_5.17M", s, n * 30)
# the peer-memory reduce + Adam kernel with every "rank" on this device (local HBM side of the kernel only: the NVLink side is
# measured on two / eight GPUs by scripts/gpu_dp_fused_check.py): world = 1 moves the same 30 B per element as adam_flat
import ctypes as C
from multimodal_sc_b200 import _lib
from multimodal_sc_b200.peer import preload_kernels
preload_kernels("cuda")
arr = lambda *ts: (C.c_void_p * len(ts))(*[x.data_ptr() for x in ts])
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
s = t(lambda: _lib.check(lib.sc_dp_reduce_adam(arr(g_), arr(p_), arr(pb), m_.abs_().data_ptr(), v_.abs_().data_ptr(), 1, 0, n,
                                               3e-4, 0.9, 0.999, 1e-8, 3, None, 0, st)))
rec("dp_reduce_adam_world1_5.17M", s, n * 30)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/microbench.json", "w"), indent=1)
