"""Native SemComPolicy: the SPEC v0 model driven kernel-by-kernel through the C ABI.

Drop-in for the oracle's `nn.Module` API (upstream /root/reference/README.md:1-2
has no code, so the API is the one `oracle/model.py` fixes): same constructor, same
`forward(cam, bev, snr_db, *, channel, seed, noise, fading, sample_offset, rician_k) -> (logits, value, aux)`,
same `act()`, same `state_dict()` keys and shapes.

B200-first layout instead of a module-per-layer autograd graph:
  * every parameter is a view into ONE flat fp32 bucket (`FlatParams.p32`) with a bf16 shadow
    for the tensor-core operands, one flat fp32 gradient bucket (the NCCL all-reduce payload)
    and flat Adam moments — the optimizer and the collective are single launches;
  * forward and backward are explicit kernel schedules (`Engine`) over a preallocated
    activation arena, no autograd tape, no allocator traffic in the step, CUDA-graph capturable;
  * weight gradients accumulate straight into the flat bucket from the wgrad GEMM epilogue
    (split-K + red.global.add.f32) — there are no per-layer grad tensors to concatenate.
`forward()` still participates in torch autograd (one custom Function around the whole engine)
so that `loss.backward()` works exactly as with the oracle.
"""
from __future__ import annotations

import os
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _lib, ops
from .spec import SpecV0

BF16, F32 = torch.bfloat16, torch.float32


# ----------------------------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------------------------
def _block_shapes(prefix: str, spec: SpecV0):
    D, R = spec.d_model, spec.d_mlp
    return [(f"{prefix}.ln1.weight", (D,)), (f"{prefix}.ln1.bias", (D,)),
            (f"{prefix}.qkv.weight", (3 * D, D)), (f"{prefix}.qkv.bias", (3 * D,)),
            (f"{prefix}.proj.weight", (D, D)), (f"{prefix}.proj.bias", (D,)),
            (f"{prefix}.ln2.weight", (D,)), (f"{prefix}.ln2.bias", (D,)),
            (f"{prefix}.fc1.weight", (R, D)), (f"{prefix}.fc1.bias", (R,)),
            (f"{prefix}.fc2.weight", (D, R)), (f"{prefix}.fc2.bias", (D,))]


def param_shapes(spec: SpecV0) -> "OrderedDict[str, tuple]":
    """name -> shape, in the oracle's state_dict order (tests assert equality with the oracle)."""
    D = spec.d_model
    out = []
    for enc, tokens, k in (("cam_enc", spec.cam_tokens, spec.cam_k), ("bev_enc", spec.bev_tokens, spec.bev_k)):
        out += [(f"{enc}.pos", (tokens, D)), (f"{enc}.patch_embed.weight", (D, k)),
                (f"{enc}.patch_embed.bias", (D,))]
        for l in range(spec.enc_layers):
            out += _block_shapes(f"{enc}.blocks.{l}", spec)
        out += [(f"{enc}.norm.weight", (D,)), (f"{enc}.norm.bias", (D,)),
                (f"{enc}.sym.weight", (spec.c_sym, D)), (f"{enc}.sym.bias", (spec.c_sym,))]
        if spec.snr_cond:                                    # SNR-conditioned encoders (oracle/model.py: SnrEmbed)
            E = spec.snr_embed_dim
            out += [(f"{enc}.snr.u", (E,)), (f"{enc}.snr.v", (E,)), (f"{enc}.snr.w", (D, E))]
    out += [("dec.pos", (spec.dec_tokens, D)), ("dec.in_proj.weight", (D, spec.c_sym)), ("dec.in_proj.bias", (D,))]
    for l in range(spec.dec_layers):
        out += _block_shapes(f"dec.blocks.{l}", spec)
    out += [("dec.norm.weight", (D,)), ("dec.norm.bias", (D,)), ("dec.recon.weight", (D, D)),
            ("dec.recon.bias", (D,)), ("ac.trunk.weight", (D, D)), ("ac.trunk.bias", (D,)),
            ("ac.head.weight", (spec.n_actions + 1, D)), ("ac.head.bias", (spec.n_actions + 1,))]
    return OrderedDict(out)


class FlatParams:
    """flat fp32 master / bf16 shadow / fp32 grad / Adam moments; 256-byte aligned slots."""
    ALIGN = 64   # elements

    def __init__(self, spec: SpecV0, device, peer: bool = False):
        """peer=True: gradients, master parameters and the bf16 shadow live in exportable device memory
        (peer.peer_empty) so that the other ranks of a data-parallel job can map them (peer.PeerReduceAdam)."""
        self.peer = bool(peer)
        self.shapes = param_shapes(spec)
        self.offsets = OrderedDict()
        off = 0
        for name, shp in self.shapes.items():
            n = 1
            for s in shp:
                n *= s
            self.offsets[name] = (off, n)
            off += (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.numel = off
        self._views = {}
        if self.peer:
            from .peer import peer_empty
            self.p32 = peer_empty(off, F32, device)
            self.p16 = peer_empty(off, BF16, device)
            self.g32 = peer_empty(off, F32, device)
        else:
            self.p32 = torch.zeros(off, device=device, dtype=F32)
            self.p16 = torch.zeros(off, device=device, dtype=BF16)
            self.g32 = torch.zeros(off, device=device, dtype=F32)
        self.m = torch.zeros(off, device=device, dtype=F32)
        self.v = torch.zeros(off, device=device, dtype=F32)

    def _view(self, buf, name):
        off, n = self.offsets[name]
        return buf[off:off + n].view(self.shapes[name])

    # The views are built once per (bucket, name): an eager small-batch forward makes ~150 of these lookups, and two
    # tensor ops per lookup cost more host time than the kernels take on the device (cfg2, B = 256).  The cache is
    # keyed by the bucket's storage, so swapping a bucket (the autograd facade's scratch gradient bucket) is safe.
    def _cached(self, key, buf, name):
        c = self._views.get(key)
        if c is None or c[0] != buf.data_ptr():
            c = self._views[key] = (buf.data_ptr(), {})
        v = c[1].get(name)
        if v is None:
            v = c[1][name] = self._view(buf, name)
        return v

    def w32(self, name): return self._cached("p32", self.p32, name)
    def w16(self, name): return self._cached("p16", self.p16, name)
    def g(self, name): return self._cached("g32", self.g32, name)

    def sync_shadow(self):
        ops.cast_f32_bf16(self.p32, self.p16)

    def n_params(self) -> int:
        return sum(n for _, n in self.offsets.values())


def _init_flat(flat: FlatParams, spec: SpecV0, std: float, seed: int):
    """same init rule (and the same host RNG stream) as oracle SemComPolicy.reset_parameters."""
    g = torch.Generator().manual_seed(seed)
    for name, shp in flat.shapes.items():
        dst = flat.w32(name)
        if name.endswith("bias"):
            dst.zero_()
        elif ".ln" in name or "norm" in name:
            dst.fill_(1.0)
        else:
            w = torch.empty(shp).normal_(0.0, 1.0, generator=g).clamp_(-2.0, 2.0) * std
            dst.copy_(w)
    flat.sync_shadow()


# ----------------------------------------------------------------------------------------------
# engine: explicit forward / backward schedules over a per-batch-size activation arena
# ----------------------------------------------------------------------------------------------
class _Arena(dict):
    """named activation buffers for one batch size.

    `cap` > 0 makes it a CAPACITY arena: every per-sample buffer is allocated for `cap` samples once and handed out as a
    view of the first B rows, so the ragged local slices of a data-parallel update (each rank processes the
    samples of the global minibatch it owns, a different count every step) reuse one set of buffers.
    `generation` counts Engine.forward calls on this arena: an autograd node that saved activations here checks it in
    backward (a later forward with the same batch size would otherwise silently corrupt the gradients)."""

    def __init__(self, device, cap: int = 0):
        super().__init__()
        self.device = device
        self.cap = cap
        self.B = 0                      # batch size of the current forward (capacity arenas)
        self.generation = 0
        self._store = {}

    def buf(self, name, shape, dtype=BF16):
        if not self.cap:
            t = self.get(name)
            if t is None:
                t = torch.empty(shape, device=self.device, dtype=dtype)
                self[name] = t
            return t
        n = 1
        for d in shape:
            n *= d
        full = n // self.B * self.cap if (self.B and n % self.B == 0) else n
        st = self._store.get(name)
        if st is None or st.numel() < max(full, n) or st.dtype != dtype:
            st = self._store[name] = torch.empty(max(full, n), device=self.device, dtype=dtype)
        t = st[:n].view(shape)
        self[name] = t
        return t


def _wgrad_ksplit(n_out: int, k_out: int, m_red: int, sms: int = 148) -> int:
    tiles = ((n_out + 127) // 128) * ((k_out + 255) // 256)
    ks = max(1, sms // tiles)
    return max(1, min(ks, (m_red + 63) // 64))


class Engine:
    def __init__(self, spec: SpecV0, flat: FlatParams, device):
        self.spec, self.flat, self.device = spec, flat, device
        self.arenas = {}
        self.sms = torch.cuda.get_device_properties(device).multi_processor_count
        # SC_FUSED_MLP=0 selects the two-GEMM MLP schedule everywhere (A/B runs)
        self.fused_mlp = os.environ.get("SC_FUSED_MLP", "1") != "0"
        # SC_FUSED_LN=0: every LayerNorm forward is its own kernel instead of a GEMM epilogue
        self.fused_ln = os.environ.get("SC_FUSED_LN", "1") != "0"
        # SC_FUSED_MLP_LN=1: the inference MLP kernel also carries the LayerNorm that follows it (sc_mlp_fwd_ln).  Off:
        # parity-green but SLOWER — the second pass over the output accumulator sits on the critical path of a single-wave
        # kernel and costs more than the LayerNorm launch it saves (forward B = 256 0.650 -> 0.660 ms, B = 1024 +6 %)
        self.fused_mlp_ln = os.environ.get("SC_FUSED_MLP_LN", "0") != "0"
        self.rician_k = spec.rician_k          # K-factor used when channel == "rician" (SemComPolicy.forward may override)
        # SC_WGRAD_STREAM=1: the four weight-gradient GEMMs of every block run on a second stream, so the HBM-bound
        # LayerNorm-backward kernels of the dgrad chain (no shared memory, no TMEM) co-run with them and the tail of
        # one GEMM-class kernel overlaps the ramp-up of the next.  Block-parity scratch sets + events keep it race-free.
        self.wgrad_stream = torch.cuda.Stream(device) if os.environ.get("SC_WGRAD_STREAM", "0") != "0" else None
        self._side_done = []          # per block of the current backward: event after its last side-stream wgrad
        self._free_arenas = {}        # batch size -> arenas returned by finished autograd nodes
        # exact (erf) GELU: the fused MLP kernels implement the tanh form only -> two-GEMM schedule, direct-store epilogue
        self.gelu_act = "gelu" if spec.gelu == "tanh" else "gelu_erf"
        if spec.gelu != "tanh":
            assert spec.gelu == "erf", f"unknown GELU form {spec.gelu!r}"
            self.fused_mlp = False
        # deterministic=True: weight gradients without split-K (one red.add per element, in stream order) — a debug /
        # test mode; the default split-K wgrads add their K-slices in arbitrary order (fp32 rounding-level run-to-run noise)
        self.deterministic = os.environ.get("SC_DETERMINISTIC", "0") != "0"
        # Small batches (cfg2, cfg4, act()): every kernel is a single wave that is bound by its own ramp-up and tail,
        # not by data (profiles/README.md).  The two encoders are independent until the channel, so the LiDAR-BEV encoder
        # runs on a second stream (a parallel branch of the CUDA graph when the forward is captured): one encoder's CTAs
        # fill the SMs the other's tail leaves idle and no kernel waits for a launch gap; backward does the same with
        # per-encoder scratch buffers.  Measured (profiles/r2_enc_overlap_ab.json): forward B = 256 0.73 -> 0.65 ms as a
        # graph, B = 1024 +5 %, the 4096-sample PPO step +0.4 % (same-box A/B, twice); above 4096 samples it stays off (the
        # second scratch set costs 1.6 GB per 4096 samples).  SC_ENC_OVERLAP=0 turns it off.
        self.enc_overlap = os.environ.get("SC_ENC_OVERLAP", "1") != "0"
        self.enc_overlap_max_batch = int(os.environ.get("SC_ENC_OVERLAP_MAX_BATCH", "4096"))
        self._enc_stream = None

    def _enc_side_stream(self, B: int):
        """the second encoder stream, or None when this forward keeps both encoders on the current stream"""
        if (not self.enc_overlap or B > self.enc_overlap_max_batch or ops._PROF is not None
                or self.wgrad_stream is not None):
            return None
        if B < 128 and not torch.cuda.is_current_stream_capturing():
            return None            # eagerly these sizes are bound by the host's launch rate: two more events per pass cost 10 %
        if self._enc_stream is None:
            self._enc_stream = torch.cuda.Stream(self.device)
        return self._enc_stream

    def arena(self, B: int, cap: int = 0) -> _Arena:
        """the shared arena of batch size B, or (cap >= B) the capacity arena for up to `cap` samples"""
        key = ("cap", cap) if cap else B
        a = self.arenas.get(key)
        if a is None:
            a = self.arenas[key] = _Arena(self.device, cap)
        return a

    def lease_arena(self, B: int) -> "_ArenaLease":
        """a private arena for one outstanding autograd node (two grad-mode forwards followed by one backward each own
        their activations); returned to the pool when the node is freed"""
        pool = self._free_arenas.setdefault(B, [])
        a = pool.pop() if pool else _Arena(self.device)
        return _ArenaLease(a, pool)

    # ---- helpers ------------------------------------------------------------------------
    def _linear_bwd(self, dy, x, wname, dx=None, bias_colsum=True, side=False, **dgrad_epi):
        """dx = dy W (optional), dW += dy^T x, db += colsum(dy).  dy [M,N] bf16, x [M,K] bf16."""
        f = self.flat
        W = f.w16(f"{wname}.weight")
        N, K = W.shape
        if dx is not None or dgrad_epi.get("out_f32") is not None:
            ops.gemm(dy, W, b_mn=True, out=dx, **dgrad_epi)
        self._wgrad(dy, x, wname, bias_colsum, side=side)

    def _wgrad(self, dy, x, wname, bias_colsum=True, side=False):
        """dW += dy^T x (split-K, red.add into the flat bucket); db += colsum(dy) as one extra MMA per K-slice.
        side=True (and a wgrad stream configured): enqueued on the side stream after everything the current stream
        has enqueued so far; the caller owns the lifetime of dy / x (block-parity scratch, Engine._block_bwd)."""
        f = self.flat
        N, K = f.shapes[f"{wname}.weight"]

        def run():
            ops.gemm(dy, x, a_mn=True, b_mn=True, out_f32=f.g(f"{wname}.weight"), atomic=True,
                     ksplit=1 if self.deterministic else _wgrad_ksplit(N, K, dy.shape[0], self.sms),
                     colsum=f.g(f"{wname}.bias") if bias_colsum else None)

        if not (side and self._side_on()):
            run()
            return
        main = torch.cuda.current_stream(self.device)
        ev = torch.cuda.Event()
        ev.record(main)
        self.wgrad_stream.wait_event(ev)
        with torch.cuda.stream(self.wgrad_stream):
            run()

    def _side_on(self) -> bool:
        # the roofline pass of bench.py (ops.profile) times launches one by one: keep them on one stream there
        return self.wgrad_stream is not None and ops._PROF is None

    def _ln_spec(self, A, wname, M, y_name, mean_name, rstd_name):
        """arguments of a LayerNorm whose parameters are `wname`.{weight,bias}: either run it (ops.layernorm_fwd)
        or hand it to the producing GEMM as its fused second output (ops.gemm(..., ln=spec))."""
        sp, f = self.spec, self.flat
        return dict(gamma=f.w32(f"{wname}.weight"), beta=f.w32(f"{wname}.bias"), eps=sp.ln_eps,
                    y=A.buf(y_name, (M, sp.d_model)), mean=A.buf(mean_name, (M,), F32), rstd=A.buf(rstd_name, (M,), F32))

    def _block_ln1(self, A, pfx, M):
        return self._ln_spec(A, f"{pfx}.ln1", M, f"{pfx}.xn1", f"{pfx}.mean1", f"{pfx}.rstd1")

    @staticmethod
    def _run_ln(x, ln):
        ops.layernorm_fwd(x, ln["gamma"], ln["beta"], ln["eps"], y=ln["y"], mean=ln["mean"], rstd=ln["rstd"])

    def _block_fwd(self, A, pfx, x_in, x_out, B, S, train, xn1_ready=False, next_ln=None):
        """one pre-LN block.  `xn1_ready`: the producer of x_in already wrote LN1(x_in) (fused epilogue).
        `next_ln`: the LayerNorm that consumes x_out (next block's LN1 or the stack's final norm); it is applied
        here, fused into the fc2 GEMM's epilogue when possible."""
        sp, f = self.spec, self.flat
        M, D, H = x_in.shape[0], sp.d_model, sp.n_heads
        k = lambda n, shape, dt=BF16: A.buf(f"{pfx}.{n}", shape, dt)
        fuse_ln = self.fused_ln and D == 256
        ln1 = self._block_ln1(A, pfx, M)
        if not xn1_ready:
            self._run_ln(x_in, ln1)
        xn1 = ln1["y"]
        qkv = k("qkv", (M, 3 * D))
        ops.gemm(xn1, f.w16(f"{pfx}.qkv.weight"), bias=f.w32(f"{pfx}.qkv.bias"), out=qkv)
        ao = k("ao", (M, D))
        ops.attn_fwd(qkv, B, S, H, out=ao, lse=k("lse", (M, H), F32))
        x1 = k("x1", (M, D))
        ln2 = self._ln_spec(A, f"{pfx}.ln2", M, f"{pfx}.xn2", f"{pfx}.mean2", f"{pfx}.rstd2")
        ops.gemm(ao, f.w16(f"{pfx}.proj.weight"), bias=f.w32(f"{pfx}.proj.bias"), resid=x_in, out=x1,
                 ln=ln2 if fuse_ln else None)
        if not fuse_ln:
            self._run_ln(x1, ln2)
        xn2 = ln2["y"]
        if not train and self.fused_mlp and D == 256:
            # inference: fc1 -> GELU -> fc2 in one kernel, the hidden activation never leaves the SM
            # ... and the LayerNorm that consumes the block's output rides on its output epilogue (one launch less per block)
            ln_fused = next_ln if (next_ln is not None and fuse_ln and self.fused_mlp_ln) else None
            ops.mlp_fwd(xn2, f.w16(f"{pfx}.fc1.weight"), f.w32(f"{pfx}.fc1.bias"), f.w16(f"{pfx}.fc2.weight"),
                        f.w32(f"{pfx}.fc2.bias"), resid=x1, out=x_out, ln=ln_fused)
            if next_ln is not None and ln_fused is None:
                self._run_ln(x_out, next_ln)
            return
        # training saves both hpre and h: two wide outputs make the fused kernel store-bound, so the forward
        # keeps the two-GEMM form (profiles/README.md, r1 fused-MLP table)
        h = k("h", (M, sp.d_mlp))
        hpre = k("hpre", (M, sp.d_mlp)) if train else None
        ops.gemm(xn2, f.w16(f"{pfx}.fc1.weight"), bias=f.w32(f"{pfx}.fc1.bias"), act=self.gelu_act, aux_out=hpre, out=h)
        ops.gemm(h, f.w16(f"{pfx}.fc2.weight"), bias=f.w32(f"{pfx}.fc2.bias"), resid=x1, out=x_out,
                 ln=next_ln if fuse_ln else None)
        if next_ln is not None and not fuse_ln:
            self._run_ln(x_out, next_ln)

    def _block_bwd(self, A, pfx, x_in, dx_out, dx_in, prev_bias_grad, B, S, tag=""):
        """dx_out: grad wrt block output [M,D] (its column sums were already added to fc2.bias grad by the
        producer).  Writes dx_in; adds colsum(dx_in) to `prev_bias_grad` (bias of whatever produced x_in)."""
        sp, f = self.spec, self.flat
        M, D, H = x_in.shape[0], sp.d_model, sp.n_heads
        k = lambda n: A[f"{pfx}.{n}"]
        # scratch is shared across the blocks of one width (`tag` separates encoders that run on two streams)
        s = lambda n, shape, dt=BF16: A.buf(f"scratch{tag}.{n}.S{shape[0] // B}", shape, dt)
        side = self._side_on()
        # With the wgrads on the side stream, the buffers they read (dhpre, dx1, dqkv) alternate between two scratch
        # sets by block parity, and the side stream never lags by more than one block: LN1-backward of block j (the
        # kernel that overwrites the dx_out of block j-1) waits for the wgrads of block j-1.
        j = len(self._side_done)
        s2 = (lambda n, shape: s(f"{n}.p{j & 1}", shape)) if side else s
        # ---- MLP ----
        dhpre = s2("dh", (M, sp.d_mlp))
        dxn2 = s("dxn", (M, D))
        if self.fused_mlp and D == 256:
            # both dgrad GEMMs in one kernel (dhpre stays on chip between them); the wgrads follow as split-K GEMMs
            ops.mlp_bwd_dgrad(dx_out, f.w16(f"{pfx}.fc1.weight"), f.w16(f"{pfx}.fc2.weight"), k("hpre"),
                              dhpre=dhpre, dx=dxn2)
            self._linear_bwd(dx_out, k("h"), f"{pfx}.fc2", bias_colsum=False, side=True)
            self._linear_bwd(dhpre, k("xn2"), f"{pfx}.fc1", side=True)
        else:
            self._linear_bwd(dx_out, k("h"), f"{pfx}.fc2", dx=dhpre, bias_colsum=False, aux_in=k("hpre"), side=True,
                             aux_in_erf=self.gelu_act == "gelu_erf")
            self._linear_bwd(dhpre, k("xn2"), f"{pfx}.fc1", dx=dxn2, side=True)
        dx1 = s2("dx1", (M, D))
        ops.layernorm_bwd(dxn2, k("x1"), k("mean2"), k("rstd2"), f.w32(f"{pfx}.ln2.weight"),
                          f.g(f"{pfx}.ln2.weight"), f.g(f"{pfx}.ln2.bias"), dres=dx_out, dx=dx1,
                          colsum=f.g(f"{pfx}.proj.bias"))
        # ---- attention ----
        dao = s("dao", (M, D))
        self._linear_bwd(dx1, k("ao"), f"{pfx}.proj", dx=dao, bias_colsum=False, side=True)
        dqkv = s2("dqkv", (M, 3 * D))
        ops.attn_bwd(k("qkv"), k("ao"), dao, k("lse"), B, S, H, dqkv=dqkv)
        dxn1 = s("dxn", (M, D))
        self._linear_bwd(dqkv, k("xn1"), f"{pfx}.qkv", dx=dxn1, side=True)
        if side:
            if j > 0:
                torch.cuda.current_stream(self.device).wait_event(self._side_done[j - 1])
            ev = torch.cuda.Event()
            ev.record(self.wgrad_stream)
            self._side_done.append(ev)
        ops.layernorm_bwd(dxn1, x_in, k("mean1"), k("rstd1"), f.w32(f"{pfx}.ln1.weight"),
                          f.g(f"{pfx}.ln1.weight"), f.g(f"{pfx}.ln1.bias"), dres=dx1, dx=dx_in,
                          colsum=prev_bias_grad)

    def _encoder_fwd(self, A, mi, enc, img, patch, B, S, z_mm, snr_db, train, row_idx):
        """one modality encoder: patchify -> patch embedding (+ LN) -> blocks -> final norm -> symbol projection into z_mm[mi]"""
        sp, f = self.spec, self.flat
        D, C = sp.d_model, sp.c_sym
        M = B * S
        tok = ops.patchify(img, patch, out=A.buf(f"{enc}.tok", (M, img.shape[1] * patch * patch)), row_idx=row_idx)
        xs = [A.buf(f"{enc}.x{l}", (M, D)) for l in range(sp.enc_layers + 1)]
        fuse_ln = self.fused_ln and D == 256
        # every LayerNorm rides on the epilogue of the GEMM that produces its input: the patch embedding feeds
        # block 0's LN1, each block's fc2 feeds the next block's LN1 / the final norm
        if sp.snr_cond:
            # SNR-conditioned encoder: the per-sample embedding rides on the row-group add of the epilogue (the same
            # path the mean-pool backward uses); that epilogue kind has no fused LayerNorm, so block 0's LN1 runs
            # as its own kernel
            emb, _ = ops.snr_embed_fwd(snr_db, f.w32(f"{enc}.snr.u"), f.w32(f"{enc}.snr.v"), f.w32(f"{enc}.snr.w"), B,
                                       emb=A.buf(f"{enc}.snr_emb", (B, D), F32),
                                       h=A.buf(f"{enc}.snr_h", (B, sp.snr_embed_dim), F32))
            ops.gemm(tok, f.w16(f"{enc}.patch_embed.weight"), bias=f.w32(f"{enc}.patch_embed.bias"),
                     rowadd=f.w16(f"{enc}.pos"), rowgrp=emb, rowgrp_size=S, rowgrp_scale=1.0, out=xs[0])
            if fuse_ln:
                self._run_ln(xs[0], self._block_ln1(A, f"{enc}.blocks.0", M))
        else:
            ops.gemm(tok, f.w16(f"{enc}.patch_embed.weight"), bias=f.w32(f"{enc}.patch_embed.bias"),
                     rowadd=f.w16(f"{enc}.pos"), out=xs[0],
                     ln=self._block_ln1(A, f"{enc}.blocks.0", M) if fuse_ln else None)
        final_ln = self._ln_spec(A, f"{enc}.norm", M, f"{enc}.t", f"{enc}.mean", f"{enc}.rstd")
        for l in range(sp.enc_layers):
            nxt = self._block_ln1(A, f"{enc}.blocks.{l + 1}", M) if l + 1 < sp.enc_layers else final_ln
            self._block_fwd(A, f"{enc}.blocks.{l}", xs[l], xs[l + 1], B, S, train, xn1_ready=fuse_ln, next_ln=nxt)
        t = final_ln["y"]
        ops.gemm(t, f.w16(f"{enc}.sym.weight"), bias=f.w32(f"{enc}.sym.bias"), out_f32=z_mm[mi].view(M, C))

    # ---- forward --------------------------------------------------------------------------
    def forward(self, cam, bev, snr_db, channel, seed, noise, fading, sample_offset, train, want_aux=False,
                seed_dev=None, row_idx=None, sample_idx=None, arena=None, need_recon=True):
        """row_idx (int64 [B], optional): the step's samples are rows row_idx of cam / bev (the minibatch gather is fused
        into the patchify pass).  sample_idx (int64 [B], optional): global sample index of every local sample for the
        channel's noise counters (default sample_offset + b).  arena: the activation arena to use (default: the shared
        arena of this batch size).  need_recon=False skips the reconstruction head (act(): only the actor-critic outputs
        are needed, L_sem is not)."""
        sp, f = self.spec, self.flat
        B = cam.shape[0] if row_idx is None else row_idx.numel()
        A = arena if arena is not None else self.arena(B)
        A.B = B
        A.generation += 1
        D, S, C = sp.d_model, sp.cam_tokens, sp.c_sym
        k_sym = sp.k_sym
        z_mm = A.buf("z_mm", (2, B, k_sym), F32)                       # modality-major symbols
        side = self._enc_side_stream(B)
        main = torch.cuda.current_stream(self.device)
        if side is not None:
            fork = torch.cuda.Event()
            fork.record(main)
            side.wait_event(fork)                                      # inputs / earlier work of this stream are visible
        # (the BEV encoder goes first so that, eagerly, its launches are already queued when the camera's start)
        for mi, (enc, img, patch) in ((1, ("bev_enc", bev, sp.bev_patch)), (0, ("cam_enc", cam, sp.cam_patch))):
            with torch.cuda.stream(side if (side is not None and mi == 1) else main):
                self._encoder_fwd(A, mi, enc, img, patch, B, S, z_mm, snr_db, train, row_idx)
        if side is not None:
            join = torch.cuda.Event()
            join.record(side)
            main.wait_event(join)
        # ---- channel ----
        y16 = A.buf("y16", (B, 2, k_sym))
        y32 = A.buf("y32", (B, 2, k_sym), F32) if want_aux else None
        scale = A.buf("scale", (B, 2), F32)
        ops.pnorm_channel_fwd(z_mm, channel, snr_db, rician_k=self.rician_k, seed=seed or 0, sample_offset=sample_offset, noise=noise,
                              fading=fading, eps=sp.pnorm_eps, y_f32=y32, y_bf16=y16, scale=scale,
                              z_strides=(k_sym, B * k_sym, B, k_sym), seed_dev=seed_dev, sample_idx=sample_idx,
                              quant_bits=sp.quant_bits, quant_clip=sp.quant_clip)
        # ---- decoder ----
        Sd = sp.dec_tokens
        Md = B * Sd
        xd = [A.buf(f"dec.x{l}", (Md, D)) for l in range(sp.dec_layers + 1)]
        fuse_ln = self.fused_ln and D == 256
        ops.gemm(y16.view(Md, C), f.w16("dec.in_proj.weight"), bias=f.w32("dec.in_proj.bias"),
                 rowadd=f.w16("dec.pos"), out=xd[0], ln=self._block_ln1(A, "dec.blocks.0", Md) if fuse_ln else None)
        final_ln = self._ln_spec(A, "dec.norm", Md, "dec.tn", "dec.mean", "dec.rstd")
        for l in range(sp.dec_layers):
            nxt = self._block_ln1(A, f"dec.blocks.{l + 1}", Md) if l + 1 < sp.dec_layers else final_ln
            self._block_fwd(A, f"dec.blocks.{l}", xd[l], xd[l + 1], B, Sd, train, xn1_ready=fuse_ln, next_ln=nxt)
        tn = final_ln["y"]
        # the reconstruction head and the mean-pool -> trunk -> heads chain both start from tn and are independent: with the
        # second stream available the three small launches of the chain run beside / in the tail of the recon GEMM
        if need_recon:
            recon = A.buf("dec.recon", (Md, D))
            if side is not None:
                fork = torch.cuda.Event()
                fork.record(main)
                side.wait_event(fork)
            with torch.cuda.stream(side if side is not None else main):
                ops.gemm(tn, f.w16("dec.recon.weight"), bias=f.w32("dec.recon.bias"), out=recon)
                if side is not None:
                    join = torch.cuda.Event()
                    join.record(side)
        feat = ops.meanpool_fwd(tn.view(B, Sd, D), out_bf16=A.buf("dec.f", (B, D)))
        # ---- heads ----
        trunk = A.buf("ac.trunk", (B, D))
        ops.gemm(feat, f.w16("ac.trunk.weight"), bias=f.w32("ac.trunk.bias"), act="tanh", out=trunk)
        ho = ops.heads_fwd(trunk, f.w32("ac.head.weight"), f.w32("ac.head.bias"),
                           out=A.buf("ac.ho", (B, sp.n_actions + 1), F32))
        if need_recon and side is not None:
            main.wait_event(join)
        return ho, A

    def l_sem(self, A, B, loss_accum, grad_coef=None):
        """loss_accum[0] += mean((recon - t)^2).  With grad_coef (= d loss / d l_sem) also writes
        d loss / d recon = 2 (recon - t) * grad_coef / numel (bf16) for the backward pass."""
        sp = self.spec
        recon = A["dec.recon"]
        dr = A.buf("dec.drecon", tuple(recon.shape)) if grad_coef is not None else None
        gs = grad_coef / recon.numel() if grad_coef is not None else 0.0
        ops.mse_fwd_bwd(recon.view(B, sp.dec_tokens, sp.d_model), A["cam_enc.t"], loss_accum, dr=dr,
                        grad_scale=gs, t1=A["bev_enc.t"], seg_rows=sp.cam_tokens)

    # ---- backward -------------------------------------------------------------------------
    def backward(self, A, B, dho, cam_unused, channel, snr_db, seed, fading, sample_offset, seed_dev=None,
                 grad_ready=None, sample_idx=None):
        """dho: fp32 [B, A+1] = d loss / d head_out.  `dec.drecon` must hold d loss / d recon (l_sem()).
        Accumulates into flat.g32.  `grad_ready(name)` is called on the compute stream as soon as the gradient
        region `name` of dp.grad_buckets() ("dec", "cam_enc", "bev_enc") is final (data-parallel overlap)."""
        sp, f = self.spec, self.flat
        D, S, C, Sd = sp.d_model, sp.cam_tokens, sp.c_sym, sp.dec_tokens
        Md = B * Sd
        self._side_done = []
        # heads + trunk
        dtrunk = ops.heads_bwd(dho, A["ac.trunk"], f.w32("ac.head.weight"), f.g("ac.head.weight"),
                               f.g("ac.head.bias"), dtrunk_pre=A.buf("ac.dtrunk", (B, D)))
        df = A.buf("dec.df", (B, D), F32)
        self._linear_bwd(dtrunk, A["dec.f"], "ac.trunk", out_f32=df)
        # recon head; d tn = d recon W_rec + (1/Sd) d f   (mean-pool backward fused as a row-group add)
        dtn = A.buf("scratch.dxn.S%d" % Sd, (Md, D))
        self._linear_bwd(A["dec.drecon"], A["dec.tn"], "dec.recon", dx=dtn, rowgrp=df, rowgrp_size=Sd,
                         rowgrp_scale=1.0 / Sd)
        L = sp.dec_layers
        dxa = A.buf("scratch.dxa.S%d" % Sd, (Md, D))
        dxb = A.buf("scratch.dxb.S%d" % Sd, (Md, D))
        ops.layernorm_bwd(dtn, A[f"dec.x{L}"], A["dec.mean"], A["dec.rstd"], f.w32("dec.norm.weight"),
                          f.g("dec.norm.weight"), f.g("dec.norm.bias"), dx=dxa,
                          colsum=f.g(f"dec.blocks.{L - 1}.fc2.bias"))
        cur, nxt = dxa, dxb
        for l in range(L - 1, -1, -1):
            prev_bias = f.g(f"dec.blocks.{l - 1}.fc2.bias") if l > 0 else f.g("dec.in_proj.bias")
            self._block_bwd(A, f"dec.blocks.{l}", A[f"dec.x{l}"], cur, nxt, prev_bias, B, Sd)
            cur, nxt = nxt, cur
        # decoder input projection: cur = d xd0
        ops.colsum(cur.view(B, Sd * D), f.g("dec.pos").view(-1))
        dy = A.buf("dy", (B, 2, sp.k_sym), F32)
        self._linear_bwd(cur, A["y16"].view(Md, C), "dec.in_proj", bias_colsum=False, out_f32=dy.view(Md, C))
        if grad_ready is not None:
            self._join_side()
            grad_ready("dec")
        # channel
        dz_mm = A.buf("dz_mm", (2, B, sp.k_sym))
        ops.pnorm_channel_bwd(dy, A["z_mm"], A["scale"], channel, snr_db, rician_k=self.rician_k, seed=seed or 0,
                              sample_offset=sample_offset, fading=fading, dz_bf16=dz_mm,
                              z_strides=(sp.k_sym, B * sp.k_sym, B, sp.k_sym), seed_dev=seed_dev, sample_idx=sample_idx,
                              quant_bits=sp.quant_bits, quant_clip=sp.quant_clip)
        # encoders (independent of each other: on two streams for small batches, see Engine.__init__)
        side = self._enc_side_stream(B) if grad_ready is None else None
        main = torch.cuda.current_stream(self.device)
        if side is not None:
            fork = torch.cuda.Event()
            fork.record(main)
            side.wait_event(fork)
        for mi, enc in ((1, "bev_enc"), (0, "cam_enc")) if side is not None else ((0, "cam_enc"), (1, "bev_enc")):
            with torch.cuda.stream(side if (side is not None and mi == 1) else main):
                self._encoder_bwd(A, mi, enc, B, dz_mm, snr_db, tag=".bev" if (side is not None and mi == 1) else "")
            if grad_ready is not None:
                self._join_side()
                grad_ready(enc)
        if side is not None:
            join = torch.cuda.Event()
            join.record(side)
            main.wait_event(join)
        self._join_side()                         # the flat gradient bucket is complete on the current stream

    def _encoder_bwd(self, A, mi, enc, B, dz_mm, snr_db, tag=""):
        """backward of one modality encoder from d z_mm[mi]; `tag` names its own scratch set (two-stream backward)"""
        sp, f = self.spec, self.flat
        D, S, C = sp.d_model, sp.cam_tokens, sp.c_sym
        M = B * S
        dz = dz_mm[mi].view(M, C)
        dt = A.buf("scratch%s.dxn.S%d" % (tag, S), (M, D))
        self._linear_bwd(dz, A[f"{enc}.t"], f"{enc}.sym", dx=dt)
        Le = sp.enc_layers
        dxa = A.buf("scratch%s.dxa.S%d" % (tag, S), (M, D))
        dxb = A.buf("scratch%s.dxb.S%d" % (tag, S), (M, D))
        ops.layernorm_bwd(dt, A[f"{enc}.x{Le}"], A[f"{enc}.mean"], A[f"{enc}.rstd"], f.w32(f"{enc}.norm.weight"),
                          f.g(f"{enc}.norm.weight"), f.g(f"{enc}.norm.bias"), dx=dxa,
                          colsum=f.g(f"{enc}.blocks.{Le - 1}.fc2.bias"))
        cur, nxt = dxa, dxb
        for l in range(Le - 1, -1, -1):
            prev_bias = f.g(f"{enc}.blocks.{l - 1}.fc2.bias") if l > 0 else f.g(f"{enc}.patch_embed.bias")
            self._block_bwd(A, f"{enc}.blocks.{l}", A[f"{enc}.x{l}"], cur, nxt, prev_bias, B, S, tag=tag)
            cur, nxt = nxt, cur
        ops.colsum(cur.view(B, S * D), f.g(f"{enc}.pos").view(-1))
        if sp.snr_cond:               # d emb[b] = sum over the sample's tokens of d x0 (mean-pool x S), then the tiny MLP
            demb = ops.meanpool_fwd(cur.view(B, S, D), out_f32=A.buf(f"{enc}.snr_demb", (B, D), F32))
            ops.snr_embed_bwd(demb, A[f"{enc}.snr_h"], snr_db, f.w32(f"{enc}.snr.w"), f.g(f"{enc}.snr.w"),
                              f.g(f"{enc}.snr.u"), f.g(f"{enc}.snr.v"), demb_scale=float(S))
        self._linear_bwd(cur, A[f"{enc}.tok"], f"{enc}.patch_embed", bias_colsum=False)

    def _join_side(self):
        if self._side_done:
            torch.cuda.current_stream(self.device).wait_event(self._side_done[-1])


# ----------------------------------------------------------------------------------------------
# nn.Module facade
# ----------------------------------------------------------------------------------------------
class _ArenaLease:
    """ownership of one activation arena by one autograd node; the arena returns to the engine's pool when the node
    (and every other holder of the lease) is gone"""

    def __init__(self, arena: _Arena, pool: list):
        self.arena, self._pool = arena, pool

    def __del__(self):
        try:
            self._pool.append(self.arena)
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass


class _PolicyFn(torch.autograd.Function):
    """one autograd node around the whole engine (generic API path; PPOAgent bypasses it).  Each node owns a private
    activation arena (lease), so several grad-mode forwards may be outstanding before their backwards run."""

    @staticmethod
    def forward(ctx, policy, cam, bev, kw, lease, *params):
        eng = policy.engine
        B = cam.shape[0]
        eng.rician_k = kw["rician_k"]
        ho, A = eng.forward(cam, bev, kw["snr_db"], kw["channel"], kw["seed"], kw["noise"], kw["fading"],
                            kw["sample_offset"], train=True, want_aux=True, arena=lease.arena)
        lsem = torch.zeros(1, device=cam.device)
        eng.l_sem(A, B, lsem)
        ctx.policy, ctx.kw, ctx.B, ctx.A, ctx.lease, ctx.generation = policy, kw, B, A, lease, A.generation
        return ho.clone(), lsem

    @staticmethod
    def backward(ctx, dho, dlsem):
        policy, kw, B, A = ctx.policy, ctx.kw, ctx.B, ctx.A
        if A.generation != ctx.generation:
            raise RuntimeError("SemComPolicy: the activation arena of this autograd node was overwritten by a later "
                               "Engine.forward (generation %d -> %d); its gradients would be wrong" %
                               (ctx.generation, A.generation))
        eng, flat = policy.engine, policy.flat
        saved = flat.g32
        flat.g32 = torch.zeros_like(saved)                  # scratch bucket for this backward
        try:
            eng.rician_k = kw["rician_k"]
            eng.l_sem(A, B, torch.zeros(1, device=dho.device), grad_coef=float(dlsem))
            eng.backward(A, B, dho.contiguous().float(), None, kw["channel"], kw["snr_db"], kw["seed"],
                         kw["fading"], kw["sample_offset"])
            grads = [flat.g(n).clone() for n in flat.shapes]
        finally:
            flat.g32 = saved
        return (None, None, None, None, None, *grads)


class SemComPolicy(nn.Module):
    def __init__(self, spec: SpecV0 = SpecV0(), device=None, init_std: float | None = None, seed: int = 0,
                 peer: bool | None = None):
        """peer: keep the flat parameter / gradient buckets in NVLink peer memory (the fused data-parallel optimizer
        step, peer.PeerReduceAdam, maps them into the other ranks).  Default: on when torch.distributed runs with more
        than one rank on the NCCL backend (one process per GPU) and SC_DP_FUSED is not 0."""
        super().__init__()
        _lib.require_device()
        self.spec = spec
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if peer is None:
            import torch.distributed as dist
            peer = (os.environ.get("SC_DP_FUSED", "1") != "0" and dist.is_available() and dist.is_initialized()
                    and dist.get_world_size() > 1 and dist.get_backend() == "nccl")
        self.flat = FlatParams(spec, self.device, peer=peer)
        _init_flat(self.flat, spec, spec.init_std if init_std is None else init_std, seed)
        self.engine = Engine(spec, self.flat, self.device)
        for name in self.flat.shapes:                       # nested parameter tree -> oracle state_dict keys
            mod = self
            parts = name.split(".")
            for part in parts[:-1]:
                if not hasattr(mod, part):
                    mod.add_module(part, nn.Module())
                mod = getattr(mod, part)
            mod.register_parameter(parts[-1], nn.Parameter(self.flat.w32(name)))

    # weights changed through torch (load_state_dict, manual edits): refresh the bf16 shadow
    def sync(self):
        self.flat.sync_shadow()

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        r = super().load_state_dict(state_dict, strict=strict, assign=False)
        self.sync()
        return r

    def _kw(self, snr_db, channel, seed, noise, fading, sample_offset, rician_k=None):
        if channel not in ("none", "awgn", "rayleigh", "rician"):
            raise ValueError(f"unknown channel kind {channel!r}")
        if channel != "none" and seed is None and noise is None:
            raise ValueError("channel needs seed= or noise=")
        if channel in ("rayleigh", "rician") and seed is None and fading is None:
            raise ValueError(f"{channel} channel needs seed= or fading=")
        if torch.is_tensor(snr_db):
            snr_db = snr_db.to(self.device, F32).contiguous()
            if snr_db.dim() == 0:
                snr_db = float(snr_db)
        return dict(snr_db=snr_db, channel=channel, seed=seed,
                    rician_k=self.spec.rician_k if rician_k is None else float(rician_k),
                    noise=None if noise is None else noise.to(self.device, F32).contiguous(),
                    fading=None if fading is None else fading.to(self.device, F32).contiguous(),
                    sample_offset=sample_offset)

    def forward(self, cam, bev, snr_db, *, channel: str = "awgn", seed=None, noise=None, fading=None,
                sample_offset: int = 0, rician_k=None):
        kw = self._kw(snr_db, channel, seed, noise, fading, sample_offset, rician_k)
        # observations may be fp32 or bf16 (the first kernel casts to bf16 either way)
        cam = cam.to(self.device, cam.dtype if cam.dtype == BF16 else F32).contiguous()
        bev = bev.to(self.device, bev.dtype if bev.dtype == BF16 else F32).contiguous()
        B, A_ = cam.shape[0], self.spec.n_actions
        if torch.is_grad_enabled():
            lease = self.engine.lease_arena(B)
            ho, lsem = _PolicyFn.apply(self, cam, bev, kw, lease, *self.parameters())
            arena = lease.arena
        else:
            self.engine.rician_k = kw["rician_k"]
            ho, arena = self.engine.forward(cam, bev, kw["snr_db"], channel, seed, kw["noise"], kw["fading"],
                                            sample_offset, train=False, want_aux=True)
            ho = ho.clone()
            lsem = torch.zeros(1, device=self.device)
            self.engine.l_sem(arena, B, lsem)
        zn = (arena["z_mm"] * arena["scale"].t().unsqueeze(-1)).permute(1, 0, 2)      # API view only
        return ho[:, :A_], ho[:, A_], {"l_sem": lsem[0], "z": zn, "y": arena["y32"].clone()}

    @torch.no_grad()
    def act(self, cam, bev, snr_db, u, **kw):
        logits, value, _ = self.forward(cam, bev, snr_db, **kw)
        ho = torch.cat([logits, value[:, None]], dim=1).contiguous()
        a, logp = ops.sample_action(ho, u.to(self.device, F32).contiguous(), self.spec.n_actions)
        return a, logp, value
