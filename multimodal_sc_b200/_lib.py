"""ctypes binding of libsemcom_sm100.so (the C ABI in include/semcom_sm100.h).

The library is the product: if it is missing or the device is not sm_100 every call raises
`SemcomError` — there is no eager / CPU fallback behind these functions.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

LIB_PATH = Path(__file__).resolve().parent / "libsemcom_sm100.so"

ABI_VERSION = 4          # SC_ABI_VERSION in include/semcom_sm100.h (3: index arguments for the fused minibatch gather /
                         # ragged data-parallel slices, quantised symbols, SNR embedding, erf GELU, sc_step_begin;
                         # 4: peer-memory helpers, sc_dp_barrier, sc_dp_reduce_adam)
SC_DP_MAX_WORLD = 16
SC_PEER_HANDLE_BYTES = 64
SC_CH = {"none": 0, "awgn": 1, "rayleigh": 2, "rician": 3}
SC_ACT = {None: 0, "none": 0, "gelu": 1, "tanh": 2, "gelu_erf": 3}


class SemcomError(RuntimeError):
    pass


class GemmEpilogue(C.Structure):
    _fields_ = [
        ("bias", C.c_void_p),
        ("rowadd", C.c_void_p),
        ("rowadd_period", C.c_int64),
        ("rowgrp", C.c_void_p),
        ("rowgrp_size", C.c_int64),
        ("rowgrp_scale", C.c_float),
        ("alpha", C.c_float),
        ("resid", C.c_void_p),
        ("aux_in", C.c_void_p),
        ("aux_out", C.c_void_p),
        ("out_bf16", C.c_void_p),
        ("out_f32", C.c_void_p),
        ("ldo", C.c_int64),
        ("act", C.c_int32),
        ("atomic", C.c_int32),
        ("rowadd_bf16", C.c_void_p),
        ("colsum", C.c_void_p),
        ("ln_out", C.c_void_p),
        ("ln_gamma", C.c_void_p),
        ("ln_beta", C.c_void_p),
        ("ln_mean", C.c_void_p),
        ("ln_rstd", C.c_void_p),
        ("ln_eps", C.c_float),
        ("aux_in_erf", C.c_int32),
    ]


_P, _I64, _I, _F, _U64, _U32 = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_uint64, C.c_uint32

# name -> argtypes (restype is int unless listed in _RESTYPES)
_PROTOS = {
    "sc_version": [],
    "sc_last_error": [],
    "sc_device_ok": [],
    "sc_launch_count": [],
    "sc_debug_set": [C.c_char_p, _I64],
    "sc_gemm_bf16": [_P, _P, _I64, _I64, _I64, _I, _I, _I64, _I64, C.POINTER(GemmEpilogue), _I, _P],
    "sc_mlp_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I, _I, _P],
    "sc_mlp_bwd_dgrad": [_P, _P, _P, _P, _P, _P, _I64, _I, _I, _P],
    "sc_mlp_fwd_ln": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _F, _P, _P, _P, _I64, _I, _I, _P],
    "sc_attn_fwd": [_P, _P, _P, _I64, _I, _I, _P],
    "sc_attn_bwd": [_P, _P, _P, _P, _P, _I64, _I, _I, _P],
    "sc_patchify_bf16": [_P, _I, _P, _I64, _I, _I, _I, _I, _P, _P],
    "sc_layernorm_fwd": [_P, _P, _P, _P, _P, _P, _I64, _I, _F, _I, _P],
    "sc_layernorm_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I, _I, _P],
    "sc_colsum_bf16": [_P, _P, _I64, _I, _I64, _P],
    "sc_pnorm_channel_fwd": [_P, _I64, _I64, _P, _P, _I64, _I64, _P, _I64, _I, _I, _F, _P, _F, _U64, _I64,
                             _P, _P, _P, _F, _P, _I, _F, _P],
    "sc_pnorm_channel_bwd": [_P, _I64, _I64, _P, _I64, _I64, _P, _P, _P, _I64, _I, _I, _F, _P, _F, _U64,
                             _I64, _P, _P, _P, _I, _F, _P],
    "sc_philox_normals": [_P, _I64, _U64, _I64, _I, _P],
    "sc_philox_words": [_P, _I64, _U32, _U64, _I64, _P],
    "sc_meanpool_fwd": [_P, _P, _P, _I64, _I, _I, _P],
    "sc_mse_fwd_bwd": [_P, _P, _P, _I, _I, _P, _P, _I64, _F, _F, _P],
    "sc_heads_fwd": [_P, _P, _P, _P, _I64, _I, _I, _P],
    "sc_heads_bwd": [_P, _P, _P, _P, _P, _P, _I64, _I, _I, _P],
    "sc_sample_action": [_P, _I, _P, _P, _P, _I64, _I, _P],
    "sc_gae": [_P, _P, _P, _P, _P, _P, _P, _I, _I64, _F, _F, _P],
    "sc_adv_normalize": [_P, _P, _I64, _I64, _F, _P],
    "sc_ppo_loss": [_P, _I, _P, _P, _P, _P, _P, _P, _P, _I64, _I, _F, _F, _F, _F, _P],
    "sc_adam_flat": [_P, _P, _P, _P, _P, _I64, _F, _F, _F, _F, _I64, _F, _P, _P],
    "sc_cast_f32_bf16": [_P, _P, _I64, _P],
    "sc_step_begin": [_P, _I64, _P, _I, _P, _P],
    "sc_gather_rows": [_P, _I64, _I64, _P, _P, _I64, _I, _P],
    "sc_snr_embed_fwd": [_P, _F, _P, _P, _P, _P, _P, _I64, _I, _I, _P],
    "sc_snr_embed_bwd": [_P, _P, _P, _F, _P, _P, _P, _P, _F, _I64, _I, _I, _P],
    "sc_peer_alloc": [C.POINTER(C.c_void_p), _I64],
    "sc_peer_free": [_P],
    "sc_peer_export": [_P, _P],
    "sc_peer_import": [_P, C.POINTER(C.c_void_p)],
    "sc_peer_close": [_P],
    "sc_dp_barrier": [C.POINTER(C.c_void_p), _P, _I, _I, _P],
    "sc_dp_reduce_adam": [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _P, _P, _I, _I, _I64,
                          _F, _F, _F, _F, _I64, _P, _I, _P],
}
_RESTYPES = {"sc_last_error": C.c_char_p, "sc_launch_count": C.c_int64}

EXPORTED = tuple(_PROTOS)

_lib = None


def _missing(name):
    def stub(*_a, **_k):
        raise SemcomError(f"{LIB_PATH.name} does not export {name}; rebuild with "
                          "`python -m multimodal_sc_b200.build --force`")
    return stub


def missing_exports():
    """names declared in include/semcom_sm100.h that the built library lacks (tests use this)."""
    lib = C.CDLL(str(LIB_PATH))
    return [n for n in _PROTOS if not hasattr(lib, n)]


def load() -> C.CDLL:
    """dlopen the library and attach prototypes.  Raises SemcomError when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise SemcomError(
            f"{LIB_PATH} is missing: build it with `python -m multimodal_sc_b200.build` "
            "(there is no CPU/eager fallback)")
    lib = C.CDLL(str(LIB_PATH))
    for name, argtypes in _PROTOS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:           # stale build: fail loudly at call time, not with a fallback
            setattr(lib, name, _missing(name))
            continue
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    if lib.sc_version() != ABI_VERSION:   # a stale .so would be called with the wrong argument lists
        raise SemcomError(f"{LIB_PATH.name} has ABI version {lib.sc_version()}, this package needs {ABI_VERSION}: "
                          "rebuild with `python -m multimodal_sc_b200.build --force`")
    _lib = lib
    return lib


def last_error() -> str:
    return (load().sc_last_error() or b"").decode()


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise SemcomError(f"{what or 'libsemcom_sm100'} failed (rc={rc}): {last_error()}")


def stream_ptr(ref=None) -> int:
    """current stream of the device the operands live on (`ref`: any operand tensor), not of the current device"""
    if ref is not None and ref.is_cuda:
        return torch.cuda.current_stream(ref.device).cuda_stream
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int | None:
    """device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def launch_count() -> int:
    return int(load().sc_launch_count())


def require_device() -> None:
    if not torch.cuda.is_available():
        raise SemcomError("no CUDA device: the sm_100a path has no CPU fallback")
    check(load().sc_device_ok(), "sc_device_ok")
