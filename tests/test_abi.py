"""CPU checks of the drop-in boundary: the C-ABI library builds, loads, exports every symbol
include/semcom_sm100.h declares, refuses to run without an sm_100 device (no fallback), and the
host-side parameter tree mirrors the oracle's state_dict."""
import re
from pathlib import Path

import pytest
import torch

from multimodal_sc_b200 import _lib
from multimodal_sc_b200.model import FlatParams, param_shapes
from multimodal_sc_b200.spec import SpecV0

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "semcom_sm100.h"


@pytest.fixture(scope="module")
def built():
    from multimodal_sc_b200.build import build_library
    return build_library()


def _declared():
    txt = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(sc_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(built):
    declared = _declared()
    assert len(declared) >= 25
    assert _lib.missing_exports() == []
    import ctypes
    lib = ctypes.CDLL(str(built))
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.EXPORTED) == declared, "ctypes prototypes out of sync with the header"


def test_version_and_no_fallback(built):
    lib = _lib.load()
    assert lib.sc_version() == _lib.ABI_VERSION == 4
    if not torch.cuda.is_available():
        assert lib.sc_device_ok() == -3                       # SC_E_ARCH: no device -> refuse
        assert "no CUDA device" in _lib.last_error() or "sm_100" in _lib.last_error()
        with pytest.raises(_lib.SemcomError):
            _lib.require_device()
        from multimodal_sc_b200 import ops
        with pytest.raises(_lib.SemcomError):                 # CPU tensors are rejected, never computed on the host
            ops.layernorm_fwd(torch.zeros(2, 256), torch.ones(256), torch.zeros(256), 1e-5)
        from multimodal_sc_b200.model import SemComPolicy
        with pytest.raises(_lib.SemcomError):
            SemComPolicy(SpecV0())


def test_param_tree_matches_oracle():
    from oracle.model import SemComPolicy as Oracle
    sd = Oracle(SpecV0()).state_dict()
    ps = param_shapes(SpecV0())
    assert list(sd.keys()) == list(ps.keys())
    assert all(tuple(sd[k].shape) == ps[k] for k in ps)


def test_flat_bucket_layout():
    fp = FlatParams(SpecV0(), "cpu")
    assert fp.n_params() == 5165866
    prev_end = 0
    for name, (off, n) in fp.offsets.items():
        assert off % FlatParams.ALIGN == 0 and off >= prev_end      # 256-byte aligned, non-overlapping
        prev_end = off + n
        assert fp.w32(name).shape == fp.shapes[name]
        assert fp.w32(name).data_ptr() == fp.p32.data_ptr() + 4 * off
        assert fp.w16(name).data_ptr() % 16 == 0                    # TMA base alignment of the bf16 shadow
    assert fp.numel >= prev_end


def test_peer_memory_has_no_host_fallback(built):
    """peer.py: exportable device memory and the fused data-parallel optimizer step exist on the GPU only"""
    from multimodal_sc_b200 import peer
    with pytest.raises(_lib.SemcomError):
        peer.peer_empty(16, torch.float32, "cpu")
    assert not peer.is_peer(torch.zeros(4))
    # the slice rule of sc_dp_reduce_adam (4-float granularity) is dp.shard_bounds on n / 4
    from multimodal_sc_b200 import dp
    n4, world = 1291504, 8
    cover = [dp.shard_bounds(n4, world, r) for r in range(world)]
    assert cover[0][0] == 0 and cover[-1][1] == n4 and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))

