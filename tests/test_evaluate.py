"""SNR-sweep evaluation tooling: properties on the oracle (CPU), parity native vs oracle (GPU)."""
import pytest
import torch

from multimodal_sc_b200.evaluate import snr_sweep
from multimodal_sc_b200.spec import SpecV0
from oracle.model import SemComPolicy as OraclePolicy

SPEC = SpecV0()


def _inputs(B, seed=0):
    g = torch.Generator().manual_seed(seed)
    cam = torch.rand(B, 3, 128, 128, generator=g)
    bev = (torch.rand(B, 1, 64, 64, generator=g) < 0.1).float() * torch.rand(B, 1, 64, 64, generator=g)
    return cam, bev


def test_sweep_properties_on_the_oracle():
    torch.set_num_threads(4)
    o = OraclePolicy(SPEC)
    o.reset_parameters(std=0.05, seed=2)
    cam, bev = _inputs(4)
    snrs = (-5.0, 5.0, 20.0, 60.0)
    res = snr_sweep(o, cam, bev, snrs, channels=("awgn", "rayleigh", "rician"), seed=9)
    assert set(res) == {"awgn", "rayleigh", "rician"}
    for ch, rows in res.items():
        assert [r["snr_db"] for r in rows] == list(snrs)
        mse = [r["sym_mse"] for r in rows]
        assert all(a > b for a, b in zip(mse, mse[1:])), (ch, mse)        # common noise realisation: monotone in SNR
        assert mse[-1] < 1e-4                                                # 60 dB: the link is transparent
        assert rows[-1]["action_agreement"] == 1.0
        assert all(0.0 < r["entropy"] <= torch.log(torch.tensor(float(SPEC.n_actions))) + 1e-5 for r in rows)
    # AWGN symbol error is the noise variance itself: sigma^2 = 10^(-snr/10) per real dimension
    assert abs(res["awgn"][1]["sym_mse"] / 10 ** (-0.5) - 1.0) < 0.05
    # fading only hurts (MMSE over a random gain) and a strong line of sight sits between Rayleigh and AWGN
    assert res["rayleigh"][1]["sym_mse"] != res["awgn"][1]["sym_mse"]


@pytest.mark.gpu
def test_sweep_native_matches_oracle(cuda_lib):
    from multimodal_sc_b200.model import SemComPolicy
    o = OraclePolicy(SPEC)
    o.reset_parameters(std=0.05, seed=2)
    n = SemComPolicy(SPEC)
    n.load_state_dict(o.state_dict())
    cam, bev = _inputs(6)
    snrs = (-5.0, 10.0, 20.0)
    ro = snr_sweep(o, cam, bev, snrs, channels=("awgn", "rician"), seed=9)
    rn = snr_sweep(n, cam.cuda(), bev.cuda(), snrs, channels=("awgn", "rician"), seed=9)
    for ch in ro:
        for a, b in zip(ro[ch], rn[ch]):
            assert abs(a["sym_mse"] - b["sym_mse"]) <= 2e-2 * a["sym_mse"] + 1e-6, (ch, a, b)
            assert abs(a["l_sem"] - b["l_sem"]) <= 2e-2 * a["l_sem"], (ch, a, b)
            assert abs(a["entropy"] - b["entropy"]) <= 2e-2 * a["entropy"], (ch, a, b)
            assert abs(a["value_mean"] - b["value_mean"]) <= 2e-2 * max(abs(a["value_mean"]), 0.25), (ch, a, b)   # values are O(0.3)
