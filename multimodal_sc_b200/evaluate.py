"""Channel-robustness evaluation (BASELINE.json configs[3]): the forward pass of a policy over an
SNR grid and one or more channel kinds, reporting the semantic distortion and the policy's behaviour at every point.

Works on anything that exposes the hot path's `forward(cam, bev, snr_db, *, channel, seed, ...)` — the native
`SemComPolicy` (sm_100a kernels) and the oracle (`oracle/model.py`) alike, which is how the sweep itself is
parity-tested.  No arithmetic of the path lives here: per grid point it is one call of the policy's own forward plus
reductions of its outputs.  Upstream has no evaluation code (/root/reference/README.md:1-2)."""
from __future__ import annotations

import torch


@torch.no_grad()
def snr_sweep(policy, cam, bev, snrs_db=(-5.0, 0.0, 5.0, 10.0, 15.0, 20.0), channels=("awgn", "rayleigh"), *,
              seed: int = 1234, rician_k=None):
    """-> {channel: [{snr_db, l_sem, sym_mse, entropy, value_mean, action_agreement}, ...]} in the order of `snrs_db`.

    l_sem             semantic distortion of the reconstruction head (the training loss term)
    sym_mse           mean squared error between the transmitted and the equalised received symbols
    entropy           mean entropy of the action distribution (nats)
    value_mean        mean critic output
    action_agreement  fraction of samples whose greedy action equals the greedy action over a clean link
                      (`channel="none"`: power normalisation only)
    The same `seed` is used at every grid point, so the noise realisation is common and the curves are smooth in SNR."""
    ref_logits, _, _ = policy(cam, bev, 0.0, channel="none", seed=seed)
    ref_action = ref_logits.float().argmax(dim=-1)
    out = {}
    for ch in channels:
        rows = []
        for snr in snrs_db:
            kw = {"rician_k": rician_k} if (ch == "rician" and rician_k is not None) else {}
            logits, value, aux = policy(cam, bev, float(snr), channel=ch, seed=seed, **kw)
            logits = logits.float()
            logp = torch.log_softmax(logits, dim=-1)
            ent = -(logp.exp() * logp).sum(-1).mean()
            rows.append({"snr_db": float(snr), "l_sem": float(aux["l_sem"]),
                         "sym_mse": float((aux["y"].float() - aux["z"].float()).pow(2).mean()),
                         "entropy": float(ent), "value_mean": float(value.float().mean()),
                         "action_agreement": float((logits.argmax(dim=-1) == ref_action).float().mean())})
        out[ch] = rows
    return out
