"""CPU restatement of the CVS mechanistic right-hand side (TEST INFRASTRUCTURE).

Follows ``data/cvs/cvs_data.py``:

* ``cvs_rhs``            -- ``dx_dt`` (``:52-91``), float64 NumPy, one trajectory or a batch
                            (last axis = 4 states ``(Pa/100, Pv/10, S, SV/100)``);
* ``CVS_CONST``          -- the shared constants of ``get_random_params`` (``:28-49``);
* ``observe``            -- ``states_trajectory_to_sample`` (``:94-103``): ``(x0, x1, f_hr(x2))``;
* ``lsoda_trajectories`` -- ``create_cvs_data`` loop (``:111-134``): ``scipy.integrate.odeint``
                            (LSODA, default tolerances) from ``x(0)=ones(4)`` at ``t=0..seq_len-1``;
* ``CvsRhsTorch``        -- the same RHS as a ``forward(t, state)`` torch module (for the
                            solver oracle and autograd reference gradients).

PINNED: ``tests/test_oracle_cvs.py`` checks ``lsoda_trajectories`` against the reference's own
golden trajectories (``data/cvs/test_latent_data.pkl`` / ``gt_test_data.pkl``, a slice of which is
committed as ``tests/golden/cvs_golden.npz``).
"""
from __future__ import annotations

import numpy as np
import torch

CVS_CONST = dict(
    f_hr_max=3.0, f_hr_min=2.0 / 3.0, r_tpr_max=2.134, r_tpr_min=0.5335, sv_mod=0.0001,
    ca=4.0, cv=111.0, k_width=0.1838, p_aset=70.0, tau=20.0,
)


def cvs_rhs(state, i_ext, r_tpr_mod, c=CVS_CONST):
    """d(state)/dt; ``state[..., 4]``, ``i_ext`` / ``r_tpr_mod`` broadcast over the leading axes."""
    state = np.asarray(state, dtype=np.float64)
    p_a = 100.0 * state[..., 0]
    p_v = 10.0 * state[..., 1]
    s = state[..., 2]
    sv = 100.0 * state[..., 3]
    f_hr = s * (c["f_hr_max"] - c["f_hr_min"]) + c["f_hr_min"]
    r_tpr = s * (c["r_tpr_max"] - c["r_tpr_min"]) + c["r_tpr_min"] - r_tpr_mod
    dva = -1.0 * (p_a - p_v) / r_tpr + sv * f_hr
    dvv = -1.0 * dva + i_ext
    dpa = dva / (c["ca"] * 100.0)
    dpv = dvv / (c["cv"] * 10.0)
    ds = (1.0 / c["tau"]) * (1.0 - 1.0 / (1 + np.exp(-1 * c["k_width"] * (p_a - c["p_aset"]))) - s)
    dsv = i_ext * c["sv_mod"] * np.ones_like(s)
    return np.stack([dpa, dpv, ds, dsv], axis=-1)


def observe(states, c=CVS_CONST):
    f_hr = states[..., 2] * (c["f_hr_max"] - c["f_hr_min"]) + c["f_hr_min"]
    return np.stack([states[..., 0], states[..., 1], f_hr], axis=-1)


def lsoda_trajectories(i_ext, r_tpr_mod, seq_len=86, delta_t=1.0):
    """(N, seq_len, 4) float64 latent trajectories, one LSODA solve per sample like the reference."""
    from scipy import integrate

    t = np.arange(0.0, stop=seq_len * delta_t, step=delta_t)
    out = np.zeros((len(i_ext), seq_len, 4))
    for n, (ie, rm) in enumerate(zip(i_ext, r_tpr_mod)):
        out[n] = integrate.odeint(lambda x, _t: cvs_rhs(x, ie, rm), np.ones(4), t)
    return out


class CvsRhsTorch(torch.nn.Module):
    """``forward(t, state)`` with per-trajectory ``i_ext (B,)`` and ``r_tpr_mod (B,)``."""

    def __init__(self, i_ext, r_tpr_mod):
        super().__init__()
        self.i_ext = i_ext
        self.r_tpr_mod = r_tpr_mod

    def forward(self, t, state):
        c = CVS_CONST
        p_a = 100.0 * state[..., 0]
        p_v = 10.0 * state[..., 1]
        s = state[..., 2]
        sv = 100.0 * state[..., 3]
        f_hr = s * (c["f_hr_max"] - c["f_hr_min"]) + c["f_hr_min"]
        r_tpr = s * (c["r_tpr_max"] - c["r_tpr_min"]) + c["r_tpr_min"] - self.r_tpr_mod
        dva = -1.0 * (p_a - p_v) / r_tpr + sv * f_hr
        dvv = -1.0 * dva + self.i_ext
        dpa = dva / (c["ca"] * 100.0)
        dpv = dvv / (c["cv"] * 10.0)
        ds = (1.0 / c["tau"]) * (1.0 - 1.0 / (1 + torch.exp(-1 * c["k_width"] * (p_a - c["p_aset"]))) - s)
        dsv = self.i_ext * c["sv_mod"] * torch.ones_like(s)
        return torch.stack([dpa, dpv, ds, dsv], dim=-1)
