"""Generate the committed golden fixtures.  Run ONCE in the build container (needs /root/reference):

    python tests/golden/make_golden.py

* ``cvs_golden.npz``      -- a slice of the reference's own golden CVS trajectories
                             (``data/cvs/test_latent_data.pkl``, ``gt_test_data.pkl``,
                             ``test_params_data.pkl``): pins the mechanistic RHS + integration.
* ``blackbox_golden.npz`` -- outputs of the reference's REAL ``OdeModel`` / ``Dynamics`` / ``OdeFunc``
                             classes (``models/blackbox_ode.py``, imported unchanged) on seeded inputs,
                             float32, CPU.  ``torchdiffeq`` itself is absent, so the solver loop under
                             those classes is ``oracle.torchdiffeq_oracle`` (parity unpinned for the loop,
                             pinned for the RHS / wiring / weight layout).
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle import shims  # noqa: E402
from oracle import torchdiffeq_oracle as tde  # noqa: E402

REF = shims.REFERENCE_ROOT


def cvs():
    lat = torch.load(f"{REF}/data/cvs/test_latent_data.pkl", weights_only=False)
    gt = torch.load(f"{REF}/data/cvs/gt_test_data.pkl", weights_only=False)
    par = torch.load(f"{REF}/data/cvs/test_params_data.pkl", weights_only=False)
    n = 24
    np.savez_compressed(os.path.join(HERE, "cvs_golden.npz"), latent=lat[:n], gt=gt[:n],
                        i_ext=par["i_ext"][:n], r_tpr_mod=par["r_tpr_mod"][:n])


def proc_like_times(T=100):
    # the proc grid is non-uniform float32 hours 0 -> 19.25 with dt in [0.193, 0.196]
    g = torch.Generator().manual_seed(7)
    dt = 0.193 + 0.003 * torch.rand(T - 1, generator=g)
    return torch.cat([torch.zeros(1), torch.cumsum(dt, 0)]).float()


def blackbox():
    bb, _ = shims.import_reference_blackbox()
    out = {}
    cases = [
        # name, B, L, H, S, times, methods
        ("cvs", 8, 15, 25, 5, torch.arange(0.0, 86.0, 1.0), ("euler", "midpoint", "rk4")),
        ("proc", 4, 50, 25, 8, proc_like_times(), ("midpoint", "rk4")),
        ("chal", 6, 15, 25, 5, torch.arange(0.0, 24.0, 1.0), ("dopri5",)),
    ]
    for name, B, L, H, S, times, methods in cases:
        torch.manual_seed(12)
        with contextlib.redirect_stdout(io.StringIO()):
            m = bb.OdeModel()
            m.init_with_params(times=times, ode_state_dim=S, latent_dim=L, ode_hidden_dim=H,
                               adjoint_solver=False, solver="midpoint", device="cpu")
        z0 = torch.randn(B, L)
        G = torch.randn(B, len(times), S)
        out[f"{name}/times"] = times.numpy()
        out[f"{name}/z"] = z0.numpy()
        out[f"{name}/G"] = G.numpy()
        for k, v in m.state_dict().items():
            if ".prod." in k or ".degr." in k:
                continue
            out[f"{name}/w/{k}"] = v.numpy()
        for method in methods:
            for adj in (False, True):
                m.solver, m.adjoint_solver = method, adj
                m.zero_grad()
                z = z0.clone().requires_grad_(True)
                if method == "dopri5":
                    x0 = m.initialize_state(z)
                    x0.retain_grad()
                    f = m.gen_dynamics(z)
                    solve = tde.odeint_adjoint if adj else tde.odeint
                    sol = solve(f, x0, times, method="dopri5", rtol=1e-5, atol=1e-6).permute(1, 0, 2)
                    out[f"{name}/{method}/{int(adj)}/accepted"] = np.array(tde.last_stats.accepted)
                    out[f"{name}/{method}/{int(adj)}/dts"] = np.array(tde.last_stats.dts)
                else:
                    sol = m.solve_ODE(z)
                (sol * G).sum().backward()
                key = f"{name}/{method}/{int(adj)}"
                if method == "dopri5" and adj:
                    # the backward pass of odeint_adjoint: one adaptive solve per output interval; its step log in the
                    # device's format (interval index, -, step size, accepted) so that the sequence can be replayed
                    rows = [(float(i), 0.0, d, 1.0 if a else 0.0) for i, acc, dts in tde.last_adjoint_intervals
                            for a, d in zip(acc, dts)]
                    out[f"{key}/backward_steps"] = np.array(rows, dtype=np.float64)
                    out[f"{key}/grad_y0"] = x0.grad.numpy() if x0.grad is not None else None
                out[f"{key}/sol"] = sol.detach().numpy()
                out[f"{key}/grad_z"] = z.grad.numpy()
                for k, p in m.named_parameters():
                    if ".prod." in k or ".degr." in k:
                        continue
                    out[f"{key}/g/{k}"] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "blackbox_golden.npz"), **out)
    print({k: v.shape for k, v in out.items() if k.endswith("sol")})


if __name__ == "__main__":
    cvs()
    blackbox()
