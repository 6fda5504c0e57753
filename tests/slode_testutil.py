"""Shared helpers for the parity tests: build oracle/product model pairs on seeded inputs."""
import contextlib
import io

import numpy as np
import torch

from oracle import slode_port


def proc_like_times(T=100, seed=7):
    g = torch.Generator().manual_seed(seed)
    dt = 0.193 + 0.003 * torch.rand(T - 1, generator=g)
    return torch.cat([torch.zeros(1), torch.cumsum(dt, 0)]).float()


SHAPES = {
    # name: (L, H, S, times)
    "cvs": (15, 25, 5, torch.arange(0.0, 86.0, 1.0)),
    "chal": (15, 25, 5, torch.arange(0.0, 142.0, 1.0)),
    "proc": (50, 25, 8, proc_like_times()),
    "small": (6, 16, 4, torch.linspace(0.0, 3.0, 17)),
    "h32": (15, 32, 5, torch.arange(0.0, 40.0, 1.0)),
    "h64": (15, 64, 5, torch.arange(0.0, 30.0, 1.0)),
    # BASELINE configs[4] width sweep: per-trajectory tables in global scratch, rolled unit loops
    "h128": (15, 128, 5, torch.arange(0.0, 20.0, 1.0)),
    "h256": (15, 256, 5, torch.arange(0.0, 12.0, 1.0)),
    "h512": (15, 512, 5, torch.arange(0.0, 9.0, 1.0)),
}


def make_oracle(shape, method, adjoint, seed=12, dtype=torch.float32):
    L, H, S, times = SHAPES[shape]
    torch.manual_seed(seed)
    m = slode_port.OdeModel(times.to(dtype), S, L, H, adjoint, method).to(dtype)
    return m


def make_oracle_like(oracle_model, dtype):
    """A copy of an oracle model (same weights, grid and solver settings) in another dtype."""
    d = oracle_model.dynamics
    m = slode_port.OdeModel(oracle_model.times.to(dtype), d.n_outputs, d.n_inputs, d.dynamics_hidden.out_features,
                            oracle_model.adjoint_solver, oracle_model.solver).to(dtype)
    m.load_state_dict({k: v.to(dtype) for k, v in oracle_model.state_dict().items()})
    return m


def make_product(oracle_model, device="cuda", layout="tbs"):
    import structured_latent_odes_b200 as slode

    d = oracle_model.dynamics
    L, H, S = d.n_inputs, d.dynamics_hidden.out_features, d.n_outputs
    with contextlib.redirect_stdout(io.StringIO()):
        p = slode.OdeModel()
        p.init_with_params(times=oracle_model.times.float().to(device), ode_state_dim=S, latent_dim=L,
                           ode_hidden_dim=H, adjoint_solver=oracle_model.adjoint_solver,
                           solver=oracle_model.solver, device=device, layout=layout)
    p.load_state_dict({k: v.float() for k, v in oracle_model.state_dict().items()})
    return p.to(device)


def run_fwd_bwd(model, z, G):
    """Returns sol (B,T,S), grad_z, {param: grad} for loss = sum(sol * G)."""
    model.zero_grad()
    z = z.clone().requires_grad_(True)
    sol = model.solve_ODE(z)
    (sol * G).sum().backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()
             if p.grad is not None and ".prod." not in k and ".degr." not in k}
    return sol.detach(), (z.grad.detach().clone() if z.grad is not None else None), grads


def rel_err(a, b):
    """max|a-b| / max|b| (relative to the tensor's scale)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = max(b.abs().max().item(), 1e-30)
    return (a - b).abs().max().item() / denom
