"""CPU port of the reference's latent-ODE modules (checker + CPU baseline; TEST INFRASTRUCTURE).

Restates, in plain PyTorch and on top of ``oracle.torchdiffeq_oracle``:

* ``Dynamics``      -- ``models/blackbox_ode.py:64-109``: ``x=[t,z]``, ``h=relu(W1 x+b1)``,
                       ``f = sigmoid(Wg h+bg) - sigmoid(Wd h+bd) * state``.  The reference evaluates
                       the hidden layer twice (``prod``/``degr`` share it, ``:84-95``); numerically
                       that is the same value, so it is evaluated once here.
* ``OdeFunc``       -- ``:50-61`` (``constants`` is a plain tensor, not a Parameter).
* ``OdeModel``      -- ``:6-47`` (``x0 = sigmoid(W2 relu(W1 z + b1) + b2)``; solver call; permute).
* ``Decoder`` / ``GaussianDecoder`` heads -- ``models/decoders.py:42-54,84-91``.

Attribute names (including the reference's mis-spellings ``dyanamics_growth`` /
``dyanmics_degradation``) are kept so a reference ``state_dict`` loads.  Pinned against the real
reference classes in ``tests/test_oracle_vs_reference.py`` (runs where ``/root/reference`` exists)
and against ``tests/golden/blackbox_*.npz`` (generated from the real classes) everywhere.
"""
from __future__ import annotations

import torch
from torch import nn

from . import torchdiffeq_oracle as tde


class Dynamics(nn.Module):
    def __init__(self, n_inputs, hidden_dim, n_outputs):
        super().__init__()
        self.n_inputs, self.n_outputs = n_inputs, n_outputs
        self.dynamics_hidden = nn.Linear(n_inputs + 1, hidden_dim)  # column 0 multiplies t
        self.dyanamics_growth = nn.Linear(hidden_dim, n_outputs)
        self.dyanmics_degradation = nn.Linear(hidden_dim, n_outputs)
        nn.init.xavier_uniform_(self.dynamics_hidden.weight)
        nn.init.xavier_uniform_(self.dyanamics_growth.weight, gain=0.5)
        nn.init.xavier_uniform_(self.dyanmics_degradation.weight, gain=1)
        # aliases only so that a reference state_dict (keys ``prod.0.weight`` ...) loads strictly
        self.prod = nn.Sequential(self.dynamics_hidden, nn.ReLU(), self.dyanamics_growth, nn.Sigmoid())
        self.degr = nn.Sequential(self.dynamics_hidden, nn.ReLU(), self.dyanmics_degradation, nn.Sigmoid())

    def forward(self, t, state, constants, n_batch):
        x = torch.cat([t.repeat([n_batch, 1]), constants], dim=1)
        h = torch.relu(self.dynamics_hidden(x))
        return torch.sigmoid(self.dyanamics_growth(h)) - torch.sigmoid(self.dyanmics_degradation(h)) * state


class OdeFunc(nn.Module):
    def __init__(self, z, dynamics):
        super().__init__()
        self.dynamics = dynamics
        self.n_batch = z.shape[0]
        self.constants = z

    def forward(self, t, state):
        return self.dynamics(t, state, self.constants, self.n_batch)


class OdeModel(nn.Module):
    def __init__(self, times, ode_state_dim, latent_dim, ode_hidden_dim, adjoint_solver, solver):
        super().__init__()
        self.times, self.adjoint_solver, self.solver = times, adjoint_solver, solver
        self.latent_to_ode_net = nn.Sequential(
            nn.Linear(latent_dim, ode_hidden_dim), nn.ReLU(),
            nn.Linear(ode_hidden_dim, ode_state_dim), nn.Sigmoid())
        self.dynamics = Dynamics(latent_dim, ode_hidden_dim, ode_state_dim)

    def solve_ODE(self, z, rtol=1e-7, atol=1e-9):
        x0 = self.latent_to_ode_net(z)
        func = OdeFunc(z, self.dynamics)
        solve = tde.odeint_adjoint if self.adjoint_solver else tde.odeint
        sol = solve(func, x0, self.times, method=self.solver, rtol=rtol, atol=atol)
        return sol.permute(1, 0, 2)


class QuantileHeads(nn.Module):
    """The three bias-free ``Linear(S->O)`` heads + ``softplus(constant_std)`` of ``Decoder``."""

    def __init__(self, ode_model, obs_dim, n_times, constant_std=1e-2):
        super().__init__()
        s = ode_model.dynamics.n_outputs
        self.ode_model = ode_model
        self.output_q50 = nn.Sequential(nn.Linear(s, obs_dim, bias=False))
        self.output_q75 = nn.Sequential(nn.Linear(s, obs_dim, bias=False))
        self.output_q25 = nn.Sequential(nn.Linear(s, obs_dim, bias=False))
        self.constant_std = nn.Parameter(torch.ones(obs_dim, n_times) * constant_std)

    def forward(self, z):
        sol = self.ode_model.solve_ODE(z)
        mu = [head(sol).permute(0, 2, 1) for head in (self.output_q75, self.output_q50, self.output_q25)]
        std = torch.ones_like(mu[0]) * nn.functional.softplus(self.constant_std)
        return sol, mu[0], mu[1], mu[2], std


class Decoder(QuantileHeads):
    """``models/decoders.py::Decoder`` constructor signature over the CPU port (checker / CPU baseline of the
    training step: ``structured_latent_odes_b200.training_cvs.MechanisticModel(decoder_cls=oracle Decoder)``)."""

    def __init__(self, config, times, latent_dim, device="cpu"):
        ode = OdeModel(times, config.ode_state_dim, latent_dim, config.ode_hidden_dim, config.adjoint_solver,
                       config.solver)
        super().__init__(ode, config.obs_dim, len(times), config.constant_std)
