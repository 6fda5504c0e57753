"""Host-side mirror of the reference's latent-ODE modules (``models/blackbox_ode.py``).

Same class names, constructor protocol, attribute names (including the reference's mis-spelt
``dyanamics_growth`` / ``dyanmics_degradation``) and ``state_dict`` keys, so that a reference
checkpoint loads and ``Decoder`` code written against the reference keeps working.  The only
behavioural difference is where ``solve_ODE`` sends the solve: to the fused sm_100a kernels through
``structured_latent_odes_b200.torchdiffeq_api`` instead of ``torchdiffeq``.

  OdeModel.init_with_params / gen_dynamics / initialize_state / solve_ODE   blackbox_ode.py:6-47
  OdeFunc.forward(t, state)                                                  blackbox_ode.py:50-61
  Dynamics.forward(t, state, constants, n_batch)                             blackbox_ode.py:64-109
"""
from __future__ import annotations

import torch
from torch import nn

from . import torchdiffeq_api as _api


class Dynamics(nn.Module):
    """``dx/dt = sigmoid(Wg h + bg) - sigmoid(Wd h + bd) * x`` with ``h = act(W1 [t, z] + b1)``."""

    def __init__(self, n_inputs, hidden_dim, n_outputs, hidden_activation=nn.Tanh):
        super().__init__()
        self.n_inputs = n_inputs
        self.n_outputs = n_outputs
        self.dynamics_hidden = nn.Linear(n_inputs + 1, hidden_dim)  # input column 0 is time
        self.dyanamics_growth = nn.Linear(hidden_dim, n_outputs)
        self.dyanmics_degradation = nn.Linear(hidden_dim, n_outputs)
        # reference initialisation (blackbox_ode.py:75,79,82)
        nn.init.xavier_uniform_(self.dynamics_hidden.weight)
        nn.init.xavier_uniform_(self.dyanamics_growth.weight, gain=0.5)
        nn.init.xavier_uniform_(self.dyanmics_degradation.weight, gain=1)
        act = hidden_activation()
        # the reference registers these two pipelines (sharing the hidden layer); kept for key parity
        self.prod = nn.Sequential(self.dynamics_hidden, act, self.dyanamics_growth, nn.Sigmoid())
        self.degr = nn.Sequential(self.dynamics_hidden, act, self.dyanmics_degradation, nn.Sigmoid())

    def forward(self, t, state, constants, n_batch):
        """Single RHS evaluation (module API; the solve itself never calls this -- it runs fused)."""
        t_col = t.to(state.dtype).reshape(1, 1).expand(n_batch, 1)
        x = t_col if constants is None else torch.cat([t_col, constants], dim=1)
        h = self.prod[1](self.dynamics_hidden(x))
        return torch.sigmoid(self.dyanamics_growth(h)) - torch.sigmoid(self.dyanmics_degradation(h)) * state


class OdeFunc(nn.Module):
    def __init__(self, z, dynamics):
        super().__init__()
        self.dynamics = dynamics
        self.n_batch = z.shape[0]
        self.constants = z  # plain tensor on purpose: not an adjoint parameter

    def forward(self, t, state):
        return self.dynamics.forward(t=t, state=state, constants=self.constants, n_batch=self.n_batch)


class OdeModel(nn.Module):
    """Two-phase construction like the reference: ``m = OdeModel(); m.init_with_params(...)``."""

    def __init__(self, *args, **kwargs):
        # the reference defers nn.Module.__init__ to init_with_params; doing it here as well is harmless
        super().__init__()
        if args or kwargs:
            self.init_with_params(*args, **kwargs)

    def init_with_params(self, times, ode_state_dim, latent_dim, ode_hidden_dim, adjoint_solver, solver, device,
                         layout="tbs"):
        super().__init__()
        self.times = times
        self.ode_state_dim = ode_state_dim
        self.latent_dim = latent_dim
        self.ode_hidden_dim = ode_hidden_dim
        self.device = device
        self.adjoint_solver = adjoint_solver
        self.solver = solver
        self.layout = layout
        self.latent_to_ode_net = nn.Sequential(
            nn.Linear(latent_dim, ode_hidden_dim), nn.ReLU(),
            nn.Linear(ode_hidden_dim, ode_state_dim), nn.Sigmoid())
        self.dynamics = Dynamics(n_inputs=latent_dim, hidden_dim=ode_hidden_dim, n_outputs=ode_state_dim,
                                 hidden_activation=nn.ReLU)

    def gen_dynamics(self, z):
        return OdeFunc(z=z, dynamics=self.dynamics)

    def initialize_state(self, z):
        return self.latent_to_ode_net(z)

    def solve_ODE(self, z):
        """(B, L) latent -> (B, T, S) latent trajectories (a permuted view, as in the reference)."""
        if self.solver in _api.FIXED_METHODS and self._x0_net_is_reference_shaped():
            # everything in two kernels: x0 = latent_to_ode_net(z), c = z W1[:,1:]^T + b1, the solve, and all
            # of their gradients in the reverse sweep
            sol = _api.solve_latent(z, self.dynamics, self.latent_to_ode_net, self.times, self.solver,
                                    self.adjoint_solver, layout=self.layout)
            return sol.permute(1, 0, 2)
        init_state = self.initialize_state(z).to(self.device)
        func = self.gen_dynamics(z=z)
        solve = _api.odeint_adjoint if self.adjoint_solver else _api.odeint
        sol = solve(func=func, y0=init_state, t=self.times, method=self.solver, layout=self.layout)
        return sol.permute(1, 0, 2)

    def solve_ODE_heads(self, z, head_weights, want_solution=False, contiguous=False):
        """``solve_ODE`` with the decoder heads fused into the solver's epilogue, under ``no_grad``:
        returns ``(mu (NQ,B,O,T), solution (B,T,S) or None)``; without ``want_solution`` the trajectories never reach
        HBM.  Falls back to NOTHING: solvers / state nets the fused kernel does not cover raise."""
        if not (self.solver in _api.FIXED_METHODS and self._x0_net_is_reference_shaped()):
            raise NotImplementedError("fused decoder heads need a fixed-grid solver and the reference's latent_to_ode_net")
        mu, sol = _api.solve_latent_heads(z, self.dynamics, self.latent_to_ode_net, self.times, self.solver,
                                          head_weights, want_solution=want_solution, layout=self.layout,
                                          contiguous=contiguous)
        return mu, (None if sol is None else sol.permute(1, 0, 2))

    def _x0_net_is_reference_shaped(self):
        n = self.latent_to_ode_net
        return (isinstance(n, nn.Sequential) and len(n) == 4 and isinstance(n[0], nn.Linear) and isinstance(n[1], nn.ReLU)
                and isinstance(n[2], nn.Linear) and isinstance(n[3], nn.Sigmoid) and n[0].bias is not None
                and n[2].bias is not None)
