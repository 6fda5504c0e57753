"""Host-side mirror of the reference's decoders (``models/decoders.py``) on top of the fused kernels.

Same class names, constructor signature ``(config, times, latent_dim, device)``, attribute names
(``ode_model``, ``output_q50`` / ``output_q75`` / ``output_q25`` / ``output_mean``, ``constant_std``), ``state_dict``
keys and return tuples as the reference:

  Decoder.forward(z)          -> solution_xt (B,T,S), mu_75, mu_50, mu_25, std   each (B, obs_dim, T)   :42-54
  GaussianDecoder.forward(z)  -> solution_xt, mean, std                                                :84-91

The latent solve runs in the fused solver kernels (``OdeModel.solve_ODE``), stored (B,T,S)-contiguous, and all
heads are produced by ONE pass over the trajectories in their final (B,O,T) layout (``slode_heads_fwd`` / ``_bwd``)
instead of one tiny-K matmul + permute per head.  ``predict`` (no gradients: reconstruction, ``multiple_samples``)
goes one step further: the heads run inside the solver kernel's epilogue (``slode_latent_fixed_heads_fwd``) and the
trajectories are written to HBM only on request.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _cabi
from .blackbox_ode import OdeModel

__all__ = ["Decoder", "GaussianDecoder", "VarianceGaussianDecoder", "decoder_heads", "multiple_samples"]


class _Heads(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sol, W):
        """sol: (B,T,S) view (any strides with unit stride in S); W: (NQ,O,S) -> mu (NQ,B,O,T)."""
        B, T, S = sol.shape
        NQ, O, _ = W.shape
        if not sol.is_cuda:
            raise RuntimeError("structured_latent_odes_b200 runs on CUDA tensors only (no CPU fallback)")
        if sol.dtype != torch.float32 or (S > 1 and sol.stride(2) != 1):
            sol = sol.to(torch.float32).contiguous()
        Wc = W.detach().to(torch.float32).contiguous()
        mu = torch.empty((NQ, B, O, T), device=sol.device, dtype=torch.float32)
        with torch.cuda.device(sol.device):
            rc = _cabi.lib().slode_heads_fwd(B, T, S, O, NQ, sol.data_ptr(), sol.stride(1), sol.stride(0), Wc.data_ptr(),
                                             mu.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _cabi.check(rc, "slode_heads_fwd")
        ctx.save_for_backward(sol, Wc)
        # one output per head (views of one buffer): autograd then hands every head's gradient over on its own, and a
        # head that does not enter the loss costs nothing -- a single stacked output made it zero-fill and copy
        # (NQ,B,O,T) once per head
        return tuple(mu.unbind(0))

    @staticmethod
    def backward(ctx, *grads):
        sol, Wc = ctx.saved_tensors
        B, T, S = sol.shape
        NQ, O, _ = Wc.shape
        gs = [None if g is None else g.to(torch.float32).contiguous() for g in grads] + [None] * (3 - NQ)
        grad_sol = torch.empty((B, T, S), device=sol.device, dtype=torch.float32)
        grad_W = torch.zeros_like(Wc)
        with torch.cuda.device(sol.device):
            rc = _cabi.lib().slode_heads_bwd_split(
                B, T, S, O, NQ, sol.data_ptr(), sol.stride(1), sol.stride(0), Wc.data_ptr(),
                *[g.data_ptr() if g is not None else None for g in gs[:3]], grad_sol.data_ptr(), grad_sol.stride(1),
                grad_sol.stride(0), grad_W.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _cabi.check(rc, "slode_heads_bwd_split")
        return grad_sol, grad_W


def decoder_heads(solution, weights):
    """``[Linear_q(solution).permute(0, 2, 1) for q]`` for bias-free ``Linear(S -> O)`` weights, in one kernel."""
    W = torch.stack(list(weights), dim=0)
    return list(_Heads.apply(solution, W))


def _make_ode_model(config, times, latent_dim, device):
    m = OdeModel()
    m.init_with_params(times=times, ode_state_dim=config.ode_state_dim, latent_dim=latent_dim,
                       ode_hidden_dim=config.ode_hidden_dim, adjoint_solver=config.adjoint_solver, solver=config.solver,
                       device=device, layout="bts")
    return m


class Decoder(nn.Module):
    def __init__(self, config, times, latent_dim, device):
        super().__init__()
        self.times = times
        self.ode_state_dim = config.ode_state_dim
        self.obs_dim = config.obs_dim
        self.latent_dim = latent_dim
        self.ode_hidden_dim = config.ode_hidden_dim
        self.ode_model = _make_ode_model(config, times, latent_dim, device)
        self.output_q50 = nn.Sequential(nn.Linear(self.ode_state_dim, self.obs_dim, bias=False))
        self.output_q75 = nn.Sequential(nn.Linear(self.ode_state_dim, self.obs_dim, bias=False))
        self.output_q25 = nn.Sequential(nn.Linear(self.ode_state_dim, self.obs_dim, bias=False))
        self.constant_std = nn.Parameter(torch.ones(self.obs_dim, len(self.times)) * config.constant_std,
                                         requires_grad=True)

    def forward(self, z):
        solution = self.ode_model.solve_ODE(z=z)
        mu_50, mu_75, mu_25 = decoder_heads(solution, (self.output_q50[0].weight, self.output_q75[0].weight,
                                                       self.output_q25[0].weight))
        std = torch.ones_like(mu_50) * nn.functional.softplus(self.constant_std)
        return solution, mu_75, mu_50, mu_25, std

    @torch.no_grad()
    def predict(self, z, want_solution=False):
        """``forward`` for callers that do not differentiate (``recon`` / posterior sampling): the heads are fused
        into the solver kernel's epilogue, the same tuple comes back with ``solution`` = ``None`` unless asked for --
        the (B,T,S) trajectories are then never written to HBM.  Bit-equal to ``forward``."""
        mu, solution = self.ode_model.solve_ODE_heads(
            z, (self.output_q50[0].weight, self.output_q75[0].weight, self.output_q25[0].weight), want_solution)
        mu_50, mu_75, mu_25 = mu.unbind(0)
        # the same values as forward's ``ones_like(mu) * softplus(constant_std)``, as a broadcast view: materialising
        # (B,O,T) copies of an (O,T) table cost a third of the whole call at 2^20 trajectories
        std = nn.functional.softplus(self.constant_std).expand(mu_50.shape)
        return solution, mu_75, mu_50, mu_25, std


class GaussianDecoder(nn.Module):
    def __init__(self, config, times, latent_dim, device):
        super().__init__()
        self.times = times
        self.ode_state_dim = config.ode_state_dim
        self.obs_dim = config.obs_dim
        self.latent_dim = latent_dim
        self.ode_hidden_dim = config.ode_hidden_dim
        self.ode_model = _make_ode_model(config, times, latent_dim, device)
        self.output_mean = nn.Sequential(nn.Linear(self.ode_state_dim, self.obs_dim, bias=False))
        self.constant_std = nn.Parameter(torch.ones(self.obs_dim, len(self.times)) * config.constant_std,
                                         requires_grad=True)

    def forward(self, z):
        solution = self.ode_model.solve_ODE(z=z)
        (mean,) = decoder_heads(solution, (self.output_mean[0].weight,))
        std = torch.ones_like(mean) * nn.functional.softplus(self.constant_std)
        return solution, mean, std

    @torch.no_grad()
    def predict(self, z, want_solution=False):
        """``forward`` under ``no_grad`` with the mean head fused into the solver kernel (see ``Decoder.predict``)."""
        mu, solution = self.ode_model.solve_ODE_heads(z, (self.output_mean[0].weight,), want_solution)
        mean = mu[0]
        std = nn.functional.softplus(self.constant_std).expand(mean.shape)
        return solution, mean, std


class VarianceGaussianDecoder(nn.Module):
    """``models/decoders.py:94-141``: a second latent ODE (``std_ode_model``) whose head gives the observation scale
    (no model of the reference uses it; mirrored for the module's completeness).  Two fused solves, two head passes."""

    def __init__(self, config, times, latent_dim, device):
        super().__init__()
        self.times = times
        self.ode_state_dim = config.ode_state_dim
        self.obs_dim = config.obs_dim
        self.latent_dim = latent_dim
        self.ode_hidden_dim = config.ode_hidden_dim
        self.ode_model = _make_ode_model(config, times, latent_dim, device)
        self.std_ode_model = _make_ode_model(config, times, latent_dim, device)
        self.output_mean = nn.Sequential(nn.Linear(self.ode_state_dim, self.obs_dim, bias=False))
        self.output_std = nn.Sequential(nn.Linear(self.ode_state_dim, self.obs_dim, bias=False))
        self.constant_std = nn.Parameter(torch.ones(self.obs_dim, len(self.times)) * config.constant_std,
                                         requires_grad=True)

    def forward(self, z):
        solution = self.ode_model.solve_ODE(z=z)
        (mean,) = decoder_heads(solution, (self.output_mean[0].weight,))
        solution_std = self.std_ode_model.solve_ODE(z=z)
        (std,) = decoder_heads(solution_std, (self.output_std[0].weight,))
        return solution, mean, std


@torch.no_grad()
def multiple_samples(decoder, z_loc, z_scale, num_samples, generator=None):
    """Posterior / prior predictive sampling in ONE launch.

    The reference draws ``num_samples`` (200) latent samples one at a time and re-solves the whole set each time
    (``training_challenge.py:174-195``, ``training_proc.py:205-223``: ``recon`` in a Python loop, results
    concatenated on a new last axis).  Trajectories are independent, so the samples are folded into the batch axis
    instead: ``z ~ Normal(z_loc, z_scale)`` for all (sample, series) pairs, one solve + one heads pass.

    Returns ``{"mu_25", "mu_50", "mu_75"}`` of shape ``(B, obs_dim, T, num_samples)`` (the reference's layout) and
    ``"z"`` of shape ``(num_samples, B, latent_dim)``.
    """
    B, L = z_loc.shape
    eps = torch.randn((num_samples, B, L), device=z_loc.device, dtype=z_loc.dtype, generator=generator)
    z = z_loc[None] + z_scale[None] * eps
    # our decoders: heads fused into the solve, no (B,T,S) round trip through HBM; any other module: its forward
    run = decoder.predict if hasattr(decoder, "predict") else decoder
    out = run(z.reshape(num_samples * B, L))
    names = ("mu_75", "mu_50", "mu_25") if len(out) == 5 else ("mu_50",)
    res = {"z": z}
    for name, mu in zip(names, out[1:1 + len(names)]):
        res[name] = mu.reshape(num_samples, B, *mu.shape[1:]).permute(1, 2, 3, 0)
    return res
