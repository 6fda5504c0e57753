"""The CVS mechanistic right-hand side as an ``odeint``-style dynamics module, solved by fused sm_100a kernels.

Reference: ``dx_dt`` (``data/cvs/cvs_data.py:52-91``), constants of ``get_random_params`` (``:28-49``),
observation map ``states_trajectory_to_sample`` (``:94-103``), generator loop ``create_cvs_data``
(``:111-134``: one scipy LSODA solve per sample, float64, ``x(0) = ones(4)``, ``t = 0..seq_len-1``).

``CvsMechanistic(i_ext, r_tpr_mod)`` has the module API of the latent ODE (``forward(t, state)``); handing it to
``structured_latent_odes_b200.odeint`` / ``odeint_adjoint`` runs the whole fixed-grid solve (and its reverse
sweep) in one kernel each.  float32 tensors take the odeint drop-in path, float64 tensors the generator path.
``options={"step_size": h}`` splits every output interval into ``dt/h`` solver steps (uniform grids whose
spacing is a multiple of ``h``; at most 16) -- the accuracy knob of :func:`generate_cvs_latents`.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _cabi

__all__ = ["CvsMechanistic", "CVS_CONSTANTS", "THETA_ORDER", "observe", "generate_cvs_latents"]

THETA_ORDER = ("f_hr_max", "f_hr_min", "r_tpr_max", "r_tpr_min", "sv_mod", "ca", "cv", "k_width", "p_aset", "tau")
CVS_CONSTANTS = dict(f_hr_max=3.0, f_hr_min=2.0 / 3.0, r_tpr_max=2.134, r_tpr_min=0.5335, sv_mod=0.0001,
                     ca=4.0, cv=111.0, k_width=0.1838, p_aset=70.0, tau=20.0)


class CvsMechanistic(nn.Module):
    """``forward(t, state)`` with per-trajectory treatments ``i_ext (B,)``, ``r_tpr_mod (B,)``.

    ``theta`` (the ten shared constants, order :data:`THETA_ORDER`) is an ``nn.Parameter`` so a mechanistic fit
    can learn it; it does not require grad unless ``learn_constants=True``.
    """

    def __init__(self, i_ext, r_tpr_mod, constants=None, learn_constants=False):
        super().__init__()
        if i_ext.ndim != 1 or i_ext.shape != r_tpr_mod.shape:
            raise ValueError("i_ext and r_tpr_mod must be (B,) tensors of equal length")
        self.i_ext = i_ext
        self.r_tpr_mod = r_tpr_mod
        c = dict(CVS_CONSTANTS)
        c.update(constants or {})
        self.theta = nn.Parameter(torch.tensor([c[k] for k in THETA_ORDER], dtype=i_ext.dtype, device=i_ext.device),
                                  requires_grad=learn_constants)

    def forward(self, t, state):
        """Eager single evaluation (module API parity; the fused solve does not call this)."""
        th = dict(zip(THETA_ORDER, self.theta.unbind()))
        p_a, p_v, s, sv = 100.0 * state[..., 0], 10.0 * state[..., 1], state[..., 2], 100.0 * state[..., 3]
        f_hr = s * (th["f_hr_max"] - th["f_hr_min"]) + th["f_hr_min"]
        r_tpr = s * (th["r_tpr_max"] - th["r_tpr_min"]) + th["r_tpr_min"] - self.r_tpr_mod
        dva = -(p_a - p_v) / r_tpr + sv * f_hr
        dpa = dva / (th["ca"] * 100.0)
        dpv = (-dva + self.i_ext) / (th["cv"] * 10.0)
        ds = (1.0 - 1.0 / (1.0 + torch.exp(-th["k_width"] * (p_a - th["p_aset"]))) - s) / th["tau"]
        dsv = self.i_ext * th["sv_mod"] * torch.ones_like(s)
        return torch.stack([dpa, dpv, ds, dsv], dim=-1)


def observe(states, constants=None):
    """(…,4) latent states -> (…,3) observables (Pa/100, Pv/10, f_HR) (``cvs_data.py:94-103``)."""
    c = dict(CVS_CONSTANTS)
    c.update(constants or {})
    f_hr = states[..., 2] * (c["f_hr_max"] - c["f_hr_min"]) + c["f_hr_min"]
    return torch.stack([states[..., 0], states[..., 1], f_hr], dim=-1)


def _ptr(x):
    return x.data_ptr()


class _CvsFixedSolve(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y0, i_ext, r_tpr_mod, theta, t, method_id, mode, substeps, layout):
        B = y0.shape[0]
        T = t.numel()
        dt = _cabi.F32 if y0.dtype == torch.float32 else _cabi.F64
        y0c, ie, rm, th = (x.detach().contiguous() for x in (y0, i_ext, r_tpr_mod, theta))
        if layout == "bts":
            sol = torch.empty((B, T, 4), device=y0.device, dtype=y0.dtype).permute(1, 0, 2)
        else:
            sol = torch.empty((T, B, 4), device=y0.device, dtype=y0.dtype)
        with torch.cuda.device(y0.device):
            rc = _cabi.lib().slode_cvs_fixed_fwd(method_id, dt, B, T, substeps, _ptr(t), _ptr(y0c), _ptr(ie), _ptr(rm),
                                                 _ptr(th), _ptr(sol), sol.stride(0), sol.stride(1),
                                                 torch.cuda.current_stream().cuda_stream)
        _cabi.check(rc, "slode_cvs_fixed_fwd")
        ctx.save_for_backward(ie, rm, th, t, sol)
        ctx.cfg = (method_id, mode, substeps, dt)
        return sol

    @staticmethod
    def backward(ctx, grad_sol):
        ie, rm, th, t, sol = ctx.saved_tensors
        method_id, mode, substeps, dt = ctx.cfg
        T, B, _ = sol.shape
        if grad_sol.dtype != sol.dtype or grad_sol.stride(2) != 1 or (T > 1 and grad_sol.stride(0) <= 0) \
                or (B > 1 and grad_sol.stride(1) <= 0):
            grad_sol = grad_sol.to(sol.dtype).contiguous()
        gy0 = torch.empty((B, 4), device=sol.device, dtype=sol.dtype)
        gie = torch.empty(B, device=sol.device, dtype=sol.dtype)
        grm = torch.empty(B, device=sol.device, dtype=sol.dtype)
        gth = torch.zeros(10, device=sol.device, dtype=sol.dtype)
        with torch.cuda.device(sol.device):
            rc = _cabi.lib().slode_cvs_fixed_bwd(
                method_id, mode, dt, B, T, substeps, _ptr(t), _ptr(ie), _ptr(rm), _ptr(th), _ptr(sol), sol.stride(0),
                sol.stride(1), _ptr(grad_sol), grad_sol.stride(0), grad_sol.stride(1), _ptr(gy0), _ptr(gie), _ptr(grm),
                _ptr(gth), torch.cuda.current_stream().cuda_stream)
        _cabi.check(rc, "slode_cvs_fixed_bwd")
        return gy0, gie, grm, gth, None, None, None, None, None


def _substeps_from_options(t, options):
    if not options:
        return 1
    extra = set(options) - {"step_size"}
    if extra:
        raise NotImplementedError(f"options {sorted(extra)} are not supported by the fused CVS solve")
    h = float(options["step_size"])
    if t.numel() < 2:
        return 1
    d = (t[1:] - t[:-1]).double()
    n = d / h
    nr = torch.round(n)
    if not bool(((n - nr).abs() < 1e-6 * nr.clamp(min=1)).all()) or not bool((nr == nr[0]).all()) or nr[0] < 1:
        raise NotImplementedError("step_size must divide a uniform output spacing (torchdiffeq's interpolating grid "
                                  "for other cases is not implemented)")
    n = int(nr[0])
    if n > 16:
        raise NotImplementedError("at most 16 solver steps per output interval")
    return n


def solve_cvs(func, y0, t, method, mode, layout, options):
    """Called by ``torchdiffeq_api._solve`` for ``CvsMechanistic`` dynamics."""
    if y0.shape[1] != 4:
        raise ValueError(f"the CVS state has 4 components, got y0 {tuple(y0.shape)}")
    if func.i_ext.shape[0] != y0.shape[0]:
        raise ValueError(f"i_ext batch {func.i_ext.shape[0]} != y0 batch {y0.shape[0]}")
    for name in ("i_ext", "r_tpr_mod", "theta"):
        x = getattr(func, name)
        if x.device != y0.device or x.dtype != y0.dtype:
            raise RuntimeError(f"CvsMechanistic.{name} must be {y0.dtype} on {y0.device}")
    substeps = _substeps_from_options(t, options)
    ie, rm = func.i_ext, func.r_tpr_mod
    if mode == _cabi.BWD_TDE_ADJOINT:
        # odeint_adjoint: only func.parameters() are adjoint parameters; plain-tensor attributes get no gradient
        ie = ie if isinstance(ie, nn.Parameter) else ie.detach()
        rm = rm if isinstance(rm, nn.Parameter) else rm.detach()
    return _CvsFixedSolve.apply(y0, ie, rm, func.theta, t, _cabi.METHODS[method], mode, substeps, layout)


def generate_cvs_latents(i_ext, r_tpr_mod, seq_len=86, delta_t=1.0, substeps=8, method="rk4"):
    """Batched replacement of the reference's LSODA loop (``create_cvs_data``): float64 latent trajectories
    ``(N, seq_len, 4)`` from ``x(0) = ones(4)`` at ``t = 0, delta_t, ...``, one kernel for the whole data set."""
    from . import torchdiffeq_api as api
    if not i_ext.is_cuda:
        raise RuntimeError("generate_cvs_latents runs on CUDA tensors only (no CPU fallback)")
    ie = i_ext.to(torch.float64)
    rm = r_tpr_mod.to(torch.float64)
    f = CvsMechanistic(ie, rm)
    t = torch.arange(seq_len, device=ie.device, dtype=torch.float64) * delta_t
    y0 = torch.ones(ie.shape[0], 4, device=ie.device, dtype=torch.float64)
    with torch.no_grad():
        sol = api.odeint(f, y0, t, method=method, options={"step_size": delta_t / substeps}, layout="bts")
    return sol.permute(1, 0, 2)
