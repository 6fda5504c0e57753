"""Pure-PyTorch restatement of the part of ``torchdiffeq`` the reference calls.

TEST INFRASTRUCTURE ONLY (checker / CPU baseline) -- never imported by the product.

The reference's hot path is ``torchdiffeq.odeint_adjoint(func=, y0=, t=, method=)`` /
``torchdiffeq.odeint(...)`` at ``models/blackbox_ode.py:40-45``.  ``torchdiffeq`` is a
third-party PyPI dependency that is NOT vendored under ``/root/reference``, is not
listed in ``requirements.txt:1-6`` (so its version is unpinned; 0.2.3-0.2.5 were current
when the paper was published) and is not installable here (no network).

    PARITY UNPINNED for the solver loop: the reference ships no tests, golden vectors
    or stored trajectories for this boundary, and the real torchdiffeq cannot be run.

This file therefore restates the published torchdiffeq 0.2.x algorithm:

* fixed-grid solvers ``euler`` / ``midpoint`` / ``rk4`` (torchdiffeq's rk4 is the
  3/8-rule ``rk4_alt_step_func``), grid == ``t`` when no ``step_size`` option is given,
  output at a grid point is exactly ``y1`` (``_linear_interp`` short-circuits ``t == t1``);
* ``dopri5``: Dormand-Prince 5(4) with FSAL, Hairer initial step, batch-global RMS error
  norm, ``safety=0.9, ifactor=10, dfactor=0.2``, float64 time carried by the controller
  and cast to ``y.dtype`` at each RHS call, 4th-order dense output fitted through the
  ``DPS_C_MID`` midpoint;
* ``odeint_adjoint``: forward under ``no_grad``; backward integrates the augmented system
  ``[vjp_t, y, adj_y, adj_params]`` from ``t[i]`` to ``t[i-1]`` with the same method (for dopri5: a
  fresh adaptive solve per output interval, error norm = torchdiffeq's mixed norm over the tuple, i.e.
  the largest per-tensor RMS of err/tol among y, adj_y and each parameter's adjoint; ``vjp_t`` is
  identically zero because the reference's time grid does not require grad), resets
  ``y`` to the stored forward value and adds ``grad_y[i-1]``; ``adjoint_params`` defaults to
  ``tuple(func.parameters())`` (so plain-tensor attributes such as ``OdeFunc.constants``
  get no gradient -- SURVEY.md F5).

It is self-validated in ``tests/test_oracle_solvers.py`` (order of convergence,
closed-form linear ODE, gradcheck in float64, scipy's Dormand-Prince step).

The module can be registered as ``sys.modules["torchdiffeq"]`` (see ``oracle/shims.py``)
so that the reference's ``models/blackbox_ode.py`` imports unchanged.
"""
from __future__ import annotations

import torch

__all__ = ["odeint", "odeint_adjoint", "FIXED_METHODS", "ADAPTIVE_METHODS", "SolverStats"]

FIXED_METHODS = ("euler", "midpoint", "rk4")
ADAPTIVE_METHODS = ("dopri5",)

_ONE_THIRD = 1.0 / 3.0
_TWO_THIRDS = 2.0 / 3.0


class SolverStats:
    """Accept/reject bookkeeping of the last adaptive solve (test aid, not in torchdiffeq)."""

    def __init__(self):
        self.n_accept = 0
        self.n_reject = 0
        self.n_rhs = 0
        self.accepted = []  # list of bools per attempted step
        self.dts = []  # dt (float64) per attempted step

    def reset(self):
        self.__init__()


last_stats = SolverStats()
# backward pass of the last odeint_adjoint call with an adaptive method: one (interval index i, accepted[], dts[])
# entry per output interval, in the order they were solved (i = T-1 .. 1); test aid, not in torchdiffeq
last_adjoint_intervals = []


# ----------------------------------------------------------------------------------------
# fixed grid
# ----------------------------------------------------------------------------------------

def _euler_step(func, t0, dt, t1, y0):
    f0 = func(t0, y0)
    return dt * f0


def _midpoint_step(func, t0, dt, t1, y0):
    half_dt = 0.5 * dt
    f0 = func(t0, y0)
    y_mid = y0 + f0 * half_dt
    return dt * func(t0 + half_dt, y_mid)


def _rk4_38_step(func, t0, dt, t1, y0):
    k1 = func(t0, y0)
    k2 = func(t0 + dt * _ONE_THIRD, y0 + dt * k1 * _ONE_THIRD)
    k3 = func(t0 + dt * _TWO_THIRDS, y0 + dt * (k2 - k1 * _ONE_THIRD))
    k4 = func(t1, y0 + dt * (k1 - k2 + k3))
    return (k1 + 3 * (k2 + k3) + k4) * dt * 0.125


_FIXED_STEP = {"euler": _euler_step, "midpoint": _midpoint_step, "rk4": _rk4_38_step}


def _linear_interp(t0, t1, y0, y1, t):
    if t == t0:
        return y0
    if t == t1:
        return y1
    slope = (t - t0) / (t1 - t0)
    return y0 + slope * (y1 - y0)


def _grid_from_step_size(t, step_size):
    start, end = t[0], t[-1]
    niters = torch.ceil((end - start) / step_size + 1).item()
    grid = torch.arange(0, niters, dtype=t.dtype, device=t.device) * step_size + start
    grid[-1] = t[-1]
    return grid


def _integrate_fixed(func, y0, t, method, step_size=None):
    step = _FIXED_STEP[method]
    grid = t if step_size is None else _grid_from_step_size(t, step_size)
    sol = [y0]
    j = 1
    y = y0
    for t0, t1 in zip(grid[:-1], grid[1:]):
        dt = t1 - t0
        y1 = y + step(func, t0, dt, t1, y)
        while j < len(t) and t1 >= t[j]:
            sol.append(_linear_interp(t0, t1, y, y1, t[j]))
            j += 1
        y = y1
    return torch.stack(sol, dim=0)


# ----------------------------------------------------------------------------------------
# dopri5
# ----------------------------------------------------------------------------------------

_DP_ALPHA = [1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0]
_DP_BETA = [
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
_DP_C_SOL = [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0]
_DP_C_ERROR = [
    35 / 384 - 1951 / 21600,
    0,
    500 / 1113 - 22642 / 50085,
    125 / 192 - 451 / 720,
    -2187 / 6784 - -12231 / 42400,
    11 / 84 - 649 / 6300,
    -1.0 / 60.0,
]
_DP_C_MID = [
    6025192743 / 30085553152 / 2,
    0,
    51252292925 / 65400821598 / 2,
    -2691868925 / 45128329728 / 2,
    187940372067 / 1594534317056 / 2,
    -1776094331 / 19743644256 / 2,
    11237099 / 235043384 / 2,
]


def _rms_norm(x):
    return x.abs().pow(2).mean().sqrt()


def _tab(vals, like):
    return torch.tensor(vals, dtype=like.dtype, device=like.device)


def _dopri5_rk_step(func, y0, f0, t0, dt, t1):
    """One Dormand-Prince attempt.  t0/dt/t1 are float64 scalars, cast to y.dtype here
    (torchdiffeq ``_runge_kutta_step``)."""
    t0 = t0.to(y0.dtype)
    dt = dt.to(y0.dtype)
    t1 = t1.to(y0.dtype)
    k = [f0]
    yi = y0
    for alpha_i, beta_i in zip(_DP_ALPHA, _DP_BETA):
        ti = t1 if alpha_i == 1.0 else t0 + alpha_i * dt
        kk = torch.stack(k, dim=-1)
        yi = y0 + kk.matmul(_tab(beta_i, y0) * dt).view_as(f0)
        k.append(func(ti, yi))
        last_stats.n_rhs += 1
    kk = torch.stack(k, dim=-1)
    y1 = yi  # FSAL: c_sol[:-1] == beta[-1] and c_sol[-1] == 0
    f1 = k[-1]
    y1_error = kk.matmul(dt * _tab(_DP_C_ERROR, y0))
    return y1, f1, y1_error, kk


def _mixed_norm(tensors):
    """torchdiffeq ``_mixed_norm``: the largest RMS norm among the tensors of a tuple state."""
    if len(tensors) == 0:
        return 0.0
    return max(_rms_norm(x) for x in tensors)


def _select_initial_step(func, t0, y0, order, rtol, atol, f0, norm=_rms_norm):
    dtype = y0.dtype
    t_dtype = t0.dtype
    t0 = t0.to(dtype)
    scale = atol + torch.abs(y0) * rtol
    d0 = norm(y0 / scale)
    d1 = norm(f0 / scale)
    if d0 < 1e-5 or d1 < 1e-5:
        h0 = torch.tensor(1e-6, dtype=dtype, device=y0.device)
    else:
        h0 = 0.01 * d0 / d1
    y1 = y0 + h0 * f0
    f1 = func(t0 + h0, y1)
    last_stats.n_rhs += 1
    d2 = norm((f1 - f0) / scale) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = torch.max(torch.tensor(1e-6, dtype=dtype, device=y0.device), h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1.0 / float(order + 1))
    # The step sequence is data for the gradient, not a differentiable quantity: the initial
    # step is detached (the path through it is O(truncation error) and is dropped here).
    return torch.min(100 * h0, h1).to(t_dtype).detach()


@torch.no_grad()  # torchdiffeq decorates this the same way: step sizes carry no gradient
def _optimal_step_size(last_step, error_ratio, safety, ifactor, dfactor, order):
    if error_ratio == 0:
        return last_step * ifactor
    if error_ratio < 1:
        dfactor = torch.ones((), dtype=last_step.dtype, device=last_step.device)
    error_ratio = error_ratio.type_as(last_step)
    exponent = torch.tensor(order, dtype=last_step.dtype, device=last_step.device).reciprocal()
    factor = torch.min(ifactor, torch.max(safety / error_ratio ** exponent, dfactor))
    return last_step * factor


def _interp_fit(y0, y1, y_mid, f0, f1, dt):
    a = 2 * dt * (f1 - f0) - 8 * (y1 + y0) + 16 * y_mid
    b = dt * (5 * f0 - 3 * f1) + 18 * y0 + 14 * y1 - 32 * y_mid
    c = dt * (f1 - 4 * f0) - 11 * y0 - 5 * y1 + 16 * y_mid
    d = dt * f0
    e = y0
    return [e, d, c, b, a]


def _interp_evaluate(coefficients, t0, t1, t):
    dtype = coefficients[0].dtype
    t0 = t0.to(dtype)
    t1 = t1.to(dtype)
    t = t.to(dtype)
    x = ((t - t0) / (t1 - t0)).to(dtype)
    total = coefficients[0] + x * coefficients[1]
    x_power = x
    for coefficient in coefficients[2:]:
        x_power = x_power * x
        total = total + x_power * coefficient
    return total


def _integrate_dopri5(func, y0, t, rtol, atol, options):
    opts = dict(options or {})
    safety = opts.pop("safety", 0.9)
    ifactor = opts.pop("ifactor", 10.0)
    dfactor = opts.pop("dfactor", 0.2)
    max_num_steps = opts.pop("max_num_steps", 2 ** 31 - 1)
    first_step = opts.pop("first_step", None)
    norm = opts.pop("norm", _rms_norm)  # torchdiffeq options["norm"]; a tuple state gets the mixed norm (_odeint_tuple)
    tdtype = torch.promote_types(opts.pop("dtype", torch.float64), y0.dtype)
    if opts:
        raise ValueError(f"unsupported dopri5 options {sorted(opts)}")
    dev = y0.device
    safety = torch.as_tensor(safety, dtype=tdtype, device=dev)
    ifactor = torch.as_tensor(ifactor, dtype=tdtype, device=dev)
    dfactor = torch.as_tensor(dfactor, dtype=tdtype, device=dev)
    order = 5

    last_stats.reset()
    t = t.to(tdtype)
    f0 = func(t[0].to(y0.dtype), y0)
    last_stats.n_rhs += 1
    if first_step is None:
        dt = _select_initial_step(func, t[0], y0, order - 1, rtol, atol, f0, norm)
    else:
        dt = torch.as_tensor(first_step, dtype=tdtype, device=dev)
    st_y, st_f, st_t0, st_t1, st_dt = y0, f0, t[0], t[0], dt
    st_interp = [y0] * 5

    sol = [y0]
    for i in range(1, len(t)):
        next_t = t[i]
        n_steps = 0
        while next_t > st_t1:
            assert n_steps < max_num_steps, "max_num_steps exceeded"
            # one adaptive attempt (torchdiffeq ``_adaptive_step``)
            a_t0, a_dt = st_t1, st_dt
            a_t1 = a_t0 + a_dt
            assert a_t0 + a_dt > a_t0, "underflow in dt {}".format(a_dt.item())
            y1, f1, y1_error, k = _dopri5_rk_step(func, st_y, st_f, a_t0, a_dt, a_t1)
            error_tol = atol + rtol * torch.max(st_y.abs(), y1.abs())
            error_ratio = norm(y1_error / error_tol).abs()
            accept = bool(error_ratio <= 1)
            last_stats.accepted.append(accept)
            last_stats.dts.append(float(a_dt.detach()))
            if accept:
                last_stats.n_accept += 1
                dt_y = a_dt.type_as(st_y)
                y_mid = st_y + k.matmul(dt_y * _tab(_DP_C_MID, st_y)).view_as(st_y)
                st_interp = _interp_fit(st_y, y1, y_mid, k[..., 0], k[..., -1], dt_y)
                st_t0, st_t1 = a_t0, a_t1
                st_y, st_f = y1, f1
            else:
                last_stats.n_reject += 1
            st_dt = _optimal_step_size(a_dt, error_ratio, safety, ifactor, dfactor, order)
            n_steps += 1
        sol.append(_interp_evaluate(st_interp, st_t0, st_t1, next_t))
    return torch.stack(sol, dim=0)


# ----------------------------------------------------------------------------------------
# public API
# ----------------------------------------------------------------------------------------

def _check_inputs(y0, t, method):
    if not torch.is_tensor(y0):
        raise TypeError("oracle odeint supports tensor y0 only")
    if method is None:
        method = "dopri5"
    if method not in FIXED_METHODS + ADAPTIVE_METHODS:
        raise ValueError(f"Invalid method {method!r}")
    if t.ndim != 1:
        raise ValueError("t must be one dimensional")
    d = t[1:] - t[:-1]
    if not (bool((d > 0).all()) or bool((d < 0).all())):
        raise ValueError("t must be strictly increasing or decreasing")
    return method


def odeint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, event_fn=None):
    """``torchdiffeq.odeint`` restatement; returns ``(len(t), *y0.shape)``."""
    if event_fn is not None:
        raise NotImplementedError("event_fn is not used by the reference")
    method = _check_inputs(y0, t, method)
    reversed_t = bool(t[0] > t[-1])
    if reversed_t:  # torchdiffeq ``_ReverseFunc``: integrate -t with the negated RHS
        t = -t
        inner = func
        func = lambda tt, yy: -inner(-tt, yy)  # noqa: E731
    if method in FIXED_METHODS:
        step_size = None if not options else options.get("step_size")
        return _integrate_fixed(func, y0, t, method, step_size)
    return _integrate_dopri5(func, y0, t, rtol, atol, options)


class _OdeintAdjoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, func, method, rtol, atol, options, n_params, y0, t, *adjoint_params):
        ctx.func = func
        ctx.method = method
        ctx.rtol, ctx.atol, ctx.options = rtol, atol, options
        with torch.no_grad():
            y = odeint(func, y0, t, rtol=rtol, atol=atol, method=method, options=options)
        ctx.save_for_backward(t, y, *adjoint_params)
        return y

    @staticmethod
    def backward(ctx, grad_y):
        func = ctx.func
        t, y, *adjoint_params = ctx.saved_tensors
        adjoint_params = tuple(adjoint_params)
        with torch.no_grad():
            aug = [torch.zeros((), dtype=y.dtype, device=y.device), y[-1], grad_y[-1]]
            aug.extend(torch.zeros_like(p) for p in adjoint_params)

            def augmented_dynamics(tt, y_aug):
                yy = y_aug[1]
                adj_y = y_aug[2]
                with torch.enable_grad():
                    # torchdiffeq evaluates func at the DETACHED time unless t itself requires grad (it never does
                    # in the reference: OdeModel.times is a plain tensor), so vjp_t is identically zero
                    tt_ = tt.detach()
                    yy = yy.detach().requires_grad_(True)
                    f = func(tt_, yy)
                    vjp_y, *vjp_params = torch.autograd.grad(
                        f, (yy,) + adjoint_params, -adj_y, allow_unused=True, retain_graph=True)
                vjp_t = torch.zeros_like(tt)
                vjp_y = torch.zeros_like(yy) if vjp_y is None else vjp_y
                vjp_params = [torch.zeros_like(p) if v is None else v
                              for p, v in zip(adjoint_params, vjp_params)]
                return (vjp_t, f, vjp_y, *vjp_params)

            del last_adjoint_intervals[:]
            for i in range(len(t) - 1, 0, -1):
                aug = _odeint_tuple(augmented_dynamics, tuple(aug), t[i - 1:i + 1].flip(0),
                                    ctx.method, ctx.rtol, ctx.atol, ctx.options)
                if ctx.method in ADAPTIVE_METHODS:
                    last_adjoint_intervals.append((i, list(last_stats.accepted), list(last_stats.dts)))
                aug = [a[1] for a in aug]
                aug[1] = y[i - 1]
                aug[2] = aug[2] + grad_y[i - 1]
            adj_y = aug[2]
            adj_params = aug[3:]
        return (None, None, None, None, None, None, adj_y, None, *adj_params)


def _odeint_tuple(func, y0_tuple, t, method, rtol, atol, options):
    """odeint over a tuple state by flattening (torchdiffeq ``_TupleFunc``)."""
    shapes = [y.shape for y in y0_tuple]
    numels = [y.numel() for y in y0_tuple]
    flat0 = torch.cat([y.reshape(-1) for y in y0_tuple])

    def unflat(v):
        out, o = [], 0
        for s, n in zip(shapes, numels):
            out.append(v[..., o:o + n].reshape(tuple(v.shape[:-1]) + tuple(s)))
            o += n
        return out

    def flat_func(tt, v):
        return torch.cat([f.reshape(-1) for f in func(tt, unflat(v))])

    # torchdiffeq wraps the norm of a tuple state so that it sees the tuple: by default the mixed norm (the largest
    # per-tensor RMS); odeint_adjoint's default_adjoint_norm max(|t|, rms(y), rms(adj_y), mixed(adj_params)) is the
    # same thing for a single-tensor y
    if method in ADAPTIVE_METHODS:
        options = dict(options or {})
        options.setdefault("norm", lambda v: _mixed_norm(unflat(v)))
    sol = odeint(flat_func, flat0, t, rtol=rtol, atol=atol, method=method, options=options)
    return unflat(sol)


def odeint_adjoint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, event_fn=None,
                   adjoint_rtol=None, adjoint_atol=None, adjoint_method=None, adjoint_options=None,
                   adjoint_params=None):
    """``torchdiffeq.odeint_adjoint`` restatement (same-method adjoint, as the reference uses it)."""
    if event_fn is not None:
        raise NotImplementedError
    if not isinstance(func, torch.nn.Module) and adjoint_params is None:
        raise ValueError("func must be an nn.Module when adjoint_params is not given")
    if adjoint_params is None:
        adjoint_params = tuple(p for p in func.parameters() if p.requires_grad)
    else:
        adjoint_params = tuple(adjoint_params)
    if (adjoint_method not in (None, method) or adjoint_options not in (None, options)
            or adjoint_rtol not in (None, rtol) or adjoint_atol not in (None, atol)):
        raise NotImplementedError("reference never sets separate adjoint solver options")
    method = _check_inputs(y0, t, method)
    return _OdeintAdjoint.apply(func, method, rtol, atol, options,
                                len(adjoint_params), y0, t, *adjoint_params)
