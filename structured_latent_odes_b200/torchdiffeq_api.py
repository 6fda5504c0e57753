"""``torchdiffeq``-shaped entry points backed by the sm_100a kernels.

The reference crosses exactly one boundary on its hot path (``models/blackbox_ode.py:40-45``)::

    sol = torchdiffeq.odeint_adjoint(func=d_states_d_t, y0=init_state, t=self.times, method=self.solver)
    sol = torchdiffeq.odeint(func=d_states_d_t, y0=init_state, t=self.times, method=self.solver)

``odeint`` / ``odeint_adjoint`` below keep that signature and the ``(len(t), *y0.shape)`` return
layout.  They are not generic solvers: ``func`` must be one of the right-hand sides this package has
a fused kernel for (the reference's ``OdeFunc`` over ``Dynamics``, or ``CvsMechanistic``); anything
else raises ``NotImplementedError`` -- there is deliberately no eager / CPU fallback.

Gradient semantics
  * ``odeint``          -> exact discrete adjoint of the unrolled solver (what autograd through
                           ``torchdiffeq.odeint`` gives, ``adjoint_solver=False``).
  * ``odeint_adjoint``  -> torchdiffeq's continuous adjoint re-discretised with the same method and
                           restarted from the stored forward state at every output time; gradients
                           go to ``y0`` and ``func.parameters()`` only -- ``OdeFunc.constants`` (z) is
                           a plain tensor and gets none through the dynamics.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _cabi

__all__ = ["odeint", "odeint_adjoint", "install_as_torchdiffeq", "is_blackbox_func", "KernelTimer", "solve_latent",
           "solve_latent_heads", "solve_fixed_from_c"]

FIXED_METHODS = ("euler", "midpoint", "rk4")

class KernelTimer:
    """Optional CUDA-event timing of the C-ABI calls (bench.py uses it for the roofline numbers).

    ``with KernelTimer() as kt: ...`` records an event pair on the launching stream around every
    forward / backward library call made inside the block; ``kt.summary()`` synchronises and returns
    ``{"fwd": [ms...], "bwd": [ms...]}``.
    """

    active = None

    def __init__(self):
        self.events = {"fwd": [], "bwd": []}

    def __enter__(self):
        KernelTimer.active = self
        return self

    def __exit__(self, *exc):
        KernelTimer.active = None

    def summary(self):
        torch.cuda.synchronize()
        return {k: [a.elapsed_time(b) for a, b in v] for k, v in self.events.items()}


class _timed:
    def __init__(self, kind):
        self.kind = kind

    def __enter__(self):
        kt = KernelTimer.active
        if kt is not None:
            self.pair = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self.pair[0].record()
        return self

    def __exit__(self, *exc):
        kt = KernelTimer.active
        if kt is not None:
            self.pair[1].record()
            kt.events[self.kind].append(self.pair)


# ----------------------------------------------------------------------------------------------
# func recognition
# ----------------------------------------------------------------------------------------------
def is_blackbox_func(func) -> bool:
    """True for the reference's ``OdeFunc`` (``models/blackbox_ode.py:50-61``) or our mirror of it."""
    dyn = getattr(func, "dynamics", None)
    return (
        dyn is not None
        and isinstance(getattr(dyn, "dynamics_hidden", None), nn.Linear)
        and isinstance(getattr(dyn, "dyanamics_growth", None), nn.Linear)
        and isinstance(getattr(dyn, "dyanmics_degradation", None), nn.Linear)
        and torch.is_tensor(getattr(func, "constants", None))
    )


def _check_blackbox(func):
    dyn = func.dynamics
    prod = getattr(dyn, "prod", None)
    if isinstance(prod, nn.Sequential) and len(prod) > 1 and not isinstance(prod[1], nn.ReLU):
        raise NotImplementedError(
            f"hidden activation {type(prod[1]).__name__}: the fused kernels implement the ReLU hidden layer "
            "that OdeModel always builds (models/blackbox_ode.py:26-27)")
    z = func.constants
    hid, gro, deg = dyn.dynamics_hidden, dyn.dyanamics_growth, dyn.dyanmics_degradation
    if z.ndim != 2 or hid.in_features != z.shape[1] + 1:
        raise ValueError(f"constants {tuple(z.shape)} do not match dynamics_hidden.in_features={hid.in_features}")
    if gro.in_features != hid.out_features or deg.in_features != hid.out_features or gro.out_features != deg.out_features:
        raise ValueError("inconsistent Dynamics layer sizes")
    if hid.bias is None or gro.bias is None or deg.bias is None:
        raise NotImplementedError("Dynamics layers without bias")
    return z, hid, gro, deg


def _check_weights(dev, what, *tensors):
    """Every weight the kernels read through a raw pointer must be float32 on the device of the batch: a model left
    on the CPU / another GPU would be an illegal address (sticky, kills the context), a .double() model silent
    garbage.  The reference raises torch's device / dtype errors here; so do we, before any launch."""
    for name, x in tensors:
        if x is None:
            continue
        if x.device != dev:
            raise RuntimeError(f"{what}: {name} is on {x.device} but the batch is on {dev}; move the module with "
                               ".to(device) (no implicit copies, no CPU fallback)")
        if x.dtype != torch.float32:
            raise TypeError(f"{what}: {name} has dtype {x.dtype}; the kernels compute in float32 like the reference")


_T_CHECKED = [None, -1]   # (weak reference to the last grid tensor that passed, its version counter)


def _check_monotone(t):
    """t strictly increasing or decreasing.  For a device tensor the test is a device->host sync, so the verdict is
    remembered for the SAME tensor object at the same version (``OdeModel.times`` is handed over unchanged on every
    solve): repeated solves on one grid do not sync."""
    import weakref
    ref, ver = _T_CHECKED
    if ref is not None and ref() is t and ver == t._version:
        return
    d = t.detach()[1:] - t.detach()[:-1]
    if not (bool((d > 0).all()) or bool((d < 0).all())):
        raise ValueError("t must be strictly increasing or decreasing")
    _T_CHECKED[0], _T_CHECKED[1] = weakref.ref(t), t._version


def _check_common(func, y0, t, method, options, event_fn):
    if event_fn is not None:
        raise NotImplementedError("event_fn is not used by the reference and is not supported")
    if not torch.is_tensor(y0):
        raise NotImplementedError("tuple states are not supported (the reference passes one (B,S) tensor)")
    if method is None:
        method = "dopri5"  # torchdiffeq's default
    if method not in _cabi.METHODS:
        raise ValueError(f"Invalid method {method!r}; supported: {sorted(_cabi.METHODS)}")
    if not y0.is_cuda:
        raise RuntimeError("structured_latent_odes_b200 runs on CUDA tensors only (no CPU fallback); "
                           f"got y0 on {y0.device}")
    from .cvs_mechanistic import CvsMechanistic
    is_cvs = isinstance(func, CvsMechanistic)
    if y0.dtype != torch.float32 and not (is_cvs and y0.dtype == torch.float64):
        raise TypeError(f"y0 must be float32 (float64 only for CvsMechanistic), got {y0.dtype}")
    if y0.ndim != 2:
        raise ValueError(f"y0 must be (B, S), got {tuple(y0.shape)}")
    if not torch.is_tensor(t) or t.ndim != 1 or t.numel() < 1:
        raise ValueError("t must be a one-dimensional tensor")
    if not t.is_floating_point():
        raise TypeError("t must be floating point")
    if t.numel() > 1:
        _check_monotone(t)
    t = t.detach().to(device=y0.device, dtype=y0.dtype).contiguous()
    if method in FIXED_METHODS and options and not is_cvs:
        raise NotImplementedError(f"options={options!r} for fixed-grid solvers (the reference passes none: "
                                  "the solver grid is t itself)")
    return t, method


# ----------------------------------------------------------------------------------------------
# blackbox MLP dynamics, fixed grid
# ----------------------------------------------------------------------------------------------
def _ptr(x):
    return x.data_ptr() if x is not None else None


def _workspace(dev, backward, method_id, mode, B, T, L, H, S, fused, rows_in_time):
    """Scratch tensor for one fixed-grid call (flip records of the reverse sweep; wide-layer tables): sized by the
    library, owned by the torch caching allocator -- stream-ordered reuse, and safe inside CUDA-graph capture."""
    n = _cabi.lib().slode_fixed_workspace_bytes(int(backward), method_id, mode, B, T, L, H, S, fused, int(rows_in_time))
    if n < 0:
        raise _cabi.SlodeError("slode_fixed_workspace_bytes: " + _cabi.lib().slode_last_error().decode("utf-8", "replace"))
    return torch.empty(n, device=dev, dtype=torch.uint8) if n > 0 else None


def _dense_tbs_strides(x):
    """(stride_t, stride_b) of a (T,B,S) tensor usable by the kernels, or None if it must be copied."""
    st, sb, ss = x.stride()
    T, B, S = x.shape
    if S > 1 and ss != 1:
        return None
    if (T > 1 and st <= 0) or (B > 1 and sb <= 0):
        return None
    return st, sb


class _MlpFixedSolve(torch.autograd.Function):
    """sol = solve(y0, c, weights, t); one forward kernel, one reverse-sweep kernel."""

    @staticmethod
    def forward(ctx, y0, c, w1t, Wg, bg, Wd, bd, t, method_id, mode, layout):
        B, S = y0.shape
        H = c.shape[1]
        T = t.numel()
        y0c, cc = y0.contiguous(), c.contiguous()
        w = [x.detach().contiguous() for x in (w1t, Wg, bg, Wd, bd)]
        if layout == "bts":  # (B,T,S)-contiguous storage, returned as a (T,B,S) view
            store = torch.empty((B, T, S), device=y0.device, dtype=torch.float32)
            sol = store.permute(1, 0, 2)
        else:
            sol = torch.empty((T, B, S), device=y0.device, dtype=torch.float32)
        st, sb = sol.stride(0), sol.stride(1)
        with torch.cuda.device(y0.device), _timed("fwd"):
            ws = _workspace(y0.device, False, method_id, mode, B, T, 0, H, S, 0, st == S)
            stream = torch.cuda.current_stream().cuda_stream
            rc = _cabi.lib().slode_mlp_fixed_fwd(method_id, B, T, H, S, _ptr(t), _ptr(cc), _ptr(y0c),
                                                 *[_ptr(x) for x in w], _ptr(sol), st, sb, _ptr(ws),
                                                 ws.numel() if ws is not None else 0, stream)
        _cabi.check(rc, "slode_mlp_fixed_fwd")
        ctx.save_for_backward(cc, *w, t, sol)
        ctx.method_id, ctx.mode = method_id, mode
        return sol

    @staticmethod
    def backward(ctx, grad_sol):
        cc, w1t, Wg, bg, Wd, bd, t, sol = ctx.saved_tensors
        T, B, S = sol.shape
        H = cc.shape[1]
        strides = _dense_tbs_strides(grad_sol)
        if strides is None or grad_sol.dtype != torch.float32:
            grad_sol = grad_sol.to(torch.float32).contiguous()
            strides = (grad_sol.stride(0), grad_sol.stride(1))
        grad_y0 = torch.empty((B, S), device=sol.device, dtype=torch.float32)
        grad_c = torch.empty((B, H), device=sol.device, dtype=torch.float32)
        grad_w = torch.zeros(H + 2 * (S * H + S), device=sol.device, dtype=torch.float32)
        with torch.cuda.device(sol.device), _timed("bwd"):
            ws = _workspace(sol.device, True, ctx.method_id, ctx.mode, B, T, 0, H, S, 0, False)
            stream = torch.cuda.current_stream().cuda_stream
            rc = _cabi.lib().slode_mlp_fixed_bwd(
                ctx.method_id, ctx.mode, B, T, H, S, _ptr(t), _ptr(cc), _ptr(w1t), _ptr(Wg), _ptr(bg), _ptr(Wd),
                _ptr(bd), _ptr(sol), sol.stride(0), sol.stride(1), _ptr(grad_sol), strides[0], strides[1],
                _ptr(grad_y0), _ptr(grad_c), _ptr(grad_w), _ptr(ws), ws.numel() if ws is not None else 0, stream)
        _cabi.check(rc, "slode_mlp_fixed_bwd")
        o = 0
        gw1t = grad_w[o:o + H]; o += H
        gWg = grad_w[o:o + S * H].view(S, H); o += S * H
        gbg = grad_w[o:o + S]; o += S
        gWd = grad_w[o:o + S * H].view(S, H); o += S * H
        gbd = grad_w[o:o + S]
        return grad_y0, grad_c, gw1t, gWg, gbg, gWd, gbd, None, None, None, None


def solve_fixed_from_c(y0, c, w1t, Wg, bg, Wd, bd, t, method, adjoint=False, layout="tbs"):
    """The solve from precomputed ``c = z W1[:,1:]^T + b1`` (B,H) and ``y0`` (B,S): the ``slode_mlp_fixed_*`` entry
    points, for hosts that already hold those (gradients flow to ``y0``, ``c`` and the five weight tensors)."""
    if method not in FIXED_METHODS:
        raise NotImplementedError(method)
    t = t.detach().to(device=y0.device, dtype=torch.float32).contiguous()
    mode = _cabi.BWD_TDE_ADJOINT if adjoint else _cabi.BWD_DISCRETE
    return _MlpFixedSolve.apply(y0, c, w1t, Wg, bg, Wd, bd, t, _cabi.METHODS[method], mode, layout)


class _LatentFixedSolve(torch.autograd.Function):
    """The fused solve: c = z W1[:,1:]^T + b1 (and optionally x0 = latent_to_ode_net(z)) computed inside the solver
    kernels, their gradients inside the reverse sweep.  ``x0net`` is (Wa, ba, Wb, bb) or four Nones (then y0 is an
    input)."""

    @staticmethod
    def forward(ctx, z, y0, W1, b1, Wg, bg, Wd, bd, Wa, ba, Wb, bb, t, method_id, mode, layout):
        B, L = z.shape
        H = W1.shape[0]
        S = Wg.shape[0]
        T = t.numel()
        fx0 = Wa is not None
        zc = z.detach().to(torch.float32).contiguous()
        w = [x.detach().contiguous() for x in (W1, b1, Wg, bg, Wd, bd)]
        x0w = [x.detach().contiguous() for x in (Wa, ba, Wb, bb)] if fx0 else [None] * 4
        y0c = None if fx0 else y0.detach().contiguous()
        if layout == "bts":
            sol = torch.empty((B, T, S), device=z.device, dtype=torch.float32).permute(1, 0, 2)
        else:
            sol = torch.empty((T, B, S), device=z.device, dtype=torch.float32)
        with torch.cuda.device(z.device), _timed("fwd"):
            ws = _workspace(z.device, False, method_id, mode, B, T, L, H, S, 2 if fx0 else 1, sol.stride(0) == S)
            rc = _cabi.lib().slode_latent_fixed_fwd(
                method_id, B, T, L, H, S, _ptr(t), _ptr(zc), *[_ptr(x) for x in w], *[_ptr(x) for x in x0w], _ptr(y0c),
                _ptr(sol), sol.stride(0), sol.stride(1), _ptr(ws), ws.numel() if ws is not None else 0,
                torch.cuda.current_stream().cuda_stream)
        _cabi.check(rc, "slode_latent_fixed_fwd")
        ctx.save_for_backward(zc, *w, t, sol, *(x0w if fx0 else []))
        ctx.cfg = (method_id, mode, fx0)
        return sol

    @staticmethod
    def backward(ctx, grad_sol):
        method_id, mode, fx0 = ctx.cfg
        zc, W1, b1, Wg, bg, Wd, bd, t, sol, *x0w = ctx.saved_tensors
        if not fx0:
            x0w = [None] * 4
        T, B, S = sol.shape
        H, L = W1.shape[0], zc.shape[1]
        strides = _dense_tbs_strides(grad_sol)
        if strides is None or grad_sol.dtype != torch.float32:
            grad_sol = grad_sol.to(torch.float32).contiguous()
            strides = (grad_sol.stride(0), grad_sol.stride(1))
        dev = sol.device
        grad_z = torch.empty((B, L), device=dev, dtype=torch.float32)
        grad_y0 = None if fx0 else torch.empty((B, S), device=dev, dtype=torch.float32)
        nbase = H + 2 * (S * H + S)
        n = nbase + H * L + H + ((H * L + H + S * H + S) if fx0 else 0)
        gp = torch.zeros(n, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev), _timed("bwd"):
            ws = _workspace(dev, True, method_id, mode, B, T, L, H, S, 2 if fx0 else 1, False)
            rc = _cabi.lib().slode_latent_fixed_bwd(
                method_id, mode, B, T, L, H, S, _ptr(t), _ptr(zc), _ptr(W1), _ptr(b1), _ptr(Wg), _ptr(bg), _ptr(Wd),
                _ptr(bd), *[_ptr(x) for x in x0w], _ptr(sol), sol.stride(0), sol.stride(1), _ptr(grad_sol), strides[0],
                strides[1], _ptr(grad_z), _ptr(grad_y0), _ptr(gp), _ptr(ws), ws.numel() if ws is not None else 0,
                torch.cuda.current_stream().cuda_stream)
        _cabi.check(rc, "slode_latent_fixed_bwd")
        o = 0

        def take(*shape):
            nonlocal o
            k = 1
            for d in shape:
                k *= d
            v = gp[o:o + k].view(*shape)
            o += k
            return v

        gw1t, gWg, gbg, gWd, gbd = take(H), take(S, H), take(S), take(S, H), take(S)
        gW1z, gb1 = take(H, L), take(H)
        gW1 = torch.cat([gw1t[:, None], gW1z], dim=1)
        gx0 = (take(H, L), take(H), take(S, H), take(S)) if fx0 else (None,) * 4
        if mode == _cabi.BWD_TDE_ADJOINT and not fx0:
            grad_z = None  # odeint_adjoint: the constants get no gradient through the dynamics
        return (grad_z, grad_y0, gW1, gb1, gWg, gbg, gWd, gbd, *gx0, None, None, None, None)


def _latent_checked(z, dynamics, x0_net, t, method, who):
    """Argument checks shared by the fused latent entry points; returns (t on the device, the ten weight tensors)."""
    if method not in FIXED_METHODS:
        raise NotImplementedError(method)
    if not z.is_cuda:
        raise RuntimeError("structured_latent_odes_b200 runs on CUDA tensors only (no CPU fallback); "
                           f"got z on {z.device}")
    hid, gro, deg = dynamics.dynamics_hidden, dynamics.dyanamics_growth, dynamics.dyanmics_degradation
    la, lb = x0_net[0], x0_net[2]
    H, S = hid.out_features, gro.out_features
    if not _cabi.lib().slode_mlp_supported(H, S):
        raise NotImplementedError(f"(ode_hidden_dim={H}, ode_state_dim={S}) has no compiled kernel; available (H,S): "
                                  f"{_cabi.supported_shapes()}. There is no generic fallback.")
    if la.out_features != H or lb.in_features != H or lb.out_features != S or la.in_features != z.shape[1]:
        raise ValueError("latent_to_ode_net layer sizes do not match the dynamics")
    if z.ndim != 2 or hid.in_features != z.shape[1] + 1:
        raise ValueError(f"z {tuple(z.shape)} does not match dynamics_hidden.in_features={hid.in_features}")
    if not z.is_floating_point():
        raise TypeError(f"z must be floating point, got {z.dtype}")
    _check_weights(z.device, who, ("dynamics_hidden.weight", hid.weight), ("dynamics_hidden.bias", hid.bias),
                   ("dyanamics_growth.weight", gro.weight), ("dyanamics_growth.bias", gro.bias),
                   ("dyanmics_degradation.weight", deg.weight), ("dyanmics_degradation.bias", deg.bias),
                   ("latent_to_ode_net.0.weight", la.weight), ("latent_to_ode_net.0.bias", la.bias),
                   ("latent_to_ode_net.2.weight", lb.weight), ("latent_to_ode_net.2.bias", lb.bias))
    t = t.detach().to(device=z.device, dtype=torch.float32).contiguous()
    return t, (hid.weight, hid.bias, gro.weight, gro.bias, deg.weight, deg.bias, la.weight, la.bias, lb.weight,
               lb.bias)


def solve_latent(z, dynamics, x0_net, t, method, adjoint, layout="tbs"):
    """Whole ``OdeModel.solve_ODE`` body in two kernels: returns ``(T,B,S)``.  ``x0_net`` is the reference's
    ``latent_to_ode_net`` Sequential(Linear, ReLU, Linear, Sigmoid)."""
    t, weights = _latent_checked(z, dynamics, x0_net, t, method, "solve_ODE")
    mode = _cabi.BWD_TDE_ADJOINT if adjoint else _cabi.BWD_DISCRETE
    return _LatentFixedSolve.apply(z, None, *weights, t, _cabi.METHODS[method], mode, layout)


@torch.no_grad()
def solve_latent_heads(z, dynamics, x0_net, t, method, head_weights, want_solution=False, layout="tbs",
                       contiguous=False):
    """``OdeModel.solve_ODE`` + the decoder heads (``models/decoders.py:43-47``, ``:85-86``) in ONE kernel, for callers
    that do not differentiate (reconstruction, posterior sampling): returns ``(mu, sol)`` with ``mu`` the
    ``(NQ,B,O,T)`` head outputs -- ``mu[q]`` is ``Linear_q(solution).permute(0,2,1)`` -- and ``sol`` the ``(T,B,S)``
    trajectories, or ``None`` unless ``want_solution`` (they are then never written to HBM).  Bit-equal to
    ``solve_latent`` followed by ``decoders.decoder_heads``.

    ``mu`` is a ``[..., :T]`` view of a buffer whose rows are padded to a multiple of eight floats (32-byte sectors:
    every row then has the same sector phase and the kernel's stores never diverge); ``contiguous=True`` asks for
    the reference's exactly contiguous ``(B,O,T)`` rows instead (same values, the slower store schedule when
    ``T % 8 != 0``)."""
    t, weights = _latent_checked(z, dynamics, x0_net, t, method, "solve_latent_heads")
    W = torch.stack([w.detach() for w in head_weights], dim=0)
    _check_weights(z.device, "solve_latent_heads", ("decoder head weights", W))
    W = W.contiguous()
    B, L = z.shape
    H, S = weights[0].shape[0], weights[2].shape[0]
    NQ, O, SW = W.shape
    if SW != S:
        raise ValueError(f"head weights {tuple(W.shape)} do not match ode_state_dim={S}")
    T = t.numel()
    method_id = _cabi.METHODS[method]
    zc = z.detach().to(torch.float32).contiguous()
    w = [x.detach().contiguous() for x in weights]
    pitch = T if contiguous else (T + 7) // 8 * 8
    mu = torch.empty((NQ, B, O, pitch), device=z.device, dtype=torch.float32)
    sol = None
    if want_solution:
        # always (T,B,S)-contiguous here: the fused kernel writes a state row per step, coalesced across the
        # trajectories; ``layout`` is a storage choice of the other entry points, the values are the same
        sol = torch.empty((T, B, S), device=z.device, dtype=torch.float32)
    with torch.cuda.device(z.device), _timed("fwd"):
        ws = _workspace(z.device, False, method_id, _cabi.BWD_DISCRETE, B, T, L, H, S, 2, False)
        rc = _cabi.lib().slode_latent_fixed_heads_fwd(
            method_id, B, T, L, H, S, _ptr(t), _ptr(zc), *[_ptr(x) for x in w], None, O, NQ, _ptr(W), _ptr(mu),
            pitch, _ptr(sol), sol.stride(0) if sol is not None else 0, sol.stride(1) if sol is not None else 0,
            _ptr(ws), ws.numel() if ws is not None else 0, torch.cuda.current_stream().cuda_stream)
    _cabi.check(rc, "slode_latent_fixed_heads_fwd")
    return mu[..., :T], sol


class SolverStats:
    """Bookkeeping of the last dopri5 solve on this process (``last_dopri5_stats``): accepted / rejected step
    counts, RHS evaluations per trajectory, and -- when ``options={"log_steps": True}`` -- the (t0, dt, accepted)
    record of every attempted step as a float64 CPU tensor."""

    def __init__(self):
        self.n_accept = self.n_reject = self.n_rhs = 0
        self.steps = None


last_dopri5_stats = SolverStats()
_DOPRI5_STATUS = {1: "underflow in dt", 2: "max_num_steps exceeded", 3: "checkpoint capacity exceeded",
                  4: "replay_steps ended before the last output time"}


class Dopri5ShardSolve:
    """One shard of a dopri5 forward solve whose batch is split over several devices / processes
    (``slode_mlp_dopri5_fwd_step``).  torchdiffeq's controller is batch-global: ``step()`` runs ONE pass over this
    shard and leaves the shard's sums of squares in ``self.out`` (device float64[2]); the caller adds ``out`` over all
    shards, hands the total back through ``self.ext`` and steps again until ``step()`` returns True.  Every shard then
    takes the accept / reject decisions and the step sizes of the unsharded solve."""

    def __init__(self, y0, c, w, t, rtol, atol, n_global, first_step=None, max_num_steps=1 << 20, log_cap=0,
                 ckpt_cap=0, layout="tbs"):
        B, S = y0.shape
        if B < 1:
            raise ValueError("a dopri5 shard may not be empty (give every rank at least one trajectory)")
        dev = y0.device
        self.args = (y0, c, w, t)
        self.B, self.S, self.H, self.T = B, S, c.shape[1], t.numel()
        self.cfg = (float(rtol), float(atol), float(first_step) if first_step is not None else -1.0, int(max_num_steps),
                    int(n_global))
        if layout == "bts":
            self.sol = torch.empty((B, self.T, S), device=dev, dtype=torch.float32).permute(1, 0, 2)
        else:
            self.sol = torch.empty((self.T, B, S), device=dev, dtype=torch.float32)
        self.ckpt = torch.empty((ckpt_cap, B, S), device=dev, dtype=torch.float32) if ckpt_cap else None
        self.log = torch.zeros((log_cap, 3), device=dev, dtype=torch.float64) if log_cap else None
        self.stats = torch.zeros(5, device=dev, dtype=torch.int64)
        self.ext = torch.zeros(2, device=dev, dtype=torch.float64)
        self.out = torch.zeros(2, device=dev, dtype=torch.float64)
        n = _cabi.lib().slode_mlp_dopri5_step_workspace_bytes(B, S)
        if n < 0:
            raise _cabi.SlodeError("slode_mlp_dopri5_step_workspace_bytes: "
                                   + _cabi.lib().slode_last_error().decode("utf-8", "replace"))
        self.ws = torch.empty(n, device=dev, dtype=torch.uint8)
        self.restart = 1
        self.done = False

    def step(self):
        y0, c, w, t = self.args
        rtol, atol, first_step, max_steps, n_global = self.cfg
        with torch.cuda.device(y0.device), _timed("fwd"):
            rc = _cabi.lib().slode_mlp_dopri5_fwd_step(
                self.B, self.T, self.H, self.S, _ptr(t), _ptr(c), _ptr(y0), *[_ptr(x) for x in w], rtol, atol, first_step,
                max_steps, n_global, self.restart, _ptr(self.ext), _ptr(self.out), _ptr(self.sol), self.sol.stride(0),
                self.sol.stride(1), _ptr(self.ckpt), self.ckpt.shape[0] if self.ckpt is not None else 0, _ptr(self.log),
                self.log.shape[0] if self.log is not None else 0, _ptr(self.stats), _ptr(self.ws), self.ws.numel(),
                torch.cuda.current_stream().cuda_stream)
        _cabi.check(rc, "slode_mlp_dopri5_fwd_step")
        self.restart = 0
        self.done = int(self.stats[4].item()) == 5   # the host has to know when to stop: one small read per pass
        return self.done

    def result(self):
        n_acc, n_rej, n_rhs, status = (int(v) for v in self.stats[:4].tolist())
        return n_acc, n_rej, n_rhs, status


def _dopri5_forward(y0, c, w, t, rtol, atol, options, layout, want_ckpt):
    """Runs slode_mlp_dopri5_fwd; returns (sol, ckpt or None, steps (n,3) float64 device tensor or None).

    ``options["shard_reducer"]`` (a callable that sums a device float64 tensor over all shards in place, e.g.
    ``lambda x: dist.all_reduce(x)``) together with ``options["global_batch"]`` switches to the sharded solve: one
    pass per launch with the batch-wide sums combined across the shards (``Dopri5ShardSolve``)."""
    opts = dict(options or {})
    first_step = opts.pop("first_step", None)
    max_num_steps = int(opts.pop("max_num_steps", 1 << 20))
    log_steps = bool(opts.pop("log_steps", False)) or want_ckpt
    replay = opts.pop("replay_steps", None)
    reducer = opts.pop("shard_reducer", None)
    n_global = opts.pop("global_batch", None)
    if replay is not None:
        replay = torch.as_tensor(replay, dtype=torch.float64).reshape(-1, 3).to(y0.device).contiguous()
    if opts:
        raise NotImplementedError(f"dopri5 options {sorted(opts)} are not supported (supported: first_step, "
                                  "max_num_steps, log_steps, replay_steps, shard_reducer + global_batch)")
    if (reducer is None) != (n_global is None):
        raise ValueError("shard_reducer and global_batch go together")
    if reducer is not None and replay is not None:
        raise NotImplementedError("replay_steps in a sharded dopri5 solve")
    B, S = y0.shape
    H = c.shape[1]
    T = t.numel()
    dev = y0.device
    cap = 64 if want_ckpt else 0
    log_cap = 4096 if log_steps else 0
    if reducer is not None:
        while True:
            sh = Dopri5ShardSolve(y0, c, w, t, rtol, atol, n_global, first_step, max_num_steps, log_cap, cap, layout)
            while not sh.step():
                sh.ext.copy_(sh.out)
                reducer(sh.ext)
            n_acc, n_rej, n_rhs, status = sh.result()
            if status == 3 or (log_cap and n_acc + n_rej > log_cap):   # identical on every shard: all of them re-run
                cap = max(2 * cap, 64) if want_ckpt else 0
                log_cap = max(2 * log_cap, n_acc + n_rej) if log_cap else 0
                continue
            if status != 0:
                raise _cabi.SlodeError(f"dopri5: {_DOPRI5_STATUS.get(status, status)} after {n_acc + n_rej} attempted steps")
            break
        last_dopri5_stats.n_accept, last_dopri5_stats.n_reject, last_dopri5_stats.n_rhs = n_acc, n_rej, n_rhs
        steps = sh.log[: n_acc + n_rej] if sh.log is not None else None
        last_dopri5_stats.steps = steps.cpu() if steps is not None else None
        return sh.sol, (sh.ckpt[:n_acc] if sh.ckpt is not None else None), steps
    if layout == "bts":
        sol = torch.empty((B, T, S), device=dev, dtype=torch.float32).permute(1, 0, 2)
    else:
        sol = torch.empty((T, B, S), device=dev, dtype=torch.float32)
    stats = torch.zeros(4, device=dev, dtype=torch.int64)
    while True:
        ckpt = torch.empty((cap, B, S), device=dev, dtype=torch.float32) if cap else None
        log = torch.zeros((log_cap, 3), device=dev, dtype=torch.float64) if log_cap else None
        with torch.cuda.device(dev), _timed("fwd"):
            rc = _cabi.lib().slode_mlp_dopri5_fwd(
                B, T, H, S, _ptr(t), _ptr(c), _ptr(y0), *[_ptr(x) for x in w], float(rtol), float(atol),
                float(first_step) if first_step is not None else -1.0, max_num_steps, _ptr(replay),
                replay.shape[0] if replay is not None else 0, _ptr(sol), sol.stride(0),
                sol.stride(1), _ptr(ckpt), cap, _ptr(log), log_cap, _ptr(stats),
                torch.cuda.current_stream().cuda_stream)
        _cabi.check(rc, "slode_mlp_dopri5_fwd")
        n_acc, n_rej, n_rhs, status = (int(v) for v in stats.tolist())  # the adaptive loop has to finish anyway
        if status == 3 or (log_cap and n_acc + n_rej > log_cap):
            cap = max(2 * cap, 64) if want_ckpt else 0
            log_cap = max(2 * log_cap, n_acc + n_rej) if log_cap else 0
            continue  # deterministic step sequence: the re-run reproduces the same steps with room for all of them
        if status != 0:
            raise _cabi.SlodeError(f"dopri5: {_DOPRI5_STATUS.get(status, status)} after {n_acc + n_rej} attempted steps")
        break
    last_dopri5_stats.n_accept, last_dopri5_stats.n_reject, last_dopri5_stats.n_rhs = n_acc, n_rej, n_rhs
    steps = log[: n_acc + n_rej] if log is not None else None
    last_dopri5_stats.steps = steps.cpu() if steps is not None else None
    return sol, (ckpt[:n_acc] if ckpt is not None else None), steps


def _check_func_tensors(y0, z, hid, gro, deg):
    if z.device != y0.device:
        raise RuntimeError(f"func.constants is on {z.device} but y0 is on {y0.device}")
    if not z.is_floating_point():
        raise TypeError(f"func.constants must be floating point, got {z.dtype}")
    _check_weights(y0.device, "odeint", ("dynamics_hidden.weight", hid.weight), ("dynamics_hidden.bias", hid.bias),
                   ("dyanamics_growth.weight", gro.weight), ("dyanamics_growth.bias", gro.bias),
                   ("dyanmics_degradation.weight", deg.weight), ("dyanmics_degradation.bias", deg.bias))


def _solve_blackbox(func, y0, t, method, mode, layout):
    z, hid, gro, deg = _check_blackbox(func)
    B, S = y0.shape
    if z.shape[0] != B:
        raise ValueError(f"constants batch {z.shape[0]} != y0 batch {B}")
    if gro.out_features != S:
        raise ValueError(f"y0 state dim {S} != Dynamics n_outputs {gro.out_features}")
    H = hid.out_features
    if not _cabi.lib().slode_mlp_supported(H, S):
        raise NotImplementedError(
            f"(ode_hidden_dim={H}, ode_state_dim={S}) has no compiled kernel; available (H,S): "
            f"{_cabi.supported_shapes()}. There is no generic fallback.")
    _check_func_tensors(y0, z, hid, gro, deg)
    # the time-invariant part of the hidden pre-activation, c = z W1[:,1:]^T + b1, is computed inside the kernels
    return _LatentFixedSolve.apply(z, y0, hid.weight, hid.bias, gro.weight, gro.bias, deg.weight, deg.bias,
                                   None, None, None, None, t, _cabi.METHODS[method], mode, layout)


class _MlpDopri5Solve(torch.autograd.Function):
    """dopri5 solve with the exact gradient of its accepted-step sequence (odeint + autograd parity)."""

    @staticmethod
    def forward(ctx, y0, c, w1t, Wg, bg, Wd, bd, t, rtol, atol, options, layout):
        w = [x.detach().contiguous() for x in (w1t, Wg, bg, Wd, bd)]
        cc = c.contiguous()
        sol, ckpt, steps = _dopri5_forward(y0.contiguous(), cc, w, t, rtol, atol, options, layout, want_ckpt=True)
        # accepted steps and the output times interpolated inside each of them (tiny, host side)
        log = steps.cpu() if steps is not None else torch.zeros((0, 3), dtype=torch.float64)
        acc = log[log[:, 2] != 0][:, :2].contiguous()
        tt = t.detach().double().cpu()
        emit, out_idx = [1], 1
        sgn = -1.0 if tt.numel() > 1 and float(tt[-1]) < float(tt[0]) else 1.0   # decreasing t: steps have dt < 0
        for t0, dt in acc.tolist():
            t1 = t0 + dt
            while out_idx < tt.numel() and sgn * float(tt[out_idx]) <= sgn * t1:
                out_idx += 1
            emit.append(out_idx)
        ctx.save_for_backward(cc, *w, t, ckpt if ckpt is not None else sol.new_zeros(0))
        ctx.acc = acc.to(y0.device)
        ctx.emit = torch.tensor(emit, dtype=torch.int32, device=y0.device)
        ctx.shape = tuple(sol.shape)
        return sol

    @staticmethod
    def backward(ctx, grad_sol):
        cc, w1t, Wg, bg, Wd, bd, t, ckpt = ctx.saved_tensors
        T, B, S = ctx.shape
        H = cc.shape[1]
        strides = _dense_tbs_strides(grad_sol)
        if strides is None or grad_sol.dtype != torch.float32:
            grad_sol = grad_sol.to(torch.float32).contiguous()
            strides = (grad_sol.stride(0), grad_sol.stride(1))
        dev = cc.device
        grad_y0 = torch.empty((B, S), device=dev, dtype=torch.float32)
        grad_c = torch.empty((B, H), device=dev, dtype=torch.float32)
        grad_w = torch.zeros(H + 2 * (S * H + S), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev), _timed("bwd"):
            rc = _cabi.lib().slode_mlp_dopri5_bwd(
                B, T, H, S, _ptr(t), _ptr(cc), _ptr(w1t), _ptr(Wg), _ptr(bg), _ptr(Wd), _ptr(bd), ctx.acc.shape[0],
                _ptr(ctx.acc), _ptr(ctx.emit), _ptr(ckpt), _ptr(grad_sol), strides[0], strides[1], _ptr(grad_y0),
                _ptr(grad_c), _ptr(grad_w), torch.cuda.current_stream().cuda_stream)
        _cabi.check(rc, "slode_mlp_dopri5_bwd")
        o = 0
        gw1t = grad_w[o:o + H]; o += H
        gWg = grad_w[o:o + S * H].view(S, H); o += S * H
        gbg = grad_w[o:o + S]; o += S
        gWd = grad_w[o:o + S * H].view(S, H); o += S * H
        gbd = grad_w[o:o + S]
        return grad_y0, grad_c, gw1t, gWg, gbg, gWd, gbd, None, None, None, None, None


class _MlpDopri5AdjointSolve(torch.autograd.Function):
    """``odeint_adjoint(..., method="dopri5")``: the forward is the plain adaptive solve; the backward is
    torchdiffeq's adjoint -- one fresh adaptive solve of [y, a, a_theta] per output interval, backwards in time,
    restarted from the stored forward state (``slode_mlp_dopri5_adjoint_bwd``).  Gradients: ``y0`` and
    ``func.parameters()``; ``func.constants`` gets none."""

    @staticmethod
    def forward(ctx, y0, z, W1, b1, Wg, bg, Wd, bd, t, rtol, atol, options, layout):
        zc = z.detach().to(torch.float32).contiguous()
        W1c, b1c = W1.detach().contiguous(), b1.detach().contiguous()
        c = torch.addmm(b1c, zc, W1c[:, 1:].t()).contiguous()
        hw = [x.detach().contiguous() for x in (Wg, bg, Wd, bd)]
        fwd_opts = {k: v for k, v in (options or {}).items() if k != "adjoint_replay_steps"}
        sol, _, _ = _dopri5_forward(y0.detach().contiguous(), c, [W1c[:, 0].contiguous()] + hw, t, rtol, atol, fwd_opts,
                                    layout, want_ckpt=False)
        ctx.save_for_backward(zc, c, W1c, *hw, t, sol)
        opts = dict(options or {})
        replay = opts.get("adjoint_replay_steps")   # (n,4) rows of a backward step log to take instead of controlling
        if replay is not None:
            replay = torch.as_tensor(replay, dtype=torch.float64).reshape(-1, 4).to(sol.device).contiguous()
        ctx.replay = replay
        ctx.cfg = (float(rtol), float(atol), int(opts.get("max_num_steps", 1 << 20)), bool(opts.get("log_steps", False)))
        return sol

    @staticmethod
    def backward(ctx, grad_sol):
        zc, c, W1, Wg, bg, Wd, bd, t, sol = ctx.saved_tensors
        rtol, atol, max_steps, log_steps = ctx.cfg
        T, B, S = sol.shape
        H, L = W1.shape[0], zc.shape[1]
        strides = _dense_tbs_strides(grad_sol)
        if strides is None or grad_sol.dtype != torch.float32:
            grad_sol = grad_sol.to(torch.float32).contiguous()
            strides = (grad_sol.stride(0), grad_sol.stride(1))
        dev = sol.device
        lib = _cabi.lib()
        n = lib.slode_mlp_dopri5_adjoint_workspace_bytes(B, L, H, S)
        if n < 0:
            _cabi.check(2, "slode_mlp_dopri5_adjoint_workspace_bytes")
        P = H * (L + 1) + H + 2 * (S * H + S)
        grad_y0 = torch.empty((B, S), device=dev, dtype=torch.float32)
        gp = torch.empty(P, device=dev, dtype=torch.float32)
        stats = torch.zeros(4, device=dev, dtype=torch.int64)
        log_cap = 1 << 16 if log_steps else 0
        log = torch.zeros((log_cap, 4), device=dev, dtype=torch.float64) if log_cap else None
        with torch.cuda.device(dev), _timed("bwd"):
            ws = torch.empty(max(n, 256), device=dev, dtype=torch.uint8)
            rc = lib.slode_mlp_dopri5_adjoint_bwd(
                B, T, L, H, S, _ptr(t), _ptr(zc), _ptr(c), _ptr(W1), _ptr(Wg), _ptr(bg), _ptr(Wd), _ptr(bd), _ptr(sol),
                sol.stride(0), sol.stride(1), _ptr(grad_sol), strides[0], strides[1], rtol, atol, max_steps,
                _ptr(ctx.replay), ctx.replay.shape[0] if ctx.replay is not None else 0, _ptr(grad_y0), _ptr(gp), _ptr(log), log_cap, _ptr(stats), _ptr(ws), ws.numel(),
                torch.cuda.current_stream().cuda_stream)
        _cabi.check(rc, "slode_mlp_dopri5_adjoint_bwd")
        n_acc, n_rej, n_rhs, status = (int(v) for v in stats.tolist())
        if status != 0:
            raise _cabi.SlodeError(f"dopri5 adjoint: {_DOPRI5_STATUS.get(status, status)} after {n_acc + n_rej} attempted steps")
        st = last_dopri5_adjoint_stats
        st.n_accept, st.n_reject, st.n_rhs = n_acc, n_rej, n_rhs
        st.steps = log[: min(n_acc + n_rej, log_cap)].cpu() if log is not None else None
        o = 0

        def take(*shape):
            nonlocal o
            k = 1
            for d in shape:
                k *= d
            v = gp[o:o + k].view(*shape)
            o += k
            return v

        gW1, gb1, gWg, gbg, gWd, gbd = take(H, L + 1), take(H), take(S, H), take(S), take(S, H), take(S)
        return grad_y0, None, gW1, gb1, gWg, gbg, gWd, gbd, None, None, None, None, None


last_dopri5_adjoint_stats = SolverStats()   # backward pass of the last odeint_adjoint(method="dopri5") call


def _solve_blackbox_dopri5(func, y0, t, rtol, atol, options, mode, layout):
    z, hid, gro, deg = _check_blackbox(func)
    B, S = y0.shape
    if z.shape[0] != B:
        raise ValueError(f"constants batch {z.shape[0]} != y0 batch {B}")
    H = hid.out_features
    if gro.out_features != S or not _cabi.lib().slode_dopri5_supported(H, S):
        raise NotImplementedError(f"(ode_hidden_dim={H}, ode_state_dim={S}) has no compiled dopri5 kernel")
    _check_func_tensors(y0, z, hid, gro, deg)
    decreasing = t.numel() > 1 and bool(t[0] > t[-1])
    needs_grad = torch.is_grad_enabled() and (y0.requires_grad or z.requires_grad
                                              or any(p.requires_grad for p in func.parameters()))
    if needs_grad:
        if mode == _cabi.BWD_TDE_ADJOINT:
            if decreasing:
                raise NotImplementedError("odeint_adjoint + dopri5 with decreasing output times (the reference always "
                                          "integrates forward); odeint (discrete gradient) takes them")
            bad = sorted(set(options or {}) & {"first_step", "shard_reducer", "global_batch"})
            if bad:
                raise NotImplementedError(f"odeint_adjoint with dopri5: options {bad} (the reference passes none)")
            return _MlpDopri5AdjointSolve.apply(y0, z, hid.weight, hid.bias, gro.weight, gro.bias, deg.weight,
                                                deg.bias, t, rtol, atol, options, layout)
        W1g = hid.weight
        cg = torch.addmm(hid.bias, z.to(torch.float32), W1g[:, 1:].t())
        return _MlpDopri5Solve.apply(y0, cg, W1g[:, 0], gro.weight, gro.bias, deg.weight, deg.bias, t, rtol, atol,
                                     options, layout)
    W1 = hid.weight.detach()
    c = torch.addmm(hid.bias.detach(), z.detach().to(torch.float32), W1[:, 1:].t()).contiguous()
    w = [x.detach().contiguous() for x in (W1[:, 0], gro.weight, gro.bias, deg.weight, deg.bias)]
    sol, _, _ = _dopri5_forward(y0.detach().contiguous(), c, w, t, rtol, atol, options, layout, want_ckpt=False)
    return sol


# ----------------------------------------------------------------------------------------------
# public API
# ----------------------------------------------------------------------------------------------
def _solve(func, y0, t, rtol, atol, method, options, event_fn, mode, layout):
    t, method = _check_common(func, y0, t, method, options, event_fn)
    if layout not in ("tbs", "bts"):
        raise ValueError("layout must be 'tbs' (torchdiffeq's) or 'bts'")
    from . import cvs_mechanistic as _cvs
    if isinstance(func, _cvs.CvsMechanistic):
        if method == "dopri5":
            raise NotImplementedError("dopri5 for the CVS mechanistic dynamics: use rk4 with options={'step_size': h}")
        return _cvs.solve_cvs(func, y0, t, method, mode, layout, options)
    if is_blackbox_func(func):
        if method == "dopri5":
            return _solve_blackbox_dopri5(func, y0, t, rtol, atol, options, mode, layout)
        return _solve_blackbox(func, y0, t, method, mode, layout)
    raise NotImplementedError(
        f"func of type {type(func).__name__} has no fused kernel; supported: OdeFunc over Dynamics "
        "(models/blackbox_ode.py) and CvsMechanistic (data/cvs/cvs_data.py dx_dt). There is no generic fallback solver.")


def odeint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, event_fn=None, layout="tbs"):
    """Drop-in for ``torchdiffeq.odeint`` on the reference's hot path; returns ``(len(t), B, S)``.

    ``layout="bts"`` stores the result (B,T,S)-contiguous (still returned as a (T,B,S) view) so that the
    ``sol.permute(1, 0, 2)`` in ``OdeModel.solve_ODE`` hands the decoder a contiguous tensor.
    """
    return _solve(func, y0, t, rtol, atol, method, options, event_fn, _cabi.BWD_DISCRETE, layout)


def odeint_adjoint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, event_fn=None,
                   adjoint_rtol=None, adjoint_atol=None, adjoint_method=None, adjoint_options=None,
                   adjoint_params=None, layout="tbs"):
    """Drop-in for ``torchdiffeq.odeint_adjoint`` as the reference calls it (same-method adjoint)."""
    if not isinstance(func, nn.Module):
        raise ValueError("func must be an instance of nn.Module to specify the adjoint parameters")
    if adjoint_method not in (None, method) or adjoint_options is not None \
            or adjoint_rtol not in (None, rtol) or adjoint_atol not in (None, atol):
        raise NotImplementedError("separate adjoint solver settings are never set by the reference")
    if adjoint_params is not None:
        want = {id(p) for p in func.parameters()}
        if {id(p) for p in adjoint_params} != want:
            raise NotImplementedError("adjoint_params other than tuple(func.parameters())")
    return _solve(func, y0, t, rtol, atol, method, options, event_fn, _cabi.BWD_TDE_ADJOINT, layout)


def install_as_torchdiffeq():
    """Register this module as ``sys.modules['torchdiffeq']`` so that the reference's
    ``import torchdiffeq`` (``models/blackbox_ode.py:3``) resolves to the B200 path unchanged."""
    import sys
    sys.modules["torchdiffeq"] = sys.modules[__name__]
