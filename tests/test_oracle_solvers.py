"""Self-validation of the torchdiffeq restatement (``oracle/torchdiffeq_oracle.py``).

torchdiffeq itself is absent and the reference pins nothing at this boundary, so the
restated solver loop is checked against mathematics: closed-form solutions, order of convergence,
scipy's Dormand-Prince pair, float64 gradcheck and the defining properties of odeint_adjoint.
"""
import math

import numpy as np
import pytest
import torch

from oracle import torchdiffeq_oracle as tde


class Linear(torch.nn.Module):
    """dy/dt = a(t) - d*y, the same affine-in-state structure as the blackbox RHS."""

    def __init__(self, d=0.7, dtype=torch.float64):
        super().__init__()
        self.d = torch.nn.Parameter(torch.tensor(d, dtype=dtype))

    def forward(self, t, y):
        return torch.cos(t) - self.d * y

    def exact(self, t, y0):
        d = self.d.detach()
        # y = C e^{-dt} + (d cos t + sin t)/(1+d^2)
        part = lambda s: (d * torch.cos(s) + torch.sin(s)) / (1 + d * d)  # noqa: E731
        return (y0 - part(t[0])) * torch.exp(-d * (t - t[0]))[:, None] + part(t)[:, None]


@pytest.mark.parametrize("method,order", [("euler", 1), ("midpoint", 2), ("rk4", 4)])
def test_fixed_grid_order_of_convergence(method, order):
    f = Linear()
    y0 = torch.tensor([[1.0, -0.5]], dtype=torch.float64).t().reshape(2, 1)
    errs = []
    for n in (20, 40, 80):
        t = torch.linspace(0.0, 2.0, n + 1, dtype=torch.float64)
        sol = tde.odeint(f, y0, t, method=method)
        assert sol.shape == (n + 1, 2, 1)
        assert torch.equal(sol[0], y0)
        exact_end = (y0[:, 0] - (f.d * math.cos(0.0) + math.sin(0.0)) / (1 + f.d ** 2)) * math.exp(-f.d.item() * 2.0) \
            + (f.d * math.cos(2.0) + math.sin(2.0)) / (1 + f.d ** 2)
        errs.append((sol[-1, :, 0] - exact_end.detach()).abs().max().item())
    p1 = math.log2(errs[0] / errs[1])
    p2 = math.log2(errs[1] / errs[2])
    assert abs(p1 - order) < 0.25 and abs(p2 - order) < 0.25, (errs, p1, p2)


def test_rk4_is_three_eighths_rule():
    """One step of torchdiffeq's rk4 on y' = y equals the 3/8-rule polynomial (same as classical RK4 for a
    linear autonomous problem) but differs from classical RK4 on a t-dependent problem."""
    t = torch.tensor([0.0, 0.5], dtype=torch.float64)
    y0 = torch.ones(1, 1, dtype=torch.float64)
    f = lambda tt, y: tt ** 3 + 0 * y  # noqa: E731  quadrature: exactness reveals the nodes
    got = tde.odeint(f, y0, t, method="rk4")[-1].item() - 1.0
    h = 0.5
    three_eighths = h / 8 * (0 + 3 * (h / 3) ** 3 + 3 * (2 * h / 3) ** 3 + h ** 3)
    assert abs(got - three_eighths) < 1e-15


def test_output_on_nonuniform_grid_and_reverse_time():
    f = Linear()
    t = torch.tensor([0.0, 0.1, 0.35, 0.4, 1.0], dtype=torch.float64)
    y0 = torch.tensor([[0.3]], dtype=torch.float64)
    sol = tde.odeint(f, y0, t, method="rk4")
    ex = f.exact(t, y0[:, 0])
    assert (sol[:, 0, :] - ex).abs().max() < 2e-4
    # decreasing t integrates backwards to the start value
    back = tde.odeint(f, sol[-1], t.flip(0), method="rk4")
    assert (back[-1] - y0).abs().max() < 5e-4


def test_step_size_option_interpolates_linearly():
    f = Linear()
    t = torch.tensor([0.0, 0.25, 1.0], dtype=torch.float64)
    y0 = torch.tensor([[0.3]], dtype=torch.float64)
    coarse = tde.odeint(f, y0, t, method="euler", options={"step_size": 0.5})
    g = tde.odeint(f, y0, torch.tensor([0.0, 0.5, 1.0], dtype=torch.float64), method="euler")
    assert torch.allclose(coarse[1], 0.5 * (g[0] + g[1]))
    assert torch.allclose(coarse[2], g[2])


@pytest.mark.parametrize("method", ["euler", "midpoint", "rk4"])
def test_fixed_grid_gradcheck(method):
    f = Linear()
    t = torch.linspace(0.0, 1.0, 6, dtype=torch.float64)
    y0 = torch.tensor([[0.3], [1.2]], dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(lambda y: tde.odeint(f, y, t, method=method), (y0,), eps=1e-6, atol=1e-8)


def test_dopri5_tolerance_scaling_and_closed_form():
    f = Linear()
    t = torch.linspace(0.0, 5.0, 11, dtype=torch.float64)
    y0 = torch.tensor([[0.3], [1.2]], dtype=torch.float64)
    ex = torch.stack([f.exact(t, y0[i])[:, 0] for i in range(2)], dim=1)[..., None]
    prev = None
    for tol in (1e-4, 1e-6, 1e-8):
        sol = tde.odeint(f, y0, t, method="dopri5", rtol=tol, atol=tol)
        err = (sol - ex).abs().max().item()
        assert err < 30 * tol, (tol, err)
        if prev is not None:
            assert err < prev
        prev = err
        assert tde.last_stats.n_accept > 0
        assert tde.last_stats.n_rhs == 2 + 6 * (tde.last_stats.n_accept + tde.last_stats.n_reject)


def test_dopri5_single_step_matches_scipy_rk45_tableau():
    """The 5th-order Dormand-Prince solution is the one scipy's RK45 uses: one attempted step from the same
    (t0, y0, h) must give the same y1 / f1.  torchdiffeq's embedded 4th-order weights (c_error, SURVEY.md
    section 8c) are the 1951/21600... variant, whose error estimate is exactly -2/3 of the textbook
    (5179/57600...) estimate scipy uses -- checked here as that ratio."""
    from scipy.integrate._ivp import rk as srk

    fnp = lambda tt, y: np.cos(tt) - 0.7 * y  # noqa: E731
    y0 = np.array([0.3, 1.2])
    h = 0.37
    f0 = fnp(0.0, y0)
    K = np.zeros((7, 2))
    y1, f1 = srk.rk_step(fnp, 0.0, y0, f0, h, srk.RK45.A, srk.RK45.B, srk.RK45.C, K)
    err = K.T @ srk.RK45.E * h
    f = Linear()
    ty0 = torch.tensor(y0)[:, None]
    dt = torch.tensor(h, dtype=torch.float64)
    t0 = torch.tensor(0.0, dtype=torch.float64)
    ty1, tf1, terr, _ = tde._dopri5_rk_step(f, ty0, f(t0, ty0), t0, dt, t0 + dt)
    assert np.allclose(ty1[:, 0].detach().numpy(), y1, rtol=0, atol=1e-15)
    assert np.allclose(tf1[:, 0].detach().numpy(), f1, rtol=0, atol=1e-15)
    assert np.allclose(terr[:, 0].detach().numpy(), -2.0 / 3.0 * err, rtol=1e-9, atol=1e-18)


def test_dopri5_controller_is_batch_global():
    """one step size for the whole batch -- adding a stiff-ish trajectory changes the accepted
    sequence seen by an easy one."""
    class F(torch.nn.Module):
        def forward(self, t, y):
            return -torch.tensor([[0.1], [25.0]], dtype=y.dtype)[: y.shape[0]] * y

    t = torch.tensor([0.0, 1.0], dtype=torch.float64)
    tde.odeint(F(), torch.ones(1, 1, dtype=torch.float64), t, method="dopri5", rtol=1e-6, atol=1e-8)
    n1 = tde.last_stats.n_accept + tde.last_stats.n_reject
    tde.odeint(F(), torch.ones(2, 1, dtype=torch.float64), t, method="dopri5", rtol=1e-6, atol=1e-8)
    n2 = tde.last_stats.n_accept + tde.last_stats.n_reject
    assert n2 > n1


def test_adjoint_semantics_params_only_and_restart():
    """odeint_adjoint: (i) plain-tensor attributes get no gradient (F5); (ii) its gradient is the continuous
    adjoint re-discretised -> differs from the discrete gradient by O(dt^2) for midpoint and converges to it."""
    torch.manual_seed(0)

    class F(torch.nn.Module):
        def __init__(self, z):
            super().__init__()
            self.w = torch.nn.Parameter(torch.tensor(0.8, dtype=torch.float64))
            self.constants = z

        def forward(self, t, y):
            return torch.sigmoid(self.w * t + self.constants) - torch.sigmoid(self.w) * y

    diffs = []
    for n in (8, 16, 32):
        t = torch.linspace(0.0, 2.0, n + 1, dtype=torch.float64)
        G = torch.cos(torch.arange(n + 1, dtype=torch.float64))[:, None, None] * (8.0 / n)
        grads = []
        for solve in (tde.odeint, tde.odeint_adjoint):
            z = torch.tensor([[0.2], [-0.4]], dtype=torch.float64, requires_grad=True)
            y0 = torch.tensor([[0.5], [0.1]], dtype=torch.float64, requires_grad=True)
            f = F(z)
            sol = solve(f, y0, t, method="midpoint")
            (sol * G).sum().backward()
            grads.append((y0.grad.clone(), f.w.grad.clone(), z.grad))
        assert grads[0][2] is not None and grads[1][2] is None
        diffs.append(max((grads[0][0] - grads[1][0]).abs().max().item(), (grads[0][1] - grads[1][1]).abs().item()))
    assert diffs[0] > 1e-6  # genuinely different gradients ...
    assert diffs[1] < diffs[0] / 2.5 and diffs[2] < diffs[1] / 2.5  # ... that agree as dt -> 0


def test_adjoint_forward_equals_odeint():
    f = Linear()
    t = torch.linspace(0.0, 1.0, 9, dtype=torch.float64)
    y0 = torch.tensor([[0.3], [1.2]], dtype=torch.float64)
    for m in ("euler", "midpoint", "rk4", "dopri5"):
        a = tde.odeint(f, y0, t, method=m, rtol=1e-6, atol=1e-8)
        b = tde.odeint_adjoint(f, y0, t, method=m, rtol=1e-6, atol=1e-8)
        assert torch.equal(a, b)


def test_invalid_inputs():
    f = Linear()
    y0 = torch.zeros(1, 1, dtype=torch.float64)
    with pytest.raises(ValueError):
        tde.odeint(f, y0, torch.tensor([0.0, 1.0, 0.5], dtype=torch.float64), method="rk4")
    with pytest.raises(ValueError):
        tde.odeint(f, y0, torch.tensor([0.0, 1.0], dtype=torch.float64), method="rk45")
    with pytest.raises(ValueError):
        tde.odeint_adjoint(lambda t, y: y, y0, torch.tensor([0.0, 1.0], dtype=torch.float64), method="rk4")


def test_dopri5_adjoint_gradient_and_mixed_norm():
    """odeint_adjoint with dopri5: (i) at tight tolerances its gradient is the true gradient (compared with autograd
    through a tight dopri5 solve: signs and structure of the augmented system); (ii) one fresh adaptive solve per
    output interval is logged; (iii) the step controller uses torchdiffeq's MIXED norm over the augmented tuple (the
    largest per-tensor RMS), not one RMS over the flattened state: a parameter adjoint whose error dominates must
    shrink the steps although it is a single element among many state components."""
    torch.manual_seed(0)

    class F(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.tensor([0.8, -0.3], dtype=torch.float64))

        def forward(self, t, y):
            return torch.sigmoid(self.w[0] * t + self.w[1]) - torch.sigmoid(self.w[0]) * y

    t = torch.linspace(0.0, 3.0, 7, dtype=torch.float64)
    G = torch.cos(torch.arange(7, dtype=torch.float64))[:, None, None]
    grads = []
    for solve in (tde.odeint, tde.odeint_adjoint):
        f = F()
        y0 = torch.tensor([[0.5], [0.1], [0.9]], dtype=torch.float64, requires_grad=True)
        sol = solve(f, y0, t, method="dopri5", rtol=1e-10, atol=1e-12)
        (sol * G).sum().backward()
        grads.append((y0.grad.clone(), f.w.grad.clone()))
    assert torch.allclose(grads[0][0], grads[1][0], rtol=1e-7, atol=1e-9)
    assert torch.allclose(grads[0][1], grads[1][1], rtol=1e-7, atol=1e-9)
    assert [i for i, _, _ in tde.last_adjoint_intervals] == [6, 5, 4, 3, 2, 1]
    assert all(sum(acc) >= 1 for _, acc, _ in tde.last_adjoint_intervals)
    # mixed norm: largest per-tensor RMS
    a, b = torch.tensor([3.0, 4.0]), torch.tensor([[10.0]])
    assert float(tde._mixed_norm((a, b))) == pytest.approx(10.0)
    assert float(tde._mixed_norm((a,))) == pytest.approx(math.sqrt(12.5))

    # a wide state (many well-resolved components) next to ONE stiff parameter adjoint: under a flattened RMS the
    # parameter's error would be diluted by 1/sqrt(n) and fewer steps taken
    class Wide(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.k = torch.nn.Parameter(torch.tensor(3.0, dtype=torch.float64))

        def forward(self, t, y):
            return -0.01 * y + 1e-3 * torch.sin(self.k * 40.0 * t)

    counts = {}
    for name in ("mixed", "flat"):
        f = Wide()
        y0 = torch.ones(64, 1, dtype=torch.float64, requires_grad=True)
        tt = torch.tensor([0.0, 1.0], dtype=torch.float64)
        if name == "flat":
            saved = tde._mixed_norm
            tde._mixed_norm = lambda ts: tde._rms_norm(torch.cat([x.reshape(-1) for x in ts]))
        try:
            sol = tde.odeint_adjoint(f, y0, tt, method="dopri5", rtol=1e-6, atol=1e-8)
            sol[-1].sum().backward()
        finally:
            if name == "flat":
                tde._mixed_norm = saved
        counts[name] = sum(len(acc) for _, acc, _ in tde.last_adjoint_intervals)
    assert counts["mixed"] > counts["flat"]
