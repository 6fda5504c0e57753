"""Parity of the sm_100a path against the CPU oracle (same seeded inputs), the committed golden fixtures, and --
at BASELINE.json's full size -- size-independent properties.  Tolerance (north_star): 1e-5 relative, fp32,
fixed step; relative = max|a-b| / max|b| over the tensor.  Every call goes through the C ABI
(``csrc/libslode_b200.so``); the oracle is only the checker."""
import os

import numpy as np
import pytest
import torch

import slode_testutil as U

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _cuda():
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")


def _compare(shape, method, adjoint, B, layout="tbs", seed=3):
    _cuda()
    L, H, S, times = U.SHAPES[shape]
    o = U.make_oracle(shape, method, adjoint)
    p = U.make_product(o, layout=layout)
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(B, L, generator=g)
    G = torch.randn(B, len(times), S, generator=g)
    so, gzo, gro = U.run_fwd_bwd(o, z, G)
    sp, gzp, grp = U.run_fwd_bwd(p, z.cuda(), G.cuda())
    assert sp.shape == so.shape == (B, len(times), S)
    assert U.rel_err(sp, so) < TOL
    assert U.rel_err(gzp, gzo) < TOL
    assert set(grp) == set(gro)
    for k in gro:
        assert U.rel_err(grp[k], gro[k]) < TOL, (k, U.rel_err(grp[k], gro[k]))


@pytest.mark.parametrize("adjoint", [False, True])
@pytest.mark.parametrize("method", ["euler", "midpoint", "rk4"])
@pytest.mark.parametrize("shape", ["cvs", "proc", "small", "h32", "h64"])
def test_fixed_grid_matches_oracle(shape, method, adjoint):
    _compare(shape, method, adjoint, B=200)


def test_long_horizon_many_relu_crossings():
    """Every hidden unit crosses zero somewhere on a long, non-uniform grid that starts at a negative time: the
    piecewise-linear evaluator must follow all of them (increasing and decreasing grids)."""
    _cuda()
    for sign in (1.0, -1.0):
        for method, adjoint in (("rk4", False), ("midpoint", True), ("euler", False)):
            o = U.make_oracle("cvs", method, adjoint)
            g = torch.Generator().manual_seed(11)
            t = sign * (torch.cumsum(torch.rand(150, generator=g) * 0.8 + 0.05, 0) - 20.0)
            o.times = t
            p = U.make_product(o)
            z = 2.0 * torch.randn(130, 15, generator=g)
            G = torch.randn(130, 150, 5, generator=g)
            so, gzo, gro = U.run_fwd_bwd(o, z, G)
            sp, gzp, grp = U.run_fwd_bwd(p, z.cuda(), G.cuda())
            assert U.rel_err(sp, so) < TOL
            assert U.rel_err(gzp, gzo) < 2 * TOL
            for k in gro:
                assert U.rel_err(grp[k], gro[k]) < 2 * TOL, (sign, method, k, U.rel_err(grp[k], gro[k]))


@pytest.mark.parametrize("shape", ["cvs", "h32", "h64"])
def test_many_crossings_per_evaluation(shape):
    """A short grid around t = 0 with small |c|: most hidden units cross zero within a few steps, several per
    evaluation and trajectory."""
    _cuda()
    L, H, S, _ = U.SHAPES[shape]
    for method, adjoint in (("midpoint", False), ("midpoint", True), ("rk4", False), ("euler", True)):
        o = U.make_oracle(shape, method, adjoint)
        o.times = torch.linspace(-3.0, 3.0, 7)
        p = U.make_product(o)
        g = torch.Generator().manual_seed(5)
        z = 0.3 * torch.randn(70, L, generator=g)
        G = torch.randn(70, 7, S, generator=g)
        so, gzo, gro = U.run_fwd_bwd(o, z, G)
        sp, gzp, grp = U.run_fwd_bwd(p, z.cuda(), G.cuda())
        assert U.rel_err(sp, so) < TOL
        if gzo is not None:
            assert U.rel_err(gzp, gzo) < 2 * TOL
        for k in gro:
            assert U.rel_err(grp[k], gro[k]) < 2 * TOL, (method, adjoint, k, U.rel_err(grp[k], gro[k]))


@pytest.mark.parametrize("adjoint", [False, True])
@pytest.mark.parametrize("method", ["euler", "midpoint", "rk4"])
@pytest.mark.parametrize("shape", ["h128", "h256", "h512"])
def test_wide_hidden_layers_match_oracle(shape, method, adjoint):
    """BASELINE configs[4] (hidden 32-512): above 64 units the per-trajectory tables (c, sorted crossing keys, unit
    status) live in caller-provided global scratch and the unit loops are rolled."""
    _compare(shape, method, adjoint, B=150)


@pytest.mark.parametrize("B", [1, 2, 127, 128, 129, 1000])
def test_ragged_batch_sizes(B):
    _compare("cvs", "rk4", False, B)
    _compare("cvs", "midpoint", True, B)


def test_challenge_length_and_bts_layout():
    _compare("chal", "midpoint", True, B=35)
    _compare("chal", "rk4", False, B=35, layout="bts")


def test_golden_fixtures(golden_dir):
    """Outputs of the reference's real classes (tests/golden/make_golden.py)."""
    _cuda()
    import structured_latent_odes_b200 as slode
    g = np.load(os.path.join(golden_dir, "blackbox_golden.npz"))
    for name, methods in (("cvs", ("euler", "midpoint", "rk4")), ("proc", ("midpoint", "rk4"))):
        times = torch.from_numpy(g[f"{name}/times"]).cuda()
        W = {k[len(name) + 3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(f"{name}/w/")}
        H, Lp1 = W["dynamics.dynamics_hidden.weight"].shape
        S = W["dynamics.dyanamics_growth.weight"].shape[0]
        for method in methods:
            for adj in (0, 1):
                m = slode.OdeModel()
                m.init_with_params(times, S, Lp1 - 1, H, bool(adj), method, "cuda")
                m.load_state_dict(W, strict=False)
                m = m.cuda()
                sol, gz, gr = U.run_fwd_bwd(m, torch.from_numpy(g[f"{name}/z"]).cuda(), torch.from_numpy(g[f"{name}/G"]).cuda())
                key = f"{name}/{method}/{adj}"
                assert U.rel_err(sol, torch.from_numpy(g[f"{key}/sol"])) < TOL
                assert U.rel_err(gz, torch.from_numpy(g[f"{key}/grad_z"])) < TOL
                for k, v in gr.items():
                    assert U.rel_err(v, torch.from_numpy(g[f"{key}/g/{k}"])) < TOL, (key, k)


def test_edge_cases_empty_single_time_and_reverse_time():
    _cuda()
    import structured_latent_odes_b200 as slode
    from oracle import torchdiffeq_oracle as tde
    o = U.make_oracle("cvs", "rk4", False)
    p = U.make_product(o)
    # empty batch
    sol = p.solve_ODE(torch.zeros(0, 15, device="cuda"))
    assert sol.shape == (0, 86, 5)
    # a single output time returns y0
    f = p.gen_dynamics(torch.randn(3, 15, device="cuda"))
    y0 = torch.rand(3, 5, device="cuda")
    one = slode.odeint(f, y0, torch.tensor([0.5], device="cuda"), method="rk4")
    assert one.shape == (1, 3, 5) and torch.equal(one[0], y0)
    # decreasing times (torchdiffeq integrates -t with the negated RHS)
    t = torch.linspace(3.0, 0.0, 13)
    z = torch.randn(5, 15)
    fo = o.dynamics
    from oracle import slode_port
    want = tde.odeint(slode_port.OdeFunc(z, fo), y0[:1].cpu().expand(5, 5), t, method="midpoint")
    got = slode.odeint(p.gen_dynamics(z.cuda()), y0[:1].expand(5, 5).contiguous(), t.cuda(), method="midpoint")
    assert U.rel_err(got, want) < TOL
    # t may live on the CPU (proc's data.times is never moved to the device)
    got2 = slode.odeint(p.gen_dynamics(z.cuda()), y0[:1].expand(5, 5).contiguous(), t, method="midpoint")
    assert torch.equal(got, got2)


def test_strided_upstream_gradient_and_double_backward_free():
    """The decoder hands back a permuted, non-contiguous grad (sol.permute(1,0,2) @ W^T): no copy is required,
    and the result equals the contiguous case."""
    _cuda()
    o = U.make_oracle("cvs", "midpoint", False)
    p = U.make_product(o)
    z = torch.randn(64, 15, device="cuda", requires_grad=True)
    Wq = torch.randn(3, 5, device="cuda")
    sol = p.solve_ODE(z)
    mu = (sol @ Wq.t()).permute(0, 2, 1)
    mu.square().sum().backward()
    g1 = z.grad.clone()
    zc = z.detach().clone().cpu().requires_grad_(True)
    (o.solve_ODE(zc) @ Wq.cpu().t()).permute(0, 2, 1).square().sum().backward()
    assert U.rel_err(g1, zc.grad) < TOL


def test_errors_on_device():
    _cuda()
    import structured_latent_odes_b200 as slode
    o = U.make_oracle("cvs", "rk4", False)
    p = U.make_product(o)
    f = p.gen_dynamics(torch.randn(4, 15, device="cuda"))
    y0 = torch.rand(4, 5, device="cuda")
    t = torch.arange(0.0, 5.0, device="cuda")
    with pytest.raises(TypeError):
        slode.odeint(f, y0.double(), t, method="rk4")
    with pytest.raises(ValueError, match="strictly"):
        slode.odeint(f, y0, torch.tensor([0.0, 1.0, 1.0], device="cuda"), method="rk4")
    with pytest.raises(ValueError, match="batch"):
        slode.odeint(f, y0[:3], t, method="rk4")
    m = slode.OdeModel()
    m.init_with_params(t, 3, 15, 7, False, "rk4", "cuda")
    with pytest.raises(NotImplementedError, match="no compiled kernel"):
        m.cuda().solve_ODE(torch.randn(4, 15, device="cuda"))
    # weights the kernels would read through raw pointers: wrong device / dtype must raise BEFORE any launch
    # (an illegal address would be sticky and kill the context for every later test)
    cpu_model = slode.OdeModel()
    cpu_model.init_with_params(t, 5, 15, 25, False, "rk4", "cuda")   # parameters left on the CPU
    with pytest.raises(RuntimeError, match="is on cpu"):
        cpu_model.solve_ODE(torch.randn(4, 15, device="cuda"))
    with pytest.raises(RuntimeError, match="is on cpu"):
        slode.odeint(cpu_model.gen_dynamics(torch.randn(4, 15, device="cuda")), y0, t, method="rk4")
    dbl = slode.OdeModel()
    dbl.init_with_params(t, 5, 15, 25, False, "midpoint", "cuda")
    dbl = dbl.cuda().double()
    with pytest.raises(TypeError, match="float64"):
        dbl.solve_ODE(torch.randn(4, 15, device="cuda"))
    with pytest.raises(TypeError, match="float64"):
        slode.odeint_adjoint(dbl.gen_dynamics(torch.randn(4, 15, device="cuda")), y0, t, method="midpoint")
    torch.cuda.synchronize()  # the context is still healthy
    assert p.solve_ODE(torch.randn(4, 15, device="cuda")).shape == (4, 86, 5)


# ------------------------------------------------------------------------------------------------------------
# full-size properties (BASELINE.json configs[1]: 2^20 trajectories x 100 times, rk4)
# ------------------------------------------------------------------------------------------------------------
def _full_model(method="rk4", adjoint=False, T=100):
    import structured_latent_odes_b200 as slode
    torch.manual_seed(12)
    m = slode.OdeModel()
    m.init_with_params(torch.arange(0.0, T, 1.0, device="cuda"), 5, 15, 25, adjoint, method, "cuda")
    return m.cuda()


def test_full_size_affine_in_the_initial_state():
    """f = A(t,z) - D(t,z) x is affine in x, so sol(a) + sol(b) - sol(c) == sol(a + b - c) up to rounding and
    <grad_y0, v> is the exact directional derivative; checked on 2^20 trajectories."""
    _cuda()
    import structured_latent_odes_b200 as slode
    B = 1 << 20
    m = _full_model()
    g = torch.Generator(device="cuda").manual_seed(12)
    z = torch.randn(B, 15, device="cuda", generator=g)
    f = m.gen_dynamics(z)
    a, b, c = (torch.rand(B, 5, device="cuda", generator=g) for _ in range(3))
    solve = lambda y: slode.odeint(f, y, m.times, method="rk4")  # noqa: E731
    lhs = solve(a) + solve(b) - solve(c)
    rhs = solve(a + b - c)
    assert U.rel_err(lhs, rhs) < TOL
    # directional derivative through the reverse sweep
    G = torch.randn(100, B, 5, device="cuda", generator=g)
    y0 = a.clone().requires_grad_(True)
    (solve(y0) * G).sum().backward()
    v = b - c
    exact = ((solve(a + v) - solve(a)).double() * G.double()).sum(dim=(0, 2))   # per trajectory
    pred = (y0.grad.double() * v.double()).sum(dim=1)
    assert (exact - pred).abs().max().item() < 2e-4 * max(exact.abs().max().item(), 1.0)


def test_full_size_batch_permutation_and_subset_consistency():
    """Trajectories are independent: solving a random subset alone gives bit-identical rows, and parameter
    gradients of the full batch equal the sum over two halves (the multi-GPU sharding identity)."""
    _cuda()
    B = 1 << 20
    m = _full_model("midpoint", adjoint=False, T=86)
    g = torch.Generator(device="cuda").manual_seed(1)
    z = torch.randn(B, 15, device="cuda", generator=g)
    full = m.solve_ODE(z)
    idx = torch.randperm(B, device="cuda", generator=g)[:4099]
    assert torch.equal(m.solve_ODE(z[idx]), full[idx])
    # oracle spot check of 64 rows of the big batch
    o = U.make_oracle("cvs", "midpoint", False)
    o.load_state_dict({k: v.cpu() for k, v in m.state_dict().items()})
    assert U.rel_err(full[idx[:64]], o.solve_ODE(z[idx[:64]].cpu())) < TOL
    G = torch.randn(B, 86, 5, device="cuda", generator=g)

    def grads(sl):
        m.zero_grad()
        (m.solve_ODE(z[sl]) * G[sl]).sum().backward()
        return torch.cat([p.grad.reshape(-1) for k, p in m.named_parameters() if ".prod." not in k and ".degr." not in k])

    whole = grads(slice(0, B))
    halves = grads(slice(0, B // 2)) + grads(slice(B // 2, B))
    assert U.rel_err(halves, whole) < 5e-5  # fp32 sums of 1e6 terms in a different order


def test_full_size_exact_bench_step_oracle_spot_check():
    """The EXACT step bench.py times (configs[1]: rk4, T=100, L15/H25/S5, seed-12 reference-init weights, 2^20
    trajectories z ~ N(0,1) from Generator(seed 12), upstream gradient G from the same generator): 512 rows of the
    full-size solve and of the full-size reverse sweep's grad_z against the CPU oracle, and the parameter gradients
    of those rows solved alone (the sweep is row-independent; parameter gradients are sums over rows)."""
    _cuda()
    B, T = 1 << 20, 100
    m = _full_model("rk4", adjoint=False, T=T)
    g = torch.Generator(device="cuda").manual_seed(12)
    z = torch.randn(B, 15, device="cuda", generator=g).requires_grad_(True)
    G = torch.randn(T, B, 5, device="cuda", generator=g).permute(1, 0, 2)   # bench.py's resident layout ("tbs")
    sol = m.solve_ODE(z)
    sol.backward(G)
    rows = torch.cat([torch.arange(0, 256, device="cuda"),                      # first tile
                      torch.randint(0, B, (192,), device="cuda", generator=g),  # anywhere
                      torch.arange(B - 64, B, device="cuda")])                  # last tile
    o = U.make_oracle("cvs", "rk4", False)
    o.times = torch.arange(0.0, T, 1.0)
    o.load_state_dict({k: v.cpu() for k, v in m.state_dict().items()})
    zo = z.detach()[rows].cpu().requires_grad_(True)
    so = o.solve_ODE(zo)
    (so * G[rows].cpu()).sum().backward()
    assert U.rel_err(sol.detach()[rows], so) < TOL
    assert U.rel_err(z.grad[rows], zo.grad) < TOL
    # parameter gradients: the same rows through the product alone
    m.zero_grad()
    zs = z.detach()[rows].clone().requires_grad_(True)
    (m.solve_ODE(zs) * G[rows]).sum().backward()
    gro = {k: p.grad for k, p in o.named_parameters() if p.grad is not None and ".prod." not in k and ".degr." not in k}
    grp = {k: p.grad for k, p in m.named_parameters() if p.grad is not None and ".prod." not in k and ".degr." not in k}
    assert set(gro) == set(grp)
    for k in gro:
        assert U.rel_err(grp[k], gro[k]) < TOL, (k, U.rel_err(grp[k], gro[k]))


@pytest.mark.parametrize("method,adjoint", [("rk4", False), ("midpoint", True)])
def test_non_finite_latents_propagate_like_torch(method, adjoint):
    """A NaN / inf latent row (an encoder whose exp-scale overflowed) must surface as non-finite trajectories of THAT
    row, as torch's relu / sigmoid give in the reference -- not be laundered into finite numbers by fmin / fmax
    clamps -- and must leave every other row untouched."""
    o = U.make_oracle("cvs", method, adjoint)
    p = U.make_product(o)
    g = torch.Generator().manual_seed(3)
    z = torch.randn(70, 15, generator=g)
    clean = p.solve_ODE(z.cuda()).detach()
    z_bad = z.clone()
    z_bad[5, 3] = float("nan")
    z_bad[40, 0] = float("inf")
    with torch.no_grad():
        want = o.solve_ODE(z_bad)
        got = p.solve_ODE(z_bad.cuda()).cpu()
    for row in (5, 40):
        assert not torch.isfinite(want[row]).all()          # the reference's arithmetic does not hide it ...
        assert not torch.isfinite(got[row]).all()           # ... and neither do the kernels
    keep = [i for i in range(70) if i not in (5, 40)]
    assert torch.equal(got[keep], clean.cpu()[keep])
