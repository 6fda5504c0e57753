"""dopri5 with torchdiffeq's batch-global controller: identical accept/reject sequence, step sizes and trajectories
against the oracle (``oracle/torchdiffeq_oracle.py``) and the committed golden fixture produced by the reference's
real classes (``tests/golden/blackbox_golden.npz``, case ``chal``)."""
import os

import numpy as np
import pytest
import torch

import slode_testutil as U
from oracle import slode_port
from oracle import torchdiffeq_oracle as tde

pytestmark = pytest.mark.gpu


def _pair(shape, seed=12):
    o = U.make_oracle(shape, "dopri5", False, seed=seed)
    p = U.make_product(o)
    return o, p


def _run(o, p, z, rtol, atol, options=None, replay=None):
    import structured_latent_odes_b200 as slode
    from structured_latent_odes_b200 import torchdiffeq_api as api
    with torch.no_grad():
        want = tde.odeint(slode_port.OdeFunc(z, o.dynamics), o.latent_to_ode_net(z), o.times, method="dopri5",
                          rtol=rtol, atol=atol, options=options)
        acc_o, dts_o, nrhs_o = list(tde.last_stats.accepted), list(tde.last_stats.dts), tde.last_stats.n_rhs
        zc = z.cuda()
        opts = dict(options or {})
        opts["log_steps"] = True
        if replay is not None:
            opts["replay_steps"] = replay
        got = slode.odeint(p.gen_dynamics(zc), p.initialize_state(zc), p.times, method="dopri5", rtol=rtol, atol=atol,
                           options=opts)
    st = api.last_dopri5_stats
    return want, got, acc_o, dts_o, nrhs_o, st


def _oracle_steps(acc, dts, t0):
    """the oracle's (accepted, dt) lists in the step_log format (t0 column is informational)"""
    rows, t = [], float(t0)
    for a, d in zip(acc, dts):
        rows.append((t, d, 1.0 if a else 0.0))
        if a:
            t = t + d
    return torch.tensor(rows, dtype=torch.float64)


@pytest.mark.parametrize("shape,B,rtol,atol", [("cvs", 8, 1e-5, 1e-6), ("chal", 35, 1e-5, 1e-6), ("proc", 78, 1e-4, 1e-5),
                                               ("small", 300, 1e-6, 1e-7), ("cvs", 1, 1e-5, 1e-6)])
def test_replaying_the_oracle_step_sequence_gives_the_oracle_trajectories(shape, B, rtol, atol):
    """Stage arithmetic, FSAL, float64 time keeping and the dense output: with the oracle's own accept/reject
    decisions and step sizes prescribed, trajectories agree to fp32 rounding (1e-5)."""
    o, p = _pair(shape)
    L = U.SHAPES[shape][0]
    z = torch.randn(B, L, generator=torch.Generator().manual_seed(5))
    want, got_free, acc_o, dts_o, nrhs_o, st = _run(o, p, z, rtol, atol)
    # Hairer's initial step: two batch-wide norms and one extra RHS evaluation, bit-for-bit the oracle's value
    assert st.steps[0, 1].item() == pytest.approx(dts_o[0], rel=1e-6)
    # free-running controller: the fp32 error estimate is rounding-noise dominated at these tolerances (err ~1e-10
    # against terms ~1e-3), so step sizes drift apart; the solutions still agree to a few rtol
    assert U.rel_err(got_free, want) < 20 * rtol
    _, got, _, _, _, st2 = _run(o, p, z, rtol, atol, options=None, replay=_oracle_steps(acc_o, dts_o, o.times[0]))
    assert [bool(a) for a in st2.steps[:, 2]] == acc_o
    assert np.array_equal(st2.steps[:, 1].numpy(), np.array(dts_o))
    assert st2.n_accept == sum(acc_o) and st2.n_reject == len(acc_o) - sum(acc_o)
    assert U.rel_err(got, want) < 1e-5


@pytest.mark.parametrize("shape,B", [("cvs", 64), ("proc", 78)])
def test_controller_matches_the_oracle_where_the_error_estimate_is_above_rounding(shape, B):
    """At rtol=1e-3 / atol=1e-4 the embedded error estimate (~1e-4) is far above fp32 rounding (~1e-9): the
    batch-global controller must then reproduce the oracle's accept/reject sequence and step sizes."""
    o, p = _pair(shape)
    L = U.SHAPES[shape][0]
    z = torch.randn(B, L, generator=torch.Generator().manual_seed(6))
    want, got, acc_o, dts_o, nrhs_o, st = _run(o, p, z, 1e-3, 1e-4)
    steps = st.steps.numpy()
    assert steps.shape[0] == len(acc_o)
    assert [bool(a) for a in steps[:, 2]] == acc_o
    assert np.allclose(steps[:, 1], np.array(dts_o), rtol=5e-3)
    assert st.n_rhs == nrhs_o
    assert U.rel_err(got, want) < 1e-4


def test_golden_fixture_from_reference_classes(golden_dir):
    """tests/golden/blackbox_golden.npz case ``chal``: dopri5 output of the reference's REAL OdeModel / OdeFunc /
    Dynamics classes with its logged step sequence."""
    import structured_latent_odes_b200 as slode
    from structured_latent_odes_b200 import torchdiffeq_api as api
    g = np.load(os.path.join(golden_dir, "blackbox_golden.npz"))
    times = torch.from_numpy(g["chal/times"]).cuda()
    W = {k[len("chal/w/"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("chal/w/")}
    m = slode.OdeModel()
    m.init_with_params(times, 5, 15, 25, False, "dopri5", "cuda")
    m.load_state_dict(W, strict=False)
    m = m.cuda()
    z = torch.from_numpy(g["chal/z"]).cuda()
    replay = _oracle_steps(list(g["chal/dopri5/0/accepted"]), list(g["chal/dopri5/0/dts"]), 0.0)
    with torch.no_grad():
        free = slode.odeint(m.gen_dynamics(z), m.initialize_state(z), times, method="dopri5", rtol=1e-5, atol=1e-6,
                            options={"log_steps": True}).permute(1, 0, 2)
        first_dt = api.last_dopri5_stats.steps[0, 1].item()
        sol = slode.odeint(m.gen_dynamics(z), m.initialize_state(z), times, method="dopri5", rtol=1e-5, atol=1e-6,
                           options={"replay_steps": replay}).permute(1, 0, 2)
    assert first_dt == pytest.approx(float(g["chal/dopri5/0/dts"][0]), rel=1e-6)
    assert U.rel_err(sol, torch.from_numpy(g["chal/dopri5/0/sol"])) < 1e-5
    assert U.rel_err(free, torch.from_numpy(g["chal/dopri5/0/sol"])) < 2e-4


def test_first_step_option_and_determinism():
    o, p = _pair("cvs")
    z = torch.randn(16, 15, generator=torch.Generator().manual_seed(9))
    want, got, acc_o, dts_o, _, st = _run(o, p, z, 1e-5, 1e-6, options={"first_step": 0.05})
    assert st.steps[0, 1].item() == pytest.approx(0.05)
    assert U.rel_err(got, want) < 2e-4  # free-running controller at a noise-dominated tolerance, see above
    _, got2, _, _, _, st2 = _run(o, p, z, 1e-5, 1e-6, options={"first_step": 0.05})
    assert torch.equal(got, got2) and torch.equal(st.steps, st2.steps)  # bitwise reproducible step sequence


def test_controller_is_batch_global_on_device():
    """A large batch takes ONE step sequence (F6): every trajectory of the batch shares the accepted times, and the
    sequence differs from the one a sub-batch would take alone."""
    o, p = _pair("cvs")
    import structured_latent_odes_b200 as slode
    from structured_latent_odes_b200 import torchdiffeq_api as api
    g = torch.Generator().manual_seed(2)
    z = torch.randn(70000, 15, generator=g).cuda()   # more pairs than one wave of resident threads
    with torch.no_grad():
        full = slode.odeint(p.gen_dynamics(z), p.initialize_state(z), p.times, method="dopri5", rtol=1e-4, atol=1e-5,
                            options={"log_steps": True})
        steps_full = api.last_dopri5_stats.steps.clone()
        sub = slode.odeint(p.gen_dynamics(z[:7]), p.initialize_state(z[:7]), p.times, method="dopri5", rtol=1e-4,
                           atol=1e-5, options={"log_steps": True})
        steps_sub = api.last_dopri5_stats.steps.clone()
    assert steps_full.shape != steps_sub.shape or not torch.equal(steps_full, steps_sub)
    assert U.rel_err(full[:, :7], sub) < 1e-3  # both within tolerance of the true solution, not bit-equal
    # oracle spot check of the big batch is exact only with the same global controller: replay its step sizes
    zc = z[:64].cpu()
    with torch.no_grad():
        ref = tde.odeint(slode_port.OdeFunc(zc, o.dynamics), o.latent_to_ode_net(zc), o.times, method="dopri5",
                         rtol=1e-9, atol=1e-11)
    assert U.rel_err(full[:, :64], ref) < 5e-4


def test_errors():
    import structured_latent_odes_b200 as slode
    o, p = _pair("cvs")
    z = torch.randn(4, 15).cuda()
    f, y0 = p.gen_dynamics(z), p.initialize_state(z).detach()
    with torch.no_grad():
        with pytest.raises(Exception, match="max_num_steps"):
            slode.odeint(f, y0, p.times, method="dopri5", rtol=1e-5, atol=1e-6, options={"max_num_steps": 3})
        back = slode.odeint(f, y0, p.times.flip(0).contiguous(), method="dopri5", rtol=1e-5, atol=1e-6)
        assert torch.equal(back[0], y0) and bool(torch.isfinite(back).all())   # decreasing times are solved (see below)
        one = slode.odeint(f, y0, p.times[:1], method="dopri5", rtol=1e-5, atol=1e-6)
        assert torch.equal(one[0], y0)
    with pytest.raises(NotImplementedError, match="odeint_adjoint with dopri5"):
        slode.odeint_adjoint(f, y0.requires_grad_(True), p.times, method="dopri5", rtol=1e-5, atol=1e-6,
                             options={"first_step": 0.1})


@pytest.mark.parametrize("shape,B,rtol,atol", [("cvs", 40, 1e-5, 1e-6), ("proc", 30, 1e-4, 1e-5), ("small", 129, 1e-6, 1e-7)])
def test_gradients_equal_autograd_through_the_oracle_on_the_same_steps(shape, B, rtol, atol):
    """odeint(dopri5) + backward == autograd through the oracle's unrolled adaptive solve.  The oracle's step
    sequence is replayed on the device so both differentiate the same computation (see the replay test)."""
    L, H, S, times = U.SHAPES[shape]
    o, p = _pair(shape)
    g = torch.Generator().manual_seed(11)
    z = torch.randn(B, L, generator=g)
    G = torch.randn(B, len(times), S, generator=g)
    o.zero_grad()
    zo = z.clone().requires_grad_(True)
    sol_o = tde.odeint(slode_port.OdeFunc(zo, o.dynamics), o.latent_to_ode_net(zo), o.times, method="dopri5", rtol=rtol,
                       atol=atol).permute(1, 0, 2)
    replay = _oracle_steps(list(tde.last_stats.accepted), list(tde.last_stats.dts), o.times[0])
    (sol_o * G).sum().backward()
    import structured_latent_odes_b200 as slode
    p.zero_grad()
    zp = z.cuda().requires_grad_(True)
    sol_p = slode.odeint(p.gen_dynamics(zp), p.initialize_state(zp), p.times, method="dopri5", rtol=rtol, atol=atol,
                         options={"replay_steps": replay}).permute(1, 0, 2)
    (sol_p * G.cuda()).sum().backward()
    assert U.rel_err(sol_p, sol_o) < 1e-5
    assert U.rel_err(zp.grad, zo.grad) < 2e-5
    go = {k: v.grad for k, v in o.named_parameters() if v.grad is not None and ".prod." not in k and ".degr." not in k}
    gp = {k: v.grad for k, v in p.named_parameters() if v.grad is not None and ".prod." not in k and ".degr." not in k}
    assert set(go) == set(gp)
    for k in go:
        assert U.rel_err(gp[k], go[k]) < 2e-5, (k, U.rel_err(gp[k], go[k]))


def test_free_running_gradient_is_consistent():
    """Without replay the device picks its own steps; the gradient must still be the exact derivative of what it
    computed: directional derivative check against a central finite difference on the same (logged) steps."""
    import structured_latent_odes_b200 as slode
    from structured_latent_odes_b200 import torchdiffeq_api as api
    o, p = _pair("cvs")
    g = torch.Generator().manual_seed(3)
    z = torch.randn(32, 15, generator=g).cuda()
    G = torch.randn(len(p.times), 32, 5, generator=g).cuda()
    y0 = p.initialize_state(z).detach().requires_grad_(True)
    f = p.gen_dynamics(z)
    sol = slode.odeint(f, y0, p.times, method="dopri5", rtol=1e-5, atol=1e-6, options={"log_steps": True})
    steps = api.last_dopri5_stats.steps.clone()
    (sol * G).sum().backward()
    v = torch.randn(32, 5, generator=torch.Generator().manual_seed(4)).cuda()
    with torch.no_grad():  # the ODE is affine in the state: a finite difference on fixed steps is exact up to rounding
        a = slode.odeint(f, y0 + v, p.times, method="dopri5", rtol=1e-5, atol=1e-6, options={"replay_steps": steps})
        b = slode.odeint(f, y0 - v, p.times, method="dopri5", rtol=1e-5, atol=1e-6, options={"replay_steps": steps})
    fd = 0.5 * ((a - b).double() * G.double()).sum()
    an = (y0.grad.double() * v.double()).sum()
    assert abs(fd - an) < 1e-4 * max(abs(fd), 1.0)


@pytest.mark.parametrize("shape,B,split", [("cvs", 300, (120, 180)), ("chal", 35, (1, 34)), ("proc", 90, (30, 30, 30))])
def test_sharded_solve_takes_the_step_sequence_of_the_unsharded_one(shape, B, split):
    """the controller is batch-global, so a trajectory-sharded solve must combine the error norm
    over the shards at every attempted step.  Several shards are stepped in lockstep here on one device (the sum of
    their ``out`` sums plays the all-reduce): accepted / rejected decisions and step sizes must be those of the
    unsharded solve, and every shard's trajectories those rows of it."""
    from structured_latent_odes_b200 import torchdiffeq_api as api
    import structured_latent_odes_b200 as slode
    o, p = _pair(shape)
    L = U.SHAPES[shape][0]
    z = torch.randn(B, L, generator=torch.Generator().manual_seed(31)).cuda()
    rtol, atol = 1e-4, 1e-5
    with torch.no_grad():
        y0 = p.initialize_state(z).contiguous()
        full = slode.odeint(p.gen_dynamics(z), y0, p.times, method="dopri5", rtol=rtol, atol=atol,
                            options={"log_steps": True})
        steps_full = api.last_dopri5_stats.steps.clone()
        d = p.dynamics
        W1 = d.dynamics_hidden.weight.detach()
        c = torch.addmm(d.dynamics_hidden.bias.detach(), z, W1[:, 1:].t()).contiguous()
        w = [x.detach().contiguous() for x in (W1[:, 0], d.dyanamics_growth.weight, d.dyanamics_growth.bias,
                                               d.dyanmics_degradation.weight, d.dyanmics_degradation.bias)]
        shards, lo = [], 0
        for n in split:
            shards.append(api.Dopri5ShardSolve(y0[lo:lo + n].contiguous(), c[lo:lo + n].contiguous(), w, p.times, rtol,
                                               atol, n_global=B, log_cap=4096))
            lo += n
        for _ in range(100000):
            done = [sh.step() for sh in shards]
            assert all(done) or not any(done)      # every shard finishes at the same pass
            if all(done):
                break
            total = shards[0].out.clone()
            for sh in shards[1:]:
                total += sh.out
            for sh in shards:
                sh.ext.copy_(total)
    for sh in shards:
        n_acc, n_rej, _, status = sh.result()
        assert status == 0
        assert torch.equal(sh.log[: n_acc + n_rej].cpu(), steps_full)     # bit-identical step log
    got = torch.cat([sh.sol for sh in shards], dim=1)
    assert U.rel_err(got, full) < 1e-6


def test_sharded_options_through_odeint_with_gradients():
    """odeint(..., options=dopri5_shard_options(...)) on a single shard (world size 1) equals the plain solve, forward
    and backward: the pass-per-launch path and the persistent kernel run the same state machine."""
    import structured_latent_odes_b200 as slode
    from structured_latent_odes_b200 import sharding
    o, p = _pair("cvs")
    g = torch.Generator().manual_seed(32)
    z = torch.randn(50, 15, generator=g).cuda()
    G = torch.randn(86, 50, 5, generator=g).cuda()
    outs = []
    for opts in (None, sharding.dopri5_shard_options(50, device="cuda")):
        p.zero_grad()
        zz = z.clone().requires_grad_(True)
        sol = slode.odeint(p.gen_dynamics(zz), p.initialize_state(zz), p.times, method="dopri5", rtol=1e-4, atol=1e-5,
                           options=opts)
        (sol * G).sum().backward()
        outs.append((sol.detach(), zz.grad.clone(), p.dynamics.dynamics_hidden.weight.grad.clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)


@pytest.mark.parametrize("shape,B", [("cvs", 8), ("small", 50)])
def test_decreasing_output_times_run_torchdiffeqs_reversed_solve(shape, B):
    """t[0] > t[-1]: torchdiffeq solves s = -t with the negated right-hand side; the kernels take negative steps on the
    caller's time axis instead.  Same Hairer step, same accept/reject sequence where the error estimate is above
    rounding (rtol 1e-3), same trajectories and -- with the oracle's steps replayed -- the same gradients through
    odeint + autograd."""
    import structured_latent_odes_b200 as slode
    from structured_latent_odes_b200 import torchdiffeq_api as api
    o, p = _pair(shape)
    L = U.SHAPES[shape][0]
    z = torch.randn(B, L, generator=torch.Generator().manual_seed(21))
    t_dec = torch.flip(o.times, [0]).contiguous()
    rtol, atol = 1e-3, 1e-4
    # oracle: its log is on the reversed axis (s0 = -t0, ds > 0)
    zo = z.clone().requires_grad_(True)
    y0o = o.latent_to_ode_net(zo)
    want = tde.odeint(slode_port.OdeFunc(zo, o.dynamics), y0o, t_dec, method="dopri5", rtol=rtol, atol=atol)
    acc_o, dts_o, nrhs_o = list(tde.last_stats.accepted), list(tde.last_stats.dts), tde.last_stats.n_rhs
    G = torch.randn(want.shape, generator=torch.Generator().manual_seed(22))
    (want * G).sum().backward()
    zc = z.cuda()
    with torch.no_grad():
        free = slode.odeint(p.gen_dynamics(zc), p.initialize_state(zc), t_dec.cuda(), method="dopri5", rtol=rtol, atol=atol,
                            options={"log_steps": True})
    st = api.last_dopri5_stats
    steps = st.steps.numpy()
    assert steps.shape[0] == len(acc_o) and [bool(a) for a in steps[:, 2]] == acc_o
    assert np.all(steps[:, 1] < 0) and np.allclose(-steps[:, 1], np.array(dts_o), rtol=5e-3)
    assert steps[0, 1] == pytest.approx(-dts_o[0], rel=1e-6) and st.n_rhs == nrhs_o
    assert U.rel_err(free, want.detach()) < 1e-4
    # replay the oracle's sequence (signed for the caller's axis) with gradients
    rows, tcur = [], float(t_dec[0])
    for a, d in zip(acc_o, dts_o):
        rows.append((tcur, -d, 1.0 if a else 0.0))
        if a:
            tcur -= d
    replay = torch.tensor(rows, dtype=torch.float64)
    zp = zc.clone().requires_grad_(True)
    got = slode.odeint(p.gen_dynamics(zp), p.initialize_state(zp), t_dec.cuda(), method="dopri5", rtol=rtol, atol=atol,
                       options={"replay_steps": replay})
    (got * G.cuda()).sum().backward()
    assert U.rel_err(got.detach(), want.detach()) < 1e-5
    assert U.rel_err(zp.grad, zo.grad) < 2e-5
    go = dict(o.named_parameters())
    for k, v in p.named_parameters():
        if ".prod." in k or ".degr." in k or v.grad is None:
            continue
        assert U.rel_err(v.grad, go[k].grad) < 2e-5, k
    # odeint_adjoint keeps refusing the reversed axis (stated limitation)
    with pytest.raises(NotImplementedError):
        slode.odeint_adjoint(p.gen_dynamics(zp), p.initialize_state(zp), t_dec.cuda(), method="dopri5", rtol=rtol, atol=atol)
