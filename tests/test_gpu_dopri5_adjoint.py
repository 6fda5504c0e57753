"""``odeint_adjoint(..., method="dopri5")`` -- the mode every shipped config reaches when ``solver`` is set to
dopri5 (``adjoint_solver=True``; models/blackbox_ode.py:40-42): torchdiffeq's adaptive backward solve of the
augmented system, one fresh solve per output interval, mixed error norm over [y, a, a_theta].

Checked against ``oracle/torchdiffeq_oracle.py::odeint_adjoint`` on the same seeded inputs: the per-interval
accept / reject sequences and step sizes where the error estimate is above fp32 rounding, and the gradients."""
import numpy as np
import pytest
import torch

import slode_testutil as U
from oracle import slode_port
from oracle import torchdiffeq_oracle as tde

pytestmark = pytest.mark.gpu


def _oracle_adjoint(o, z, G, rtol, atol, dtype=torch.float32):
    od = U.make_oracle_like(o, dtype) if dtype != torch.float32 else o
    od.zero_grad()
    zo = z.to(dtype)
    y0 = od.latent_to_ode_net(zo).detach().requires_grad_(True)
    f = slode_port.OdeFunc(zo, od.dynamics)
    sol = tde.odeint_adjoint(f, y0, od.times, method="dopri5", rtol=rtol, atol=atol)
    (sol * G.to(dtype)).sum().backward()
    grads = {k: v.grad.detach().clone() for k, v in od.dynamics.named_parameters()
             if v.grad is not None and k.split(".")[0] in ("dynamics_hidden", "dyanamics_growth", "dyanmics_degradation")}
    return sol.detach(), y0.grad.detach().clone(), grads, [(i, list(a), list(d)) for i, a, d in tde.last_adjoint_intervals]


def _device_adjoint(p, z, G, rtol, atol, log=True):
    import structured_latent_odes_b200 as slode
    from structured_latent_odes_b200 import torchdiffeq_api as api
    p.zero_grad()
    zc = z.cuda()
    y0 = p.initialize_state(zc).detach().requires_grad_(True)
    f = p.gen_dynamics(zc)
    sol = slode.odeint_adjoint(f, y0, p.times, method="dopri5", rtol=rtol, atol=atol,
                               options={"log_steps": True} if log else None)
    (sol * G.cuda()).sum().backward()
    grads = {k: v.grad.detach().clone() for k, v in p.dynamics.named_parameters()
             if v.grad is not None and k.split(".")[0] in ("dynamics_hidden", "dyanamics_growth", "dyanmics_degradation")}
    return sol.detach(), y0.grad.detach().clone(), grads, api.last_dopri5_adjoint_stats


def _split(steps):
    """device step log (interval, s0, ds, accepted) -> [(interval, accepted[], ds[])] in solve order"""
    out = []
    for iv, _, ds, acc in steps.tolist():
        if not out or out[-1][0] != int(iv):
            out.append((int(iv), [], []))
        out[-1][1].append(bool(acc))
        out[-1][2].append(ds)
    return out


def _replay_rows(iv_o):
    """the oracle's per-interval (accepted, dt) lists in the device's backward step-log format"""
    rows = []
    for i, acc, dts in iv_o:
        for a, d in zip(acc, dts):
            rows.append((float(i), 0.0, d, 1.0 if a else 0.0))
    return torch.tensor(rows, dtype=torch.float64)


@pytest.mark.parametrize("shape,B", [("cvs", 8), ("chal", 35), ("proc", 40), ("small", 200), ("cvs", 1), ("h64", 5)])
def test_gradients_match_the_oracle_on_the_oracle_step_sequence(shape, B):
    """Stage arithmetic of the augmented system, the four parameter-adjoint combinations, FSAL, float64 time keeping,
    the dense output at the interval ends and the restart from the stored forward states: with the oracle's backward
    step sizes and decisions prescribed, the gradients agree to fp32 rounding.  Free-running, every interval starts
    from Hairer's initial step (two mixed norms over the whole augmented state), which must equal the oracle's."""
    L, H, S, times = U.SHAPES[shape]
    o = U.make_oracle(shape, "dopri5", True)
    p = U.make_product(o)
    g = torch.Generator().manual_seed(21)
    z = torch.randn(B, L, generator=g)
    G = torch.randn(len(times), B, S, generator=g)
    rtol, atol = 1e-3, 1e-4
    sol_o, gy_o, gr_o, iv_o = _oracle_adjoint(o, z, G, rtol, atol)
    # free-running controller
    sol_p, gy_p, gr_p, st = _device_adjoint(p, z, G, rtol, atol)
    iv_p = _split(st.steps)
    assert [i for i, _, _ in iv_p] == [i for i, _, _ in iv_o] == list(range(len(times) - 1, 0, -1))
    # (d2 of the initial-step rule is a difference quotient of two nearby right-hand sides over a tiny probe step,
    # scaled by 1 / (atol + rtol |a_theta|): while the parameter adjoints are still small, the fp32 summation noise of
    # the two batch sums it subtracts is amplified by 1 / (atol h0) and moves h1 = (0.01 / max(d1, d2))^(1/5) by up to
    # ~15 % in single intervals; the typical interval agrees to 1e-3)
    for (_, _, dp), (_, _, do) in zip(iv_p, iv_o):
        assert dp[0] == pytest.approx(do[0], rel=0.5)
    assert np.median([abs(dp[0] / do[0] - 1) for (_, _, dp), (_, _, do) in zip(iv_p, iv_o)]) < 1e-3
    free = max([U.rel_err(gy_p, gy_o)] + [U.rel_err(gr_p[k], gr_o[k]) for k in gr_o])
    # After that small first step the embedded error estimate is ~1e-5 of the tolerance -- the fp32 rounding level of
    # its own terms -- and the growth factor 0.9 / ratio^(1/5) carries that noise (second step equal to 0.2-1 %, the
    # third, which overshoots the interval and is only interpolated, to tens of percent; rarely an extra rejection),
    # so later steps are not compared; the gradients of the two step sequences agree within rtol (observed 0.05-0.4 rtol).
    assert free < rtol, free
    # the oracle's own backward step sequence replayed
    import structured_latent_odes_b200 as slode
    from structured_latent_odes_b200 import torchdiffeq_api as api
    p.zero_grad()
    zc = z.cuda()
    y0 = p.initialize_state(zc).detach().requires_grad_(True)
    sol = slode.odeint_adjoint(p.gen_dynamics(zc), y0, p.times, method="dopri5", rtol=rtol, atol=atol,
                               options={"log_steps": True, "adjoint_replay_steps": _replay_rows(iv_o)})
    (sol * G.cuda()).sum().backward()
    st = api.last_dopri5_adjoint_stats
    iv_r = _split(st.steps)
    assert [a for _, a, _ in iv_r] == [a for _, a, _ in iv_o]
    assert st.n_accept == sum(sum(a) for _, a, _ in iv_o)
    # sol itself comes from the free-running forward solve (its own tests: test_gpu_dopri5.py)
    assert U.rel_err(sol.permute(1, 0, 2), sol_o.permute(1, 0, 2)) < 1e-4
    assert U.rel_err(y0.grad, gy_o) < 5e-5, U.rel_err(y0.grad, gy_o)   # the oracle itself runs in fp32 here
    gr = {k: v.grad for k, v in p.dynamics.named_parameters()
          if v.grad is not None and k.split(".")[0] in ("dynamics_hidden", "dyanamics_growth", "dyanmics_degradation")}
    assert set(gr) == set(gr_o) and len(gr_o) == 6
    for k in gr_o:
        assert U.rel_err(gr[k], gr_o[k]) < 5e-5, (k, U.rel_err(gr[k], gr_o[k]))


@pytest.mark.parametrize("shape,B", [("chal", 35), ("cvs", 130)])
def test_default_tolerances_against_the_float64_oracle(shape, B):
    """torchdiffeq's defaults (rtol 1e-7, atol 1e-9 -- what the reference's call passes): in fp32 the controller is
    driven by rounding noise there, so step sequences are not comparable; the gradients must still be those of the
    float64 oracle at the same tolerances."""
    L, H, S, times = U.SHAPES[shape]
    o = U.make_oracle(shape, "dopri5", True)
    p = U.make_product(o)
    g = torch.Generator().manual_seed(22)
    z = torch.randn(B, L, generator=g)
    G = torch.randn(len(times), B, S, generator=g)
    _, gy_o, gr_o, _ = _oracle_adjoint(o, z, G, 1e-7, 1e-9, dtype=torch.float64)
    _, gy_p, gr_p, st = _device_adjoint(p, z, G, 1e-7, 1e-9)
    assert st.n_accept >= len(times) - 1
    # (observed 1e-5 .. 5.5e-5 depending on the order the parameter adjoints are summed in: at these tolerances the
    # fp32 controller's decisions are rounding noise, and the accepted steps differ run configuration to run configuration)
    assert U.rel_err(gy_p, gy_o) < 5e-5
    for k in gr_o:
        assert U.rel_err(gr_p[k], gr_o[k]) < 1e-4, (k, U.rel_err(gr_p[k], gr_o[k]))


def test_reference_model_path_constants_get_no_gradient_and_runs_are_bitwise_reproducible():
    """Through OdeModel.solve_ODE as the reference calls it (adjoint_solver=True, solver='dopri5'): z gets its
    gradient only through latent_to_ode_net, every parameter of both nets gets one, and two runs give
    bit-identical step logs and gradients (fixed-order reductions, no atomics)."""
    from structured_latent_odes_b200 import torchdiffeq_api as api
    o = U.make_oracle("cvs", "dopri5", True)
    p = U.make_product(o)
    g = torch.Generator().manual_seed(23)
    z = torch.randn(300, 15, generator=g)
    G = torch.randn(300, 86, 5, generator=g)
    runs = []
    for _ in range(2):
        sol, gz, grads = U.run_fwd_bwd(p, z.cuda(), G.cuda())
        runs.append((sol, gz, grads))
    assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1])
    for k in runs[0][2]:
        assert torch.equal(runs[0][2][k], runs[1][2][k]), k
    assert any(k.startswith("latent_to_ode_net") for k in runs[0][2]) and any(k.startswith("dynamics") for k in runs[0][2])
    # oracle with the same semantics (rtol/atol defaults of the reference's call)
    o.zero_grad()
    sol_o, gz_o, gr_o = U.run_fwd_bwd(U.make_oracle_like(o, torch.float64), z.double(), G.double())
    assert U.rel_err(runs[0][0], sol_o) < 1e-4   # free-running fp32 forward over 85 time units, values up to ~60
    assert U.rel_err(runs[0][1], gz_o) < 1e-4, U.rel_err(runs[0][1], gz_o)
    for k in gr_o:
        assert U.rel_err(runs[0][2][k], gr_o[k]) < 1e-4, (k, U.rel_err(runs[0][2][k], gr_o[k]))


def test_raw_c_abi_and_errors():
    import structured_latent_odes_b200 as slode
    from structured_latent_odes_b200 import _cabi
    lib = _cabi.lib()
    assert lib.slode_mlp_dopri5_adjoint_workspace_bytes(0, 15, 25, 5) == 0
    assert lib.slode_mlp_dopri5_adjoint_workspace_bytes(100, 15, 25, 5) > 0
    assert lib.slode_mlp_dopri5_adjoint_workspace_bytes(100, 15, 25, 7) == -1     # state dim not compiled
    assert b"ode_state_dim" in lib.slode_last_error()
    o = U.make_oracle("cvs", "dopri5", True)
    p = U.make_product(o)
    z = torch.randn(4, 15).cuda()
    y0 = p.initialize_state(z).detach().requires_grad_(True)
    # max_attempts of the backward solve (raw call on the tensors of a forward solve)
    sol = slode.odeint_adjoint(p.gen_dynamics(z), y0, p.times, method="dopri5", rtol=1e-3, atol=1e-4)
    d = p.dynamics
    W1 = d.dynamics_hidden.weight.detach().contiguous()
    c = torch.addmm(d.dynamics_hidden.bias.detach(), z, W1[:, 1:].t()).contiguous()
    w = [x.detach().contiguous() for x in (d.dyanamics_growth.weight, d.dyanamics_growth.bias,
                                           d.dyanmics_degradation.weight, d.dyanmics_degradation.bias)]
    T, B, S = sol.shape
    gs = torch.ones_like(sol)
    n = lib.slode_mlp_dopri5_adjoint_workspace_bytes(B, 15, 25, S)
    ws = torch.empty(n, dtype=torch.uint8, device="cuda")
    gy0, gp = torch.empty(B, S, device="cuda"), torch.empty(25 * 16 + 25 + 2 * (S * 25 + S), device="cuda")
    stats = torch.zeros(4, dtype=torch.int64, device="cuda")
    args = lambda cap, wsn: (B, T, 15, 25, S, p.times.data_ptr(), z.data_ptr(), c.data_ptr(), W1.data_ptr(),
                             *[x.data_ptr() for x in w], sol.data_ptr(), sol.stride(0), sol.stride(1), gs.data_ptr(),
                             gs.stride(0), gs.stride(1), 1e-3, 1e-4, cap, None, 0, gy0.data_ptr(), gp.data_ptr(), None, 0,
                             stats.data_ptr(), ws.data_ptr(), wsn, torch.cuda.current_stream().cuda_stream)
    assert lib.slode_mlp_dopri5_adjoint_bwd(*args(3, n)) == 0
    assert stats.tolist()[3] == 2 and stats.tolist()[0] + stats.tolist()[1] == 3      # stopped: max_attempts
    assert lib.slode_mlp_dopri5_adjoint_bwd(*args(1 << 20, n)) == 0
    assert stats.tolist()[3] == 0 and stats.tolist()[0] >= T - 1
    assert lib.slode_mlp_dopri5_adjoint_bwd(*args(1 << 20, n - 256)) != 0             # workspace too small
    assert b"workspace" in lib.slode_last_error()
    # T == 1: the gradient of sol[0] = y0 is the cotangent itself
    y1 = p.initialize_state(z).detach().requires_grad_(True)
    one = slode.odeint_adjoint(p.gen_dynamics(z), y1, p.times[:1], method="dopri5")
    one.sum().backward()
    assert torch.equal(y1.grad, torch.ones_like(y1))


def test_golden_fixture_from_reference_classes(golden_dir):
    """tests/golden/blackbox_golden.npz case ``chal/dopri5/1``: ``odeint_adjoint(..., method="dopri5")`` under the
    reference's REAL OdeModel / OdeFunc / Dynamics classes (tests/golden/make_golden.py), with the step logs of its
    forward and of its backward solves.  Both sequences replayed on the device: trajectories, dL/dx0 and the six
    parameter gradients of func.parameters() equal the fixture's."""
    import os
    import structured_latent_odes_b200 as slode
    g = np.load(os.path.join(golden_dir, "blackbox_golden.npz"))
    times = torch.from_numpy(g["chal/times"]).cuda()
    W = {k[len("chal/w/"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("chal/w/")}
    m = slode.OdeModel()
    m.init_with_params(times, 5, 15, 25, True, "dopri5", "cuda")
    m.load_state_dict(W, strict=False)
    m = m.cuda()
    z = torch.from_numpy(g["chal/z"]).cuda()
    G = torch.from_numpy(g["chal/G"]).cuda()
    rows, t = [], float(times[0])
    for a, d in zip(g["chal/dopri5/1/accepted"], g["chal/dopri5/1/dts"]):
        rows.append((t, float(d), 1.0 if a else 0.0))
        t = t + float(d) if a else t
    y0 = m.initialize_state(z).detach().requires_grad_(True)
    sol = slode.odeint_adjoint(m.gen_dynamics(z), y0, times, method="dopri5", rtol=1e-5, atol=1e-6,
                               options={"replay_steps": torch.tensor(rows, dtype=torch.float64),
                                        "adjoint_replay_steps": torch.from_numpy(g["chal/dopri5/1/backward_steps"])})
    (sol.permute(1, 0, 2) * G).sum().backward()
    assert U.rel_err(sol.permute(1, 0, 2), torch.from_numpy(g["chal/dopri5/1/sol"])) < 1e-5
    assert U.rel_err(y0.grad, torch.from_numpy(g["chal/dopri5/1/grad_y0"])) < 5e-5
    for k, p_ in m.dynamics.named_parameters():
        if ".prod." in k or ".degr." in k:
            continue
        want = torch.from_numpy(g[f"chal/dopri5/1/g/dynamics.{k}"])
        assert U.rel_err(p_.grad, want) < 5e-5, (k, U.rel_err(p_.grad, want))
