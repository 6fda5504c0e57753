"""Direct calls into the C ABI with raw device pointers (what a non-PyTorch host would do)."""
import ctypes

import pytest
import torch

import slode_testutil as U
from structured_latent_odes_b200 import _cabi

pytestmark = pytest.mark.gpu


def test_raw_pointer_call_matches_the_python_api():
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    o = U.make_oracle("cvs", "rk4", False)
    p = U.make_product(o)
    B, T, L, H, S = 300, 86, 15, 25, 5
    z = torch.randn(B, L, device="cuda")
    want = p.solve_ODE(z).detach().permute(1, 0, 2).contiguous()
    d = p.dynamics
    W1 = d.dynamics_hidden.weight.detach()
    c = torch.addmm(d.dynamics_hidden.bias.detach(), z, W1[:, 1:].t()).contiguous()
    y0 = p.initialize_state(z).detach().contiguous()
    w1t = W1[:, 0].contiguous()
    Wg, bg = d.dyanamics_growth.weight.detach().contiguous(), d.dyanamics_growth.bias.detach().contiguous()
    Wd, bd = d.dyanmics_degradation.weight.detach().contiguous(), d.dyanmics_degradation.bias.detach().contiguous()
    sol = torch.empty(T, B, S, device="cuda")
    L_ = _cabi.lib()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    rc = L_.slode_mlp_fixed_fwd(_cabi.METHOD_RK4, B, T, H, S, p.times.data_ptr(), c.data_ptr(), y0.data_ptr(),
                                w1t.data_ptr(), Wg.data_ptr(), bg.data_ptr(), Wd.data_ptr(), bd.data_ptr(),
                                sol.data_ptr(), B * S, S, None, 0, ctypes.c_void_p(s.cuda_stream))
    assert rc == 0, L_.slode_last_error()
    s.synchronize()
    # the Python API runs the FUSED entry point (c and x0 computed inside the kernel): same numbers to rounding
    assert U.rel_err(sol, want) < 1e-6
    assert L_.slode_query(_cabi.Q_FWD_LAUNCHES) >= 1
    # narrow layers keep their tables in shared memory: the forward needs no workspace, the reverse sweep its records
    assert L_.slode_fixed_workspace_bytes(0, _cabi.METHOD_RK4, 0, B, T, L, H, S, 2, 0) == 0
    assert L_.slode_fixed_workspace_bytes(1, _cabi.METHOD_RK4, 0, B, T, L, H, S, 2, 0) > 0
    assert L_.slode_fixed_workspace_bytes(1, _cabi.METHOD_RK4, 0, B, T, L, 7, 3, 2, 0) == -1
    # the fused entry point with raw pointers: bit-identical to the Python API
    n0, n2 = p.latent_to_ode_net[0], p.latent_to_ode_net[2]
    sol2 = torch.empty(T, B, S, device="cuda")
    rc = L_.slode_latent_fixed_fwd(_cabi.METHOD_RK4, B, T, L, H, S, p.times.data_ptr(), z.data_ptr(),
                                   W1.contiguous().data_ptr(), d.dynamics_hidden.bias.detach().data_ptr(),
                                   Wg.data_ptr(), bg.data_ptr(), Wd.data_ptr(), bd.data_ptr(),
                                   n0.weight.detach().data_ptr(), n0.bias.detach().data_ptr(),
                                   n2.weight.detach().data_ptr(), n2.bias.detach().data_ptr(), None,
                                   sol2.data_ptr(), B * S, S, None, 0,
                                   ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, L_.slode_last_error()
    torch.cuda.synchronize()
    assert torch.equal(sol2, want)


@pytest.mark.parametrize("method,adjoint", [("rk4", False), ("midpoint", True), ("euler", False)])
def test_c_based_entry_points_match_the_fused_path(method, adjoint):
    """slode_mlp_fixed_fwd / _bwd (precomputed c and y0, gradients to c / y0 / weights) against the fused
    slode_latent_fixed_* path through autograd: same trajectories and the same gradients for every parameter."""
    from structured_latent_odes_b200 import torchdiffeq_api as api
    o = U.make_oracle("cvs", method, adjoint)
    p = U.make_product(o)
    g = torch.Generator().manual_seed(1)
    z = torch.randn(150, 15, generator=g).cuda()
    G = torch.randn(150, 86, 5, generator=g).cuda()
    sol_f, gz_f, gr_f = U.run_fwd_bwd(p, z, G)
    p.zero_grad()
    zz = z.clone().requires_grad_(True)
    d = p.dynamics
    W1 = d.dynamics_hidden.weight
    zc = zz.detach() if adjoint else zz   # odeint_adjoint: no gradient to z through the dynamics
    c = torch.addmm(d.dynamics_hidden.bias, zc, W1[:, 1:].t())
    y0 = p.initialize_state(zz)
    sol = api.solve_fixed_from_c(y0, c, W1[:, 0], d.dyanamics_growth.weight, d.dyanamics_growth.bias,
                                 d.dyanmics_degradation.weight, d.dyanmics_degradation.bias, p.times, method,
                                 adjoint=adjoint).permute(1, 0, 2)
    (sol * G).sum().backward()
    assert U.rel_err(sol, sol_f) < 2e-6
    assert U.rel_err(zz.grad, gz_f) < 1e-5
    for k, v in p.named_parameters():
        if ".prod." in k or ".degr." in k:
            continue
        assert U.rel_err(v.grad, gr_f[k]) < 1e-5, k


@pytest.mark.parametrize("shape,method,mode", [("cvs", "rk4", 0), ("cvs", "midpoint", 1), ("proc", "rk4", 0),
                                               ("h64", "midpoint", 0), ("h128", "euler", 0)])
@pytest.mark.parametrize("B", [1, 130])
def test_no_write_outside_the_callers_buffers(shape, method, mode, B):
    """compute-sanitizer is closed on this GPU pool, so out-of-bounds WRITES are hunted with guard bands: every
    output and the workspace are carved out of one arena with canary words on both sides (and between them), the
    fused entry points run on raw pointers, and the canaries must be intact afterwards.  Ragged B exercises the
    masked tail threads."""
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    Ld, H, S, times = U.SHAPES[shape]
    T = len(times)
    o = U.make_oracle(shape, method, bool(mode))
    p = U.make_product(o)
    d, n0, n2 = p.dynamics, p.latent_to_ode_net[0], p.latent_to_ode_net[2]
    lib = _cabi.lib()
    mid = _cabi.METHODS[method]
    ws_f = lib.slode_fixed_workspace_bytes(0, mid, mode, B, T, Ld, H, S, 2, 0)
    ws_b = lib.slode_fixed_workspace_bytes(1, mid, mode, B, T, Ld, H, S, 2, 0)
    assert ws_f >= 0 and ws_b > 0
    n_par = H + 2 * (S * H + S) + 2 * (H * Ld + H) + S * H + S
    GUARD = 256  # floats
    r64 = lambda n: (n + 63) // 64 * 64  # noqa: E731  (views stay 256-byte aligned like allocator blocks)
    sizes = {"sol": r64(T * B * S), "ws_f": r64((ws_f + 3) // 4), "ws_b": r64((ws_b + 3) // 4), "gz": r64(B * Ld),
             "gp": r64(n_par)}
    used = {"sol": T * B * S, "gz": B * Ld, "gp": n_par}
    total = GUARD + sum(v + GUARD for v in sizes.values())
    arena = torch.full((total,), 12345.678, device="cuda")
    off, view = GUARD, {}
    for k, v in sizes.items():
        view[k] = arena[off:off + v]
        off += v + GUARD
    view["gp"][:n_par].zero_()
    z = torch.randn(B, Ld, device="cuda")
    G = torch.randn(T, B, S, device="cuda")
    w = [x.detach().contiguous() for x in (d.dynamics_hidden.weight, d.dynamics_hidden.bias, d.dyanamics_growth.weight,
                                           d.dyanamics_growth.bias, d.dyanmics_degradation.weight,
                                           d.dyanmics_degradation.bias, n0.weight, n0.bias, n2.weight, n2.bias)]
    s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.slode_latent_fixed_fwd(mid, B, T, Ld, H, S, p.times.data_ptr(), z.data_ptr(), *[x.data_ptr() for x in w], None,
                                    view["sol"].data_ptr(), B * S, S, view["ws_f"].data_ptr() if ws_f else None, ws_f, s)
    assert rc == 0, lib.slode_last_error()
    rc = lib.slode_latent_fixed_bwd(mid, mode, B, T, Ld, H, S, p.times.data_ptr(), z.data_ptr(), *[x.data_ptr() for x in w],
                                    view["sol"].data_ptr(), B * S, S, G.data_ptr(), B * S, S, view["gz"].data_ptr(), None,
                                    view["gp"].data_ptr(), view["ws_b"].data_ptr(), ws_b, s)
    assert rc == 0, lib.slode_last_error()
    torch.cuda.synchronize()
    # canaries: everything that is not inside a view (for the outputs: not inside the part the call owns)
    mask = torch.ones(total, dtype=torch.bool, device="cuda")
    off = GUARD
    for k, v in sizes.items():
        mask[off:off + used.get(k, v)] = False
        off += v + GUARD
    assert bool((arena[mask] == 12345.678).all()), "a kernel wrote outside the buffers it was given"
    # and the results are the Python API's
    sol_ref = p.solve_ODE(z).detach().permute(1, 0, 2)
    assert torch.equal(view["sol"][:T * B * S].view(T, B, S), sol_ref)
    assert torch.isfinite(view["gz"][:B * Ld]).all() and torch.isfinite(view["gp"][:n_par]).all()
    # a workspace that is too small is refused, not overrun
    rc = lib.slode_latent_fixed_bwd(mid, mode, B, T, Ld, H, S, p.times.data_ptr(), z.data_ptr(), *[x.data_ptr() for x in w],
                                    view["sol"].data_ptr(), B * S, S, G.data_ptr(), B * S, S, view["gz"].data_ptr(), None,
                                    view["gp"].data_ptr(), view["ws_b"].data_ptr(), ws_b - 1, s)
    assert rc == 1 and b"workspace" in lib.slode_last_error()
    rc = lib.slode_latent_fixed_bwd(mid, mode, B, T, Ld, H, S, p.times.data_ptr(), z.data_ptr(), *[x.data_ptr() for x in w],
                                    view["sol"].data_ptr(), B * S, S, G.data_ptr(), B * S, S, view["gz"].data_ptr(), None,
                                    view["gp"].data_ptr(), view["ws_b"].data_ptr() + 4, ws_b, s)
    assert rc == 1 and b"aligned" in lib.slode_last_error()


@pytest.mark.parametrize("shape,method,O,NQ,B,pad,shift", [("cvs", "rk4", 3, 3, 130, 0, 0), ("cvs", "midpoint", 3, 3, 33, 2, 0),
                                                          ("chal", "euler", 4, 1, 65, 5, 3), ("proc", "midpoint", 4, 3, 40, 4, 1),
                                                          ("small", "rk4", 8, 3, 1, 7, 5), ("h128", "rk4", 2, 2, 50, 1, 2)])
def test_fused_heads_entry_raw_pointers_any_pitch_and_alignment(shape, method, O, NQ, B, pad, shift):
    """slode_latent_fixed_heads_fwd on raw pointers: row pitch = T + pad (sector phases of every kind, incl. odd ones),
    a mu base that is only 4-byte aligned (shift floats into the arena), sol skipped -- the values are those of
    slode_latent_fixed_fwd + torch heads, the padding columns and the guard bands around mu stay untouched."""
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    Ld, H, S, times = U.SHAPES[shape]
    T = len(times)
    o = U.make_oracle(shape, method, False)
    p = U.make_product(o)
    d, n0, n2 = p.dynamics, p.latent_to_ode_net[0], p.latent_to_ode_net[2]
    lib = _cabi.lib()
    mid = _cabi.METHODS[method]
    g = torch.Generator(device="cuda").manual_seed(B + pad)
    z = torch.randn(B, Ld, device="cuda", generator=g)
    W = torch.randn(NQ, O, S, device="cuda", generator=g)
    with torch.no_grad():
        sol = p.solve_ODE(z)                                           # (B,T,S)
        want = torch.stack([(sol @ W[q].t()).permute(0, 2, 1) for q in range(NQ)], 0)   # (NQ,B,O,T)
    P = T + pad
    GUARD, CANARY = 256, 12345.678
    n_mu = NQ * B * O * P
    arena = torch.full((2 * GUARD + n_mu + 8,), CANARY, device="cuda")
    mu = arena[GUARD + shift:GUARD + shift + n_mu]
    ws_n = lib.slode_fixed_workspace_bytes(0, mid, 0, B, T, Ld, H, S, 2, 0)
    ws = torch.empty(max(ws_n, 1), dtype=torch.uint8, device="cuda")
    w = [x.detach().contiguous() for x in (d.dynamics_hidden.weight, d.dynamics_hidden.bias, d.dyanamics_growth.weight,
                                           d.dyanamics_growth.bias, d.dyanmics_degradation.weight,
                                           d.dyanmics_degradation.bias, n0.weight, n0.bias, n2.weight, n2.bias)]
    s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.slode_latent_fixed_heads_fwd(mid, B, T, Ld, H, S, p.times.data_ptr(), z.data_ptr(), *[x.data_ptr() for x in w],
                                          None, O, NQ, W.data_ptr(), mu.data_ptr(), P, None, 0, 0,
                                          ws.data_ptr() if ws_n else None, ws_n, s)
    assert rc == 0, lib.slode_last_error()
    torch.cuda.synchronize()
    got = mu.view(NQ, B, O, P)
    assert U.rel_err(got[..., :T], want) < 2e-6
    if pad:
        assert bool((got[..., T:] == CANARY).all()), "padding columns of the rows were written"
    assert bool((arena[:GUARD + shift] == CANARY).all()) and bool((arena[GUARD + shift + n_mu:] == CANARY).all())
    # argument checks: pitch below T, too many outputs
    assert lib.slode_latent_fixed_heads_fwd(mid, B, T, Ld, H, S, p.times.data_ptr(), z.data_ptr(), *[x.data_ptr() for x in w],
                                            None, O, NQ, W.data_ptr(), mu.data_ptr(), T - 1, None, 0, 0, None, 0, s) == 1
    assert lib.slode_latent_fixed_heads_fwd(mid, B, T, Ld, H, S, p.times.data_ptr(), z.data_ptr(), *[x.data_ptr() for x in w],
                                            None, 9, NQ, W.data_ptr(), mu.data_ptr(), P, None, 0, 0, None, 0, s) == 1
