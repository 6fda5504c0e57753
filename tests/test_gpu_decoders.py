"""Decoder mirrors (a9: the boundary consumer of the solve) against the oracle's QuantileHeads, which is itself
pinned on the reference's real ``Decoder`` class (tests/test_oracle_golden.py)."""
import contextlib
import io

import pytest
import torch

import slode_testutil as U
from oracle import shims, slode_port

pytestmark = pytest.mark.gpu


def _cfg(shape, method, adjoint, O):
    L, H, S, times = U.SHAPES[shape]
    return shims._Munch(obs_dim=O, system_input_dim=0, ode_state_dim=S, ode_hidden_dim=H, adjoint_solver=adjoint,
                        solver=method, constant_std=1e-2, seq_len=len(times))


@pytest.mark.parametrize("shape,O,method,adjoint", [("cvs", 3, "midpoint", True), ("chal", 4, "rk4", False),
                                                    ("proc", 4, "midpoint", True)])
def test_decoder_matches_oracle_heads_and_gradients(shape, O, method, adjoint):
    import structured_latent_odes_b200 as slode
    L, H, S, times = U.SHAPES[shape]
    o_ode = U.make_oracle(shape, method, adjoint)
    torch.manual_seed(4)
    heads = slode_port.QuantileHeads(o_ode, O, len(times))
    dec = slode.Decoder(_cfg(shape, method, adjoint, O), times.cuda(), L, "cuda")
    sd = {"ode_model." + k: v for k, v in o_ode.state_dict().items()}
    sd.update({k: v for k, v in heads.state_dict().items() if not k.startswith("ode_model.")})
    dec.load_state_dict(sd)
    dec = dec.cuda()
    g = torch.Generator().manual_seed(8)
    B = 77
    z = torch.randn(B, L, generator=g)
    G = [torch.randn(B, O, len(times), generator=g) for _ in range(3)]

    zo = z.clone().requires_grad_(True)
    sol_o, q75_o, q50_o, q25_o, std_o = heads(zo)
    (q75_o * G[0] + q50_o * G[1] + q25_o * G[2]).sum().backward()
    zp = z.cuda().requires_grad_(True)
    sol_p, q75_p, q50_p, q25_p, std_p = dec(zp)
    assert q50_p.shape == (B, O, len(times)) and sol_p.shape == (B, len(times), S)
    (q75_p * G[0].cuda() + q50_p * G[1].cuda() + q25_p * G[2].cuda()).sum().backward()
    for a, b in ((sol_p, sol_o), (q75_p, q75_o), (q50_p, q50_o), (q25_p, q25_o), (std_p, std_o)):
        assert U.rel_err(a, b) < 1e-5
    assert U.rel_err(zp.grad, zo.grad) < 1e-5
    go = dict(heads.named_parameters())
    for k, p in dec.named_parameters():
        if ".prod." in k or ".degr." in k or p.grad is None:
            continue
        ref = go[k].grad
        assert ref is not None, k
        assert U.rel_err(p.grad, ref) < 1e-5, (k, U.rel_err(p.grad, ref))


def test_gaussian_decoder_and_raw_heads():
    import structured_latent_odes_b200 as slode
    L, H, S, times = U.SHAPES["cvs"]
    dec = slode.GaussianDecoder(_cfg("cvs", "midpoint", True, 3), times.cuda(), L, "cuda").cuda()
    z = torch.randn(33, L, device="cuda")
    sol, mean, std = dec(z)
    want = (sol @ dec.output_mean[0].weight.t()).permute(0, 2, 1)
    assert U.rel_err(mean, want) < 1e-6 and std.shape == mean.shape
    # heads on a (T,B,S)-contiguous solution (strided view) and ragged sizes
    x = torch.randn(7, 130, 5, device="cuda").permute(1, 0, 2).requires_grad_(True)
    W = [torch.randn(4, 5, device="cuda", requires_grad=True) for _ in range(2)]
    mus = slode.decoder_heads(x, W)
    ref = [(x @ w.t()).permute(0, 2, 1) for w in W]
    G = [torch.randn_like(m) for m in mus]
    gx, gw0, gw1 = torch.autograd.grad(sum((m * g).sum() for m, g in zip(mus, G)), [x] + W)
    rx, rw0, rw1 = torch.autograd.grad(sum((m * g).sum() for m, g in zip(ref, G)), [x] + W)
    for a, b in zip(mus + [gx, gw0, gw1], ref + [rx, rw0, rw1]):
        assert U.rel_err(a, b) < 1e-5


def test_state_dict_keys_equal_reference_decoder():
    """Against the reference's REAL decoder classes (unmodified copies shipped in baseline/_ref)."""
    import structured_latent_odes_b200 as slode
    from baseline import make_ref, ref_shims
    from oracle import torchdiffeq_oracle
    if not make_ref.available():
        pytest.fail("baseline/_ref is missing: run __graft_entry__.build() in the build container")
    _, dec_ref = ref_shims.import_real(torchdiffeq_oracle)
    L, H, S, times = U.SHAPES["cvs"]
    cfg = _cfg("cvs", "midpoint", True, 3)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = dec_ref.Decoder(config=cfg, latent_dim=L, times=times, device="cpu")
        gref = dec_ref.GaussianDecoder(config=cfg, latent_dim=L, times=times, device="cpu")
    assert sorted(ref.state_dict()) == sorted(slode.Decoder(cfg, times, L, "cpu").state_dict())
    assert sorted(gref.state_dict()) == sorted(slode.GaussianDecoder(cfg, times, L, "cpu").state_dict())
    with contextlib.redirect_stdout(io.StringIO()):
        vref = dec_ref.VarianceGaussianDecoder(config=cfg, latent_dim=L, times=times, device="cpu")
    ours = slode.VarianceGaussianDecoder(cfg, times.cuda(), L, "cuda")
    assert sorted(vref.state_dict()) == sorted(ours.state_dict())
    # same weights -> same outputs as the reference's real class over the oracle solver (it runs on the CPU)
    ours.load_state_dict(vref.state_dict())
    ours = ours.cuda()
    z = torch.randn(9, L)
    with torch.no_grad():
        sol_r, mean_r, std_r = vref(z)
        sol_o, mean_o, std_o = ours(z.cuda())
    for a, b in ((sol_o, sol_r), (mean_o, mean_r), (std_o, std_r)):
        assert U.rel_err(a, b) < 1e-5


def test_multiple_samples_equals_the_reference_style_loop():
    """f3: num_samples posterior draws in one launch == the reference's loop of separate solves on the same draws."""
    import structured_latent_odes_b200 as slode
    L, H, S, times = U.SHAPES["chal"]
    dec = slode.Decoder(_cfg("chal", "midpoint", True, 4), times.cuda(), L, "cuda").cuda()
    B, K = 35, 12
    loc, scale = torch.randn(B, L, device="cuda"), 0.1 + torch.rand(B, L, device="cuda")
    res = slode.multiple_samples(dec, loc, scale, K, generator=torch.Generator(device="cuda").manual_seed(1))
    assert res["mu_50"].shape == (B, 4, len(times), K) and res["z"].shape == (K, B, L)
    with torch.no_grad():
        for k in (0, 5, K - 1):
            _, q75, q50, q25, _ = dec(res["z"][k])
            assert torch.equal(res["mu_50"][..., k], q50) and torch.equal(res["mu_75"][..., k], q75)
            assert torch.equal(res["mu_25"][..., k], q25)


@pytest.mark.parametrize("S,O,NQ,T,B", [(5, 3, 3, 100, 130), (8, 4, 3, 100, 37), (4, 2, 1, 16, 300), (5, 3, 2, 86, 65),
                                        (5, 8, 3, 40, 50)])
def test_heads_vector_and_scalar_paths_against_torch(S, O, NQ, T, B):
    """The heads kernels take four time points per thread when T % 4 == 0 on (B,T,S)-contiguous rows (16-byte loads)
    and one otherwise; NQ*O up to 12 is unrolled, above that rolled.  Both against plain torch, including a batch
    slice that starts in the middle of the buffer and a head whose output does not enter the loss."""
    import structured_latent_odes_b200 as slode
    g = torch.Generator(device="cuda").manual_seed(S * 100 + T)
    full = torch.randn(B + 5, T, S, device="cuda", generator=g)
    x = full[3:3 + B].detach().requires_grad_(True)          # view with an offset: rows stay contiguous
    W = [torch.randn(O, S, device="cuda", generator=g).requires_grad_(True) for _ in range(NQ)]
    mus = slode.decoder_heads(x, W)
    ref = [(x @ w.t()).permute(0, 2, 1) for w in W]
    used = list(range(NQ)) if NQ < 3 else [0, 2]             # with three heads, leave the middle one out of the loss
    G = {q: torch.randn(B, O, T, device="cuda", generator=g) for q in used}
    outs = torch.autograd.grad(sum((mus[q] * G[q]).sum() for q in used), [x] + [W[q] for q in used])
    refs = torch.autograd.grad(sum((ref[q] * G[q]).sum() for q in used), [x] + [W[q] for q in used])
    for a, b in zip(list(mus) + list(outs), list(ref) + list(refs)):
        assert U.rel_err(a, b) < 1e-5


@pytest.mark.parametrize("shape,O,method,B", [("cvs", 3, "midpoint", 77), ("cvs", 3, "rk4", 300), ("chal", 4, "euler", 1),
                                              ("chal", 4, "rk4", 129), ("proc", 4, "midpoint", 45),
                                              ("h64", 3, "rk4", 40)])
@pytest.mark.parametrize("want_solution", [False, True])
def test_predict_fuses_the_heads_into_the_solve_bit_equal(shape, O, method, B, want_solution):
    """f2: ``Decoder.predict`` (heads in the solver kernel's epilogue, trajectories not written unless asked for) gives
    exactly the tuple ``forward`` gives with two kernels; T = 86 / 142 / 100 cover full and partial 16-byte row groups
    and unaligned rows."""
    import structured_latent_odes_b200 as slode
    L, H, S, times = U.SHAPES[shape]
    torch.manual_seed(B)
    dec = slode.Decoder(_cfg(shape, method, False, O), times.cuda(), L, "cuda").cuda()
    z = torch.randn(B, L, device="cuda")
    with torch.no_grad():
        sol, q75, q50, q25, std = dec(z)
    fsol, f75, f50, f25, fstd = dec.predict(z, want_solution=want_solution)
    assert torch.equal(f75, q75) and torch.equal(f50, q50) and torch.equal(f25, q25) and torch.equal(fstd, std)
    if want_solution:
        assert fsol.shape == sol.shape and torch.equal(fsol, sol)
    else:
        assert fsol is None
    # predict hands out views of rows padded to whole 32-byte sectors; the reference's contiguous rows (every row at
    # its own sector phase: the lanes of a warp store at different steps) give the same numbers
    W3 = (dec.output_q50[0].weight, dec.output_q75[0].weight, dec.output_q25[0].weight)
    mu_c, _ = dec.ode_model.solve_ODE_heads(z, W3, contiguous=True)
    assert mu_c.is_contiguous() and torch.equal(mu_c[0], q50) and torch.equal(mu_c[1], q75) and torch.equal(mu_c[2], q25)
    assert f50.stride(-2) % 8 == 0 and f50.stride(-1) == 1
    # and against the oracle heads on the oracle solve
    o_ode = U.make_oracle(shape, method, False)
    o_ode.load_state_dict({k[len("ode_model."):]: v.cpu() for k, v in dec.state_dict().items()
                           if k.startswith("ode_model.")})
    with torch.no_grad():
        want = (o_ode.solve_ODE(z.cpu()) @ dec.output_q50[0].weight.cpu().t()).permute(0, 2, 1)
    assert U.rel_err(f50, want) < 1e-5


def test_gaussian_predict_and_fused_heads_argument_checks():
    import structured_latent_odes_b200 as slode
    L, H, S, times = U.SHAPES["cvs"]
    dec = slode.GaussianDecoder(_cfg("cvs", "rk4", True, 3), times.cuda(), L, "cuda").cuda()
    z = torch.randn(50, L, device="cuda")
    with torch.no_grad():
        sol, mean, std = dec(z)
    fsol, fmean, fstd = dec.predict(z)
    assert fsol is None and torch.equal(fmean, mean) and torch.equal(fstd, std)
    empty = dec.predict(z[:0], want_solution=True)
    assert empty[1].shape == (0, 3, len(times)) and empty[0].shape == (0, len(times), S)
    with pytest.raises(ValueError):
        dec.ode_model.solve_ODE_heads(z, (torch.randn(3, S + 1, device="cuda"),))
    with pytest.raises(RuntimeError):
        dec.ode_model.solve_ODE_heads(z, (torch.randn(3, S),))          # head weights left on the CPU
    dop = slode.GaussianDecoder(_cfg("cvs", "dopri5", False, 3), times.cuda(), L, "cuda").cuda()
    with pytest.raises(NotImplementedError):
        dop.predict(z)


def test_predict_on_a_grid_longer_than_the_staged_one():
    """Grids above 1024 points are read from global memory by the fused kernel (shorter ones are staged in shared
    memory): same bits as the two-kernel path either way; T = 1100 also ends on a partial sector."""
    import structured_latent_odes_b200 as slode
    L, H, S, _ = U.SHAPES["cvs"]
    for T in (1100, 1024):
        times = torch.linspace(0.0, 30.0, T)
        torch.manual_seed(T)
        dec = slode.Decoder(_cfg("cvs", "euler", False, 3), times.cuda(), L, "cuda").cuda()
        z = torch.randn(70, L, device="cuda")
        with torch.no_grad():
            sol, q75, q50, q25, _ = dec(z)
        fsol, f75, f50, f25, _ = dec.predict(z, want_solution=True)
        assert torch.equal(f50, q50) and torch.equal(f75, q75) and torch.equal(f25, q25) and torch.equal(fsol, sol)
