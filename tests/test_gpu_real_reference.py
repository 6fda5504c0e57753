"""The drop-in claim on hardware: the reference's REAL, unmodified ``models/blackbox_ode.py`` and
``models/decoders.py`` (byte-for-byte copies under ``baseline/_ref``, see ``baseline/make_ref.py``) run on the B200
with ``import torchdiffeq`` resolving to this package (what ``install_as_torchdiffeq()`` registers), and are
compared -- solution and every gradient -- with the same real classes on the CPU over the oracle's torchdiffeq
restatement.  Tolerance 1e-5 relative (north_star, fp32 fixed step)."""
import contextlib
import hashlib
import io
import json
import os

import pytest
import torch

import slode_testutil as U
from baseline import make_ref, ref_shims

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _need():
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    if not make_ref.available():
        pytest.fail("baseline/_ref is missing: run __graft_entry__.build() in the build container "
                    "(it copies the reference's hot-path sources; the directory ships with the gpurun snapshot)")


def _pair():
    """(reference modules bound to the product, reference modules bound to the CPU oracle)"""
    import structured_latent_odes_b200 as slode
    from structured_latent_odes_b200 import torchdiffeq_api
    from oracle import torchdiffeq_oracle
    import sys
    old = sys.modules.get("torchdiffeq")
    try:
        slode.install_as_torchdiffeq()
        assert sys.modules["torchdiffeq"] is torchdiffeq_api
        gpu = ref_shims.import_real(sys.modules["torchdiffeq"])
    finally:  # other tests bind the name to the oracle
        if old is None:
            sys.modules.pop("torchdiffeq", None)
        else:
            sys.modules["torchdiffeq"] = old
    cpu = ref_shims.import_real(torchdiffeq_oracle)
    assert gpu[0].torchdiffeq is torchdiffeq_api and cpu[0].torchdiffeq is torchdiffeq_oracle
    return gpu, cpu


def test_copies_are_unmodified():
    _need()
    man = json.load(open(os.path.join(make_ref.DEST, "MANIFEST.json")))
    for rel, sha in man["sha256"].items():
        with open(os.path.join(make_ref.DEST, rel), "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == sha, rel
    assert "models/blackbox_ode.py" in man["sha256"] and "models/decoders.py" in man["sha256"]


def _real_ode(mod, times, L, H, S, adjoint, method, device):
    with contextlib.redirect_stdout(io.StringIO()):
        m = mod.OdeModel()
        m.init_with_params(times=times.to(device), ode_state_dim=S, latent_dim=L, ode_hidden_dim=H,
                           adjoint_solver=adjoint, solver=method, device=device)
    return m.to(device)


@pytest.mark.parametrize("shape", ["cvs", "proc"])
@pytest.mark.parametrize("method,adjoint", [("midpoint", True), ("rk4", False), ("midpoint", False), ("rk4", True),
                                            ("euler", True)])
def test_real_odemodel_on_gpu_matches_real_odemodel_on_cpu(shape, method, adjoint):
    """models/blackbox_ode.py:36-47 executed verbatim; only `torchdiffeq` differs between the two sides."""
    _need()
    (bb_gpu, _), (bb_cpu, _) = _pair()
    L, H, S, times = U.SHAPES[shape]
    torch.manual_seed(12)
    ref = _real_ode(bb_cpu, times, L, H, S, adjoint, method, "cpu")
    dut = _real_ode(bb_gpu, times, L, H, S, adjoint, method, "cuda")
    dut.load_state_dict(ref.state_dict())
    g = torch.Generator().manual_seed(21)
    B = 96
    z = torch.randn(B, L, generator=g)
    G = torch.randn(B, len(times), S, generator=g)
    so, gzo, gro = U.run_fwd_bwd(ref, z, G)
    sp, gzp, grp = U.run_fwd_bwd(dut, z.cuda(), G.cuda())
    assert sp.shape == so.shape == (B, len(times), S)
    assert U.rel_err(sp, so) < TOL
    assert (gzo is None) == (gzp is None)
    assert U.rel_err(gzp, gzo) < TOL
    assert set(grp) == set(gro)
    for k in gro:
        assert U.rel_err(grp[k], gro[k]) < TOL, (k, U.rel_err(grp[k], gro[k]))


@pytest.mark.parametrize("adjoint,method", [(True, "midpoint"), (False, "rk4")])
def test_real_decoder_on_gpu_matches_real_decoder_on_cpu(adjoint, method):
    """models/decoders.py:42-54 executed verbatim on the GPU: solution, three quantile heads, std and all gradients."""
    _need()
    (_, dec_gpu), (_, dec_cpu) = _pair()
    L, H, S, times = U.SHAPES["cvs"]
    cfg = ref_shims.Munch(obs_dim=3, system_input_dim=0, ode_state_dim=S, ode_hidden_dim=H, adjoint_solver=adjoint,
                          solver=method, constant_std=1e-2, seq_len=len(times))
    torch.manual_seed(12)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = dec_cpu.Decoder(config=cfg, latent_dim=L, times=times, device="cpu")
        dut = dec_gpu.Decoder(config=cfg, latent_dim=L, times=times.cuda(), device="cuda").cuda()
    dut.load_state_dict(ref.state_dict())
    g = torch.Generator().manual_seed(22)
    z = torch.randn(64, L, generator=g)
    y = torch.rand(64, 3, len(times), generator=g)

    def run(model, z, y):
        model.zero_grad()
        z = z.clone().requires_grad_(True)
        sol, q75, q50, q25, std = model.forward(z)
        loss = ((q50 - y).abs() / std).mean() + (q75 - y).square().mean() + (q25 - y).square().mean() + sol.mean()
        loss.backward()
        grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()
                 if p.grad is not None and ".prod." not in k and ".degr." not in k}
        return (sol, q75, q50, q25, std), z.grad, grads

    outs_o, gzo, gro = run(ref, z, y)
    outs_p, gzp, grp = run(dut, z.cuda(), y.cuda())
    for a, b in zip(outs_p, outs_o):
        assert a.shape == b.shape and U.rel_err(a, b) < TOL
    assert U.rel_err(gzp, gzo) < TOL
    assert set(grp) == set(gro)
    for k in gro:
        assert U.rel_err(grp[k], gro[k]) < 2 * TOL, (k, U.rel_err(grp[k], gro[k]))
