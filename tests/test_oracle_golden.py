"""The CPU port of the reference classes (``oracle/slode_port.py``) against

* ``tests/golden/blackbox_golden.npz`` -- produced by the reference's REAL ``OdeModel`` / ``OdeFunc`` /
  ``Dynamics`` classes (``models/blackbox_ode.py``, imported unchanged by ``tests/golden/make_golden.py``);
* the real classes themselves when ``/root/reference`` is present (build container only).
"""
import contextlib
import io
import os

import numpy as np
import pytest
import torch

from oracle import shims, slode_port
from oracle import torchdiffeq_oracle as tde

CASES = [("cvs", m, a) for m in ("euler", "midpoint", "rk4") for a in (0, 1)] + \
        [("proc", m, a) for m in ("midpoint", "rk4") for a in (0, 1)] + [("chal", "dopri5", a) for a in (0, 1)]


def _load(golden_dir):
    return np.load(os.path.join(golden_dir, "blackbox_golden.npz"))


def _port_from_golden(g, name, method, adj):
    times = torch.from_numpy(g[f"{name}/times"])
    W = {k[len(name) + 3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(f"{name}/w/")}
    H, Lp1 = W["dynamics.dynamics_hidden.weight"].shape
    S = W["dynamics.dyanamics_growth.weight"].shape[0]
    m = slode_port.OdeModel(times, S, Lp1 - 1, H, bool(adj), method)
    missing = m.load_state_dict(W, strict=False)
    assert not missing.unexpected_keys
    assert all(".prod." in k or ".degr." in k for k in missing.missing_keys)
    return m


@pytest.mark.parametrize("name,method,adj", CASES)
def test_port_reproduces_reference_class_outputs(golden_dir, name, method, adj):
    g = _load(golden_dir)
    m = _port_from_golden(g, name, method, adj)
    z = torch.from_numpy(g[f"{name}/z"]).requires_grad_(True)
    G = torch.from_numpy(g[f"{name}/G"])
    kw = dict(rtol=1e-5, atol=1e-6) if method == "dopri5" else {}
    sol = m.solve_ODE(z, **kw)
    (sol * G).sum().backward()
    key = f"{name}/{method}/{adj}"
    assert np.array_equal(sol.detach().numpy(), g[f"{key}/sol"])
    # The reference evaluates the hidden layer twice (prod / degr share it); autograd then sums the two
    # branches in a different order than the port's single evaluation: gradients agree to f32 rounding.  In
    # dopri5 adjoint mode that rounding also moves the backward solve's adaptive step sizes (the error norm spans the parameter adjoints too): accept/reject flips -> 5e-3.
    scale = 5e-3 if (method == "dopri5" and adj) else 2e-6
    tol = dict(rtol=0, atol=scale * max(1.0, float(np.abs(g[f"{key}/grad_z"]).max())))
    assert np.allclose(z.grad.numpy(), g[f"{key}/grad_z"], **tol)
    for k, p in m.named_parameters():
        if ".prod." in k or ".degr." in k:
            continue
        want = g[f"{key}/g/{k}"]
        assert np.allclose(p.grad.numpy(), want, rtol=0, atol=scale * max(1.0, float(np.abs(want).max()))), k
    if method == "dopri5":
        if not adj:  # in adjoint mode last_stats belongs to the last backward sub-solve
            assert np.array_equal(np.array(tde.last_stats.accepted), g[f"{key}/accepted"])
            assert np.allclose(np.array(tde.last_stats.dts), g[f"{key}/dts"], rtol=1e-12)

def test_default_adjoint_drops_grad_z_through_dynamics(golden_dir):
    """F5: with odeint_adjoint z only gets gradient through x0 = latent_to_ode_net(z)."""
    g = _load(golden_dir)
    assert not np.allclose(g["cvs/midpoint/0/grad_z"], g["cvs/midpoint/1/grad_z"], atol=1e-4)
    m = _port_from_golden(g, "cvs", "midpoint", 1)
    z = torch.from_numpy(g["cvs/z"]).requires_grad_(True)
    G = torch.from_numpy(g["cvs/G"])
    x0 = m.latent_to_ode_net(z)
    f = slode_port.OdeFunc(z, m.dynamics)
    sol = tde.odeint_adjoint(f, x0, m.times, method="midpoint").permute(1, 0, 2)
    gx0, = torch.autograd.grad((sol * G).sum(), x0, retain_graph=True)
    gz_via_x0, = torch.autograd.grad(x0, z, gx0)
    assert np.allclose(gz_via_x0.numpy(), g["cvs/midpoint/1/grad_z"], rtol=0, atol=1e-5)


@pytest.mark.skipif(not shims.reference_available(), reason="/root/reference only exists in the build container")
@pytest.mark.parametrize("method,adj", [("midpoint", True), ("rk4", False)])
def test_port_equals_real_reference_classes(method, adj):
    bb, dec = shims.import_reference_blackbox()
    times = torch.arange(0.0, 20.0, 1.0)
    torch.manual_seed(3)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = bb.OdeModel()
        ref.init_with_params(times=times, ode_state_dim=5, latent_dim=15, ode_hidden_dim=25, adjoint_solver=adj,
                             solver=method, device="cpu")
    port = slode_port.OdeModel(times, 5, 15, 25, adj, method)
    port.load_state_dict(ref.state_dict())  # strict: same keys, incl. prod.* / degr.* aliases
    z = torch.randn(7, 15)
    a = ref.solve_ODE(z)
    b = port.solve_ODE(z)
    assert a.shape == b.shape == (7, 20, 5)
    assert torch.equal(a, b)


@pytest.mark.skipif(not shims.reference_available(), reason="/root/reference only exists in the build container")
def test_quantile_heads_equal_reference_decoder():
    bb, dec = shims.import_reference_blackbox()
    times = torch.arange(0.0, 12.0, 1.0)
    cfg = shims._Munch(obs_dim=3, system_input_dim=0, ode_state_dim=5, ode_hidden_dim=25, adjoint_solver=False,
                       solver="midpoint", constant_std=1e-2, seq_len=12)
    torch.manual_seed(4)
    with contextlib.redirect_stdout(io.StringIO()):
        try:
            ref = dec.Decoder(config=cfg, latent_dim=15, times=times, device="cpu")
        except TypeError:
            pytest.skip("Decoder constructor signature differs")
    port_ode = slode_port.OdeModel(times, 5, 15, 25, False, "midpoint")
    port_ode.load_state_dict(ref.ode_model.state_dict())
    heads = slode_port.QuantileHeads(port_ode, 3, 12)
    heads.output_q50.load_state_dict(ref.output_q50.state_dict())
    heads.output_q75.load_state_dict(ref.output_q75.state_dict())
    heads.output_q25.load_state_dict(ref.output_q25.state_dict())
    z = torch.randn(5, 15)
    out_ref = ref.forward(z)
    sol, q75, q50, q25, std = heads(z)
    flat = [o for o in out_ref if torch.is_tensor(o)]
    assert any(torch.equal(o, q50) for o in flat if o.shape == q50.shape)
    assert any(torch.equal(o, sol) for o in flat if o.shape == sol.shape)
