"""Host-side mirror of the reference interface (``structured_latent_odes_b200``): names, shapes, state_dict
keys and error behaviour -- everything that can be checked without a GPU."""
import contextlib
import io

import pytest
import torch

import structured_latent_odes_b200 as slode
from oracle import shims, slode_port


def _model(L=15, H=25, S=5, T=10, adjoint=True, solver="midpoint"):
    m = slode.OdeModel()
    m.init_with_params(times=torch.arange(0.0, T, 1.0), ode_state_dim=S, latent_dim=L, ode_hidden_dim=H,
                       adjoint_solver=adjoint, solver=solver, device="cpu")
    return m


def test_two_phase_construction_and_state_dict_keys_match_the_port():
    m = _model()
    port = slode_port.OdeModel(m.times, 5, 15, 25, True, "midpoint")
    assert sorted(m.state_dict()) == sorted(port.state_dict())
    m.load_state_dict(port.state_dict())


@pytest.mark.skipif(not shims.reference_available(), reason="/root/reference only exists in the build container")
def test_state_dict_round_trips_with_the_real_reference_class():
    bb, _ = shims.import_reference_blackbox()
    with contextlib.redirect_stdout(io.StringIO()):
        ref = bb.OdeModel()
        ref.init_with_params(times=torch.arange(0.0, 10.0), ode_state_dim=5, latent_dim=15, ode_hidden_dim=25,
                             adjoint_solver=True, solver="midpoint", device="cpu")
    m = _model()
    m.load_state_dict(ref.state_dict())
    ref.load_state_dict(m.state_dict())
    # the eager module API (one RHS evaluation) agrees with the reference's
    z = torch.randn(3, 15)
    x = torch.rand(3, 5)
    t = torch.tensor(0.7)
    assert torch.allclose(m.gen_dynamics(z)(t, x), ref.gen_dynamics(z)(t, x), atol=1e-7)
    assert torch.equal(m.initialize_state(z), ref.initialize_state(z))


def test_func_recognition():
    m = _model()
    f = m.gen_dynamics(torch.randn(4, 15))
    assert slode.is_blackbox_func(f)
    assert not slode.is_blackbox_func(torch.nn.Linear(2, 2))
    pf = slode_port.OdeFunc(torch.randn(4, 15), slode_port.Dynamics(15, 25, 5))
    assert slode.is_blackbox_func(pf)  # any module with the reference's attribute names qualifies


def test_no_cpu_fallback():
    m = _model()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.solve_ODE(torch.randn(4, 15))


def test_validation_errors_match_torchdiffeq_behaviour():
    m = _model()
    f = m.gen_dynamics(torch.randn(4, 15))
    y0 = torch.rand(4, 5)
    t = torch.arange(0.0, 5.0)
    with pytest.raises(ValueError, match="Invalid method"):
        slode.odeint(f, y0, t, method="rk45")
    with pytest.raises(NotImplementedError):
        slode.odeint(f, (y0, y0), t, method="rk4")
    with pytest.raises(NotImplementedError):
        slode.odeint(f, y0, t, method="rk4", event_fn=lambda t, y: y)
    with pytest.raises(ValueError, match="nn.Module"):
        slode.odeint_adjoint(lambda t, y: y, y0, t, method="rk4")
    with pytest.raises(NotImplementedError):
        slode.odeint_adjoint(f, y0, t, method="rk4", adjoint_method="euler")


def test_install_as_torchdiffeq():
    import sys
    old = sys.modules.get("torchdiffeq")
    try:
        slode.install_as_torchdiffeq()
        import torchdiffeq
        assert torchdiffeq.odeint is slode.odeint and torchdiffeq.odeint_adjoint is slode.odeint_adjoint
    finally:
        if old is None:
            sys.modules.pop("torchdiffeq", None)
        else:
            sys.modules["torchdiffeq"] = old


def test_tanh_hidden_activation_is_rejected_not_silently_wrong():
    from structured_latent_odes_b200 import torchdiffeq_api as api
    dyn = slode.Dynamics(15, 25, 5)  # default hidden_activation=nn.Tanh like the reference's signature
    f = slode.OdeFunc(torch.randn(2, 15), dyn)
    with pytest.raises(NotImplementedError, match="ReLU"):
        api._check_blackbox(f)
