"""baseline/_ref recipe (CPU side): the copied reference sources are unmodified and import behind the shims."""
import contextlib
import io

import pytest
import torch

from baseline import make_ref, ref_shims
from oracle import slode_port, torchdiffeq_oracle


@pytest.mark.skipif(not make_ref.available(), reason="baseline/_ref is created by build() where /root/reference exists")
def test_real_classes_from_baseline_ref_equal_the_port():
    bb, dec = ref_shims.import_real(torchdiffeq_oracle)
    times = torch.arange(0.0, 15.0, 1.0)
    torch.manual_seed(5)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = bb.OdeModel()
        ref.init_with_params(times=times, ode_state_dim=5, latent_dim=15, ode_hidden_dim=25, adjoint_solver=False,
                             solver="rk4", device="cpu")
    port = slode_port.OdeModel(times, 5, 15, 25, False, "rk4")
    port.load_state_dict(ref.state_dict())
    z = torch.randn(9, 15)
    assert torch.equal(ref.solve_ODE(z), port.solve_ODE(z))
    assert hasattr(dec, "Decoder") and hasattr(dec, "GaussianDecoder")


def test_make_ref_is_idempotent_and_reports_availability():
    dest = make_ref.make()
    assert (dest is not None) == make_ref.available()
