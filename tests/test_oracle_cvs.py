"""Pin the CVS mechanistic oracle (``oracle/cvs_mech.py``) on the reference's own golden trajectories
(``data/cvs/test_latent_data.pkl`` / ``gt_test_data.pkl`` / ``test_params_data.pkl``; a slice is committed as
``tests/golden/cvs_golden.npz`` by ``tests/golden/make_golden.py``)."""
import os

import numpy as np
import torch

from oracle import cvs_mech
from oracle import torchdiffeq_oracle as tde


def _golden(golden_dir):
    return np.load(os.path.join(golden_dir, "cvs_golden.npz"))


def test_lsoda_reproduces_reference_golden_latents(golden_dir):
    g = _golden(golden_dir)
    n = 12
    lat = cvs_mech.lsoda_trajectories(g["i_ext"][:n], g["r_tpr_mod"][:n])
    assert lat.shape == (n, 86, 4)
    assert np.abs(lat - g["latent"][:n]).max() < 1e-9
    assert np.abs(cvs_mech.observe(lat) - g["gt"][:n]).max() < 1e-9


def test_initial_state_and_treatment_values(golden_dir):
    g = _golden(golden_dir)
    assert np.array_equal(g["latent"][:, 0, :], np.ones((24, 4)))
    assert set(np.unique(g["i_ext"])) <= {-2.0, 0.0}
    assert set(np.unique(g["r_tpr_mod"])) <= {0.0, 0.5}


def test_torch_rhs_equals_numpy_rhs():
    rng = np.random.default_rng(0)
    x = rng.uniform(0.3, 1.5, size=(16, 4))
    ie = rng.choice([0.0, -2.0], size=16)
    rm = rng.choice([0.0, 0.5], size=16)
    want = cvs_mech.cvs_rhs(x, ie, rm)
    got = cvs_mech.CvsRhsTorch(torch.tensor(ie), torch.tensor(rm))(torch.tensor(0.0), torch.tensor(x))
    assert np.abs(got.numpy() - want).max() < 1e-15


def test_rk4_on_torch_rhs_tracks_lsoda_golden(golden_dir):
    """Fixed-step 3/8 RK4 at dt=1 (the grid the CUDA path uses) stays within 1e-4 of the LSODA goldens --
    this is the accuracy the f32 kernel is later held to against the f64 oracle."""
    g = _golden(golden_dir)
    f = cvs_mech.CvsRhsTorch(torch.tensor(g["i_ext"]), torch.tensor(g["r_tpr_mod"]))
    t = torch.arange(0.0, 86.0, 1.0, dtype=torch.float64)
    sol = tde.odeint(f, torch.ones(24, 4, dtype=torch.float64), t, method="rk4").permute(1, 0, 2)
    assert (sol.numpy() - g["latent"]).__abs__().max() < 1e-4
