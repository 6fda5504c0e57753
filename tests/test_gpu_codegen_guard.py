"""Code-generation guard: the (64,5) fixed-grid translation unit -- the instantiation with the largest
register-resident sorting network and spills -- is built twice, at the default ptxas -O3 and at -O1
(``_build.GUARD_LIB``).  Both builds must produce the same trajectories and gradients on the device; a difference is a
ptxas code-generation problem (round 1 met one in exactly such an instantiation), not an algorithmic one."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

_SCRIPT = r"""
import sys, torch
sys.path.insert(0, {tests!r}); sys.path.insert(0, {root!r})
import slode_testutil as U
out = {{}}
for method, adjoint in (("midpoint", False), ("midpoint", True), ("rk4", False), ("euler", False)):
    L, H, S, times = U.SHAPES["h64"]
    o = U.make_oracle("h64", method, adjoint)
    p = U.make_product(o)
    g = torch.Generator().manual_seed(17)
    z = torch.randn(300, L, generator=g)
    G = torch.randn(300, len(times), S, generator=g)
    sol, gz, grads = U.run_fwd_bwd(p, z.cuda(), G.cuda())
    out[(method, adjoint)] = (sol.cpu(), gz.cpu(), {{k: v.cpu() for k, v in grads.items()}})
torch.save(out, {dst!r})
"""


def _run(lib, dst):
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    env = dict(os.environ)
    if lib:
        env["SLODE_B200_LIB"] = lib
    else:
        env.pop("SLODE_B200_LIB", None)
    code = _SCRIPT.format(tests=here, root=root, dst=dst)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return torch.load(dst)


def test_h64_unit_gives_the_same_results_at_ptxas_O1_and_O3(tmp_path):
    from structured_latent_odes_b200 import _build
    if not os.path.isfile(_build.GUARD_LIB):
        pytest.fail(f"{_build.GUARD_LIB} is missing: run __graft_entry__.build()")
    a = _run(None, str(tmp_path / "o3.pt"))
    b = _run(_build.GUARD_LIB, str(tmp_path / "o1.pt"))
    assert a.keys() == b.keys()
    for key in a:
        (sa, za, ga), (sb, zb, gb) = a[key], b[key]
        # trajectories: same instruction semantics at both levels -> bit-equal; the parameter gradients are summed with
        # atomics across warps / blocks, so their last bits depend on timing
        assert torch.equal(sa, sb), key
        assert torch.equal(za, zb), key
        for k in ga:
            err = (ga[k] - gb[k]).abs().max().item() / max(gb[k].abs().max().item(), 1e-30)
            assert err < 1e-5, (key, k, err)
