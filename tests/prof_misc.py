"""One launch each of the kernels beside the fixed-grid pair (for ncu): CVS mechanistic fwd/bwd at 2^20 x 100, the
decoder heads fwd/bwd at 2^20 x 100, the dopri5 forward and its odeint_adjoint backward at the challenge shape.
python tests/prof_misc.py [cvs|heads|dopri5|predict]   (predict: Decoder.predict, heads fused into the solver kernel, 2^20 x 100 rk4)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch

import structured_latent_odes_b200 as slode

what = sys.argv[1] if len(sys.argv) > 1 else "cvs"
dev = "cuda"
torch.manual_seed(12)
if what == "cvs":
    B, T = 1 << 20, 100
    ie = torch.where(torch.rand(B, device=dev) < 0.5, -2.0, 0.0)
    rm = torch.where(torch.rand(B, device=dev) < 0.5, 0.5, 0.0)
    f = slode.CvsMechanistic(ie, rm, learn_constants=True)
    y0 = torch.ones(B, 4, device=dev, requires_grad=True)
    slode.odeint(f, y0, torch.arange(0.0, T, 1.0, device=dev), method="rk4").backward(torch.randn(T, B, 4, device=dev))
elif what == "heads":
    B, T, S, O = 1 << 20, 100, 5, 3
    sol = torch.randn(B, T, S, device=dev, requires_grad=True)
    W = [torch.randn(O, S, device=dev, requires_grad=True) for _ in range(3)]
    mu = slode.decoder_heads(sol, W)
    (mu[0].sum() + mu[1].sum() + mu[2].sum()).backward()
elif what == "predict":
    import types
    B, T = 1 << 20, 100
    cfg = types.SimpleNamespace(obs_dim=3, ode_state_dim=5, ode_hidden_dim=25, adjoint_solver=False, solver="rk4",
                                constant_std=1e-2)
    dec = slode.Decoder(cfg, torch.arange(T, dtype=torch.float32, device=dev), 15, dev).to(dev)
    dec.predict(torch.randn(B, 15, device=dev))
else:
    import slode_testutil as U
    o = U.make_oracle("chal", "dopri5", True)
    p = U.make_product(o)
    B = 7000
    z = torch.randn(B, 15, device=dev)
    y0 = p.initialize_state(z).detach().requires_grad_(True)
    sol = slode.odeint_adjoint(p.gen_dynamics(z), y0, p.times, method="dopri5", rtol=1e-5, atol=1e-6)
    sol.backward(torch.randn_like(sol))
    from structured_latent_odes_b200 import torchdiffeq_api as api
    st = api.last_dopri5_adjoint_stats
    print("adjoint backward: accepted", st.n_accept, "rejected", st.n_reject, "rhs evaluations per trajectory", st.n_rhs)
torch.cuda.synchronize()
print("done", what)
