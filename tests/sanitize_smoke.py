"""Smoke-sized forward + reverse sweep of every compiled shape and gradient mode (plus the decoder heads, the CVS
mechanistic kernels and dopri5), meant to run under compute-sanitizer on the GPU box:

    compute-sanitizer --tool memcheck  python tests/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tests/sanitize_smoke.py

(one tool per gpurun call; logs are committed under profiles/).  Shapes are tiny so that the instrumented run takes
seconds; ragged batch sizes exercise the masked tail threads."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch

import slode_testutil as U
import structured_latent_odes_b200 as slode

only = sys.argv[1:]  # optional shape filter
torch.manual_seed(0)
n = 0
for shape in ["cvs", "proc", "small", "h32", "h64", "h128", "h512"]:
    if only and shape not in only:
        continue
    L, H, S, times = U.SHAPES[shape]
    times = times[:7]
    for method in ["euler", "midpoint", "rk4"]:
        for adjoint in [False, True]:
            o = U.make_oracle(shape, method, adjoint)
            o.times = times
            p = U.make_product(o)
            for B in (37, 130):
                z = torch.randn(B, L, device="cuda")
                G = torch.randn(B, len(times), S, device="cuda")
                U.run_fwd_bwd(p, z, G)
                n += 1
    # (B,T,S) storage and the decoder heads
    cfg = type("C", (dict,), {"__getattr__": dict.__getitem__})(obs_dim=3, system_input_dim=0, ode_state_dim=S,
                                                                 ode_hidden_dim=H, adjoint_solver=True, solver="midpoint",
                                                                 constant_std=1e-2, seq_len=len(times))
    if S in (4, 5, 8):
        dec = slode.Decoder(cfg, times.cuda(), L, "cuda").cuda()
        out = dec(torch.randn(33, L, device="cuda", requires_grad=True))
        sum(o_.sum() for o_ in out).backward()
        n += 1
if not only or "cvs_mech" in only:
    f = slode.CvsMechanistic(torch.where(torch.rand(50, device="cuda") > 0.5, 0.0, -2.0),
                             torch.where(torch.rand(50, device="cuda") > 0.5, 0.0, 0.5)).cuda()
    y0 = torch.ones(50, 4, device="cuda", requires_grad=True)
    sol = slode.odeint(f, y0, torch.arange(0.0, 6.0, device="cuda"), method="rk4")
    sol.sum().backward()
    n += 1
if not only or "dopri5" in only:
    o = U.make_oracle("cvs", "dopri5", False)
    o.times = torch.arange(0.0, 4.0)
    p = U.make_product(o)
    z = torch.randn(40, 15, device="cuda")
    U.run_fwd_bwd(p, z, torch.randn(40, 4, 5, device="cuda"))
    n += 1
torch.cuda.synchronize()
print(f"sanitize_smoke: {n} cases ran")
