"""Randomised parity sweep against the oracle (diagnostic; run on the GPU box: python tests/stress_parity.py [n])."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import random
import torch
import slode_testutil as U
from oracle import slode_port

n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = random.Random(0)
worst = 0.0
for it in range(n_iter):
    L, H, S = rng.choice([(15, 25, 5), (50, 25, 8), (6, 16, 4), (15, 32, 5), (9, 64, 5)])
    T = rng.choice([2, 3, 5, 17, 60])
    B = rng.choice([1, 2, 3, 63, 64, 65, 255, 257, 700])
    method = rng.choice(["euler", "midpoint", "rk4"])
    adjoint = rng.choice([False, True])
    layout = rng.choice(["tbs", "bts"])
    scale = rng.choice([0.3, 1.0, 3.0])     # wider weights -> more relu gate flips inside the window
    kind = rng.choice(["uniform", "ragged", "reverse"])
    g = torch.Generator().manual_seed(it)
    if kind == "uniform":
        times = torch.arange(0.0, T, 1.0) * rng.choice([0.1, 1.0])
    else:
        times = torch.cat([torch.zeros(1), torch.cumsum(0.05 + torch.rand(T - 1, generator=g), 0)])
        if kind == "reverse":
            times = times.flip(0).contiguous()
    torch.manual_seed(it)
    o = slode_port.OdeModel(times, S, L, H, adjoint, method)
    with torch.no_grad():
        o.dynamics.dynamics_hidden.weight.mul_(scale)
        o.dynamics.dynamics_hidden.bias.normal_(0, 0.5 * scale)
    p = U.make_product(o, layout=layout)
    z = torch.randn(B, L, generator=g)
    G = torch.randn(B, T, S, generator=g)
    so, gzo, gro = U.run_fwd_bwd(o, z, G)
    sp, gzp, grp = U.run_fwd_bwd(p, z.cuda(), G.cuda())
    errs = [U.rel_err(sp, so), U.rel_err(gzp, gzo)] + [U.rel_err(grp[k], gro[k]) for k in gro]
    w = max(errs)
    flag = ""
    if w >= 2e-5:  # is the fp32 oracle itself that far from float64 here (cancellation)?  then it is not ours
        o64 = slode_port.OdeModel(times.double(), S, L, H, adjoint, method).double()
        o64.load_state_dict({k: v.double() for k, v in o.state_dict().items()})
        s64, gz64, gr64 = U.run_fwd_bwd(o64, z.double(), G.double())
        ref_noise = max([U.rel_err(so, s64), U.rel_err(gzo, gz64)] + [U.rel_err(gro[k], gr64[k]) for k in gro])
        ours = max([U.rel_err(sp, s64), U.rel_err(gzp, gz64)] + [U.rel_err(grp[k], gr64[k]) for k in gro])
        flag = f"   (vs f64: ours {ours:.1e}, fp32 oracle {ref_noise:.1e})" + ("" if ours < 3 * ref_noise else "   <-- LOOK")
    worst = max(worst, w)
    print(f"{it:3d} L{L} H{H} S{S} T{T} B{B} {method:8s} adj={int(adjoint)} {layout} {kind:8s} scale={scale}: worst {w:.2e}{flag}", flush=True)
print("worst overall", worst)
