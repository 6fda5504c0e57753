"""Prints parity numbers of the CUDA path against the fp32 and fp64 oracle (diagnostic, run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import slode_testutil as U

torch.manual_seed(0)
for shape in ["cvs", "proc", "small", "h32"]:
    for method in ["euler", "midpoint", "rk4"]:
        for adjoint in [False, True]:
            L, H, S, times = U.SHAPES[shape]
            B = 200
            o32 = U.make_oracle(shape, method, adjoint)
            o64 = U.make_oracle(shape, method, adjoint, dtype=torch.float64)
            o64.load_state_dict({k: v.double() for k, v in o32.state_dict().items()})
            p = U.make_product(o32)
            g = torch.Generator().manual_seed(3)
            z = torch.randn(B, L, generator=g)
            G = torch.randn(B, len(times), S, generator=g)
            s32, gz32, gr32 = U.run_fwd_bwd(o32, z, G)
            s64, gz64, gr64 = U.run_fwd_bwd(o64, z.double(), G.double())
            sp, gzp, grp = U.run_fwd_bwd(p, z.cuda(), G.cuda())
            line = f"{shape:5s} {method:8s} adj={int(adjoint)} sol: p-o32 {U.rel_err(sp, s32):.2e} p-o64 {U.rel_err(sp, s64):.2e} o32-o64 {U.rel_err(s32, s64):.2e}"
            if gz64 is not None:
                line += f" | gz: p-o32 {U.rel_err(gzp, gz32):.2e} p-o64 {U.rel_err(gzp, gz64):.2e} o32-o64 {U.rel_err(gz32, gz64):.2e}"
            worst = max(gr64, key=lambda k: U.rel_err(grp[k], gr64[k]))
            line += f" | worst param {worst.split('.')[-2][:12]}.{worst.split('.')[-1][0]}: p-o32 {U.rel_err(grp[worst], gr32[worst]):.2e} p-o64 {U.rel_err(grp[worst], gr64[worst]):.2e} o32-o64 {U.rel_err(gr32[worst], gr64[worst]):.2e}"
            print(line, flush=True)
