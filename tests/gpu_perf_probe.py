"""Quick device timing of the fixed-grid fwd / bwd kernels (diagnostic; bench.py is the real harness)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import structured_latent_odes_b200 as slode

def run(B, T, L, H, S, method, adjoint, layout="tbs", reps=5):
    dev = "cuda"
    torch.manual_seed(12)
    m = slode.OdeModel(); m.init_with_params(torch.arange(0., T, 1., device=dev), S, L, H, adjoint, method, dev, layout=layout)
    m = m.to(dev)
    z = torch.randn(B, L, device=dev)
    G = torch.randn(B, T, S, device=dev) if layout == "bts" else torch.randn(T, B, S, device=dev).permute(1, 0, 2)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    best = [1e9, 1e9]
    for r in range(reps + 2):
        m.zero_grad()
        zz = z.clone().requires_grad_(True)
        torch.cuda.synchronize()
        ev[0].record()
        sol = m.solve_ODE(zz)
        ev[1].record()
        sol.backward(G)
        ev[2].record()
        torch.cuda.synchronize()
        if r >= 2:
            best[0] = min(best[0], ev[0].elapsed_time(ev[1])); best[1] = min(best[1], ev[1].elapsed_time(ev[2]))
    steps = B * (T - 1)
    print(json.dumps(dict(B=B, T=T, L=L, H=H, S=S, method=method, adjoint=adjoint, layout=layout, fwd_ms=round(best[0], 3),
                          bwd_ms=round(best[1], 3), traj_steps_per_s=steps / ((best[0] + best[1]) * 1e-3))), flush=True)

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    for method in ["rk4", "midpoint", "euler"]:
        for adjoint in [False, True]:
            run(B, 100, 15, 25, 5, method, adjoint)
    run(B, 100, 15, 25, 5, "rk4", False, layout="bts")
    run(B // 4, 100, 50, 25, 8, "rk4", False)
