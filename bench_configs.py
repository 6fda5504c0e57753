#!/usr/bin/env python
"""Secondary measurements over the other BASELINE.json configs (the contract line is bench.py; this script writes
one JSON line per case, for profiles/):

  configs[0]  CVS default: midpoint, odeint_adjoint semantics, B=128, T=86
  configs[1]  mechanistic CVS variant: 2^20 trajectories x 100 times, rk4, f32 (HBM-bound: GB/s reported)
  configs[2]  challenge shape, dopri5 (batch-global controller), B in {35, 7000, 2^20}
  configs[3]  proc shape (L=50, S=8, non-uniform t), B = 312 wells x {1, 200, 4096} samples, midpoint + rk4

Timing: CUDA events, best of `reps` after 2 warm-ups; every tensor is device resident.
"""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import structured_latent_odes_b200 as slode  # noqa: E402
from structured_latent_odes_b200 import torchdiffeq_api as api  # noqa: E402

dev = "cuda"


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    best = 1e30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(reps):
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def blackbox(name, B, T, L, H, S, method, adjoint, times=None, rtol=1e-5, atol=1e-6, grad=True):
    torch.manual_seed(12)
    t = torch.arange(0.0, T, 1.0, device=dev) if times is None else times.to(dev)
    m = slode.OdeModel()
    m.init_with_params(t, S, L, H, adjoint, method, dev)
    m = m.to(dev)
    z = torch.randn(B, L, device=dev)
    G = torch.randn(T, B, S, device=dev).permute(1, 0, 2)

    def fwd_bwd():
        m.zero_grad(set_to_none=True)
        if method == "dopri5":
            f = m.gen_dynamics(z)
            sol = slode.odeint(f, m.initialize_state(z), t, method="dopri5", rtol=rtol, atol=atol).permute(1, 0, 2)
        else:
            sol = m.solve_ODE(z)
        sol.backward(G)

    def fwd_only():
        with torch.no_grad():
            if method == "dopri5":
                slode.odeint(m.gen_dynamics(z), m.initialize_state(z), t, method="dopri5", rtol=rtol, atol=atol)
            else:
                m.solve_ODE(z)

    ms_f = timed(fwd_only)
    out = {"case": name, "B": B, "T": T, "L": L, "H": H, "S": S, "method": method,
           "gradient": "odeint_adjoint" if adjoint else "discrete", "fwd_ms": round(ms_f, 4),
           "fwd_traj_steps_per_s": B * (T - 1) / (ms_f * 1e-3)}
    if method == "dopri5":
        st = api.last_dopri5_stats
        out.update(rtol=rtol, atol=atol, accepted=st.n_accept, rejected=st.n_reject, rhs_evals=st.n_rhs)
    if grad:
        ms = timed(fwd_bwd)
        out.update(fwd_bwd_ms=round(ms, 4), fwd_bwd_traj_steps_per_s=B * (T - 1) / (ms * 1e-3))
    print(json.dumps(out), flush=True)


def cvs_mech(B, T=100):
    g = torch.Generator(device=dev).manual_seed(12)
    ie = torch.where(torch.rand(B, device=dev, generator=g) < 0.5, -2.0, 0.0)
    rm = torch.where(torch.rand(B, device=dev, generator=g) < 0.5, 0.5, 0.0)
    f = slode.CvsMechanistic(ie, rm, learn_constants=True)
    t = torch.arange(0.0, T, 1.0, device=dev)
    y0 = torch.ones(B, 4, device=dev, requires_grad=True)
    G = torch.randn(T, B, 4, device=dev)

    def fwd():
        with torch.no_grad():
            slode.odeint(f, y0, t, method="rk4")

    def fwd_bwd():
        f.zero_grad(set_to_none=True)
        y0.grad = None
        slode.odeint(f, y0, t, method="rk4").backward(G)

    ms_f, ms = timed(fwd), timed(fwd_bwd)
    bytes_f = B * 4 * (T * 4 + 4 + 2)
    bytes_b = B * 4 * (2 * T * 4 + 4 + 4)
    print(json.dumps({"case": "configs[1] mechanistic CVS rk4 f32", "B": B, "T": T, "fwd_ms": round(ms_f, 4),
                      "fwd_bwd_ms": round(ms, 4), "fwd_GBps": bytes_f / (ms_f * 1e-3) / 1e9,
                      "bwd_GBps": bytes_b / ((ms - ms_f) * 1e-3) / 1e9,
                      "fwd_bwd_traj_steps_per_s": B * (T - 1) / (ms * 1e-3)}), flush=True)

    def gen():
        slode.generate_cvs_latents(ie[:1000], rm[:1000])
    print(json.dumps({"case": "CVS generator (f64 rk4 x8 substeps), 1000 x 86 like cvs_data.py", "ms": round(timed(gen), 4)}),
          flush=True)


def heads(B, T=100, S=5, O=3, NQ=3):
    """Decoder heads (models/decoders.py:45-47): all NQ quantile heads in one pass, forward and backward."""
    sol = torch.randn(B, T, S, device=dev, requires_grad=True)
    W = [torch.randn(O, S, device=dev, requires_grad=True) for _ in range(NQ)]
    G = [torch.randn(B, O, T, device=dev) for _ in range(NQ)]

    def fwd():
        with torch.no_grad():
            slode.decoder_heads(sol, W)

    def fwd_bwd():
        sol.grad = None
        for w in W:
            w.grad = None
        mu = slode.decoder_heads(sol, W)
        torch.autograd.backward(mu, G)

    ms_f, ms = timed(fwd), timed(fwd_bwd)
    bytes_f = B * T * 4 * (S + NQ * O)
    bytes_b = B * T * 4 * (2 * S + NQ * O)
    print(json.dumps({"case": f"decoder heads, {NQ} heads x obs_dim {O}", "B": B, "T": T, "fwd_ms": round(ms_f, 4),
                      "fwd_bwd_ms": round(ms, 4), "fwd_GBps": bytes_f / (ms_f * 1e-3) / 1e9,
                      "bwd_GBps": bytes_b / ((ms - ms_f) * 1e-3) / 1e9}), flush=True)


def predict(B, T=100, L=15, H=25, S=5, O=3, method="rk4", gaussian=False, times=None, layout="bts"):
    """Reconstruction / posterior sampling (SURVEY f2 + f3): Decoder.forward under no_grad (solver kernel + heads
    kernel, the trajectories written and read back) against Decoder.predict (heads in the solver's epilogue, the
    trajectories never in HBM)."""
    import types
    cfg = types.SimpleNamespace(obs_dim=O, ode_state_dim=S, ode_hidden_dim=H, adjoint_solver=False, solver=method,
                                constant_std=1e-2)
    t = (torch.arange(T, dtype=torch.float32) if times is None else times).to(dev)
    torch.manual_seed(12)
    dec = (slode.GaussianDecoder if gaussian else slode.Decoder)(cfg, t, L, dev).to(dev)
    dec.ode_model.layout = layout
    z = torch.randn(B, L, device=dev, generator=torch.Generator(device=dev).manual_seed(12))

    def two():
        with torch.no_grad():
            dec(z)

    ms2 = timed(two)
    ms1 = timed(lambda: dec.predict(z))
    ms1s = timed(lambda: dec.predict(z, want_solution=True))
    heads_w = ((dec.output_mean[0].weight,) if gaussian else
               (dec.output_q50[0].weight, dec.output_q75[0].weight, dec.output_q25[0].weight))
    ms_k = timed(lambda: dec.ode_model.solve_ODE_heads(z, heads_w))                       # the fused call alone
    ms_kc = timed(lambda: dec.ode_model.solve_ODE_heads(z, heads_w, contiguous=True))     # unpadded (B,O,T) rows
    with torch.no_grad():
        ms_s = timed(lambda: dec.ode_model.solve_ODE(z))                                  # the solve alone
    nq = 1 if gaussian else 3
    print(json.dumps({"case": f"predict (no_grad decoder), {nq} heads x obs_dim {O}, (L,H,S)=({L},{H},{S}), {method}, "
                              f"sol layout {layout}", "B": B, "T": T,
                      "two_kernels_ms": round(ms2, 4), "fused_ms": round(ms1, 4), "fused_with_sol_ms": round(ms1s, 4),
                      "solve_heads_call_ms": round(ms_k, 4), "solve_heads_call_contiguous_rows_ms": round(ms_kc, 4),
                      "solve_only_call_ms": round(ms_s, 4),
                      "fused_out_GBps": B * T * 4 * nq * O / (ms1 * 1e-3) / 1e9,
                      "trajectory_steps_per_s_fused": B * (T - 1) / (ms1 * 1e-3)}), flush=True)


def tensor_core_proxy(B, Hw, S=5, T=100):
    """BASELINE configs[4] asks where the dense head contraction should move from the FMA pipe to tcgen05.  The kernels
    do not run that contraction at all any more (piecewise-linear heads: O(S) per evaluation + O(S) per relu crossing),
    so the question becomes: how long would the dense contraction ALONE take on the tensor cores?  Measured with
    cuBLAS as the tensor-core proxy (library code, used here only as a yardstick): per MLP evaluation one
    [B x H] . [H x 16] product (2S = 10 head outputs padded to 16), TF32 (x3 for the error-compensated 3xTF32 split a
    1e-5 parity needs) and plain fp32, times the 3(T-1)+1 evaluations of an rk4 solve.  The hidden activations
    relu(w1t t + c) would still have to be produced per evaluation (B x H FMAs and a B x H x 4 byte operand write): the
    figure is a LOWER bound on a tensor-core forward."""
    A = torch.randn(B, Hw, device=dev)
    W = torch.randn(Hw, 16, device=dev)
    evals = 3 * (T - 1) + 1
    out = {"case": f"configs[4] dense-head contraction on tensor cores (cuBLAS proxy), H={Hw}", "B": B, "H": Hw,
           "evals_per_rk4_solve": evals}
    for name, tf32 in (("tf32", True), ("fp32", False)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        ms = timed(lambda: torch.matmul(A, W))
        out[f"gemm_{name}_ms"] = round(ms, 4)
    torch.backends.cuda.matmul.allow_tf32 = False
    out["dense_heads_3xtf32_ms_per_solve_lower_bound"] = round(3 * out["gemm_tf32_ms"] * evals, 2)
    out["dense_heads_tf32_relaxed_ms_per_solve_lower_bound"] = round(out["gemm_tf32_ms"] * evals, 2)
    out["dense_heads_fp32_ms_per_solve_lower_bound"] = round(out["gemm_fp32_ms"] * evals, 2)
    print(json.dumps(out), flush=True)


def predict_cases(big):
    predict(big)
    predict(big, layout="tbs")
    predict(big, gaussian=True)
    predict(big, method="midpoint")
    predict(35 * 200, T=142, O=4, method="midpoint")                    # challenge: 35 series x 200 posterior samples
    tt = torch.cat([torch.zeros(1), torch.cumsum(0.193 + 0.003 * torch.rand(99, generator=torch.Generator().manual_seed(7)), 0)])
    predict(312 * 200, L=50, S=8, O=4, method="midpoint", times=tt)     # proc: 312 wells x 200 samples
    predict(312 * 4096, L=50, S=8, O=4, method="midpoint", times=tt)


if __name__ == "__main__":
    big = 1 << 20
    if "predict" in sys.argv[1:]:
        predict_cases(big)
        sys.exit(0)
    blackbox("configs[0] CVS default", 128, 86, 15, 25, 5, "midpoint", True)
    blackbox("configs[0] CVS full train set", 810, 86, 15, 25, 5, "midpoint", True)
    cvs_mech(big)
    heads(big)
    heads(big, NQ=1)
    predict_cases(big)
    for B in (35, 7000, big):
        blackbox("configs[2] challenge dopri5", B, 142, 15, 25, 5, "dopri5", False, rtol=1e-5, atol=1e-6)
    blackbox("configs[2] challenge dopri5 torchdiffeq default tol", 7000, 142, 15, 25, 5, "dopri5", False, rtol=1e-7, atol=1e-9,
             grad=False)
    blackbox("configs[2] challenge midpoint (shipped solver)", 7000, 142, 15, 25, 5, "midpoint", True)
    tt = torch.cat([torch.zeros(1), torch.cumsum(0.193 + 0.003 * torch.rand(99, generator=torch.Generator().manual_seed(7)), 0)])
    for ns in (1, 200, 4096):
        for method in ("midpoint", "rk4"):
            blackbox(f"configs[3] proc 312 wells x {ns} samples", 312 * ns, 100, 50, 25, 8, method, method == "midpoint", times=tt)
    for Hw in (16, 25, 32, 64, 128, 256, 512):  # configs[4]: hidden-width sweep (every compiled width)
        Sw = 4 if Hw == 16 else 5
        for B in (1 << 10, 1 << 16, big):
            blackbox(f"configs[4] width sweep H={Hw}", B, 100, 15, Hw, Sw, "rk4", False)
        if Hw >= 64:
            tensor_core_proxy(big, Hw)
    blackbox("configs[1] blackbox rk4 (bench.py workload)", big, 100, 15, 25, 5, "rk4", False)
    blackbox("configs[1] blackbox midpoint adjoint", big, 100, 15, 25, 5, "midpoint", True)
