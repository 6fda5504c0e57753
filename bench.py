#!/usr/bin/env python
"""Benchmark of the SLODE latent-ODE hot path on B200 (contract: see the task statement / DESIGN.md section 5).

    python bench.py --gpus 1 --steps K --warmup W                      # this repo's sm_100a path
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus N --steps K --warmup W     # reference algorithm on the host cores

Workload (BASELINE.json configs[1], blackbox variant, SURVEY.md section 8d "Config 2"):
    2^20 trajectories per GPU x 100 output times, fp32, rk4 (torchdiffeq's 3/8 rule), L=15 H=25 S=5,
    forward solve + reverse sweep (exact discrete adjoint == torchdiffeq.odeint + autograd).
metric = ODE trajectory-steps/s (fwd+bwd); one trajectory-step = one trajectory advanced across one output interval.

One JSON line on stdout (rank 0).  ``value``: inputs resident in HBM.  ``e2e``: the same solve through the public
API with the step's inputs (latents z and observations y) in pinned HOST memory, copied in inside the timed
region, and the loss + flat parameter gradients read back.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "ode_trajectory_steps_per_s_fwd_bwd"
UNIT = "trajectory-steps/s"
L, H, S, O, T = 15, 25, 5, 3, 100
# fp32 FMA-pipe peak: 148 SMs x 128 lanes x 2 flop x 1.965 GHz; the FFMA2 micro-benchmark in
# profiles/r01/fp32_pipes_microbench.jsonl measures 74.0 TFLOP/s (127.2 of 128 FMA/clk/SM) on this pool's B200.
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12


def algorithmic_flops(method: str, T: int, H: int, S: int):
    """(forward, backward) fp32 flop per TRAJECTORY of the work the path needs:
    hidden layer with the time-invariant part hoisted (2H per MLP evaluation), both heads (4HS), the sigmoid's add
    (2S), f = A - D x per stage (2S), the Butcher combinations; MLP evaluations per step: euler 1, midpoint 2,
    rk4 3 (+1 per solve: the 3/8 rule's last stage time is the next step's first).  Backward = stage recompute +
    the adjoint recurrences + the prefix-sum weight-gradient bookkeeping."""
    steps = T - 1
    stages = {"euler": 1, "midpoint": 2, "rk4": 4}[method]
    evals = {"euler": steps, "midpoint": 2 * steps, "rk4": 3 * steps + 1}[method]
    mlp = 2 * H + 4 * H * S + 2 * S
    combine = {"euler": 2 * S, "midpoint": 4 * S, "rk4": 16 * S}[method]
    prologue = 2 * L * H * 2 + 2 * H * S  # c = z W1z^T + b1, x0 net (once per trajectory)
    fwd = evals * mlp + steps * (stages * 2 * S + combine) + prologue
    # per stage: cotangents of the head pre-activations (10S), P/Q prefix sums (8S), adjoint update (6S);
    # per hidden unit: two snapshots (gate flip + end of sweep) of 2 * (4*2S + 3*2S) flop.
    # ... plus the fused epilogue: dz and the outer products dc z^T, da z^T, db relu(ha)^T (once per trajectory)
    bwd = fwd + steps * stages * 24 * S + H * 2 * (14 * 2 * S) + 2 * (4 * L * H + 2 * H * S)
    return float(fwd), float(bwd)


def algorithmic_bytes(T: int, H: int, S: int, method: str = "rk4", ckpt: bool = False):
    """(forward, backward) HBM bytes per trajectory of the fused path: fwd reads z (L), writes sol (T*S);
    bwd reads sol + grad_sol + z, writes grad_z.  With evaluation checkpoints the forward also writes, and the
    reverse sweep reads, 2S floats per MLP evaluation."""
    evals = {"euler": T - 1, "midpoint": 2 * (T - 1), "rk4": 3 * (T - 1) + 1}[method]
    ck = 4.0 * evals * 2 * S if ckpt else 0.0
    return 4.0 * (L + T * S) + ck, 4.0 * (2 * T * S + 2 * L) + ck


def checkpointed_bwd_flops(method: str, T: int, H: int, S: int):
    """Reverse sweep with evaluation checkpoints: hidden-layer gates (2H per evaluation), stage recompute from the
    stored sigmoids, adjoint recurrences, prefix sums, epilogue -- no head products."""
    steps = T - 1
    stages = {"euler": 1, "midpoint": 2, "rk4": 4}[method]
    evals = {"euler": steps, "midpoint": 2 * steps, "rk4": 3 * steps + 1}[method]
    combine = {"euler": 2 * S, "midpoint": 4 * S, "rk4": 16 * S}[method]
    return float(evals * 2 * H + steps * (stages * 2 * S + combine) + steps * stages * 24 * S + H * 2 * (14 * 2 * S)
                 + 2 * (4 * L * H + 2 * H * S))



def pl_flops(method: str, T: int, H: int, S: int):
    """(forward, backward) fp32 flop per trajectory of the piecewise-linear formulation the kernels run (DESIGN.md
    section 4): per evaluation the heads are alpha*t + beta (2 * 2S) and the sigmoid adds/merged reciprocals
    (4 * 2S); per relu crossing (at most H per solve, counted as H) one rank-one update of (alpha, beta)
    (2 * 2 * 2S) and once per trajectory the dense initial coefficients (2 * 2 * 2S * H)."""
    steps = T - 1
    stages = {"euler": 1, "midpoint": 2, "rk4": 4}[method]
    evals = {"euler": steps, "midpoint": 2 * steps, "rk4": 3 * steps + 1}[method]
    combine = {"euler": 2 * S, "midpoint": 4 * S, "rk4": 16 * S}[method]
    prologue = 2 * L * H * 2 + 2 * H * S
    pl = evals * (4 * S + 8 * S) + H * 8 * S + 8 * S * H + 2 * H
    fwd = pl + steps * (stages * 2 * S + combine) + prologue
    bwd = fwd + steps * stages * 24 * S + H * 2 * (14 * 2 * S) + 2 * (4 * L * H + 2 * H * S)
    return float(fwd), float(bwd)


def sfu_ops(method: str, T: int, S: int):
    """MUFU lane-operations per trajectory and solve: 2S sigmoids per evaluation, each one ex2 and -- two
    denominators sharing one reciprocal -- half an rcp."""
    evals = {"euler": T - 1, "midpoint": 2 * (T - 1), "rk4": 3 * (T - 1) + 1}[method]
    return float(evals * 3 * S)


XU_PEAK_TOPS = 148 * 16 * 1.965e9 / 1e12  # 16 MUFU lanes per SM and clock

# ---------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
            0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index: int, period=0.01):
        self.samples, self.reasons, self.power = [], set(), []
        self.stop_flag = threading.Event()
        self.period = period
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, repr(e)
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop_flag.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    self.nv, "nvmlDeviceGetCurrentClocksEventReasons") else self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.BITS.items():
                    if r & bit:
                        self.reasons.add(name)
                self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            self.stop_flag.wait(self.period)

    def __enter__(self):
        if self.nv is not None:
            self.thread.start()
        return self

    def __exit__(self, *exc):
        self.stop_flag.set()
        if self.nv is not None:
            self.thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "power_w_max": round(max(self.power), 1) if self.power else None}


# ---------------------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------------------
def seeded_port_model(method, adjoint):
    """CPU legs only (cpu_baseline / --impl reference): the oracle port of the reference classes with
    reference-architecture weights under torch.manual_seed(12) (config_cvs.py:28)."""
    from oracle import slode_port
    torch.manual_seed(12)
    times = torch.arange(0.0, T, 1.0)
    return slode_port.OdeModel(times, S, L, H, adjoint, method)


def cpu_reference_step(model, heads_w, z, y):
    """One fwd+bwd of the reference algorithm on the host: x0 net, torchdiffeq-style solve, q50 head, MSE, backward."""
    model.zero_grad()
    sol = model.solve_ODE(z)
    loss = ((sol @ heads_w.t()) - y).square().mean()
    loss.backward()
    return float(loss.detach())


def time_cpu_reference(method, adjoint, B, reps, warmup=1):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = seeded_port_model(method, adjoint)
    g = torch.Generator().manual_seed(12)
    z = torch.randn(B, L, generator=g)
    y = torch.rand(B, T, O, generator=g)
    Wq = torch.randn(O, S, generator=g) * 0.3
    for _ in range(warmup):
        cpu_reference_step(model, Wq, z, y)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu_reference_step(model, Wq, z, y)
        times.append(time.perf_counter() - t0)
    return times, cores


def run_reference_arm(args):
    """The reference's own algorithm (oracle port of models/blackbox_ode.py + torchdiffeq restatement; the real
    package cannot be installed: torchdiffeq / pyro are absent and there is no network) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.ref_batch
    times, cores = time_cpu_reference(args.method, args.adjoint, B, args.steps, warmup=max(1, min(args.warmup, 2)))
    total = sum(times)
    value = B * (T - 1) * len(times) / total
    sample = f"{B} of 2^20 trajectories per step, T={T}, {args.method}, fwd+bwd, {len(times)} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, B_per_gpu=B, n=1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, B_per_gpu, n):
    return {"workload": "configs[1] blackbox: 2^20 trajectories/GPU x 100 obs times, fp32 rk4(3/8) fwd+bwd",
            "trajectories_per_gpu": B_per_gpu, "trajectories_total": B_per_gpu * n, "obs_times": T, "latent_dim": L, "sol_layout": getattr(args, "layout", "tbs"),
            "ode_hidden_dim": H, "ode_state_dim": S, "solver": args.method,
            "gradient": "odeint_adjoint emulation" if args.adjoint else "discrete adjoint (odeint + autograd parity)",
            "mlp_evaluation": "piecewise-linear heads (alpha t + beta per trajectory, updated at relu crossings)",
            "reverse_sweep": ("reads the forward's evaluation checkpoints (12.5 GB per 2^20 x 100 solve)"
                              if (getattr(args, "eval_ckpt", False) and not args.adjoint) else "re-evaluates the MLP"),
            "parallelism": f"trajectory-sharded x{n}, one flat all-reduce of parameter gradients",
            "l2": "inputs larger than L2 (sol / grad_sol are 2.1 GB each per GPU vs 126 MB L2)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1 << 20, help="trajectories per GPU")
    ap.add_argument("--method", default="rk4", choices=["euler", "midpoint", "rk4"])
    ap.add_argument("--adjoint", action="store_true", help="odeint_adjoint gradient semantics instead of discrete")
    ap.add_argument("--ref-batch", type=int, default=8192, help="trajectories per CPU step (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=16)
    ap.add_argument("--eval-ckpt", action="store_true",
                    help="reverse sweep reads the forward's evaluation checkpoints instead of re-evaluating the MLP")
    ap.add_argument("--layout", default="tbs", choices=["tbs", "bts"],
                    help="storage of the resident step's solution: (T,B,S) torchdiffeq's, or (B,T,S) the decoder's")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch.distributed as dist
    import structured_latent_odes_b200 as slode
    from structured_latent_odes_b200 import _cabi, sharding
    from structured_latent_odes_b200.torchdiffeq_api import KernelTimer
    from structured_latent_odes_b200 import torchdiffeq_api as _api_cfg
    _api_cfg.EVAL_CHECKPOINTS = bool(args.eval_ckpt)  # the library default (None) picks by state width

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner on the C-level stdout when the communicator is created: send that to
        # stderr so that stdout carries exactly the one JSON line
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    _cabi.lib()

    B = args.batch
    torch.manual_seed(12)  # reference seed (config_cvs.py:28); identical weights on every rank
    model = slode.OdeModel()
    model.init_with_params(times=torch.arange(0.0, T, 1.0, device=dev), ode_state_dim=S, latent_dim=L,
                           ode_hidden_dim=H, adjoint_solver=args.adjoint, solver=args.method, device=dev,
                           layout=args.layout)
    model = model.to(dev)
    params = [p for k, p in model.named_parameters()]
    reducer = sharding.FlatGradReducer(params)

    g = torch.Generator(device=dev).manual_seed(12 + rank)
    z = torch.randn(B, L, device=dev, generator=g)
    # upstream dL/dsol, resident, in the same storage layout as the solution
    G = (torch.randn(B, T, S, device=dev, generator=g) if args.layout == "bts"
         else torch.randn(T, B, S, device=dev, generator=g).permute(1, 0, 2))
    launches = [0]
    lib = _cabi.lib()

    def step_resident():
        model.zero_grad(set_to_none=True)
        sol = model.solve_ODE(z)
        sol.backward(G)
        reducer.reduce()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident()
    sync_all()
    launches[0] = lib.slode_query(_cabi.Q_TOTAL_LAUNCHES)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks, KernelTimer() as kt:
        sync_all()
        e0.record()
        for _ in range(args.steps):
            step_resident()
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        ktimes = kt.summary()
    tmax = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total = float(tmax.item())
    value = world * B * (T - 1) * args.steps / (ms_total * 1e-3)
    n_launches = lib.slode_query(_cabi.Q_TOTAL_LAUNCHES) - launches[0]

    # ---- end to end: pinned host inputs -> loss + gradients back on the host --------------------------------
    gc = torch.Generator().manual_seed(100 + rank)
    z_host = torch.randn(B, L, generator=gc).pin_memory()
    y_host = torch.rand(B, O, T, generator=gc).pin_memory()   # observations as batch_to_device hands them over
    Wq = (torch.randn(O, S, generator=torch.Generator().manual_seed(7)) * 0.3).to(dev)
    nchunk = max(1, args.e2e_chunks)
    bounds = [sharding.shard_bounds(B, i, nchunk) for i in range(nchunk)]
    copy_stream = torch.cuda.Stream()
    out_host = torch.empty(1 + reducer.numel, dtype=torch.float32).pin_memory()
    zbuf = [torch.empty(bounds[0][1] - bounds[0][0], L, device=dev) for _ in range(2)]
    ybuf = [torch.empty(bounds[0][1] - bounds[0][0], O, T, device=dev) for _ in range(2)]

    def step_e2e():
        """Chunked, double-buffered: chunk k+1 is copied on the copy stream while chunk k is solved."""
        model.layout = "bts"  # the decoder heads read the solution (B,T,S)-contiguous
        model.zero_grad(set_to_none=True)
        main = torch.cuda.current_stream()
        loss_acc = torch.zeros((), device=dev)
        ready = [None, None]
        free = [None, None]

        def issue(i):
            lo, hi = bounds[i]
            s = i & 1
            with torch.cuda.stream(copy_stream):
                if free[s] is not None:
                    copy_stream.wait_event(free[s])
                zbuf[s][: hi - lo].copy_(z_host[lo:hi], non_blocking=True)
                ybuf[s][: hi - lo].copy_(y_host[lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            ready[s] = ev

        issue(0)
        for i in range(nchunk):
            lo, hi = bounds[i]
            s = i & 1
            if i + 1 < nchunk:
                issue(i + 1)
            main.wait_event(ready[s])
            zc, yc = zbuf[s][: hi - lo], ybuf[s][: hi - lo]
            sol = model.solve_ODE(zc)
            (mu,) = slode.decoder_heads(sol, (Wq,))          # q50 head, (n,O,T) like Decoder.forward
            loss = (mu - yc).square().sum() / (B * T * O)
            loss.backward()
            loss_acc += loss.detach()
            ev = torch.cuda.Event()
            ev.record(main)
            free[s] = ev
        flat = reducer.reduce()
        out_host[:1].copy_(loss_acc.reshape(1), non_blocking=True)
        out_host[1:].copy_(flat, non_blocking=True)
        main.synchronize()  # the caller reads the loss: the step ends when it is on the host

    for _ in range(3):
        step_e2e()
    sync_all()
    model.layout = args.layout
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    sync_all()
    t_e2e = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * B * (T - 1) * args.steps / (float(t_e2e.item()) * 1e-3)
    h2d = z_host.numel() * 4 + y_host.numel() * 4
    d2h = out_host.numel() * 4

    if rank == 0:
        from structured_latent_odes_b200 import torchdiffeq_api as _api
        ckpt = bool(_api.EVAL_CHECKPOINTS) and not args.adjoint
        ff, fb = pl_flops(args.method, T, H, S)
        if ckpt:
            fb = checkpointed_bwd_flops(args.method, T, H, S)
        xu_f = sfu_ops(args.method, T, S)
        xu_b = 0.0 if ckpt else xu_f
        bf, bb = algorithmic_bytes(T, H, S, args.method, ckpt)
        fwd_ms = sum(ktimes["fwd"]) / max(len(ktimes["fwd"]), 1)
        bwd_ms = sum(ktimes["bwd"]) / max(len(ktimes["bwd"]), 1)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        traffic_map = {}
        try:
            traffic_map = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:
            pass
        achieved_b = B * fb / (bwd_ms * 1e-3) / 1e12
        achieved_f = B * ff / (fwd_ms * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, B, world),
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": float(t_e2e.item()) / args.steps,
                    "what": f"pinned z (B,{L}) + observations (B,{O},{T}) -> solve -> q50 head (slode_heads) + MSE -> backward -> "
                            f"loss + {reducer.numel} parameter gradients on the host; {nchunk} double-buffered chunks"},
            "gpu_launches": n_launches,
        }
        hbm_src = "measured" if peaks else "fallback"
        rl_b = {"kernel": "mlp_fixed_bwd_kernel (reverse sweep" + (", evaluation checkpoints)" if ckpt else ")"),
                "ms_per_launch": bwd_ms, "algorithmic_flop_per_launch": B * fb, "algorithmic_bytes_per_launch": B * bb,
                "fp32": {"achieved": achieved_b, "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved_b / FP32_PEAK_TFLOPS},
                "hbm": {"achieved": B * bb / (bwd_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": B * bb / (bwd_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src},
                "xu": {"achieved": B * xu_b / (bwd_ms * 1e-3) / 1e12, "peak": XU_PEAK_TOPS, "unit": "T MUFU lane-op/s",
                       "frac": B * xu_b / (bwd_ms * 1e-3) / 1e12 / XU_PEAK_TOPS, "sfu_ops_per_launch": B * xu_b}}
        rl_f = {"kernel": "mlp_fixed_fwd_kernel", "ms_per_launch": fwd_ms, "algorithmic_flop_per_launch": B * ff,
                "algorithmic_bytes_per_launch": B * bf,
                "fp32": {"achieved": achieved_f, "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved_f / FP32_PEAK_TFLOPS},
                "hbm": {"achieved": B * bf / (fwd_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": B * bf / (fwd_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src},
                "xu": {"achieved": B * xu_f / (fwd_ms * 1e-3) / 1e12, "peak": XU_PEAK_TOPS, "unit": "T MUFU lane-op/s",
                       "frac": B * xu_f / (fwd_ms * 1e-3) / 1e12 / XU_PEAK_TOPS, "sfu_ops_per_launch": B * xu_f}}
        fp32_src = ("148 SM x 128 lanes x 2 x 1.965 GHz; FFMA2 micro-benchmark measured 74.0 "
                    "(profiles/r01/fp32_pipes_microbench.jsonl); MEASURED_PEAKS.json has no fp32 entry")

        xu_src = "148 SM x 16 MUFU lanes x 1.965 GHz (no MUFU entry in MEASURED_PEAKS.json)"

        def flat(r, traffic_key):
            """contract shape: the roofline that binds the kernel on top (largest fraction of its peak among HBM
            bytes, fp32 FMA flop and XU/MUFU operations), the other two beside it"""
            bound = max(("hbm", "fp32", "xu"), key=lambda k: r[k]["frac"])
            top = r[bound]
            out = {"bound": bound, "kernel": r["kernel"], "achieved": top["achieved"], "peak": top["peak"],
                   "unit": top["unit"], "frac": top["frac"], "traffic": traffic_map.get(traffic_key),
                   "peak_source": {"hbm": top.get("peak_source"), "fp32": fp32_src, "xu": xu_src}[bound],
                   "ms_per_launch": r["ms_per_launch"], "algorithmic_flop_per_launch": r["algorithmic_flop_per_launch"],
                   "algorithmic_bytes_per_launch": r["algorithmic_bytes_per_launch"], "note": r.get("note")}
            for k in ("hbm", "fp32", "xu"):
                if k != bound:
                    out[k] = r[k]
            return out

        # measured context for the fractions above (ncu --set full of the same kernels, profiles/r01/INDEX.md)
        rl_b["note"] = ("re-evaluating sweep: bound by none of the three rooflines -- warp-instruction issue/latency "
                        "(ncu: 2.10e9 warp instructions, 0.41 issued per cycle and scheduler at 2 warps per scheduler, "
                        "254 registers; XU 24 %, FMA 22 %, DRAM 20 % of peak)" if not ckpt else
                        "checkpointed sweep: HBM read stream of the stored evaluations (ncu: DRAM 59 % of peak, "
                        "long_scoreboard 23 % of stall samples)")
        rl_f["note"] = ("piecewise-linear heads: 15 MUFU lane-operations per trajectory and evaluation; ncu: XU pipe "
                        "55 % of peak, FMA 21 %, mio_throttle + short_scoreboard 1.3 stall cycles per issue")
        tag = f"{args.method}_{int(args.adjoint)}" + ("_ckpt" if ckpt else "")
        dominant_is_bwd = bwd_ms >= fwd_ms
        line["roofline"] = flat(rl_b if dominant_is_bwd else rl_f, ("bwd_" if dominant_is_bwd else "fwd_") + tag)
        line["roofline_other"] = flat(rl_f if dominant_is_bwd else rl_b, ("fwd_" if dominant_is_bwd else "bwd_") + tag)
        if not args.no_cpu_baseline and world == 1:
            reps = 3
            times, cores = time_cpu_reference(args.method, args.adjoint, args.ref_batch, reps)
            v = args.ref_batch * (T - 1) * reps / sum(times)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{args.ref_batch} of 2^20 trajectories, T={T}, {args.method}, fwd+bwd "
                                              f"(x0 net, solve, q50 head + MSE, backward), best-effort all host threads, "
                                              f"{reps} reps after 1 warm-up"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
