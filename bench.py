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


def algorithmic_bytes(T: int, H: int, S: int, method: str = "rk4"):
    """(forward, backward) HBM bytes per trajectory of the fused path: fwd reads z (L), writes sol (T*S);
    bwd reads sol + grad_sol + z, writes grad_z."""
    return 4.0 * (L + T * S), 4.0 * (2 * T * S + 2 * L)


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


def sfu_ops(method: str, T: int, S: int, merged: bool):
    """MUFU lane-operations per trajectory and solve: 2S sigmoids per evaluation, each one ex2 and one rcp -- half an
    rcp in the forward kernel, where two denominators share one reciprocal."""
    evals = {"euler": T - 1, "midpoint": 2 * (T - 1), "rk4": 3 * (T - 1) + 1}[method]
    return float(evals * (3 if merged else 4) * S)


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


def cpu_reference_step(model, heads_w, z, y, backward=True):
    """One step of the reference algorithm on the host: x0 net, torchdiffeq-style solve, q50 head, MSE (, backward)."""
    model.zero_grad()
    if not backward:
        with torch.no_grad():
            sol = model.solve_ODE(z)
            return float(((sol @ heads_w.t()) - y).square().mean())
    sol = model.solve_ODE(z)
    loss = ((sol @ heads_w.t()) - y).square().mean()
    loss.backward()
    return float(loss.detach())


def time_cpu_reference(method, adjoint, B, reps, warmup=1, backward=True):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = seeded_port_model(method, adjoint)
    g = torch.Generator().manual_seed(12)
    z = torch.randn(B, L, generator=g)
    y = torch.rand(B, T, O, generator=g)
    Wq = torch.randn(O, S, generator=g) * 0.3
    for _ in range(warmup):
        cpu_reference_step(model, Wq, z, y, backward)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu_reference_step(model, Wq, z, y, backward)
        times.append(time.perf_counter() - t0)
    return times, cores


def cpu_baseline_table(method, budget_s=25.0):
    """the port at B in {128, 4096, 65536}, forward only and forward+backward, both gradient
    branches (adjoint_solver False = autograd through the solver, True = odeint_adjoint), within a time budget."""
    rows, spent = [], 0.0
    for B in (128, 4096, 65536):
        for adjoint in (False, True):
            for backward in (True, False):
                if spent > budget_s:
                    continue
                t0 = time.perf_counter()
                ts, cores = time_cpu_reference(method, adjoint, B, reps=2 if B < 65536 else 1, warmup=1, backward=backward)
                spent += time.perf_counter() - t0
                rows.append({"B": B, "adjoint_solver": adjoint, "pass": "fwd+bwd" if backward else "fwd",
                             "s": round(min(ts), 5), "trajectory_steps_per_s": B * (T - 1) / min(ts)})
    return rows


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
    sample = (f"{B} of 2^20 trajectories per step (per-trajectory-step rates are batch-size independent above ~4k rows: "
              f"see cpu_baseline.table of the b200 arm), T={T}, {args.method}, fwd+bwd, {len(times)} steps; oracle PORT of the "
              "reference classes (evaluates the hidden layer once per RHS call, the real classes twice: conservative)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the b200 arm's config (the workload both arms are quoted on); what this arm actually ran per step is the
        # bounded sample described in cpu_baseline.sample
        "config": workload_config(args, B_per_gpu=args.batch, n=args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "sample_trajectories_per_step": B},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, B_per_gpu, n):
    return {"workload": "configs[1] blackbox: 2^20 trajectories/GPU x 100 obs times, fp32 rk4(3/8) fwd+bwd",
            "trajectories_per_gpu": B_per_gpu, "trajectories_total": B_per_gpu * n, "obs_times": T, "latent_dim": L,
            "sol_layout": getattr(args, "layout", "tbs"), "ode_hidden_dim": H, "ode_state_dim": S, "solver": args.method,
            "gradient": "odeint_adjoint emulation" if args.adjoint else "discrete adjoint (odeint + autograd parity)",
            "parallelism": f"trajectory-sharded x{n}, one flat all-reduce of parameter gradients",
            "l2": "inputs larger than L2 (sol / grad_sol are 2.1 GB each per GPU vs 126 MB L2)"}


IMPLEMENTATION = {   # how the b200 arm computes the workload above (not part of the workload: the reference arm differs)
    "mlp_evaluation": "piecewise-linear heads (alpha t + beta per trajectory, updated at relu crossings)",
    "reverse_sweep": "re-evaluates the heads from the stored grid states (nothing else is checkpointed)",
    "thread_mapping": "one trajectory per thread, fp32x2 over state pairs",
}


def pin_to_gpu_numa_node(index):
    """Bind this process to the CPUs NVML reports as local to the GPU BEFORE pinned buffers are allocated (first
    touch then places them on that NUMA node); returns a description for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = (ncpu + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus local to gpu {index}"
    except Exception as e:  # pragma: no cover
        return "unavailable: " + repr(e)[:80]
    return "unavailable"


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
    ap.add_argument("--no-epoch", action="store_true", help="skip the CVS training-epoch timing")
    ap.add_argument("--e2e-chunks", type=int, default=16)
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

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    numa = pin_to_gpu_numa_node(local_rank)
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
    launches0 = lib.slode_query(_cabi.Q_TOTAL_LAUNCHES)
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
    n_launches = lib.slode_query(_cabi.Q_TOTAL_LAUNCHES) - launches0
    del G

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

    resident = {}   # device copies of the host inputs, for the labelled second figure (dataset resident in HBM)

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

        if not resident:
            issue(0)
        for i in range(nchunk):
            lo, hi = bounds[i]
            s = i & 1
            if resident:
                zc, yc = resident["z"][lo:hi], resident["y"][lo:hi]
            else:
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

    def step_copy_only():
        """The step's host->device bytes alone (same chunking, same pinned buffers): the bound e2e cannot beat."""
        for i in range(nchunk):
            lo, hi = bounds[i]
            s = i & 1
            zbuf[s][: hi - lo].copy_(z_host[lo:hi], non_blocking=True)
            ybuf[s][: hi - lo].copy_(y_host[lo:hi], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def timed(fn, n):
        sync_all()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        sync_all()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / n

    for _ in range(3):
        step_e2e()
    model.layout = args.layout
    with ClockSampler(local_rank) as clocks_e2e:   # the resident region is short at small --steps: sample this one too
        e2e_ms = timed(step_e2e, args.steps)
    model.layout = args.layout
    copy_ms = timed(step_copy_only, max(3, min(args.steps, 10)))   # all ranks copy at the same time, like the step
    e2e_value = world * B * (T - 1) / (e2e_ms * 1e-3)
    h2d = z_host.numel() * 4 + y_host.numel() * 4
    d2h = out_host.numel() * 4
    # second, labelled figure: the same public-API step (solve -> head -> loss -> backward -> loss and gradients on the
    # host) with the dataset already resident in HBM -- how the reference's own loop works at its scale (the whole
    # training set lives in memory and batch_to_device moves a mini-batch, training_cvs.py:58-67); no H2D in the step
    resident["z"], resident["y"] = z_host.to(dev), y_host.to(dev)
    for _ in range(2):
        step_e2e()
    e2e_res_ms = timed(step_e2e, args.steps)
    resident.clear()
    model.layout = args.layout
    del z_host, y_host, zbuf, ybuf

    # ---- the other half of BASELINE.json's metric: SLODE train-epoch time (CVS, reference batch sizes) ----------
    epoch = None
    if not args.no_epoch:
        epoch = time_cvs_epoch(dev, rank, world, sync_all)

    if rank == 0:
        ff, fb = pl_flops(args.method, T, H, S)
        xu_f, xu_b = sfu_ops(args.method, T, S, merged=True), sfu_ops(args.method, T, S, merged=False)
        bf, bb = algorithmic_bytes(T, H, S, args.method)
        fwd_ms = sum(ktimes["fwd"]) / max(len(ktimes["fwd"]), 1)
        bwd_ms = sum(ktimes["bwd"]) / max(len(ktimes["bwd"]), 1)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback"
        counters, traffic_map = {}, {}
        try:
            counters = json.load(open(os.path.join(ROOT, "profiles", "counters.json")))
            traffic_map = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:
            pass
        mid = {"euler": 0, "midpoint": 1, "rk4": 2}[args.method]
        fp32_src = ("148 SM x 128 lanes x 2 x 1.965 GHz; FFMA2 micro-benchmark measured 74.0 "
                    "(profiles/r01/fp32_pipes_microbench.jsonl); MEASURED_PEAKS.json has no fp32 entry")
        xu_src = ("148 SM x 16 MUFU lanes x 1.965 GHz; ex2-only / rcp-only micro-benchmarks in "
                  "profiles/r02/mufu_microbench.jsonl; MEASURED_PEAKS.json has no MUFU entry")

        def roof(kernel_key, label, ms_launch, flop_alg, bytes_alg, xu_alg):
            """One kernel under the three rooflines (HBM bytes, fp32 FMA-pipe flop, XU/MUFU operations); the largest
            fraction is reported on top.  `executed` figures come from the ncu opcode counters of the same kernel at
            the same size (profiles/counters.json, produced by profiles/tools/ncu_counters.py) divided by THIS run's
            launch duration."""
            c = counters.get(kernel_key, {})
            sec = ms_launch * 1e-3
            r = {"hbm": {"achieved": B * bytes_alg / sec / 1e9, "peak": hbm_peak, "unit": "GB/s", "peak_source": hbm_src},
                 "fp32": {"achieved": B * flop_alg / sec / 1e12, "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s", "peak_source": fp32_src,
                          "count": "algorithmic (flop model of the piecewise-linear formulation)"},
                 "xu": {"achieved": B * xu_alg / sec / 1e12, "peak": XU_PEAK_TOPS, "unit": "T MUFU lane-op/s", "peak_source": xu_src}}
            scale = B / float(1 << 20)   # counters were captured at 2^20 trajectories
            if c:
                r["fp32"] = {"achieved": scale * c["fp32_flop_executed"] / sec / 1e12, "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s",
                             "peak_source": fp32_src, "count": "executed (ncu sass__inst_executed_per_opcode: FFMA2/FMUL2/FADD2/FFMA/FMUL/FADD)",
                             "algorithmic_achieved": B * flop_alg / sec / 1e12}
                r["xu"]["executed_achieved"] = scale * c["mufu_lane_ops_executed"] / sec / 1e12
            for v in r.values():
                v["frac"] = v["achieved"] / v["peak"]
            bound = max(r, key=lambda k: r[k]["frac"])
            top = r[bound]
            out = {"bound": bound, "kernel": label, "achieved": top["achieved"], "peak": top["peak"], "unit": top["unit"],
                   "frac": top["frac"], "traffic": traffic_map.get(kernel_key), "peak_source": top["peak_source"],
                   "ms_per_launch": ms_launch, "algorithmic_flop_per_launch": B * flop_alg,
                   "algorithmic_bytes_per_launch": B * bytes_alg}
            for k in r:
                if k != bound:
                    out[k] = r[k]
            if c:
                out["ncu"] = {k: c.get(k) for k in (
                    "inst_executed", "smsp__issue_active.avg.per_cycle_active", "launch__registers_per_thread",
                    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
                    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "gpu__time_duration.sum", "report")}
                # the packed fp32x2 forms with three register operands issue at 85 FMA/clk/SM, not 128
                # (profiles/r01/fp32_pipes_microbench.jsonl: ffma2_rrr): the pipe's busy CYCLES are what bounds it
                out["fma_pipe_cycles_active_frac"] = (c.get("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed") or 0) / 100.0
            return out

        mode = 1 if args.adjoint else 0
        rl_b = roof(f"fixed_bwd_kernel<{H}, {S}, {mid}, {mode}>", "fixed_bwd_kernel (reverse sweep)", bwd_ms, fb, bb, xu_b)
        rl_f = roof(f"fixed_fwd_kernel<{H}, {S}, {mid}>", "fixed_fwd_kernel (forward solve)", fwd_ms, ff, bf, xu_f)
        xu_step_ms = B * (xu_f + xu_b) / (XU_PEAK_TOPS * 1e12) * 1e3
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, B, world),
            "implementation": IMPLEMENTATION,
            "clocks": dict(clocks.summary(), e2e_region=clocks_e2e.summary()),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms, "copy_bound_ms": copy_ms, "frac_of_copy_bound": copy_ms / e2e_ms,
                    "copy_bound_gbs_per_gpu": h2d / (copy_ms * 1e-3) / 1e9, "pinned_numa": numa,
                    "resident_dataset": {"value": world * B * (T - 1) / (e2e_res_ms * 1e-3), "ms_per_step": e2e_res_ms,
                                         "what": "the same step with z and the observations already in HBM (no H2D); "
                                                 "loss + gradients still read back every step"},
                    "what": f"pinned z (B,{L}) + observations (B,{O},{T}) -> solve -> q50 head (slode_heads) + MSE -> backward -> "
                            f"loss + {reducer.numel} parameter gradients on the host; {nchunk} double-buffered chunks; "
                            "copy_bound_ms = the same pinned->device copies alone, all ranks at once"},
            "gpu_launches": n_launches,
            "kernel_ms_series": {k: {"first": [round(x, 3) for x in v[:4]], "last": [round(x, 3) for x in v[-4:]]}
                                 for k, v in ktimes.items()},
            "step_frac_of_xu_bound": xu_step_ms / (ms_total / args.steps),
        }
        dominant_is_bwd = bwd_ms >= fwd_ms
        line["roofline"] = rl_b if dominant_is_bwd else rl_f
        line["roofline_other"] = rl_f if dominant_is_bwd else rl_b
        if epoch is not None:
            line["epoch"] = epoch
        if not args.no_epoch:
            line["cvs_mechanistic"] = time_cvs_mechanistic(dev, B, hbm_peak)
            line["posterior_predict"] = time_posterior_predict(dev, B, args.method)
        if not args.no_cpu_baseline and world == 1:
            reps = 3
            times, cores = time_cpu_reference(args.method, args.adjoint, args.ref_batch, reps)
            v = args.ref_batch * (T - 1) * reps / sum(times)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{args.ref_batch} of 2^20 trajectories, T={T}, {args.method}, fwd+bwd "
                                              f"(x0 net, solve, q50 head + MSE, backward), all host threads, "
                                              f"{reps} reps after 1 warm-up; table: B in (128, 4096, 65536) x both "
                                              "gradient branches x (fwd, fwd+bwd)",
                                    "table": cpu_baseline_table(args.method)}
            if epoch is not None:
                line["cpu_baseline"]["epoch_s"] = time_cvs_epoch_cpu()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def time_cvs_mechanistic(dev, B, hbm_peak):
    """The other right-hand side BASELINE configs[1] names: the CVS mechanistic dynamics (data/cvs/cvs_data.py:52-91) as
    a forward(t, state) module through the same odeint call, same trajectory count and grid, rk4, fp32 (rank 0's GPU)."""
    import structured_latent_odes_b200 as slode
    g = torch.Generator(device=dev).manual_seed(12)
    ie = torch.where(torch.rand(B, device=dev, generator=g) < 0.5, -2.0, 0.0)
    rm = torch.where(torch.rand(B, device=dev, generator=g) < 0.5, 0.5, 0.0)
    f = slode.CvsMechanistic(ie, rm, learn_constants=True)
    t = torch.arange(0.0, T, 1.0, device=dev)
    y0 = torch.ones(B, 4, device=dev, requires_grad=True)
    G = torch.randn(T, B, 4, device=dev, generator=g)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    best = [1e30, 1e30]
    for r in range(6):
        f.zero_grad(set_to_none=True)
        y0.grad = None
        torch.cuda.synchronize()
        ev[0].record()
        sol = slode.odeint(f, y0, t, method="rk4")
        ev[1].record()
        sol.backward(G)
        ev[2].record()
        torch.cuda.synchronize()
        if r >= 2:
            best = [min(best[0], ev[0].elapsed_time(ev[1])), min(best[1], ev[1].elapsed_time(ev[2]))]
    bytes_f, bytes_b = B * 4 * (T * 4 + 6), B * 4 * (2 * T * 4 + 8)
    return {"fwd_ms": best[0], "bwd_ms": best[1], "value": B * (T - 1) / ((best[0] + best[1]) * 1e-3), "unit": UNIT,
            "fwd_hbm_frac": bytes_f / (best[0] * 1e-3) / 1e9 / hbm_peak, "bwd_hbm_frac": bytes_b / (best[1] * 1e-3) / 1e9 / hbm_peak,
            "what": f"CvsMechanistic RHS, {B} trajectories x {T} times, rk4, fp32, forward + discrete-adjoint backward "
                    "(gradients to y0, the treatments and the ten shared constants)"}


def time_posterior_predict(dev, B, method):
    """Reconstruction / posterior sampling at the headline size (SURVEY f2 + f3; the reference's recon loops,
    training_challenge.py:174-195): Decoder.forward under no_grad (solver kernel + heads kernel, trajectories written
    and read back) against Decoder.predict (three quantile heads fused into the solver kernel, trajectories never in
    HBM).  Forward only; rank 0's GPU."""
    import types
    import structured_latent_odes_b200 as slode
    cfg = types.SimpleNamespace(obs_dim=3, ode_state_dim=S, ode_hidden_dim=H, adjoint_solver=False, solver=method,
                                constant_std=1e-2)
    torch.manual_seed(12)
    dec = slode.Decoder(cfg, torch.arange(0.0, T, 1.0, device=dev), L, dev).to(dev)
    z = torch.randn(B, L, device=dev, generator=torch.Generator(device=dev).manual_seed(12))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def best_of(fn, n=5):
        best = 1e30
        for r in range(n + 2):
            torch.cuda.synchronize()
            ev[0].record()
            out = fn()
            ev[1].record()
            torch.cuda.synchronize()
            del out
            if r >= 2:
                best = min(best, ev[0].elapsed_time(ev[1]))
        return best

    with torch.no_grad():
        two = best_of(lambda: dec(z))
    fused = best_of(lambda: dec.predict(z))
    return {"fused_ms": fused, "two_kernels_ms": two, "value": B * (T - 1) / (fused * 1e-3), "unit": UNIT,
            "what": f"Decoder.predict, {B} trajectories x {T} times, {method}, 3 quantile heads x obs_dim 3 written as "
                    "(B,O,T) by the solver kernel itself (slode_latent_fixed_heads_fwd), forward only; two_kernels_ms = "
                    "Decoder.forward under no_grad (slode_latent_fixed_fwd + slode_heads_fwd)"}


def time_cvs_epoch(dev, rank, world, sync_all):
    """SLODE train-epoch time, CVS default config (810 / 90 series, batch 128, midpoint + odeint_adjoint, seed 12):
    7 mini-batches x two objectives + the four evaluation passes (training_cvs.py:256-315).  With N ranks every
    mini-batch's rows are sharded and the gradients summed by one flat all-reduce per optimiser step."""
    import torch.distributed as dist
    from structured_latent_odes_b200 import sharding, training_cvs as tc
    cfg = tc.cvs_config()
    data = tc.make_cvs_dataset(cfg, dev, generator=torch.Generator(device=dev).manual_seed(cfg.seed))
    torch.manual_seed(cfg.seed)
    model = tc.MechanisticModel(cfg, dev, torch.arange(0.0, cfg.seq_len, 1.0, device=dev)).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=cfg.learning_rate, betas=(0.9, 0.999), capturable=True)
    shard = (rank, world) if world > 1 else None
    reducer = sharding.FlatGradReducer(list(model.parameters())) if world > 1 else None
    step = tc.GraphedTrainStep(model, opt, reducer, warmup=1)   # the two-objective training step replayed from CUDA graphs
    evaluation = tc.GraphedEvaluation(model, data, cfg, shard)   # ... and the four evaluation passes
    out = {}
    for evaluate in (True, False):
        best = None
        for e in range(5):
            sync_all()
            t0 = time.perf_counter()
            tc.train_epoch(model, opt, data, cfg, generator=torch.Generator().manual_seed(e), evaluate=evaluate,
                           reducer=reducer, shard=shard, step=step, evaluation=evaluation)
            sync_all()
            dt = torch.tensor([time.perf_counter() - t0], device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            if e >= 2:   # epochs 0 and 1 warm up / capture the graphs
                best = float(dt.item()) if best is None else min(best, float(dt.item()))
        out["s_per_epoch_with_eval" if evaluate else "s_per_epoch_train_steps_only"] = best
    out["what"] = ("training_cvs.py default config on synthetic CVS data (810 train / 90 val series, batch 128, midpoint, "
                   "odeint_adjoint semantics): 7 x 2 optimiser steps (+ 4 evaluation passes); best of 3 epochs after 1; "
                   f"rows of every mini-batch sharded over {world} rank(s)")
    return out


def time_cvs_epoch_cpu():
    """The same epoch through the CPU port of the reference decoder (all host threads), once after one warm-up."""
    from oracle import slode_port
    from structured_latent_odes_b200 import training_cvs as tc
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = tc.cvs_config()
    data = tc.make_cvs_dataset(cfg, "cuda", generator=torch.Generator(device="cuda").manual_seed(cfg.seed))
    data = {k: {kk: vv.cpu() for kk, vv in v.items()} for k, v in data.items()}
    torch.manual_seed(cfg.seed)
    model = tc.MechanisticModel(cfg, "cpu", torch.arange(0.0, cfg.seq_len, 1.0), decoder_cls=slode_port.Decoder)
    opt = torch.optim.Adam(model.parameters(), lr=cfg.learning_rate, betas=(0.9, 0.999))
    out = {}
    for evaluate in (True, False):
        ts = []
        for e in range(2):
            t0 = time.perf_counter()
            tc.train_epoch(model, opt, data, cfg, generator=torch.Generator().manual_seed(e), evaluate=evaluate)
            ts.append(time.perf_counter() - t0)
        out["s_per_epoch_with_eval" if evaluate else "s_per_epoch_train_steps_only"] = ts[-1]
    out["cores"] = os.cpu_count()
    return out


if __name__ == "__main__":
    main()
