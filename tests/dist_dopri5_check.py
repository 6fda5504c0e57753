"""Multi-GPU check of the trajectory-sharded dopri5 solve (diagnostic, run on the GPU box):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_dopri5_check.py

Every rank solves its shard with ``options=sharding.dopri5_shard_options(...)`` (one pass per launch, the batch-wide
norms all-reduced over NCCL in between), differentiates it, and sums the parameter gradients with the flat
all-reduce; rank 0 also runs the UNSHARDED solve on the whole batch and compares step logs, trajectories and
gradients.  Prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import torch.distributed as dist

import slode_testutil as U
import structured_latent_odes_b200 as slode
from structured_latent_odes_b200 import sharding
from structured_latent_odes_b200 import torchdiffeq_api as api


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 7000
    o = U.make_oracle("chal", "dopri5", False)
    p = U.make_product(o, device=dev)
    g = torch.Generator().manual_seed(41)
    z = torch.randn(B, 15, generator=g).to(dev)
    G = torch.randn(142, B, 5, generator=g).to(dev)
    lo, hi = sharding.shard_bounds(B, rank, world)
    rtol, atol = 1e-5, 1e-6
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    zs = z[lo:hi].clone().requires_grad_(True)
    opts = sharding.dopri5_shard_options(hi - lo, device=dev, options={"log_steps": True})
    p.zero_grad()
    torch.cuda.synchronize(); dist.barrier()
    ev[0].record()
    sol = slode.odeint(p.gen_dynamics(zs), p.initialize_state(zs), p.times, method="dopri5", rtol=rtol, atol=atol,
                       options=opts)
    (sol * G[:, lo:hi]).sum().backward()
    red = sharding.FlatGradReducer(p.parameters())
    flat = red.reduce().clone()
    ev[1].record()
    torch.cuda.synchronize()
    steps = api.last_dopri5_stats.steps.clone()
    gathered = [torch.zeros_like(steps.to(dev)) for _ in range(world)]
    dist.all_gather(gathered, steps.to(dev))
    same_log_on_all_ranks = all(torch.equal(gathered[0], x) for x in gathered)
    out = None
    if rank == 0:
        p.zero_grad()
        zf = z.clone().requires_grad_(True)
        full = slode.odeint(p.gen_dynamics(zf), p.initialize_state(zf), p.times, method="dopri5", rtol=rtol, atol=atol,
                            options={"log_steps": True})
        (full * G).sum().backward()
        steps_full = api.last_dopri5_stats.steps.clone()
        flat_full = torch.cat([q.grad.reshape(-1) for q in p.parameters() if q.requires_grad])
        out = {"check": "sharded dopri5 over NCCL vs unsharded", "world": world, "B": B, "attempted_steps": int(steps.shape[0]),
               "step_log_identical_to_unsharded": bool(torch.equal(steps, steps_full)),
               "step_log_identical_on_all_ranks": bool(same_log_on_all_ranks),
               "sol_rel_err_rank0_rows": U.rel_err(sol, full[:, lo:hi]),
               "param_grad_rel_err": U.rel_err(flat, flat_full),
               "sharded_fwd_bwd_ms": ev[0].elapsed_time(ev[1])}
    dist.barrier()
    dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out), flush=True)
        assert out["step_log_identical_to_unsharded"] and out["step_log_identical_on_all_ranks"]
        assert out["sol_rel_err_rank0_rows"] < 1e-6 and out["param_grad_rel_err"] < 1e-5


if __name__ == "__main__":
    main()
