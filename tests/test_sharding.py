"""Multi-GPU plumbing on CPU: trajectory sharding and the single flat-buffer gradient all-reduce, exercised with
the gloo backend at world_size 2 (the N>1 path of bench.py / a sharded training step)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from structured_latent_odes_b200 import sharding


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 128, 1_048_576, 1_000_003):
        for w in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import slode_port  # the oracle stands in for the solve on CPU
        torch.manual_seed(12)  # same weights on every rank
        times = torch.arange(0.0, 12.0, 1.0)
        model = slode_port.OdeModel(times, 5, 15, 25, False, "midpoint")
        g = torch.Generator().manual_seed(3)
        z = torch.randn(10, 15, generator=g)
        G = torch.randn(10, 12, 5, generator=g)
        zs, Gs = sharding.shard_rows(z, rank, world), sharding.shard_rows(G, rank, world)
        (model.solve_ODE(zs) * Gs).sum().backward()
        reducer = sharding.FlatGradReducer(model.parameters())
        flat = reducer.reduce().clone()
        # one collective, identical result on every rank
        gathered = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        assert all(torch.equal(gathered[0], x) for x in gathered)
        if rank == 0:
            torch.save({k: p.grad.clone() for k, p in model.named_parameters()}, out)
    finally:
        dist.destroy_process_group()


def test_flat_grad_reducer_world2_equals_single_process(tmp_path):
    out = str(tmp_path / "grads.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    sharded = torch.load(out)
    from oracle import slode_port
    torch.manual_seed(12)
    times = torch.arange(0.0, 12.0, 1.0)
    model = slode_port.OdeModel(times, 5, 15, 25, False, "midpoint")
    g = torch.Generator().manual_seed(3)
    z = torch.randn(10, 15, generator=g)
    G = torch.randn(10, 12, 5, generator=g)
    (model.solve_ODE(z) * G).sum().backward()
    for k, p in model.named_parameters():
        assert torch.allclose(sharded[k], p.grad, rtol=1e-5, atol=1e-6), k


def test_reducer_without_process_group_is_identity():
    lin = torch.nn.Linear(3, 2)
    lin(torch.ones(1, 3)).sum().backward()
    before = [p.grad.clone() for p in lin.parameters()]
    flat = sharding.FlatGradReducer(lin.parameters()).reduce()
    assert flat.numel() == 8
    assert all(torch.equal(a, p.grad) for a, p in zip(before, lin.parameters()))


def _dopri5_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import slode_port, torchdiffeq_oracle as tde
        torch.manual_seed(12)
        times = torch.arange(0.0, 12.0, 1.0)
        model = slode_port.OdeModel(times, 5, 15, 25, False, "dopri5")
        z = torch.randn(11, 15, generator=torch.Generator().manual_seed(3))
        zs = sharding.shard_rows(z, rank, world)
        opts = sharding.dopri5_shard_options(zs.shape[0])
        assert opts["global_batch"] == 11
        probe = torch.tensor([1.0 + rank, 10.0], dtype=torch.float64)
        opts["shard_reducer"](probe)
        assert probe.tolist() == [sum(1.0 + r for r in range(world)), 10.0 * world]

        # the protocol the device path follows (Dopri5ShardSolve): every batch-wide RMS norm of the controller is the
        # all-reduced sum of squares over the shards divided by the GLOBAL element count
        def global_rms(x):
            s = (x.double() ** 2).sum().reshape(1)
            opts["shard_reducer"](s)
            return (s[0] / (opts["global_batch"] * 5)).sqrt().to(x.dtype)

        with torch.no_grad():
            sol = tde.odeint(slode_port.OdeFunc(zs, model.dynamics), model.latent_to_ode_net(zs), times, method="dopri5",
                             rtol=1e-4, atol=1e-5, options={"norm": global_rms})
        torch.save({"sol": sol, "acc": list(tde.last_stats.accepted), "dts": list(tde.last_stats.dts)}, f"{out}.{rank}")
    finally:
        dist.destroy_process_group()


def test_sharded_dopri5_protocol_world2_takes_the_unsharded_step_sequence(tmp_path):
    """dopri5's controller is batch-global: sharding the batch needs one scalar all-reduce per norm.  With
    that in place both ranks take exactly the unsharded accept / reject sequence; the device path
    (torchdiffeq_api.Dopri5ShardSolve + sharding.dopri5_shard_options) follows the same protocol and is compared with
    the unsharded device solve in tests/test_gpu_dopri5.py."""
    out = str(tmp_path / "dopri5")
    mp.spawn(_dopri5_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    from oracle import slode_port, torchdiffeq_oracle as tde
    torch.manual_seed(12)
    times = torch.arange(0.0, 12.0, 1.0)
    model = slode_port.OdeModel(times, 5, 15, 25, False, "dopri5")
    z = torch.randn(11, 15, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        full = tde.odeint(slode_port.OdeFunc(z, model.dynamics), model.latent_to_ode_net(z), times, method="dopri5",
                          rtol=1e-4, atol=1e-5)
    acc, dts = list(tde.last_stats.accepted), list(tde.last_stats.dts)
    parts = [torch.load(f"{out}.{r}") for r in range(2)]
    for p in parts:
        assert p["acc"] == acc
        rel = ((torch.tensor(p["dts"]) - torch.tensor(dts)).abs() / torch.tensor(dts)).max().item()
        assert rel < 2e-2, rel   # matmul kernels differ with the batch size: fp32 rounding noise of the error estimate
    assert torch.allclose(torch.cat([p["sol"] for p in parts], dim=1), full, rtol=1e-4, atol=1e-5)
