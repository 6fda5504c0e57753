"""f1 on the device: the training step over the fused decoder equals the same step over the CPU oracle decoder."""
import pytest
import torch

import slode_testutil as U
from oracle import slode_port
from structured_latent_odes_b200 import training_cvs as tc

pytestmark = pytest.mark.gpu


def _pair(cfg):
    torch.manual_seed(7)
    times = torch.arange(0.0, cfg.seq_len, 1.0)
    cpu = tc.MechanisticModel(cfg, "cpu", times, decoder_cls=slode_port.Decoder)
    gpu = tc.MechanisticModel(cfg, "cuda", times.cuda())
    gpu.load_state_dict(cpu.state_dict())
    return cpu, gpu.cuda()


def test_dataset_generator_shapes_and_ranges():
    cfg = tc.cvs_config()
    data = tc.make_cvs_dataset(cfg, "cuda", generator=torch.Generator(device="cuda").manual_seed(12))
    assert {k: v["observations"].shape[0] for k, v in data.items()} == {"train": 810, "val": 90, "test": 100}
    o = data["train"]["observations"]
    assert o.shape[1:] == (3, 86) and o.dtype == torch.float32
    assert float(o.min()) >= -1e-6 and float(o.max()) <= 1 + 1e-6
    assert set(data["train"]["iext"].unique().tolist()) <= {0.0, 1.0}


@pytest.mark.parametrize("solver,adjoint", [("midpoint", True), ("rk4", False)])
def test_losses_gradients_and_optimisation_match_the_cpu_oracle_model(solver, adjoint):
    cfg = tc.cvs_config(solver=solver, adjoint_solver=adjoint)
    cpu, gpu = _pair(cfg)
    g = torch.Generator().manual_seed(2)
    obs = torch.rand(40, 3, 86, generator=g)
    iext, rtpr = (torch.rand(40, 1, generator=g) > 0.5).float(), (torch.rand(40, 1, generator=g) > 0.5).float()
    eps = torch.randn(40, 15, generator=g)
    lc = cpu.loss_basic(obs, iext, rtpr, eps=eps)
    lg = gpu.loss_basic(obs.cuda(), iext.cuda(), rtpr.cuda(), eps=eps.cuda())
    assert abs(float(lc.detach()) - float(lg.detach())) < 1e-5 * abs(float(lc.detach()))
    lc.backward()
    lg.backward()
    gc = dict(cpu.named_parameters())
    for k, p in gpu.named_parameters():
        if ".prod." in k or ".degr." in k:
            continue
        if gc[k].grad is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
        else:
            assert U.rel_err(p.grad, gc[k].grad) < 2e-5, (k, U.rel_err(p.grad, gc[k].grad))
    # a few optimiser steps with identical noise
    oc = torch.optim.Adam(cpu.parameters(), lr=1e-3)
    og = torch.optim.Adam(gpu.parameters(), lr=1e-3)
    batch = {"observations": obs, "iext": iext, "rtpr": rtpr}
    bg = {k: v.cuda() for k, v in batch.items()}
    for step in range(3):
        torch.manual_seed(100 + step)
        e1, e2 = torch.randn(40, 15), torch.randn(40, 10)
        for m, o, b, dev in ((cpu, oc, batch, "cpu"), (gpu, og, bg, "cuda")):
            for fn, e in ((m.loss_basic, e1), (m.loss_aux, e2)):
                o.zero_grad(set_to_none=True)
                fn(b["observations"], b["iext"], b["rtpr"], eps=e.to(dev)).backward()
                o.step()
    pc = dict(cpu.named_parameters())
    for k, p in gpu.named_parameters():
        assert U.rel_err(p, pc[k]) < 1e-4, k


def test_epoch_runs_on_generated_data():
    cfg = tc.cvs_config()
    torch.manual_seed(12)
    times = torch.arange(0.0, cfg.seq_len, 1.0, device="cuda")
    model = tc.MechanisticModel(cfg, "cuda", times).cuda()
    data = tc.make_cvs_dataset(cfg, "cuda", generator=torch.Generator(device="cuda").manual_seed(12))
    opt = torch.optim.Adam(model.parameters(), lr=cfg.learning_rate, betas=(0.9, 0.999))
    l0, stats = tc.train_epoch(model, opt, data, cfg, generator=torch.Generator().manual_seed(0))
    for _ in range(4):
        l1, stats = tc.train_epoch(model, opt, data, cfg, generator=torch.Generator().manual_seed(0), evaluate=False)
    assert torch.isfinite(l1).all() and float(l1[0]) < float(l0[0])
    assert set(stats) == set() and l1.shape == (2,)
