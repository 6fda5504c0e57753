"""The fixed-grid entry points keep no cross-call state: a training step is capturable into a CUDA graph, and a
replayed step equals the eager step bit for bit on the solver's side (same kernels, same inputs)."""
import pytest
import torch

import slode_testutil as U

pytestmark = pytest.mark.gpu


def test_solve_and_reverse_sweep_replay_from_a_cuda_graph():
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    o = U.make_oracle("cvs", "midpoint", True)
    p = U.make_product(o)
    g = torch.Generator().manual_seed(2)
    z = torch.randn(128, 15, generator=g).cuda().requires_grad_(True)
    G = torch.randn(128, 86, 5, generator=g).cuda()

    def step():
        p.zero_grad(set_to_none=True)
        z.grad = None
        sol = p.solve_ODE(z)
        (sol * G).sum().backward()
        return sol.detach(), z.grad, [q.grad for q in p.parameters() if q.grad is not None]

    for _ in range(3):  # eager warm-up (also resolves the launch plans once)
        sol_e, gz_e, gp_e = step()
    sol_e, gz_e, gp_e = sol_e.clone(), gz_e.clone(), [x.clone() for x in gp_e]
    graph = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.graph(graph):
        sol_g, gz_g, gp_g = step()
    for _ in range(2):
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(sol_g, sol_e)
    assert U.rel_err(gz_g, gz_e) < 1e-6          # block-level float atomics: summation order may differ
    for a, b in zip(gp_g, gp_e):
        assert U.rel_err(a, b) < 1e-5


def test_graphed_training_step_trains_like_the_eager_one():
    from structured_latent_odes_b200 import training_cvs as tc
    cfg = tc.cvs_config(data_size=300, mini_batch_size=64)
    dev = "cuda"
    data = tc.make_cvs_dataset(cfg, dev, generator=torch.Generator(device=dev).manual_seed(1))
    times = torch.arange(0.0, cfg.seq_len, 1.0, device=dev)
    losses = {}
    for graphed in (False, True):
        torch.manual_seed(5)
        model = tc.MechanisticModel(cfg, dev, times).to(dev)
        opt = torch.optim.Adam(model.parameters(), lr=cfg.learning_rate, capturable=True)
        step = tc.GraphedTrainStep(model, opt, warmup=2) if graphed else None
        out = []
        for e in range(4):
            torch.manual_seed(100 + e)
            l, _ = tc.train_epoch(model, opt, data, cfg, generator=torch.Generator().manual_seed(e), evaluate=False, step=step)
            out.append(l)
        losses[graphed] = torch.stack(out)
        assert torch.isfinite(losses[graphed]).all()
    # the same data order and initial weights; the sampled eps differ (graph replays advance the generator differently)
    assert losses[True][-1, 0] < losses[True][0, 0]                       # it trains
    assert (losses[True][-1] - losses[False][-1]).abs().max() < 0.25 * losses[False][-1].abs().max()


def test_graphed_evaluation_passes_match_the_eager_ones():
    from structured_latent_odes_b200 import training_cvs as tc
    cfg = tc.cvs_config(data_size=300, mini_batch_size=64)
    dev = "cuda"
    data = tc.make_cvs_dataset(cfg, dev, generator=torch.Generator(device=dev).manual_seed(1))
    torch.manual_seed(5)
    model = tc.MechanisticModel(cfg, dev, torch.arange(0.0, cfg.seq_len, 1.0, device=dev)).to(dev)
    ev = tc.GraphedEvaluation(model, data, cfg)
    eager = ev()            # first call: eager
    ev()                    # second call: capture + replay
    replay = ev()
    assert set(replay) == {"val_post", "val_prior", "train_post", "train_prior"}
    for k in replay:
        a, b = replay[k]["elbo"], eager[k]["elbo"]
        assert torch.isfinite(a).all()
        # different draws of the latent samples; same data, same weights
        assert (a - b).abs().max() < 0.2 * b.abs().max(), (k, a, b)
