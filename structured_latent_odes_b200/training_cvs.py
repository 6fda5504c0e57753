"""Pyro-free SLODE training step / epoch for the CVS model.

The reference trains ``models/mechanistic_cvs.py::MechanisticModel`` with two Pyro ``SVI`` objects
(``training_cvs.py:226-249``): ``loss_basic = SVI(model, guide)`` and ``loss_aux = SVI(model_meta, guide_meta)``,
both ``Trace_ELBO(num_particles=1)`` on one shared ``pyro.optim.Adam``.  Pyro is not installable here, so this
module restates exactly that objective in plain PyTorch; everything that is not the latent-ODE solve (the Conv1d
encoder, the small prior / classifier MLPs, the likelihood arithmetic) stays ordinary torch ops, and the decoder
-- the hot path -- is ``structured_latent_odes_b200.Decoder`` (fused solver kernels + fused heads).

    PARITY UNPINNED for the ELBO restatement itself: no Pyro to run, no stored losses in the reference.  What IS
    checked (tests/test_training_cvs.py, tests/test_gpu_training_cvs.py): every term against
    ``torch.distributions``; and that the GPU model and the same model over the CPU oracle decoder produce the
    same losses, gradients and parameters after optimisation steps from the same seeds.

Restated pieces, with the reference lines they follow:

* ``EncoderCONV``                    models/encoder_conv.py:17-51  (Conv1d -> AvgPool1d(stride 1) -> Linear -> tanh -> loc, exp scale)
* prior / classifier MLPs            models/encoder_mlp.py:60-167 as instantiated at models/mechanistic_cvs.py:64-100
* ``model`` + ``guide``  (loss_basic) models/mechanistic_cvs.py:105-238: z ~ q(z|x) reparameterised; log p(z_eps) + log p(z_iext|iext)
                                     + log p(z_rtpr|rtpr) - log q(z|x) + asymmetric-Laplace quantile likelihood for
                                     (mu_50, tau=.5), (mu_75, .5+d), (mu_25, .5-d): weight tau where x >= mu, 1 - tau elsewhere
* ``model_meta`` (loss_aux)          :240-270: z_cls sampled from the encoder's Normal inside the MODEL (so its own log-density
                                     is part of the objective) + aux_loss_multiplier * Bernoulli label log-likelihoods
* ``classifier`` / ``recon``         :277-323
* ``run_batch`` / ``input_pred_stats`` / epoch                     training_cvs.py:44-157, 256-331
* synthetic data                     data/cvs/cvs_data.py:24-183 (batched generator instead of the LSODA loop), labels as
                                     utils/ODE_dataset.py:50-51, unit-segment normalisation utils/ODE_dataset.py:196-209
"""
from __future__ import annotations

import math
import types

import torch
from torch import nn

__all__ = ["cvs_config", "EncoderCONV", "MechanisticModel", "make_cvs_dataset", "run_batch", "input_pred_stats",
           "train_epoch"]


def cvs_config(**overrides):
    """``data/cvs/config_cvs.py:6-52`` (the keys the model and the loop read)."""
    c = dict(seq_len=86, delta_t=1.0, obs_dim=3, iext_dim=1, rtpr_dim=1, z_iext_dim=5, z_rtpr_dim=5, z_epsilon_dim=5,
             u_hidden_dim=25, aux_loss_multiplier=46.0, seed=12, mini_batch_size=128, n_filters=10, filter_size=10,
             pool_size=5, cnn_hidden_dim=50, ode_state_dim=5, ode_hidden_dim=25, system_input_dim=2, learning_rate=1e-3,
             num_particles=1, adjoint_solver=True, solver="midpoint", constant_std=1e-2, quantile_diff=0.475,
             noise_std=0.05, data_size=1000)
    c.update(overrides)
    return types.SimpleNamespace(**c)


class EncoderCONV(nn.Module):
    def __init__(self, n_channels, n_filters, filter_size, pool_size, n_time, latent_dim, hidden_dim):
        super().__init__()
        n_pool = n_time - (filter_size - 1) - (pool_size - 1)
        self.conv = nn.Conv1d(n_channels, n_filters, filter_size)
        nn.init.orthogonal_(self.conv.weight)
        self.pool = nn.AvgPool1d(pool_size, stride=1)
        self.lin = nn.Linear(n_pool * n_filters, hidden_dim)
        nn.init.orthogonal_(self.lin.weight)
        self.act = nn.Tanh()
        self.z_loc = nn.Linear(hidden_dim, latent_dim)
        self.z_scale = nn.Sequential(nn.Linear(hidden_dim, latent_dim))

    def forward(self, x):
        x = self.pool(self.conv(x))
        x = self.act(self.lin(x.reshape(x.size(0), -1)))
        return self.z_loc(x), torch.exp(self.z_scale(x))


def _hidden_linear(n_in, n_out):
    lin = nn.Linear(n_in, n_out)
    lin.weight.data.normal_(0, 0.001)  # EncoderMLP initialises hidden layers this way (encoder_mlp.py:91-92)
    lin.bias.data.normal_(0, 0.001)
    return lin


class _Classifier(nn.Module):
    """EncoderMLP([z_dim, u_hidden, 1], Softplus, output Sigmoid)."""

    def __init__(self, z_dim, hidden, out):
        super().__init__()
        self.net = nn.Sequential(_hidden_linear(z_dim, hidden), nn.Softplus(), nn.Linear(hidden, out), nn.Sigmoid())

    def forward(self, z):
        return self.net(z)


class _Prior(nn.Module):
    """EncoderMLP([u_dim, [z_dim, z_dim]], output activations [None, Exp]): no hidden layer, two linear heads."""

    def __init__(self, u_dim, z_dim):
        super().__init__()
        self.loc = nn.Linear(u_dim, z_dim)
        self.log_scale = nn.Linear(u_dim, z_dim)

    def forward(self, u):
        return self.loc(u), torch.exp(self.log_scale(u))


def _bernoulli_logp(probs, y):
    """sum of Bernoulli(probs).log_prob(y) as pyro / torch.distributions compute it: probs are clamped to
    [eps, 1 - eps] (``probs_to_logits``) so a saturated sigmoid gives a finite value instead of 0 * -inf = NaN."""
    eps = torch.finfo(probs.dtype).eps
    p = probs.clamp(eps, 1.0 - eps)
    return -torch.nn.functional.binary_cross_entropy(p, y, reduction="sum")


def _sample_normal(loc, scale):
    """loc + scale * eps.  (torch.normal(loc, scale) validates ``scale >= 0`` with a device->host sync, which also
    makes it illegal inside CUDA-graph capture.)"""
    return loc + scale * torch.randn_like(loc)


def _normal_logp(x, loc, scale):
    return (-0.5 * ((x - loc) / scale) ** 2 - torch.log(scale) - 0.5 * math.log(2 * math.pi)).sum()


class MechanisticModel(nn.Module):
    """Plain-torch ``MechanisticModel`` (CVS).  ``decoder_cls(config, times, latent_dim, device)`` defaults to the
    fused ``structured_latent_odes_b200.Decoder``; the tests pass a CPU oracle decoder with the same interface."""

    def __init__(self, config, device, times, decoder_cls=None):
        super().__init__()
        c = self.config = config
        self.device = device
        self.times = times
        self.latent_dim = c.z_iext_dim + c.z_rtpr_dim + c.z_epsilon_dim
        self.q_iext_given_z_iext = _Classifier(c.z_iext_dim, c.u_hidden_dim, c.iext_dim)
        self.q_rtpr_given_z_rtpr = _Classifier(c.z_rtpr_dim, c.u_hidden_dim, c.rtpr_dim)
        self.encoder = EncoderCONV(c.obs_dim, c.n_filters, c.filter_size, c.pool_size, len(times), self.latent_dim,
                                   c.cnn_hidden_dim)
        self.p_z_iext_given_iext = _Prior(c.iext_dim, c.z_iext_dim)
        self.p_z_rtprs_given_rtprs = _Prior(c.rtpr_dim, c.z_rtpr_dim)
        if decoder_cls is None:
            from .decoders import Decoder as decoder_cls
        self.decoder = decoder_cls(config=c, times=times, latent_dim=self.latent_dim, device=device)

    # ---- likelihood ------------------------------------------------------------------------------------------
    @staticmethod
    def quantile_loglik(obs, mu, std, tau):
        """sum of the six masked Laplace sites of compute_likelihood (:180-211): weight tau where x >= mu."""
        w = torch.where(obs >= mu, float(tau), float(1.0 - tau))  # python scalars: no host tensor (graph-capturable)
        return (w * (-torch.log(2.0 * std) - (obs - mu).abs() / std)).sum()

    # ---- the two objectives (negative ELBOs summed over the batch, as Trace_ELBO returns them) -----------------
    def loss_basic(self, observations, iext, rtpr, eps=None):
        c = self.config
        loc_z, scale_z = self.encoder(observations)
        eps = torch.randn_like(loc_z) if eps is None else eps
        z = loc_z + scale_z * eps                                    # guide sample, reparameterised
        log_q = _normal_logp(z, loc_z, scale_z)
        zi, zr, ze = z[:, :c.z_iext_dim], z[:, c.z_iext_dim:c.z_iext_dim + c.z_rtpr_dim], z[:, -c.z_epsilon_dim:]
        li, si = self.p_z_iext_given_iext(iext)
        lr, sr = self.p_z_rtprs_given_rtprs(rtpr)
        log_p = (_normal_logp(ze, torch.zeros_like(ze), torch.ones_like(ze)) + _normal_logp(zi, li, si)
                 + _normal_logp(zr, lr, sr))
        _, mu_75, mu_50, mu_25, std = self.decoder(torch.cat((zi, zr, ze), dim=1))
        d = c.quantile_diff
        for mu, tau in ((mu_50, 0.5), (mu_75, 0.5 + d), (mu_25, 0.5 - d)):
            log_p = log_p + self.quantile_loglik(observations, mu, std, tau)
        return -(log_p - log_q)

    def loss_aux(self, observations, iext, rtpr, eps=None):
        c = self.config
        loc_z, scale_z = self.encoder(observations)
        n = c.z_iext_dim + c.z_rtpr_dim
        loc, scale = loc_z[:, :n], scale_z[:, :n]
        eps = torch.randn_like(loc) if eps is None else eps
        z = loc + scale * eps                                        # model-side sample sites z_*_cls
        log_p = _normal_logp(z, loc, scale)
        a_i = self.q_iext_given_z_iext(z[:, :c.z_iext_dim])
        a_r = self.q_rtpr_given_z_rtpr(z[:, c.z_iext_dim:])
        log_p = log_p + float(c.aux_loss_multiplier) * (_bernoulli_logp(a_i, iext) + _bernoulli_logp(a_r, rtpr))
        return -log_p

    # ---- evaluation helpers --------------------------------------------------------------------------------
    @torch.no_grad()
    def classifier(self, observations):
        c = self.config
        loc_z, scale_z = self.encoder(observations)
        z = _sample_normal(loc_z, scale_z)
        a_i = self.q_iext_given_z_iext(z[:, :c.z_iext_dim])
        a_r = self.q_rtpr_given_z_rtpr(z[:, c.z_iext_dim:c.z_iext_dim + c.z_rtpr_dim])
        return {"iext": (a_i > 0.5).float(), "rtpr": (a_r > 0.5).float()}

    @torch.no_grad()
    def recon(self, observations, iext, rtpr, is_post):
        c = self.config
        if is_post:
            loc_z, scale_z = self.encoder(observations)
            z = _sample_normal(loc_z, scale_z)
        else:
            B = observations.shape[0]
            ze = torch.randn(B, c.z_epsilon_dim, device=observations.device)
            zi = _sample_normal(*self.p_z_iext_given_iext(iext))
            zr = _sample_normal(*self.p_z_rtprs_given_rtprs(rtpr))
            z = torch.cat((zi, zr, ze), dim=1)
        solution_xt, mu_75, mu_50, mu_25, std = self.decoder(z)
        return {"l1": (mu_50 - observations).abs().mean(), "solution_xt": solution_xt, "mu_75": mu_75, "mu_50": mu_50,
                "mu_25": mu_25, "std": std, "z": z}


# --------------------------------------------------------------------------------------------------------------
# data
# --------------------------------------------------------------------------------------------------------------
def make_cvs_dataset(config, device, generator=None):
    """Synthetic CVS data like ``data/cvs/cvs_data.py::make_dataset`` but generated on the device in one launch
    (``generate_cvs_latents``).  Returns dict split -> {observations (N,O,T) in [0,1], iext (N,1), rtpr (N,1)} with the
    reference's 810 / 90 / 100 split of ``data_size`` = 1000."""
    from .cvs_mechanistic import generate_cvs_latents, observe
    n = config.data_size
    g = generator
    i_ext = torch.where(torch.rand(n, generator=g, device=device) > 0.5, 0.0, -2.0)
    r_mod = torch.where(torch.rand(n, generator=g, device=device) > 0.5, 0.0, 0.5)
    lat = generate_cvs_latents(i_ext, r_mod, seq_len=config.seq_len, delta_t=config.delta_t)
    raw = observe(lat).float()
    noisy = raw + config.noise_std * torch.randn(raw.shape, generator=g, device=device)
    n_train_all = int(round(n * 0.9))
    n_train = int(round(n_train_all * 0.9))
    mn = noisy[:n_train_all].amin(dim=(0, 1))
    mx = noisy[:n_train_all].amax(dim=(0, 1))
    obs = ((noisy - mn) / (mx - mn)).permute(0, 2, 1).contiguous()       # (N, O, T) as batch_to_device hands it over
    iext = (i_ext >= 0).float()[:, None]
    rtpr = (r_mod > 0).float()[:, None]
    cut = {"train": slice(0, n_train), "val": slice(n_train, n_train_all), "test": slice(n_train_all, n)}
    return {k: {"observations": obs[s], "iext": iext[s], "rtpr": rtpr[s]} for k, s in cut.items()}


def batches(split, batch_size, shuffle, generator=None, shard=None):
    """Mini-batches of a split.  ``shard=(rank, world_size)`` hands this rank its contiguous row range of EVERY
    mini-batch (same permutation on every rank: same generator seed), so that the ranks' summed gradients equal the
    single-process step (the losses are sums over batch rows)."""
    n = split["observations"].shape[0]
    from .sharding import shard_bounds
    if not shuffle:  # plain row ranges: no index tensor, no host->device copy (graph-capturable)
        for lo in range(0, n, batch_size):
            hi = min(lo + batch_size, n)
            if shard is not None:
                a, b = shard_bounds(hi - lo, shard[0], shard[1])
                lo, hi = lo + a, lo + b
            yield {k: v[lo:hi] for k, v in split.items()}
        return
    idx = torch.randperm(n, generator=generator)
    for lo in range(0, n, batch_size):
        sel = idx[lo:lo + batch_size]
        if shard is not None:
            a, b = shard_bounds(sel.numel(), shard[0], shard[1])
            sel = sel[a:b]
        sel = sel.to(split["observations"].device)
        yield {k: v[sel] for k, v in split.items()}


# --------------------------------------------------------------------------------------------------------------
# loop
# --------------------------------------------------------------------------------------------------------------
def run_batch(model, optimizer, batch, reducer=None):
    """Two SVI steps on one shared Adam (training_cvs.py:147-157).  ``reducer`` (sharding.FlatGradReducer) sums the
    gradients over ranks before each optimiser step when the batch rows are sharded."""
    out = []
    for loss_fn in (model.loss_basic, model.loss_aux):
        optimizer.zero_grad(set_to_none=True)
        loss = loss_fn(batch["observations"], batch["iext"], batch["rtpr"])
        loss.backward()
        if reducer is not None:
            reducer.reduce()
        optimizer.step()
        out.append(loss.detach() / batch["observations"].shape[0])
    return out


class GraphedTrainStep:
    """``run_batch`` captured into a CUDA graph per mini-batch size and replayed (the latency
    path).  At the reference's batch sizes (128 rows) a training step is ~50 small launches and launch latency is
    all there is; the solver kernels keep no cross-call state and take their scratch from the torch allocator, so the
    whole two-objective step (forward, reverse sweep, flat all-reduce, Adam) is capturable.  The first ``warmup``
    steps of every batch size run eagerly (they are real training steps), the next one is captured, all later ones
    are replays.  The optimiser must be constructed with ``capturable=True``."""

    def __init__(self, model, optimizer, reducer=None, warmup=3):
        self.model, self.opt, self.reducer, self.warmup = model, optimizer, reducer, warmup
        self.seen, self.graphs = {}, {}

    def __call__(self, batch):
        n = batch["observations"].shape[0]
        if n not in self.graphs:
            self.seen[n] = self.seen.get(n, 0) + 1
            if self.seen[n] <= self.warmup:
                return run_batch(self.model, self.opt, batch, self.reducer)
            self._capture(n, batch)
        graph, static, losses = self.graphs[n]
        for k, v in static.items():
            v.copy_(batch[k])
        graph.replay()
        return [l.clone() / n for l in losses]

    def _capture(self, n, batch):
        static = {k: v.clone() for k, v in batch.items()}
        graph = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        torch.cuda.synchronize()
        with torch.cuda.graph(graph):
            losses = []
            for loss_fn in (self.model.loss_basic, self.model.loss_aux):
                self.opt.zero_grad(set_to_none=True)
                loss = loss_fn(static["observations"], static["iext"], static["rtpr"])
                loss.backward()
                if self.reducer is not None:
                    self.reducer.reduce()
                self.opt.step()
                losses.append(loss.detach())
        self.graphs[n] = (graph, static, losses)


class GraphedEvaluation:
    """The four evaluation passes of an epoch (val / train x posterior / prior, training_cvs.py:270-315) captured as
    ONE CUDA graph: they only read the resident dataset and the current weights, so the graph needs no inputs.
    First call eager (warm-up), second call captures, later calls replay."""

    def __init__(self, model, data, config, shard=None):
        self.model, self.data, self.config, self.shard = model, data, config, shard
        self.calls, self.graph, self.static = 0, None, None

    def _passes(self):
        m, d, bs, sh = self.model, self.data, self.config.mini_batch_size, self.shard
        return {"val_post": input_pred_stats(m, d["val"], True, shard=sh),
                "val_prior": input_pred_stats(m, d["val"], False, shard=sh),
                "train_post": input_pred_stats(m, d["train"], True, batch_size=bs, shard=sh),
                "train_prior": input_pred_stats(m, d["train"], False, batch_size=bs, shard=sh)}

    def __call__(self):
        self.calls += 1
        if self.calls == 1:
            return self._passes()
        if self.graph is None:
            self.graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(self.graph):
                self.static = self._passes()
        self.graph.replay()
        return {k: {kk: vv.clone() for kk, vv in v.items()} for k, v in self.static.items()}


@torch.no_grad()
def input_pred_stats(model, split, is_post, batch_size=None, shard=None):
    """One evaluation pass (training_cvs.py:44-144): both losses forward only, recon, classifier accuracy.  With
    ``shard`` every rank evaluates its rows of each batch (the caller sums the statistics over ranks)."""
    n = split["observations"].shape[0]
    tot = [0.0, 0.0]
    l1 = 0.0
    hit_i = hit_r = 0.0
    for b in batches(split, batch_size or n, shuffle=False, shard=shard):
        if b["observations"].shape[0] == 0:
            continue
        o, i, r = b["observations"], b["iext"], b["rtpr"]
        tot[0] += model.loss_basic(o, i, r) / o.shape[0]
        tot[1] += model.loss_aux(o, i, r) / o.shape[0]
        res = model.recon(o, i, r, is_post)
        l1 += res["l1"]
        pred = model.classifier(o)
        hit_i += (pred["iext"] == i).float().sum()
        hit_r += (pred["rtpr"] == r).float().sum()
    return {"iext": hit_i / n, "rtpr": hit_r / n, "l1": l1 / n, "elbo": torch.stack([torch.as_tensor(t) for t in tot])}


def train_epoch(model, optimizer, data, config, generator=None, evaluate=True, reducer=None, shard=None, step=None,
                evaluation=None):
    """One reference epoch: the mini-batch loop, then the four evaluation passes (training_cvs.py:256-315).
    Multi-GPU: ``shard=(rank, world_size)`` + ``reducer`` (one flat all-reduce of the gradients per optimiser step).
    ``step`` / ``evaluation``: a ``GraphedTrainStep`` / ``GraphedEvaluation`` to replay the training steps / the four
    evaluation passes from CUDA graphs instead of launching them eagerly."""
    do = step if step is not None else (lambda b: run_batch(model, optimizer, b, reducer))
    losses = [do(b)
              for b in batches(data["train"], config.mini_batch_size, shuffle=True, generator=generator, shard=shard)]
    stats = {}
    if evaluate and evaluation is not None:
        stats = evaluation()
    elif evaluate:
        stats["val_post"] = input_pred_stats(model, data["val"], True, shard=shard)
        stats["val_prior"] = input_pred_stats(model, data["val"], False, shard=shard)
        stats["train_post"] = input_pred_stats(model, data["train"], True, batch_size=config.mini_batch_size, shard=shard)
        stats["train_prior"] = input_pred_stats(model, data["train"], False, batch_size=config.mini_batch_size, shard=shard)
    return torch.stack([torch.stack(l) for l in losses]).mean(0), stats
