"""The Pyro-free restatement of the CVS training objective (f1) on CPU, over the oracle decoder: every term against
``torch.distributions`` / the reference's masked-Laplace formulation, gradient routing of the two objectives, and a
few optimisation steps."""
import torch
from torch import distributions as D

from oracle import slode_port
from structured_latent_odes_b200 import training_cvs as tc


def _model(T=24, seed=0):
    cfg = tc.cvs_config(seq_len=T)
    torch.manual_seed(seed)
    times = torch.arange(0.0, T, 1.0)
    return cfg, tc.MechanisticModel(cfg, "cpu", times, decoder_cls=slode_port.Decoder)


def _batch(cfg, B=6, seed=1):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(B, cfg.obs_dim, cfg.seq_len, generator=g), (torch.rand(B, 1, generator=g) > 0.5).float(),
            (torch.rand(B, 1, generator=g) > 0.5).float())


def test_quantile_loglik_equals_the_reference_masked_laplace_sites():
    """compute_likelihood (mechanistic_cvs.py:180-211): per channel, Laplace(pred, scale) on the entries with
    x < mu scaled by (1 - tau), and on the entries with x >= mu scaled by tau."""
    g = torch.Generator().manual_seed(0)
    obs, mu = torch.rand(5, 3, 11, generator=g), torch.rand(5, 3, 11, generator=g)
    std = 0.1 + torch.rand(5, 3, 11, generator=g)
    for tau in (0.5, 0.975, 0.025):
        want = 0.0
        for ch in range(3):
            t, p, s = obs[:, ch], mu[:, ch], std[:, ch]
            ge = t.ge(p)
            for mask, scale in ((~ge, 1 - tau), (ge, tau)):
                want = want + scale * D.Laplace(p[mask], s[mask]).log_prob(t[mask]).sum()
        got = tc.MechanisticModel.quantile_loglik(obs, mu, std, tau)
        assert torch.allclose(got, want, rtol=1e-5)


def test_loss_basic_is_the_single_sample_negative_elbo():
    cfg, m = _model()
    obs, iext, rtpr = _batch(cfg)
    eps = torch.randn(6, 15, generator=torch.Generator().manual_seed(3))
    got = m.loss_basic(obs, iext, rtpr, eps=eps)
    loc, scale = m.encoder(obs)
    z = loc + scale * eps
    log_q = D.Normal(loc, scale).log_prob(z).sum()
    zi, zr, ze = z[:, :5], z[:, 5:10], z[:, 10:]
    log_p = D.Normal(0.0, 1.0).log_prob(ze).sum() + D.Normal(*m.p_z_iext_given_iext(iext)).log_prob(zi).sum() \
        + D.Normal(*m.p_z_rtprs_given_rtprs(rtpr)).log_prob(zr).sum()
    _, q75, q50, q25, std = m.decoder(z)
    for mu, tau in ((q50, 0.5), (q75, 0.975), (q25, 0.025)):
        w = torch.where(obs >= mu, tau, 1 - tau)
        log_p = log_p + (w * D.Laplace(mu, std).log_prob(obs)).sum()
    assert torch.allclose(got, -(log_p - log_q), rtol=1e-5)


def test_loss_aux_and_gradient_routing():
    cfg, m = _model()
    obs, iext, rtpr = _batch(cfg)
    eps = torch.randn(6, 10, generator=torch.Generator().manual_seed(4))
    got = m.loss_aux(obs, iext, rtpr, eps=eps)
    loc, scale = m.encoder(obs)
    z = loc[:, :10] + scale[:, :10] * eps
    want = -(D.Normal(loc[:, :10], scale[:, :10]).log_prob(z).sum()
             + 46.0 * (D.Bernoulli(m.q_iext_given_z_iext(z[:, :5])).log_prob(iext).sum()
                       + D.Bernoulli(m.q_rtpr_given_z_rtpr(z[:, 5:])).log_prob(rtpr).sum()))
    assert torch.allclose(got, want, rtol=1e-5)
    got.backward()
    touched = {k.split(".")[0] for k, p in m.named_parameters() if p.grad is not None}
    assert touched == {"encoder", "q_iext_given_z_iext", "q_rtpr_given_z_rtpr"}
    m.zero_grad(set_to_none=True)
    m.loss_basic(obs, iext, rtpr).backward()
    touched = {k.split(".")[0] for k, p in m.named_parameters() if p.grad is not None}
    assert touched == {"encoder", "p_z_iext_given_iext", "p_z_rtprs_given_rtprs", "decoder"}


def test_encoder_conv_matches_reference_layer_sizes():
    enc = tc.EncoderCONV(3, 10, 10, 5, 86, 15, 50)
    assert enc.lin.in_features == (86 - 9 - 4) * 10  # models/encoder_conv.py:25-27
    loc, scale = enc(torch.rand(4, 3, 86))
    assert loc.shape == scale.shape == (4, 15) and bool((scale > 0).all())
    w = enc.conv.weight.reshape(10, -1)
    assert torch.allclose(w @ w.t(), torch.eye(10), atol=1e-5)  # orthogonal init


def test_a_few_steps_reduce_both_objectives():
    cfg, m = _model(T=20)
    obs, iext, rtpr = _batch(cfg, B=16)
    batch = {"observations": obs, "iext": iext, "rtpr": rtpr}
    opt = torch.optim.Adam(m.parameters(), lr=cfg.learning_rate, betas=(0.9, 0.999))
    torch.manual_seed(0)
    first = [float(x) for x in tc.run_batch(m, opt, batch)]
    for _ in range(25):
        last = [float(x) for x in tc.run_batch(m, opt, batch)]
    assert last[0] < first[0] and last[1] < first[1]
    stats = tc.input_pred_stats(m, batch, is_post=True)
    assert set(stats) == {"iext", "rtpr", "l1", "elbo"} and stats["elbo"].shape == (2,)


def test_bernoulli_term_is_finite_when_the_classifier_saturates():
    """y*log(a) + (1-y)*log1p(-a) is NaN for a == 1.0 / 0.0 in fp32; pyro's Bernoulli(probs) clamps."""
    from structured_latent_odes_b200.training_cvs import _bernoulli_logp
    a = torch.tensor([[1.0], [0.0], [1.0], [0.0], [0.3]], requires_grad=True)
    y = torch.tensor([[1.0], [0.0], [0.0], [1.0], [1.0]])
    got = _bernoulli_logp(a, y)
    want = D.Bernoulli(probs=a.detach()).log_prob(y).sum()
    assert torch.isfinite(got) and torch.allclose(got, want, rtol=1e-6)
    got.backward()
    assert torch.isfinite(a.grad).all()
