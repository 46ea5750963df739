"""The oracle against the committed golden vectors (outputs of the real reference).

Runs everywhere (CPU only, no /root/reference needed)."""
import numpy as np
import pytest
import torch

from oracle import oac_oracle as orc
from tests.util import synth_batch, synth_eps, max_abs, rel_err
from tests import golden_util as gu


def _hidden(meta, start):
    return tuple(int(h) for h in meta[start:])


def test_sac_small_full_weights():
    g = gu.load("sac_small.npz")
    O, A, B, n_steps, seed = [int(v) for v in g['meta'][:5]]
    st = orc.SACState(O, A, hidden=_hidden(g['meta'], 5))
    for name, net in st.nets().items():
        gu.set_net(net, gu.net_from(g, 'init/' + name))
    for s in range(n_steps):
        batch = synth_batch(B, O, A, seed=10 + s)
        eps = synth_eps(2, B, A, seed=100 + s)
        out = orc.sac_step(st, batch, eps[0], eps[1], mode="A")
        d = g['diag'][s]
        assert abs(float(out['qf1_loss']) - d[0]) <= 1e-5 * abs(d[0])
        assert abs(float(out['qf2_loss']) - d[1]) <= 1e-5 * abs(d[1])
        # logged "Policy Loss" is (log_pi - q_new).mean(), without alpha (trainer/trainer.py:236)
        logged = float((out['log_pi'] - out['q_new']).mean())
        assert abs(logged - d[2]) <= 1e-5 * abs(d[2]) + 1e-6
        assert abs(float(out['alpha']) - d[3]) <= 1e-7
        assert abs(float(out['log_pi'].mean()) - d[4]) <= 1e-5
        assert abs(float(out['q_target'].mean()) - d[5]) <= 1e-5
    for name, net in st.nets().items():
        ref = gu.net_from(g, 'final/' + name)
        for k in net:
            assert max_abs(net[k], ref[k]) <= 2e-6, (name, k)
    assert max_abs(st.log_alpha['log_alpha'], g['final/log_alpha']) <= 1e-7


@pytest.mark.parametrize("name", ["sac_riverswim.npz", "sac_humanoid.npz"])
def test_sac_seeded_digests(name):
    g = gu.load(name)
    O, A, B, n_steps, seed = [int(v) for v in g['meta'][:5]]
    torch.manual_seed(seed)
    st = orc.SACState(O, A, hidden=_hidden(g['meta'], 5))
    d0 = gu.digest(st.policy['fc0.weight'])
    if not np.allclose(d0, g['init/policy/fc0.weight'], rtol=0, atol=0):
        pytest.skip("torch RNG stream differs from the one the golden was made with")
    for s in range(n_steps):
        batch = synth_batch(B, O, A, seed=10 + s)
        eps = synth_eps(2, B, A, seed=100 + s)
        out = orc.sac_step(st, batch, eps[0], eps[1], mode="A")
        d = g['diag'][s]
        assert abs(float(out['qf1_loss']) - d[0]) <= 1e-5 * abs(d[0])
        assert abs(float(out['alpha']) - d[3]) <= 1e-7
        assert abs(float(out['q_target'].mean()) - d[5]) <= 1e-5
    for nname, net in st.nets().items():
        for k in net:
            gu.assert_digest_close(net[k], g['final/%s/%s' % (nname, k)], 1e-6, nname + '/' + k)


@pytest.mark.parametrize("name", ["poac_shared_small.npz", "poac_shared_counts_small.npz",
                                  "poac_separate_small.npz"])
def test_poac_small(name):
    g = gu.load(name)
    O, A, B, n_steps, seed, P, share, counts = [int(v) for v in g['meta'][:8]]
    st = orc.ParticleState(O, A, n_estimators=P, share_layers=bool(share), hidden=_hidden(g['meta'], 8),
                           counts=bool(counts), q_min=0.0, q_max=500.0)
    gu.set_net(st.policy, gu.net_from(g, 'init/policy'))
    for i in range(len(st.qfs)):
        gu.set_net(st.qfs[i], gu.net_from(g, 'init/qf%d' % i))
        gu.set_net(st.tfs[i], gu.net_from(g, 'init/tf%d' % i))
    for s in range(n_steps):
        batch = synth_batch(B, O, A, seed=20 + s, counts=bool(counts))
        eps = synth_eps(2, B, A, seed=200 + s)
        orc.poac_step(st, batch, eps[0], eps[1])
    for k, v in gu.net_from(g, 'final/policy').items():
        assert max_abs(st.policy[k], v) <= 2e-6, k
    for i in range(len(st.qfs)):
        for k, v in gu.net_from(g, 'final/qf%d' % i).items():
            assert max_abs(st.qfs[i][k], v) <= 5e-5, (i, k)
        for k, v in gu.net_from(g, 'final/tf%d' % i).items():
            assert max_abs(st.tfs[i][k], v) <= 5e-5, (i, k)
    assert max_abs(st.log_alpha['log_alpha'], g['final/log_alpha']) <= 1e-7


@pytest.mark.parametrize("name", ["goac_shared_small.npz", "goac_shared_counts_small.npz",
                                  "goac_separate_small.npz"])
def test_goac_small(name):
    g = gu.load(name)
    O, A, B, n_steps, seed, share, counts = [int(v) for v in g['meta'][:7]]
    st = orc.GaussianState(O, A, share_layers=bool(share), hidden=_hidden(g['meta'], 7),
                           counts=bool(counts), q_min=0.0, q_max=500.0)
    names = ['policy', 'target_policy', 'q', 'q_target'] + ([] if share else ['std', 'std_target'])
    for n in names:
        gu.set_net(getattr(st, n), gu.net_from(g, 'init/' + n))
    for s in range(n_steps):
        orc.goac_step(st, synth_batch(B, O, A, seed=30 + s, counts=bool(counts)))
    for n in names:
        tol = 2e-6 if 'policy' in n else 5e-5
        for k, v in gu.net_from(g, 'final/' + n).items():
            assert max_abs(getattr(st, n)[k], v) <= tol, (n, k)


def test_explore_twin_small():
    g = gu.load("explore_twin_small.npz")
    O, A, n_obs, seed = [int(v) for v in g['meta'][:4]]
    hidden = _hidden(g['meta'], 4)
    pol, q1, q2 = orc.init_policy(O, A, hidden), orc.init_q(O, A, hidden), orc.init_q(O, A, hidden)
    for net, n in ((pol, 'policy'), (q1, 'qf1'), (q2, 'qf2')):
        gu.set_net(net, gu.net_from(g, 'init/' + n))
    for i in range(n_obs):
        ob = torch.from_numpy(g['obs'][i]).float()
        ac, mu_E, _ = orc.explore(ob, pol, [q1, q2], float(g['beta_UB']), float(g['delta']),
                                  eps_sample=torch.from_numpy(g['eps_sample'][i]))
        assert max_abs(ac, g['action'][i]) <= 2e-6
        mu, _, _ = orc.explore(ob, pol, [q1, q2], float(g['beta_UB']), float(g['delta']), deterministic=True)
        assert rel_err(mu, g['mu_E_deterministic'][i]) <= 2e-6
        # KAT: the shift sits on the KL ball of radius delta
        _, _, _, _, std, _ = orc.policy_forward(pol, ob[None], None, True)
        mu_T = orc.policy_forward(pol, ob[None], None, True)[1][0]
        kl = 0.5 * float((((mu_E - mu_T) / std[0]) ** 2).sum())
        assert abs(kl - float(g["delta"])) <= 5e-2 * float(g["delta"])  # 1e-5 in the denominator matters for tiny grads


def test_explore_ensemble_small():
    g = gu.load("explore_ensemble_small.npz")
    O, A, n_obs, seed, P = [int(v) for v in g['meta'][:5]]
    hidden = _hidden(g['meta'], 5)
    pol, q = orc.init_policy(O, A, hidden), orc.init_q(O, A, hidden, P)
    gu.set_net(pol, gu.net_from(g, 'init/policy'))
    gu.set_net(q, gu.net_from(g, 'init/qf0'))
    for i in range(n_obs):
        ob = torch.from_numpy(g['obs'][i]).float()
        ac, _, _ = orc.explore(ob, pol, [q], float(g['beta_UB']), float(g['delta']), share_layers=True,
                               eps_sample=torch.from_numpy(g['eps_sample'][i]))
        assert max_abs(ac, g['action'][i]) <= 2e-6


def test_replay_counts_golden():
    g = gu.load("replay_counts.npz")
    O, A, N, T, B = [int(v) for v in g['meta']]
    rb = orc.ReplayBufferCount(N, O, A)
    nb = 0
    for t in range(T):
        rb.add_sample(g['stream/obs'][t], g['stream/act'][t], g['stream/rew'][t],
                      g['stream/nobs'][t], g['stream/term'][t])
        if nb < int(g['n_batches']) and t == int(g['batch%d/t' % nb]):
            np.random.seed(t)
            idx = rb.draw_indices(B)
            assert np.array_equal(idx, g['batch%d/indices' % nb])
            b = rb.gather(idx)
            for k, v in b.items():
                ref = g['batch%d/%s' % (nb, k)]
                assert v.dtype == ref.dtype and np.array_equal(v, ref), k
            nb += 1
    assert nb == int(g['n_batches'])
    assert [rb._top, rb._size] == list(g['final/top_size'])
    assert np.array_equal(rb._counts, g['final/counts'])


def test_known_answers():
    """Reference-independent KATs."""
    # (iv) Polyak with tau=1 copies
    a = {'w': torch.randn(5)}
    b = {'w': torch.randn(5)}
    orc.soft_update(a, b, 1.0)
    assert torch.equal(a['w'], b['w'])
    # (iii) Adam step 1 moves every weight by ~lr*sign(g)
    p = {'w': torch.zeros(7)}
    opt = orc.Adam(p, lr=3e-4)
    gr = torch.tensor([1., -2., 3e-3, -4e-5, 5., -6., 7.])
    opt.step({'w': gr})
    assert torch.allclose(p['w'], -3e-4 * torch.sign(gr), rtol=1e-3)
    # (v) log_prob of a=tanh(z)
    torch.manual_seed(0)
    pol = orc.init_policy(6, 4, (16, 16))
    obs = torch.randn(9, 6)
    eps = torch.randn(9, 4)
    a_, mean, log_std, lp, std, z = orc.policy_forward(pol, obs, eps)
    ref = torch.distributions.Normal(mean, std).log_prob(z) - torch.log(1 - a_ * a_ + 1e-6)
    assert torch.allclose(lp, ref.sum(1, keepdim=True), atol=1e-5)
    assert abs(orc.norm_ppf(0.95) - 1.6448536269514722) < 1e-9
