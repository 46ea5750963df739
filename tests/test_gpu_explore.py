"""Fused optimistic-exploration kernel against the reference's golden actions and the oracle."""
import numpy as np
import pytest
import torch

from oracle import oac_oracle as orc
from tests.util import rel_err, max_abs
from tests import golden_util as gu
from tests.gpu_util import Box, producers, load_net

pytestmark = pytest.mark.gpu


def test_explore_twin_golden():
    from oac_explore_b200.optimistic_exploration import get_optimistic_exploration_action
    g = gu.load("explore_twin_small.npz")
    O, A, n_obs, seed = [int(v) for v in g['meta'][:4]]
    H = int(g['meta'][4])
    pp, qp = producers(O, A, H)
    pol, q1, q2 = pp(), qp(), qp()
    for net, n in ((pol, 'policy'), (q1, 'qf1'), (q2, 'qf2')):
        load_net(net, gu.net_from(g, 'init/' + n))
    hp = dict(beta_UB=float(g['beta_UB']), delta=float(g['delta']), share_layers=False)
    for i in range(n_obs):
        ac, info = get_optimistic_exploration_action(g['obs'][i], policy=pol, qfs=[q1, q2], hyper_params=hp,
                                                     eps=g['eps_sample'][i])
        assert info == {} and ac.dtype == np.float32 and ac.shape == (A,)
        assert max_abs(ac, g['action'][i]) <= 1e-5
        mu, _ = get_optimistic_exploration_action(g['obs'][i], policy=pol, qfs=[q1, q2], hyper_params=hp,
                                                  deterministic=True)
        assert rel_err(mu, g['mu_E_deterministic'][i]) <= 1e-5


def test_explore_ensemble_golden():
    from oac_explore_b200.optimistic_exploration import get_optimistic_exploration_action
    g = gu.load("explore_ensemble_small.npz")
    O, A, n_obs, seed, P = [int(v) for v in g['meta'][:5]]
    H = int(g['meta'][5])
    pp, qp = producers(O, A, H, q_out=P)
    pol, q = pp(), qp()
    load_net(pol, gu.net_from(g, 'init/policy'))
    load_net(q, gu.net_from(g, 'init/qf0'))
    hp = dict(beta_UB=float(g['beta_UB']), delta=float(g['delta']), share_layers=True)
    for i in range(n_obs):
        ac, _ = get_optimistic_exploration_action(g['obs'][i], policy=pol, qfs=[q], hyper_params=hp,
                                                  eps=g['eps_sample'][i])
        assert max_abs(ac, g['action'][i]) <= 1e-5


@pytest.mark.parametrize("O,A,H", [(376, 17, 256), (1, 1, 256)])
def test_explore_vs_oracle_full_shapes(O, A, H):
    from oac_explore_b200.optimistic_exploration import explore_batch
    torch.manual_seed(4)
    pp, qp = producers(O, A, H)
    pol, q1, q2 = pp(), qp(), qp()
    torch.manual_seed(4)
    opol, oq1, oq2 = orc.init_policy(O, A, (H, H)), orc.init_q(O, A, (H, H)), orc.init_q(O, A, (H, H))
    rng = np.random.RandomState(0)
    n = 9
    obs = rng.randn(n, O)
    eps = rng.randn(n, A).astype(np.float32)
    hp = dict(beta_UB=4.66, delta=23.53, share_layers=False)
    ac, mu_E, grad = explore_batch(obs, pol, [q1, q2], hp, eps=eps)
    for i in range(n):
        a_ref, mu_ref, g_ref = orc.explore(torch.from_numpy(obs[i]).float(), opol, [oq1, oq2], 4.66, 23.53,
                                           eps_sample=torch.from_numpy(eps[i]))
        assert rel_err(grad[i], g_ref) <= 1e-5
        assert rel_err(mu_E[i], mu_ref) <= 1e-5
        assert max_abs(ac[i], a_ref) <= 1e-5
        # KAT: the shift lies on the KL ball  1/2 (mu_E-mu_T)^T Sigma^-1 (mu_E-mu_T) = delta
        _, mu_T, _, _, std, _ = orc.policy_forward(opol, torch.from_numpy(obs[i]).float()[None], None, True)
        kl = 0.5 * float((((torch.from_numpy(mu_E[i]) - mu_T[0]) / std[0]) ** 2).sum())
        assert abs(kl - 23.53) <= 0.05 * 23.53
    # device-noise mode: actions are valid tanh samples and differ between calls
    a1, _, _ = explore_batch(obs, pol, [q1, q2], hp)
    a2, _, _ = explore_batch(obs, pol, [q1, q2], hp)
    assert np.all(np.abs(a1) <= 1) and not np.array_equal(a1, a2)


def test_policy_and_q_forward_vs_oracle():
    O, A, H, n = 23, 6, 64, 37
    torch.manual_seed(9)
    pp, qp = producers(O, A, H, q_out=3)
    pol, q = pp(), qp()
    torch.manual_seed(9)
    opol, oq = orc.init_policy(O, A, (H, H)), orc.init_q(O, A, (H, H), 3)
    obs, act = torch.randn(n, O), torch.rand(n, A) * 2 - 1
    out = q(obs.cuda(), act.cuda()).cpu()
    assert rel_err(out, orc.q_forward(oq, obs, act)) <= 1e-5
    a, mean, log_std, lp, std, pre = [t.cpu() for t in pol(obs.cuda(), deterministic=True)]
    ra, rmean, rls, rlp, rstd, rpre = orc.policy_forward(opol, obs, None, True)
    for x, y in ((a, ra), (mean, rmean), (log_std, rls), (std, rstd)):
        assert rel_err(x, y) <= 1e-5
    assert torch.equal(lp, torch.zeros_like(a))
    # stochastic call: internally consistent with the oracle given the same pre-tanh sample
    a, mean, log_std, lp, std, pre = [t.cpu() for t in pol(obs.cuda(), return_log_prob=True)]
    eps = (pre - mean) / std
    ra, _, _, rlp, _, _ = orc.policy_forward(opol, obs, eps)
    assert rel_err(a, ra) <= 1e-5 and rel_err(lp, rlp) <= 1e-4
    act_np, info = pol.get_action(obs[0].numpy().astype(np.float64), deterministic=True)
    assert act_np.shape == (A,) and info == {}
    assert max_abs(act_np, ra.new_tensor(orc.policy_forward(opol, obs[:1], None, True)[0][0])) <= 1e-5


@pytest.mark.parametrize("mode", ["twin", "ensemble", "deterministic"])
def test_cluster_and_per_cta_kernels_agree(mode):
    """Up to 16 observations run on the 8-CTA-cluster latency kernel, more on the one-CTA-per-observation kernel: the
    same 40 observations through both (one call of 40, five calls of 8) must agree to fp32 summation-order noise."""
    from oac_explore_b200.optimistic_exploration import explore_batch
    O, A, H = 376, 17, 256
    torch.manual_seed(7)
    if mode == "ensemble":
        pp, qp = producers(O, A, H, q_out=5)
        pol, qfs = pp(), [qp()]
        hp = dict(beta_UB=4.66, delta=20.53, share_layers=True)
    else:
        pp, qp = producers(O, A, H)
        pol, qfs = pp(), [qp(), qp()]
        hp = dict(beta_UB=4.66, delta=23.53, share_layers=False)
    det = mode == "deterministic"
    rng = np.random.RandomState(3)
    obs = rng.randn(40, O)
    eps = rng.randn(40, A).astype(np.float32)
    big = explore_batch(obs, pol, qfs, hp, deterministic=det, eps=eps)
    for i in range(0, 40, 8):
        small = explore_batch(obs[i:i + 8], pol, qfs, hp, deterministic=det, eps=eps[i:i + 8])
        for a, b in zip(big, small):
            assert max_abs(a[i:i + 8], b) <= 2e-5, (mode, i, max_abs(a[i:i + 8], b))


# ---- round 2: the ``trainer=`` branch, per-critic exp flags, Humanoid-size ensembles, per-seed batched exploration ----
class _PredictTwin(object):
    """SACTrainer-shaped object: ``predict(obs, action, upper_bound=, beta_UB=)`` over ``qfs[0:2]`` (trainer/trainer.py:105-123)."""
    def __init__(self, qfs):
        self.qfs = qfs

    def predict(self, obs, action, upper_bound=True, beta_UB=4.46, both_values=False):
        raise AssertionError("the fused kernel evaluates predict's formula itself")


class _PredictQuantile(_PredictTwin):
    """ParticleTrainer-shaped: sorted particle ``delta_index`` (trainer/particle_trainer_oac.py:147-167)."""
    def __init__(self, qfs, delta_index):
        self.qfs, self.delta_index = qfs, delta_index

    def predict(self, obs, action, all_particles=False, upper_bound=True, beta_UB=None):
        raise AssertionError


class _PredictGaussian(object):
    """GaussianTrainer-shaped: predict(obs, action, std=True) has no ``upper_bound`` (trainer/gaussian_trainer.py:161)."""
    def __init__(self, qfs):
        self.qfs = qfs

    def predict(self, obs, action, std=True):
        raise AssertionError


@pytest.mark.parametrize("tag", ['sac', 'poac_shared', 'poac_separate', 'goac', 'goac_sep'])
def test_explore_trainer_branch_golden(tag):
    """optimistic_exploration.py:38-39 with ``trainer=`` given (SACTrainer.predict -> twin, ParticleTrainer.predict ->
    OAC_EXPLORE_QUANTILE), and the G-OAC critics (exp on the std output; for the separate mean / std nets only the
    second critic is exponentiated) against actions the real reference produced."""
    from oac_explore_b200.optimistic_exploration import get_optimistic_exploration_action
    from tests.test_oracle_golden_r2 import explore_case_nets
    g = gu.load("explore_trainer_small.npz")
    O, A, n_obs = [int(v) for v in g['meta'][:3]]
    H = int(g['meta'][3])
    from oac_explore_b200.networks import get_policy_producer, get_q_producer
    pp = get_policy_producer(O, A, [H, H])

    def make_q(heads):
        return get_q_producer(O, A, [H, H], output_size=heads)()

    pol, qfs, names, kw = explore_case_nets(g, tag, pp, make_q)
    load_net(pol, gu.net_from(g, tag + '/policy'))
    for q, n in zip(qfs, names):
        load_net(q, gu.net_from(g, tag + '/' + n))
    trainer = None
    if tag == 'sac':
        trainer = _PredictTwin(qfs)
    elif tag.startswith('poac'):
        trainer = _PredictQuantile(qfs, kw['trainer']['delta_index'])
    elif tag == 'goac':
        qfs[0].positive = [False, True]
    else:
        qfs[1].positive = True
    hp = dict(beta_UB=float(g[tag + '/beta_UB']), delta=float(g[tag + '/delta']), share_layers=kw.get('share_layers', False))
    for i in range(n_obs):
        # with trainer= the reference ignores the qfs argument: pass garbage to prove we do too
        ac, _ = get_optimistic_exploration_action(g['obs'][i], policy=pol, qfs=qfs if trainer is None else [qfs[0]],
                                                  trainer=trainer, hyper_params=hp, eps=g[tag + '/eps_sample'][i])
        assert max_abs(ac, g[tag + '/action'][i]) <= 1e-5, (tag, i, max_abs(ac, g[tag + '/action'][i]))


def test_explore_gaussian_trainer_raises_like_the_reference():
    from oac_explore_b200.optimistic_exploration import get_optimistic_exploration_action
    O, A, H = 5, 2, 16
    pp, qp = producers(O, A, H, q_out=2)
    pol, q = pp(), qp()
    with pytest.raises(TypeError):
        get_optimistic_exploration_action(np.zeros(O), policy=pol, qfs=[q], trainer=_PredictGaussian([q]),
                                          hyper_params=dict(beta_UB=4.66, delta=20.0, share_layers=True))


@pytest.mark.parametrize("mode", ["ensemble", "quantile", "goac"])
def test_explore_humanoid_p10_vs_oracle(mode):
    """Humanoid shapes, P = 10 heads on a shared trunk (the layout reproduce_p-oac.sh uses): the ensemble mean + beta *
    unbiased-std branch, the ParticleTrainer quantile route, and the G-OAC 2-head critic, against the oracle."""
    from oac_explore_b200.optimistic_exploration import explore_batch
    O, A, H = 376, 17, 256
    P = 2 if mode == "goac" else 10
    torch.manual_seed(21)
    pp, qp = producers(O, A, H, q_out=P)
    pol, q = pp(), qp()
    torch.manual_seed(21)
    opol, oq = orc.init_policy(O, A, (H, H)), orc.init_q(O, A, (H, H), P)
    # spread the heads like a trained particle critic (init biases are linspace(q_min, q_max), particle_trainer_oac.py:75)
    bias = torch.linspace(0., 5., P)
    oq['last_fc.bias'].copy_(bias)
    q.load_state_dict({k: v for k, v in oq.items()})
    rng = np.random.RandomState(5)
    n = 6
    obs = rng.randn(n, O)
    eps = rng.randn(n, A).astype(np.float32)
    hp = dict(beta_UB=4.66, delta=20.53, share_layers=True)
    kw, trainer = dict(share_layers=True), None
    if mode == "quantile":
        trainer = _PredictQuantile([q], 9)
        kw['trainer'] = dict(kind='particle', delta_index=9, share_layers=True)
    if mode == "goac":
        q.positive = [False, True]
        kw['positives'] = [[False, True]]
    ac, mu_E, grad = explore_batch(obs, pol, [q], hp, trainer=trainer, eps=eps)
    for i in range(n):
        a_ref, mu_ref, g_ref = orc.explore(torch.from_numpy(obs[i]).float(), opol, [oq], 4.66, 20.53,
                                           eps_sample=torch.from_numpy(eps[i]), **kw)
        assert rel_err(grad[i], g_ref) <= 2e-5, (mode, i, rel_err(grad[i], g_ref))
        assert rel_err(mu_E[i], mu_ref) <= 1e-5
        assert max_abs(ac[i], a_ref) <= 1e-5


def test_exploration_noise_follows_torch_seed_and_policy_counter():
    """the Philox key of the sampling noise derives from torch.initial_seed() and the call counter lives on the
    policy object -- different --seed => different noise, same seed + same call index => same noise, and a pickled
    policy resumes its stream instead of restarting it."""
    from oac_explore_b200.optimistic_exploration import get_optimistic_exploration_action
    O, A, H = 11, 3, 32
    hp = dict(beta_UB=4.66, delta=23.53, share_layers=False)
    ob = np.random.RandomState(0).randn(O)

    def run(seed, n_calls):
        torch.manual_seed(seed)
        pp, qp = producers(O, A, H)
        torch.manual_seed(99)                      # same weights every time
        pol, q1, q2 = pp(), qp(), qp()
        torch.manual_seed(seed)
        return [get_optimistic_exploration_action(ob, policy=pol, qfs=[q1, q2], hyper_params=hp)[0] for _ in range(n_calls)], pol

    a0, pol0 = run(1, 3)
    a1, _ = run(1, 3)
    a2, _ = run(2, 3)
    for x, y in zip(a0, a1):
        assert np.array_equal(x, y)
    assert not np.array_equal(a0[0], a2[0])
    assert not np.array_equal(a0[0], a0[1])
    assert pol0._explore_calls == 3


def test_group_explorer_rollout_equals_per_seed_calls():
    """ONE launch serves the current observations of every seed of a SACSeedGroup with that seed's own policy
    and critics (path_collector.py:214-232 vectorised over seeds).  A rollout-shaped loop over toy environments: the
    batched actions equal the per-seed get_optimistic_exploration_action calls and the oracle, for 1 and 2 envs per seed,
    through the double-buffered submit / collect ring."""
    from oac_explore_b200.seed_group import SACSeedGroup
    from oac_explore_b200.optimistic_exploration import explore_batch
    O, A, H, B = 376, 17, 256, 256
    ids = [5, 9, 2, 14, 7, 3, 11, 8, 1, 0, 4, 6, 10, 12, 13, 15, 16, 17, 18, 19]       # 20 seeds: the per-CTA kernel
    grp = SACSeedGroup(ids, O, A, hidden=H, batch=B, gemm_path=0)
    hp = dict(beta_UB=4.66, delta=23.53, share_layers=False)
    ex = grp.explorer(hp, max_obs=2 * len(ids))
    rng = np.random.RandomState(1)
    S = len(ids)
    obs = rng.randn(S, O)                                   # "env.reset()" of every seed's environment
    for step in range(3):
        eps = rng.randn(S, A).astype(np.float32)
        t = ex.submit(obs, eps=eps)                         # one launch for all seeds
        acts, mu = ex.collect(t)
        for s in (0, 7, 19):
            n = grp.nets[s]
            a_one, mu_one, _ = explore_batch(obs[s:s + 1], n['policy'], [n['qf1'], n['qf2']], hp, eps=eps[s:s + 1])
            assert max_abs(acts[s], a_one[0]) <= 2e-5 and max_abs(mu[s], mu_one[0]) <= 2e-5
        if step == 0:
            sid = ids[3]
            torch.manual_seed(sid)
            st = orc.SACState(O, A, hidden=(H, H))
            a_ref, _, _ = orc.explore(torch.from_numpy(obs[3]).float(), st.policy, [st.qf1, st.qf2], 4.66, 23.53,
                                      eps_sample=torch.from_numpy(eps[3]))
            assert max_abs(acts[3], a_ref) <= 1e-4
        obs = obs + 0.1 * np.tanh(acts @ rng.randn(A, O))   # "env.step(a)": next observation depends on the action
    # two environments per seed, submitted as two half batches kept in flight together (double buffering)
    slots = np.repeat(np.arange(S), 2)
    obs2 = rng.randn(2 * S, O)
    eps2 = rng.randn(2 * S, A).astype(np.float32)
    t0 = ex.submit(obs2[:S], slots=slots[:S], eps=eps2[:S])
    t1 = ex.submit(obs2[S:], slots=slots[S:], eps=eps2[S:])
    with pytest.raises(RuntimeError):
        ex.submit(obs2[:S], slots=slots[:S])               # ring of depth 2 is full
    a0, _ = ex.collect(t0)
    a1, _ = ex.collect(t1)
    both = np.concatenate([a0, a1])
    for i in (0, 1, 2 * S - 1, S):
        n = grp.nets[slots[i]]
        a_one, _, _ = explore_batch(obs2[i:i + 1], n['policy'], [n['qf1'], n['qf2']], hp, eps=eps2[i:i + 1])
        assert max_abs(both[i], a_one[0]) <= 2e-5
    # device noise: valid tanh samples, different between seeds
    a = ex.actions(np.repeat(obs[:1], S, axis=0))
    assert np.all(np.abs(a) <= 1) and not np.array_equal(a[0], a[1])
