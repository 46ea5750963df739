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
