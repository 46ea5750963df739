"""The oracle against the round-2 golden vectors (outputs of the real reference, tests/golden/make_golden_r2.py):
K = 20 steps at Humanoid shapes, BASELINE config 1 on real riverswim transitions, and the ``trainer=`` branch of
the exploration function.  CPU only; /root/reference is not needed."""
import numpy as np
import pytest
import torch

from oracle import oac_oracle as orc
from tests.util import synth_batch, synth_eps, max_abs, rel_err
from tests import golden_util as gu


def _seeded_state(g):
    O, A, B, n_steps, seed = [int(v) for v in g['meta'][:5]]
    hidden = tuple(int(h) for h in g['meta'][-2:])
    torch.manual_seed(seed)
    st = orc.SACState(O, A, hidden=hidden)
    if not np.allclose(gu.digest(st.policy['fc0.weight']), g['init/policy/fc0.weight'], rtol=0, atol=0):
        pytest.skip("torch RNG stream differs from the one the golden was made with")
    return st, O, A, B, n_steps


def check_diag(out, d, tol):
    """d = [QF1 Loss, QF2 Loss, Policy Loss (logged, no alpha), Alpha, Log Pis Mean, Q Targets Mean, Q1 Mean, Q2 Mean,
    Policy mu Mean, Policy log std Mean] of the reference (trainer/trainer.py:243-279)."""
    assert abs(float(out['qf1_loss']) - d[0]) <= tol * abs(d[0])
    assert abs(float(out['qf2_loss']) - d[1]) <= tol * abs(d[1])
    logged = float((out['log_pi'] - out['q_new']).mean())
    assert abs(logged - d[2]) <= tol * abs(d[2]) + 1e-6
    assert abs(float(out['alpha']) - d[3]) <= 1e-6
    assert abs(float(out['log_pi'].mean()) - d[4]) <= tol * abs(d[4]) + 1e-6
    assert abs(float(out['q_target'].mean()) - d[5]) <= tol * abs(d[5]) + 1e-6
    assert abs(float(out['q1_pred'].mean()) - d[6]) <= tol * abs(d[6]) + 1e-6
    assert abs(float(out['policy_mean'].mean()) - d[8]) <= 1e-6
    assert abs(float(out['policy_log_std'].mean()) - d[9]) <= 1e-6


def test_sac_humanoid_k20():
    g = gu.load("sac_humanoid_k20.npz")
    st, O, A, B, n_steps = _seeded_state(g)
    assert n_steps == 20
    for s in range(n_steps):
        batch = synth_batch(B, O, A, seed=10 + s)
        eps = synth_eps(2, B, A, seed=100 + s)
        out = orc.sac_step(st, batch, eps[0], eps[1], mode="A")
        check_diag(out, g['diag'][s], 2e-5)
    for nname, net in st.nets().items():
        for k in net:
            gu.assert_digest_close(net[k], g['final/%s/%s' % (nname, k)], 2e-6, nname + '/' + k)


def riverswim_replay(g, rb):
    for t in range(g['stream/obs'].shape[0]):
        rb.add_sample(g['stream/obs'][t], g['stream/act'][t], g['stream/rew'][t], g['stream/nobs'][t],
                      g['stream/term'][t], env_info={})


def test_sac_riverswim_real_data():
    """BASELINE config 1: real transitions of envs/river_swim_continuous.py collected by the reference's own
    optimistic exploration, 20 updates through random_batch -> np_to_pytorch_batch -> train_from_torch."""
    g = gu.load("sac_riverswim_real.npz")
    st, O, A, B, n_steps = _seeded_state(g)
    assert (O, A) == (1, 1)
    rb = orc.ReplayBuffer(10000, O, A)
    riverswim_replay(g, rb)
    assert rb.num_steps_can_sample() == 2000
    # real riverswim data: rewards are 0 or small/large, observations live in [0, 25]
    assert g['stream/obs'].min() >= 0 and g['stream/obs'].max() <= 25 and set(np.unique(g['stream/rew'])) <= {0.0, 5e-4, 1.0}
    for s in range(n_steps):
        np.random.seed(1000 + s)
        idx = rb.draw_indices(B)
        assert np.array_equal(idx, g['indices'][s])
        batch = orc.np_to_torch_batch(rb.gather(idx))
        eps = synth_eps(2, B, A, seed=300 + s)
        out = orc.sac_step(st, batch, eps[0], eps[1], mode="A")
        check_diag(out, g['diag'][s], 2e-5)
    for nname, net in st.nets().items():
        for k in net:
            gu.assert_digest_close(net[k], g['final/%s/%s' % (nname, k)], 2e-6, nname + '/' + k)
        assert max_abs(net['last_fc.weight'], g['final_full/%s/last_fc.weight' % nname]) <= 2e-6
    assert max_abs(st.log_alpha['log_alpha'], g['final/log_alpha']) <= 1e-6


EXPLORE_CASES = ['sac', 'poac_shared', 'poac_separate', 'goac', 'goac_sep']


def explore_case_nets(g, tag, make_policy, make_q):
    """(policy, qfs, kwargs of orc.explore) of one sub-case of explore_trainer_small.npz."""
    O, A = int(g['meta'][0]), int(g['meta'][1])
    pol = make_policy()
    kw = {}
    if tag == 'sac':
        names, heads = ['qf1', 'qf2'], 1
        kw = dict(trainer=dict(kind='sac'))
    elif tag.startswith('poac'):
        P, share, delta_index = [int(v) for v in g[tag + '/meta']]
        names, heads = (['qf0'], P) if share else (['qf%d' % i for i in range(P)], 1)
        kw = dict(trainer=dict(kind='particle', delta_index=delta_index, share_layers=bool(share)), share_layers=bool(share))
    elif tag == 'goac':
        names, heads = ['qf0'], 2
        kw = dict(share_layers=True, positives=[[False, True]])
    else:
        names, heads = ['qf0', 'qf1'], 1
        kw = dict(share_layers=False, positives=[False, True])
    qfs = [make_q(heads) for _ in names]
    return pol, qfs, names, kw


@pytest.mark.parametrize("tag", EXPLORE_CASES)
def test_explore_trainer_branch(tag):
    g = gu.load("explore_trainer_small.npz")
    O, A, n_obs = [int(v) for v in g['meta'][:3]]
    hidden = tuple(int(h) for h in g['meta'][3:])
    pol, qfs, names, kw = explore_case_nets(g, tag, lambda: orc.init_policy(O, A, hidden),
                                            lambda heads: orc.init_q(O, A, hidden, heads))
    gu.set_net(pol, gu.net_from(g, tag + '/policy'))
    for q, n in zip(qfs, names):
        gu.set_net(q, gu.net_from(g, tag + '/' + n))
    for i in range(n_obs):
        ob = torch.from_numpy(g['obs'][i]).float()
        ac, _, _ = orc.explore(ob, pol, qfs, float(g[tag + '/beta_UB']), float(g[tag + '/delta']),
                               eps_sample=torch.from_numpy(g[tag + '/eps_sample'][i]), **kw)
        assert max_abs(ac, g[tag + '/action'][i]) <= 2e-6, (tag, i)


def test_explore_gaussian_trainer_raises_like_the_reference():
    O, A, hidden = 5, 2, (8, 8)
    pol, q = orc.init_policy(O, A, hidden), orc.init_q(O, A, hidden, 2)
    with pytest.raises(TypeError):
        orc.explore(torch.zeros(O), pol, [q], 4.66, 20.0, trainer=dict(kind='gaussian'))


def test_tf32_model_rounding():
    """round_tf32 = cvt.rna.tf32.f32: 10 explicit mantissa bits, nearest, ties away from zero."""
    x = torch.tensor([1.0, 1.0 + 2 ** -11, 1.0 + 2 ** -11 + 2 ** -20, -1.0 - 2 ** -11, 1.0 + 2 ** -12, 3.14159])
    r = orc.round_tf32(x)
    assert r.tolist() == [1.0, 1.0 + 2 ** -10, 1.0 + 2 ** -10, -1.0 - 2 ** -10, 1.0, 3.140625]
    assert torch.all((r.view(torch.int32) & 0x1FFF) == 0)
    # the model only touches fp32: the fp64 arbitration oracle is unaffected
    with orc.tf32_mode("all"):
        a = torch.randn(4, 8, dtype=torch.float64)
        w = torch.randn(3, 8, dtype=torch.float64)
        assert torch.equal(orc.linear(a, w, torch.zeros(3, dtype=torch.float64)), a @ w.t())
        # fp32: products of rounded operands, exact in fp32
        a32, w32 = a.float(), w.float()
        y = orc.linear(a32, w32, torch.zeros(3))
        assert torch.allclose(y, orc.round_tf32(a32) @ orc.round_tf32(w32).t(), rtol=0, atol=1e-6)
    assert orc._TF32["mode"] is None
