"""Pins oracle/oac_oracle.py against the UNMODIFIED reference, imported in-process.

Runs only where /root/reference exists (the build container); skipped elsewhere.
The committed golden vectors (tests/golden/) carry the same comparison to boxes
without the reference -- see tests/test_oracle_golden.py.
"""
import numpy as np
import pytest
import torch

from oracle import oac_oracle as orc
from oracle import ref_import as ri
from tests.util import synth_batch, synth_eps, rel_err, max_abs

pytestmark = [
    pytest.mark.reference,
    pytest.mark.skipif(not ri.reference_available(), reason="reference not present"),
]

SHAPES = [(1, 1, 256), (376, 17, 256), (11, 3, 32)]


def _sd(net):
    return {k: v.detach().clone() for k, v in net.state_dict().items()}


def _cmp_net(ours, ref_net, tol, what):
    sd = ref_net.state_dict()
    for k in ours:
        e = max_abs(ours[k], sd[k])
        assert e <= tol, "%s %s max|diff|=%g" % (what, k, e)


@pytest.mark.parametrize("O,A,B", SHAPES)
def test_init_stream_matches(O, A, B):
    ref = ri.load_reference()
    _, ac_space = ri.make_spaces(O, A)
    pp, qp = ri.make_producers(O, A)
    torch.manual_seed(3)
    tr = ref.trainer.SACTrainer(pp, qp, action_space=ac_space)
    torch.manual_seed(3)
    st = orc.SACState(O, A)
    for name, net in st.nets().items():
        _cmp_net(net, getattr(tr, name), 0.0, name)


@pytest.mark.parametrize("O,A,B", SHAPES)
@pytest.mark.parametrize("auto_alpha", [True, False])
def test_sac_step_mode_a(O, A, B, auto_alpha):
    ref = ri.load_reference()
    _, ac_space = ri.make_spaces(O, A)
    pp, qp = ri.make_producers(O, A)
    torch.manual_seed(0)
    tr = ref.trainer.SACTrainer(pp, qp, action_space=ac_space, policy_lr=3e-4, qf_lr=3e-4,
                                soft_target_tau=5e-3, use_automatic_entropy_tuning=auto_alpha)
    ri.mode_a(tr)
    torch.manual_seed(0)
    st = orc.SACState(O, A, use_automatic_entropy_tuning=auto_alpha)
    n_steps = 5
    for s in range(n_steps):
        batch = synth_batch(B, O, A, seed=10 + s)
        eps = synth_eps(2, B, A, seed=100 + s)
        with ri.injected_noise(eps):
            tr._need_to_update_eval_statistics = True
            tr.train_from_torch(dict(batch))
        out = orc.sac_step(st, batch, eps[0], eps[1], mode="A")
        es = tr.eval_statistics
        assert abs(float(out['qf1_loss']) - float(es['QF1 Loss'])) <= 1e-5 * abs(float(es['QF1 Loss'])) + 1e-7
        assert abs(float(out['qf2_loss']) - float(es['QF2 Loss'])) <= 1e-5 * abs(float(es['QF2 Loss'])) + 1e-7
        assert abs(float(out['log_pi'].mean()) - float(es['Log Pis Mean'])) <= 1e-5
        assert abs(float(out['q_target'].mean()) - float(es['Q Targets Mean'])) <= 1e-5
        if auto_alpha:
            assert abs(float(out['alpha']) - es['Alpha']) <= 1e-7
    for name, net in st.nets().items():
        # a few ulp of fp32 rounding per step (summation order / Adam formula form)
        _cmp_net(net, getattr(tr, name), 2e-6, name)


def test_sac_mode_b_differs_from_a():
    """Modes A and B differ only in the policy update, by O(lr)."""
    O, A, B = 376, 17, 256
    sts = []
    for mode in "AB":
        torch.manual_seed(0)
        st = orc.SACState(O, A)
        batch = synth_batch(B, O, A, seed=10)
        eps = synth_eps(2, B, A, seed=100)
        out = orc.sac_step(st, batch, eps[0], eps[1], mode=mode)
        sts.append((st, out))
    for k in sts[0][0].qf1:
        assert max_abs(sts[0][0].qf1[k], sts[1][0].qf1[k]) == 0.0
    e = rel_err(sts[0][1]['grad_policy']['fc0.weight'], sts[1][1]['grad_policy']['fc0.weight'])
    assert 1e-6 < e < 1e-2


@pytest.mark.parametrize("share_layers", [True, False])
@pytest.mark.parametrize("counts", [False, True])
def test_poac_step(share_layers, counts):
    O, A, B, P = 23, 5, 64, 6
    ref = ri.load_reference()
    _, ac_space = ri.make_spaces(O, A)
    pp, qp = ri.make_producers(O, A, q_out=P if share_layers else 1)
    kw = dict(policy_lr=3e-4, qf_lr=3e-4, soft_target_tau=5e-3, use_automatic_entropy_tuning=True,
              delta=0.95, q_min=0.0, q_max=500.0)
    torch.manual_seed(1)
    tr = ref.particle_trainer_oac.ParticleTrainer(pp, qp, n_estimators=P, action_space=ac_space,
                                                  share_layers=share_layers, counts=counts,
                                                  deterministic=False, **kw)
    ri.mode_a(tr)
    torch.manual_seed(1)
    st = orc.ParticleState(O, A, n_estimators=P, share_layers=share_layers, counts=counts, **kw)
    assert st.delta_index == tr.delta_index
    for s in range(4):
        batch = synth_batch(B, O, A, seed=20 + s, counts=counts)
        eps = synth_eps(2, B, A, seed=200 + s)
        with ri.injected_noise(eps):
            tr.train_from_torch(dict(batch))
        orc.poac_step(st, batch, eps[0], eps[1])
    _cmp_net(st.policy, tr.policy, 2e-6, "policy")
    for i in range(len(st.qfs)):
        _cmp_net(st.qfs[i], tr.qfs[i], 5e-5, "qf%d" % i)   # values ~500: 5e-5 abs = 1e-7 rel
        _cmp_net(st.tfs[i], tr.tfs[i], 5e-5, "tf%d" % i)
    assert max_abs(st.log_alpha['log_alpha'], tr.log_alpha.detach()) <= 1e-7


@pytest.mark.parametrize("share_layers", [True, False])
def test_poac_std_soft_update(share_layers):
    """trainer/particle_trainer_oac.py:210-219: targets keep the current particle spread and move the mean."""
    O, A, B, P = 23, 5, 64, 6
    ref = ri.load_reference()
    _, ac_space = ri.make_spaces(O, A)
    pp, qp = ri.make_producers(O, A, q_out=P if share_layers else 1)
    kw = dict(policy_lr=3e-4, qf_lr=3e-4, soft_target_tau=5e-3, use_automatic_entropy_tuning=True,
              delta=0.95, q_min=0.0, q_max=500.0, std_soft_update=True, std_soft_update_prob=0.3)
    torch.manual_seed(1)
    tr = ref.particle_trainer_oac.ParticleTrainer(pp, qp, n_estimators=P, action_space=ac_space,
                                                  share_layers=share_layers, deterministic=False, **kw)
    ri.mode_a(tr)
    torch.manual_seed(1)
    st = orc.ParticleState(O, A, n_estimators=P, share_layers=share_layers, **kw)
    for s in range(4):
        batch = synth_batch(B, O, A, seed=20 + s)
        eps = synth_eps(2, B, A, seed=200 + s)
        with ri.injected_noise(eps):
            tr.train_from_torch(dict(batch))
        orc.poac_step(st, batch, eps[0], eps[1])
    _cmp_net(st.policy, tr.policy, 2e-6, "policy")
    for i in range(len(st.qfs)):
        _cmp_net(st.qfs[i], tr.qfs[i], 5e-5, "qf%d" % i)
        _cmp_net(st.tfs[i], tr.tfs[i], 5e-5, "tf%d" % i)


@pytest.mark.parametrize("share_layers", [True, False])
@pytest.mark.parametrize("counts", [False, True])
def test_goac_step(share_layers, counts):
    O, A, B = 23, 5, 64
    ref = ri.load_reference()
    _, ac_space = ri.make_spaces(O, A)
    pp, qp = ri.make_producers(O, A, q_out=2 if share_layers else 1)
    kw = dict(policy_lr=3e-4, qf_lr=3e-4, std_lr=3e-5, soft_target_tau=5e-3, delta=0.95,
              q_min=0.0, q_max=500.0)
    torch.manual_seed(2)
    tr = ref.gaussian_trainer.GaussianTrainer(pp, qp, n_estimators=2, action_space=ac_space,
                                              share_layers=share_layers, counts=counts, **kw)
    torch.manual_seed(2)
    st = orc.GaussianState(O, A, share_layers=share_layers, counts=counts, **kw)
    assert abs(st.standard_bound - tr.standard_bound) < 1e-12
    for s in range(4):
        batch = synth_batch(B, O, A, seed=30 + s, counts=counts)
        tr.train_from_torch(dict(batch))
        orc.goac_step(st, batch)
    _cmp_net(st.policy, tr.policy, 2e-6, "policy")
    _cmp_net(st.target_policy, tr.target_policy, 2e-6, "target_policy")
    _cmp_net(st.q, tr.q, 5e-5, "q")
    _cmp_net(st.q_target, tr.q_target, 5e-5, "q_target")
    if not share_layers:
        _cmp_net(st.std, tr.std, 5e-5, "std")
        _cmp_net(st.std_target, tr.std_target, 5e-5, "std_target")


@pytest.mark.parametrize("O,A", [(376, 17), (1, 1)])
def test_explore_twin(O, A):
    ref = ri.load_reference()
    _, ac_space = ri.make_spaces(O, A)
    pp, qp = ri.make_producers(O, A)
    torch.manual_seed(4)
    tr = ref.trainer.SACTrainer(pp, qp, action_space=ac_space)
    torch.manual_seed(4)
    st = orc.SACState(O, A)
    hp = dict(beta_UB=4.66, delta=23.53, share_layers=False)
    rng = np.random.RandomState(0)
    for i in range(5):
        ob = rng.randn(O)
        torch.manual_seed(50 + i)
        ac_ref, _ = ref.optimistic_exploration.get_optimistic_exploration_action(
            ob, policy=tr.policy, qfs=tr.qfs, hyper_params=hp)
        # the reference consumes two draws of size A: rsample (discarded), then sample
        torch.manual_seed(50 + i)
        torch.normal(torch.zeros(A), torch.ones(A))
        eps2 = torch.normal(torch.zeros(A), torch.ones(A))
        ac, mu_E, _ = orc.explore(torch.from_numpy(ob).float(), st.policy, [st.qf1, st.qf2],
                                  4.66, 23.53, eps_sample=eps2)
        assert ac_ref.dtype == np.float32 and ac_ref.shape == (A,)
        assert max_abs(ac, ac_ref) <= 2e-6
        mu_ref, _ = ref.optimistic_exploration.get_optimistic_exploration_action(
            ob, policy=tr.policy, qfs=tr.qfs, hyper_params=hp, deterministic=True)
        mu, _, _ = orc.explore(torch.from_numpy(ob).float(), st.policy, [st.qf1, st.qf2],
                               4.66, 23.53, deterministic=True)
        assert rel_err(mu, mu_ref) <= 2e-6


def test_explore_ensemble_shared():
    O, A, P = 23, 5, 6
    ref = ri.load_reference()
    _, ac_space = ri.make_spaces(O, A)
    pp, qp = ri.make_producers(O, A, q_out=P)
    torch.manual_seed(5)
    tr = ref.particle_trainer_oac.ParticleTrainer(pp, qp, n_estimators=P, action_space=ac_space,
                                                  share_layers=True, q_min=0., q_max=500., deterministic=False)
    torch.manual_seed(5)
    st = orc.ParticleState(O, A, n_estimators=P, share_layers=True, q_min=0., q_max=500.)
    hp = dict(beta_UB=4.66, delta=20.53, share_layers=True)
    rng = np.random.RandomState(1)
    for i in range(3):
        ob = rng.randn(O)
        torch.manual_seed(60 + i)
        ac_ref, _ = ref.optimistic_exploration.get_optimistic_exploration_action(
            ob, policy=tr.policy, qfs=tr.qfs, hyper_params=hp)
        torch.manual_seed(60 + i)
        torch.normal(torch.zeros(A), torch.ones(A))
        eps2 = torch.normal(torch.zeros(A), torch.ones(A))
        ac, _, _ = orc.explore(torch.from_numpy(ob).float(), st.policy, st.qfs, 4.66, 20.53,
                               share_layers=True, eps_sample=eps2)
        assert max_abs(ac, ac_ref) <= 2e-6


def test_replay_buffer_matches():
    ref = ri.load_reference()
    O, A, N = 7, 3, 50
    ob_space, ac_space = ri.make_spaces(O, A)
    for cls_ref, cls_orc in ((ref.replay_buffer.ReplayBuffer, orc.ReplayBuffer),
                             (ref.replay_buffer.ReplayBufferCount, orc.ReplayBufferCount)):
        rb = cls_ref(N, ob_space, ac_space)
        ob = cls_orc(N, O, A)
        rng = np.random.RandomState(0)
        for t in range(130):  # wraps the ring twice
            s = dict(observation=rng.randn(O), action=rng.rand(A), reward=rng.randn(),
                     next_observation=rng.randn(O), terminal=bool(rng.rand() < 0.1))
            rb.add_sample(env_info={}, **s)
            ob.add_sample(**s)
            if t % 17 == 5:
                np.random.seed(t)
                b1 = rb.random_batch(16)
                np.random.seed(t)
                b2 = ob.random_batch(16)
                assert set(b1) == set(b2)
                for k in b1:
                    assert b1[k].dtype == b2[k].dtype and np.array_equal(b1[k], b2[k]), k
        assert rb._top == ob._top and rb._size == ob._size
