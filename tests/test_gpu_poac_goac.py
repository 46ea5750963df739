"""CUDA ParticleTrainer (P-OAC) and GaussianTrainer (G-OAC) against the reference's golden
weights (small shapes) and the oracle (Humanoid shapes)."""
import numpy as np
import pytest
import torch

from oracle import oac_oracle as orc
from tests.util import synth_batch, synth_eps, rel_err, max_abs
from tests import golden_util as gu
from tests.gpu_util import Box, producers, load_net, net_cpu

pytestmark = pytest.mark.gpu

KW = dict(policy_lr=3e-4, qf_lr=3e-4, soft_target_tau=5e-3, delta=0.95, q_min=0.0, q_max=500.0)


def close(ours, ref, what, rtol=1e-5, atol=6e-6):
    r, m = rel_err(ours, ref), max_abs(ours, ref)
    assert r <= rtol or m <= atol, (what, r, m)


def make_poac(O, A, H, P, share, counts):
    from oac_explore_b200.particle_trainer_oac import ParticleTrainer
    pp, qp = producers(O, A, H, q_out=P if share else 1)
    return ParticleTrainer(pp, qp, n_estimators=P, action_space=Box(A), share_layers=share, counts=counts,
                           deterministic=False, use_automatic_entropy_tuning=True, **KW)


def make_goac(O, A, H, share, counts):
    from oac_explore_b200.gaussian_trainer import GaussianTrainer
    pp, qp = producers(O, A, H, q_out=2 if share else 1)
    return GaussianTrainer(pp, qp, n_estimators=2, action_space=Box(A), share_layers=share, counts=counts,
                           std_lr=3e-5, **KW)


def poac_step_pair(tr, st, batch, eps):
    # reference noise order: eps[0] -> next_obs draw, eps[1] -> obs draw
    orc.poac_step(st, batch, eps[0], eps[1])
    tr.inject_noise(eps_obs=eps[1], eps_next=eps[0])
    tr.train_from_torch({k: v.cuda() for k, v in batch.items()})


@pytest.mark.parametrize("name", ["poac_shared_small.npz", "poac_shared_counts_small.npz",
                                  "poac_separate_small.npz"])
def test_poac_golden(name):
    g = gu.load(name)
    O, A, B, n_steps, seed, P, share, counts = [int(v) for v in g['meta'][:8]]
    H = int(g['meta'][8])
    tr = make_poac(O, A, H, P, bool(share), bool(counts))
    assert tr.delta_index == P - 1
    load_net(tr.policy, gu.net_from(g, 'init/policy'))
    for i in range(len(tr.qfs)):
        load_net(tr.qfs[i], gu.net_from(g, 'init/qf%d' % i))
        load_net(tr.tfs[i], gu.net_from(g, 'init/tf%d' % i))
    for s in range(n_steps):
        batch = synth_batch(B, O, A, seed=20 + s, counts=bool(counts))
        eps = synth_eps(2, B, A, seed=200 + s)
        tr.inject_noise(eps_obs=eps[1], eps_next=eps[0])
        tr.train_from_torch({k: v.cuda() for k, v in batch.items()})
    ours = net_cpu(tr.policy)
    for k, v in gu.net_from(g, 'final/policy').items():
        close(ours[k], v, ('policy', k))
    for i in range(len(tr.qfs)):
        oq, ot = net_cpu(tr.qfs[i]), net_cpu(tr.tfs[i])
        for k, v in gu.net_from(g, 'final/qf%d' % i).items():
            close(oq[k], v, ('qf', i, k), atol=1e-4)
        for k, v in gu.net_from(g, 'final/tf%d' % i).items():
            close(ot[k], v, ('tf', i, k), atol=1e-4)
    assert max_abs(tr.log_alpha.cpu(), g['final/log_alpha']) <= 1e-6
    assert 'QF0 Loss' in tr.eval_statistics and 'Policy Loss' in tr.eval_statistics


@pytest.mark.parametrize("share,counts", [(True, False), (True, True), (False, False)])
def test_poac_humanoid_vs_oracle(share, counts):
    O, A, B, H, P = 376, 17, 256, 256, 10
    torch.manual_seed(1)
    tr = make_poac(O, A, H, P, share, counts)
    torch.manual_seed(1)
    st = orc.ParticleState(O, A, n_estimators=P, share_layers=share, counts=counts, q_min=0., q_max=500.)
    for s in range(2):
        batch = synth_batch(B, O, A, seed=20 + s, counts=counts)
        poac_step_pair(tr, st, batch, synth_eps(2, B, A, seed=200 + s))
    for k, v in st.policy.items():
        close(net_cpu(tr.policy)[k], v, ('policy', k))
    for i in range(len(st.qfs)):
        oq, ot = net_cpu(tr.qfs[i]), net_cpu(tr.tfs[i])
        for k, v in st.qfs[i].items():
            close(oq[k], v, ('qf', i, k), atol=1e-4)
        for k, v in st.tfs[i].items():
            close(ot[k], v, ('tf', i, k), atol=1e-4)
    # sortedness property of the logged particles (domain invariant)
    e = tr._engine
    sq = e.io_view(e.lay.off_q_pred, (B, P)).cpu()
    assert torch.all(sq[:, 1:] >= sq[:, :-1])


@pytest.mark.parametrize("name", ["goac_shared_small.npz", "goac_shared_counts_small.npz",
                                  "goac_separate_small.npz"])
def test_goac_golden(name):
    g = gu.load(name)
    O, A, B, n_steps, seed, share, counts = [int(v) for v in g['meta'][:7]]
    H = int(g['meta'][7])
    tr = make_goac(O, A, H, bool(share), bool(counts))
    names = ['policy', 'target_policy', 'q', 'q_target'] + ([] if share else ['std', 'std_target'])
    for n in names:
        load_net(getattr(tr, n), gu.net_from(g, 'init/' + n))
    for s in range(n_steps):
        batch = synth_batch(B, O, A, seed=30 + s, counts=bool(counts))
        tr.train_from_torch({k: v.cuda() for k, v in batch.items()})
    for n in names:
        ours = net_cpu(getattr(tr, n))
        for k, v in gu.net_from(g, 'final/' + n).items():
            close(ours[k], v, (n, k), atol=6e-6 if 'policy' in n else 1e-4)
    for key in ('QF Loss', 'STD Loss', 'Policy Loss', 'Q STD Target Mean'):
        assert key in tr.eval_statistics


@pytest.mark.parametrize("share,counts", [(True, False), (True, True), (False, False)])
def test_goac_humanoid_vs_oracle(share, counts):
    O, A, B, H = 376, 17, 256, 256
    torch.manual_seed(2)
    tr = make_goac(O, A, H, share, counts)
    torch.manual_seed(2)
    st = orc.GaussianState(O, A, share_layers=share, counts=counts, q_min=0., q_max=500.)
    for s in range(2):
        batch = synth_batch(B, O, A, seed=30 + s, counts=counts)
        orc.goac_step(st, batch)
        tr.train_from_torch({k: v.cuda() for k, v in batch.items()})
    names = ['policy', 'target_policy', 'q', 'q_target'] + ([] if share else ['std', 'std_target'])
    for n in names:
        ours = net_cpu(getattr(tr, n))
        for k, v in getattr(st, n).items():
            close(ours[k], v, (n, k), atol=6e-6 if 'policy' in n else 1e-4)


@pytest.mark.parametrize("share", [True, False])
def test_poac_std_soft_update_vs_oracle(share):
    """std_soft_update (trainer/particle_trainer_oac.py:210-219): the oracle's restatement is pinned on the
    live reference (tests/test_oracle_vs_reference.py::test_poac_std_soft_update); here the CUDA step against it."""
    from oac_explore_b200.particle_trainer_oac import ParticleTrainer
    O, A, B, H, P = 376, 17, 256, 256, 10
    pp, qp = producers(O, A, H, q_out=P if share else 1)
    torch.manual_seed(6)
    tr = ParticleTrainer(pp, qp, n_estimators=P, action_space=Box(A), share_layers=share, deterministic=False,
                         use_automatic_entropy_tuning=True, std_soft_update=True, std_soft_update_prob=0.3, **KW)
    torch.manual_seed(6)
    st = orc.ParticleState(O, A, n_estimators=P, share_layers=share, q_min=0., q_max=500., policy_lr=3e-4, qf_lr=3e-4,
                           soft_target_tau=5e-3, std_soft_update=True, std_soft_update_prob=0.3)
    for s in range(3):
        batch = synth_batch(B, O, A, seed=40 + s)
        eps = synth_eps(2, B, A, seed=400 + s)
        o = None
        o = orc.poac_step(st, batch, eps[0], eps[1])
        tr.inject_noise(eps_obs=eps[1], eps_next=eps[0])
        tr.train_from_torch({k: v.cuda() for k, v in batch.items()})
        e = tr._engine
        assert rel_err(e.io_view(e.lay.off_q_target, (B, P)).cpu().t(), o['q_target'][:, :, 0]) <= 1e-5
    for i in range(len(st.qfs)):
        ours = net_cpu(tr.qfs[i])
        for k, v in st.qfs[i].items():
            close(ours[k], v, ('qf', i, k), atol=6e-5)          # head biases ~500: 1e-7 relative
    ours = net_cpu(tr.policy)
    for k, v in st.policy.items():
        close(ours[k], v, ('policy', k))
