"""On-device diagnostics (oac_trainer_stats), snapshot round trips of all three trainers and a BatchRLAlgorithm-shaped
epoch (rl_algorithm.py:141-239) through the drop-in classes."""
import pickle
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import oac_oracle as orc
from tests.util import synth_batch, synth_eps, rel_err, max_abs
from tests.gpu_util import Box, producers, net_cpu
from tests.test_gpu_sac import make_trainer, NETS
from tests.test_gpu_poac_goac import make_poac, make_goac

pytestmark = pytest.mark.gpu


def quad(name, x):
    """utils/eval_util.py:69-113 create_stats_ordered_dict on an ndarray."""
    return OrderedDict([(name + ' Mean', np.mean(x)), (name + ' Std', np.std(x)), (name + ' Max', np.max(x)),
                        (name + ' Min', np.min(x))])


def io_np(tr, name, shape):
    return tr._io(name, shape).cpu().numpy()


def close(d, ref):
    assert list(d.keys()) == list(ref.keys()), (list(d.keys()), list(ref.keys()))
    for k in ref:
        assert abs(float(d[k]) - float(ref[k])) <= 1e-5 * abs(float(ref[k])) + 1e-6, (k, d[k], ref[k])


def test_sac_stats_vector_matches_reference_definitions():
    """Keys, ORDER and values of trainer/trainer.py:230-279, computed by one kernel, against numpy on the same tensors."""
    O, A, B, H = 376, 17, 256, 256
    torch.manual_seed(0)
    tr = make_trainer(O, A, H)
    batch = synth_batch(B, O, A, seed=3)
    tr.train_from_torch({k: v.cuda() for k, v in batch.items()})
    e = tr._engine
    qp, qt, qn = io_np(tr, 'off_q_pred', (B, 2)), io_np(tr, 'off_q_target', (B, 2))[:, :1], io_np(tr, 'off_q_new', (B, 2))
    lp = io_np(tr, 'off_log_pi', (3 * B,))[:B, None]
    mean, log_std = io_np(tr, 'off_mean', (3 * B, A))[:B], io_np(tr, 'off_log_std', (3 * B, A))[:B]
    q1, q2 = qp[:, :1], qp[:, 1:]
    ref = OrderedDict()
    stack = np.stack([q1, q2], axis=0)
    ref['QF mean'], ref['QF std'] = np.mean(stack, axis=0).mean(), np.std(stack, axis=0).mean()
    ref['QF1 Loss'], ref['QF2 Loss'] = np.mean((q1 - qt) ** 2), np.mean((q2 - qt) ** 2)
    ref['Q Loss'] = ref['QF1 Loss'] + ref['QF2 Loss']
    ref['Policy Loss'] = np.mean(lp - np.minimum(qn[:, :1], qn[:, 1:]))
    for n, x in (('Q1 Predictions', q1), ('Q2 Predictions', q2), ('Q Targets', qt), ('Log Pis', lp), ('Policy mu', mean),
                 ('Policy log std', log_std)):
        ref.update(quad(n, x))
    sc = e.scalars().cpu().numpy()
    ref['Alpha'], ref['Alpha Loss'] = sc[0], sc[1]
    close(tr.eval_statistics, ref)
    assert e.n_stats == 32 and len(tr.STAT_KEYS) == 32


def test_poac_goac_stats_vectors():
    O, A, B, H, P = 24, 4, 64, 32, 5
    torch.manual_seed(1)
    tr = make_poac(O, A, H, P, True, False)
    batch = synth_batch(B, O, A, seed=4)
    tr.train_from_torch({k: v.cuda() for k, v in batch.items()})
    qs = io_np(tr, 'off_q_pred', (B, P)).T[:, :, None]
    tg = io_np(tr, 'off_q_target', (B, P)).T[:, :, None]
    qn = io_np(tr, 'off_q_new', (B, P))
    lp = io_np(tr, 'off_log_pi', (3 * B,))[:B, None]
    mean, log_std = io_np(tr, 'off_mean', (3 * B, A))[:B], io_np(tr, 'off_log_std', (3 * B, A))[:B]
    alpha = float(tr._engine.scalars().cpu()[0])
    ref = OrderedDict()
    ref['QF mean'], ref['QF std'] = np.mean(qs, axis=0).mean(), np.std(qs, axis=0).mean()
    for i in range(P):
        ref['QF%d Loss' % i] = np.mean((qs[i] - tg[i]) ** 2)
        ref.update(quad('Q%dPredictions' % i, qs[i]))
        ref.update(quad('Q%dTargets' % i, tg[i]))
    ref['Policy Loss'] = np.mean(alpha * lp - qn.min(axis=1, keepdims=True))
    ref.update(quad('Policy mu', mean))
    ref.update(quad('Policy log std', log_std))
    close(tr.eval_statistics, ref)

    torch.manual_seed(2)
    tg_ = make_goac(O, A, H, True, False)
    tg_.train_from_torch({k: v.cuda() for k, v in batch.items()})
    pred, tgt, qn = io_np(tg_, 'off_q_pred', (B, 2)), io_np(tg_, 'off_q_target', (B, 2)), io_np(tg_, 'off_q_new', (B, 2))
    mean, log_std = io_np(tg_, 'off_mean', (3 * B, A))[2 * B:], io_np(tg_, 'off_log_std', (3 * B, A))[2 * B:]
    q, s, qt, s_t = pred[:, :1], pred[:, 1:], tgt[:, :1], tgt[:, 1:]
    ref = OrderedDict()
    ref['QF mean'], ref['QF std'], ref['QF Loss'] = np.mean(q), np.mean(s), np.mean((q - qt) ** 2)
    ref.update(quad('Q Predictions', q)); ref.update(quad('Q Target', qt))
    ref['STD Loss'] = np.mean((s - s_t) ** 2)
    ref.update(quad('Q STD Predictions', s)); ref.update(quad('Q STD Target', s_t))
    ref['Policy Loss'] = np.mean(qn[:, :1] + tg_.standard_bound * qn[:, 1:])
    ref.update(quad('Policy mu', mean)); ref.update(quad('Policy log std', log_std))
    close(tg_.eval_statistics, ref)


def test_group_stats_is_the_full_vector_per_seed():
    from oac_explore_b200.seed_group import SACSeedGroup, STAT_NAMES, N_STATS
    O, A, B, H = 376, 17, 256, 256
    ids = [4, 1, 6]
    grp = SACSeedGroup(ids, O, A, hidden=H, batch=B, gemm_path=0)
    singles = []
    for slot, sid in enumerate(ids):
        torch.manual_seed(sid)
        singles.append(make_trainer(O, A, H))
        batch = synth_batch(B, O, A, seed=50 + sid)
        eps = synth_eps(2, B, A, seed=60 + sid)
        grp.load_batch(slot, batch)
        grp.inject_noise(slot, eps[0], eps[1])
        singles[-1].inject_noise(eps[0], eps[1])
        singles[-1].train_from_torch({k: v.cuda() for k, v in batch.items()})
    grp.step(external_eps=True)
    st = grp.stats().cpu().numpy()
    assert st.shape == (3, N_STATS) and N_STATS == 32 and STAT_NAMES[2] == 'QF1 Loss' and STAT_NAMES[30] == 'Alpha'
    for slot in range(3):
        es = singles[slot].eval_statistics
        for i, k in enumerate(STAT_NAMES):
            assert st[slot, i] == np.float32(es[k]), (slot, k)       # same kernel, same inputs: bit-identical


@pytest.mark.parametrize("kind", ["poac_shared", "poac_separate", "goac_shared", "goac_separate"])
def test_particle_and_gaussian_snapshot_roundtrip(kind):
    """get_snapshot -> pickle -> restore_from_snapshot into a fresh trainer (main.py:286-300 --load_from): the restored
    trainer continues bit-identically (weights, Adam moments and step counts, targets, log_alpha)."""
    O, A, B, H, P = 11, 3, 32, 32, 4
    share = kind.endswith("shared")
    mk = (lambda: make_poac(O, A, H, P, share, False)) if kind.startswith("poac") else (lambda: make_goac(O, A, H, share, False))
    torch.manual_seed(3)
    a = mk()
    for s in range(2):
        batch = synth_batch(B, O, A, seed=70 + s)
        eps = synth_eps(2, B, A, seed=80 + s)
        if kind.startswith("poac"):
            a.inject_noise(eps_obs=eps[1], eps_next=eps[0])
        a.train_from_torch({k: v.cuda() for k, v in batch.items()})
    snap = pickle.loads(pickle.dumps(a.get_snapshot()))
    expect = ({'policy_state_dict', 'policy_optim_state_dict', 'qfs_state_dicts', 'qfs_optims_state_dicts',
               'target_qfs_state_dicts', 'eval_statistics', '_n_train_steps_total', '_need_to_update_eval_statistics'})
    assert expect <= set(snap.keys())
    torch.manual_seed(12345)                       # different init: everything must come from the snapshot
    b = mk()
    b.restore_from_snapshot(snap)
    assert b._n_train_steps_total == 2
    for s in range(2):
        batch = synth_batch(B, O, A, seed=90 + s)
        eps = synth_eps(2, B, A, seed=95 + s)
        for t in (a, b):
            if kind.startswith("poac"):
                t.inject_noise(eps_obs=eps[1], eps_next=eps[0])
            t.train_from_torch({k: v.cuda() for k, v in batch.items()})
    for na, nb in zip(a.networks, b.networks):
        sa, sb = net_cpu(na), net_cpu(nb)
        for k in sa:
            assert torch.equal(sa[k], sb[k]), (kind, k)
    assert torch.equal(a._engine.adam_m.cpu(), b._engine.adam_m.cpu())
    assert torch.equal(a._engine.adam_v.cpu(), b._engine.adam_v.cpu())
    assert torch.equal(a.log_alpha.cpu(), b.log_alpha.cpu())


def test_batch_rl_algorithm_shaped_epoch():
    """The calls BatchRLAlgorithm makes in one epoch (rl_algorithm.py:141-239): add_paths -> N x (random_batch ->
    train_data['buffer'] = buffer -> trainer.train) -> end_epoch on every object -> get_diagnostics -> get_snapshot
    (pickled) -> restore, with exploration through get_optimistic_exploration_action in between."""
    from oac_explore_b200.replay_buffer import ReplayBuffer
    from oac_explore_b200.optimistic_exploration import get_optimistic_exploration_action
    O, A, H, B = 11, 3, 32, 32
    torch.manual_seed(0)
    np.random.seed(0)
    tr = make_trainer(O, A, H)
    rb = ReplayBuffer(500, Box(O), Box(A))
    hp = dict(should_use=True, beta_UB=4.66, delta=23.53, share_layers=False)
    rng = np.random.RandomState(0)

    def collect(n_paths, T):
        paths = []
        for _ in range(n_paths):
            o = rng.randn(O)
            p = dict(observations=[], actions=[], rewards=[], next_observations=[], terminals=[], agent_infos=[], env_infos=[])
            for t in range(T):
                a, info = get_optimistic_exploration_action(o, policy=tr.policy, qfs=tr.qfs, trainer=None, hyper_params=hp)
                no = o + 0.1 * rng.randn(O)
                p['observations'].append(o); p['actions'].append(a); p['rewards'].append(np.array([rng.randn()]))
                p['next_observations'].append(no); p['terminals'].append(np.array([t == T - 1]))
                p['agent_infos'].append(info); p['env_infos'].append({})
                o = no
            paths.append({k: (np.array(v) if k not in ('agent_infos', 'env_infos') else v) for k, v in p.items()})
        return paths

    for epoch in range(2):
        rb.add_paths(collect(3, 20))
        for _ in range(5):
            train_data = rb.random_batch(B)
            train_data['buffer'] = rb
            tr.train(train_data)
        diag = OrderedDict()
        diag.update(rb.get_diagnostics()); diag.update(tr.get_diagnostics())
        assert diag['size'] == 60 * (epoch + 1) and 'QF1 Loss' in diag and 'Policy log std Min' in diag
        assert np.isfinite([float(v) for v in diag.values()]).all()
        snap = dict(trainer=tr.get_snapshot(), replay_buffer=rb.get_snapshot(), epoch=epoch)
        blob = pickle.dumps(snap)
        tr.end_epoch(epoch); rb.end_epoch(epoch)
        assert tr._need_to_update_eval_statistics
    # resume (main.py:286-300)
    snap = pickle.loads(blob)
    torch.manual_seed(777)
    tr2 = make_trainer(O, A, H)
    tr2.restore_from_snapshot(snap['trainer'])
    rb2 = ReplayBuffer(500, Box(O), Box(A))
    rb2.restore_from_snapshot(snap['replay_buffer'])
    assert rb2.num_steps_can_sample() == 120 and tr2._n_train_steps_total == 10
    for n in NETS:
        x, y = net_cpu(getattr(tr, n)), net_cpu(getattr(tr2, n))
        for k in x:
            assert torch.equal(x[k], y[k])
    np.random.seed(5); b1 = rb.random_batch(B)
    np.random.seed(5); b2 = rb2.random_batch(B)
    assert torch.equal(torch.as_tensor(b1['observations']).cpu().float(), torch.as_tensor(b2['observations']).cpu().float())
