"""Longer runs of the fp32 path: K = 20 steps at Humanoid shapes against the oracle AND against the
reference's own 20-step record, the reference-generated 3-step digests, and BASELINE config 1 on real riverswim data."""
import numpy as np
import pytest
import torch

from oracle import oac_oracle as orc
from tests.util import synth_batch, synth_eps, rel_err, max_abs
from tests import golden_util as gu
from tests.gpu_util import Box, net_cpu
from tests.test_gpu_sac import make_trainer, NETS

pytestmark = pytest.mark.gpu

LR = 3e-4


def seeded_trainer(g, **kw):
    O, A, B, n_steps, seed = [int(v) for v in g['meta'][:5]]
    H = int(g['meta'][-1])
    torch.manual_seed(seed)
    tr = make_trainer(O, A, H, **kw)
    if not np.allclose(gu.digest(net_cpu(tr.policy)['fc0.weight']), g['init/policy/fc0.weight'], rtol=0, atol=0):
        pytest.skip("torch RNG stream differs from the one the golden was made with")
    return tr, O, A, B, n_steps, H


def check_diag(es, d, tol):
    """es: our eval_statistics; d: the reference's [QF1 Loss, QF2 Loss, Policy Loss, Alpha, Log Pis Mean, Q Targets Mean,
    Q1 Predictions Mean, Q2 Predictions Mean, Policy mu Mean, Policy log std Mean] (trainer/trainer.py:243-279)."""
    keys = ['QF1 Loss', 'QF2 Loss', 'Policy Loss', 'Alpha', 'Log Pis Mean', 'Q Targets Mean', 'Q1 Predictions Mean',
            'Q2 Predictions Mean', 'Policy mu Mean', 'Policy log std Mean']
    for k, ref in zip(keys, d):
        assert abs(float(es[k]) - ref) <= tol * abs(ref) + 2e-6, (k, float(es[k]), ref)


def test_sac_humanoid_k20_vs_reference_record_and_oracle():
    """K = 20 updates, fp32 path.  Per-step diagnostics within 2e-5 of the REFERENCE's record (the same bound the oracle
    itself meets against it), final weights: reference digests, and element-wise against the oracle within rel 1e-5 or
    the 2*lr floor of sign-ambiguous Adam elements."""
    g = gu.load("sac_humanoid_k20.npz")
    tr, O, A, B, n_steps, H = seeded_trainer(g)
    torch.manual_seed(int(g['meta'][4]))
    st = orc.SACState(O, A, hidden=(H, H))
    assert n_steps == 20
    for s in range(n_steps):
        batch = synth_batch(B, O, A, seed=10 + s)
        eps = synth_eps(2, B, A, seed=100 + s)
        out = orc.sac_step(st, batch, eps[0], eps[1])
        tr.inject_noise(eps[0], eps[1])
        tr._need_to_update_eval_statistics = True
        tr.train_from_torch({k: v.cuda() for k, v in batch.items()})
        check_diag(tr.eval_statistics, g['diag'][s], 2e-5)
        assert abs(tr.eval_statistics['QF1 Loss'] - float(out['qf1_loss'])) <= 2e-5 * abs(float(out['qf1_loss']))
    worst = 0.0
    for n in NETS:
        ours = net_cpu(getattr(tr, n))
        for k, v in getattr(st, n).items():
            r, m = rel_err(ours[k], v), max_abs(ours[k], v)
            worst = max(worst, r)
            assert r <= 1e-5 or m <= 2 * LR, (n, k, r, m)
            gu.assert_digest_close(ours[k], g['final/%s/%s' % (n, k)], 5e-6, n + '/' + k)
    assert worst <= 1e-4, worst        # norm-wise, every tensor, after 20 steps
    assert max_abs(tr.log_alpha.cpu(), g['final/log_alpha']) <= 2e-6


@pytest.mark.parametrize("name", ["sac_humanoid.npz", "sac_riverswim.npz"])
def test_sac_reference_digests_on_gpu(name):
    """The reference-generated 3-step records (seeded construction, digests of the final weights) consumed directly by
    the CUDA path, not only by the oracle."""
    g = gu.load(name)
    tr, O, A, B, n_steps, H = seeded_trainer(g)
    for s in range(n_steps):
        batch = synth_batch(B, O, A, seed=10 + s)
        eps = synth_eps(2, B, A, seed=100 + s)
        tr.inject_noise(eps[0], eps[1])
        tr._need_to_update_eval_statistics = True
        tr.train_from_torch({k: v.cuda() for k, v in batch.items()})
        check_diag(tr.eval_statistics, g['diag'][s], 1e-5)
    for n in NETS:
        ours = net_cpu(getattr(tr, n))
        for k in ours:
            gu.assert_digest_close(ours[k], g['final/%s/%s' % (n, k)], 2e-6, n + '/' + k)


def test_riverswim_real_data_through_the_drop_in_api():
    """BASELINE config 1 on REAL data: the transitions the reference collected on envs/river_swim_continuous.py with its
    own optimistic exploration go through OUR ReplayBuffer.add_sample -> random_batch (same np.random stream ->
    identical indices) -> SACTrainer.train, 20 updates; diagnostics and final weights against the reference's record."""
    from oac_explore_b200.replay_buffer import ReplayBuffer
    g = gu.load("sac_riverswim_real.npz")
    tr, O, A, B, n_steps, H = seeded_trainer(g)
    rb = ReplayBuffer(10000, Box(O, 0.0, 25.0), Box(A))
    T = g['stream/obs'].shape[0]
    half = T // 2
    for t in range(half):                                           # add_sample path ...
        rb.add_sample(g['stream/obs'][t], g['stream/act'][t], g['stream/rew'][t], g['stream/nobs'][t],
                      g['stream/term'][t], env_info={})
    rb.add_path(dict(observations=g['stream/obs'][half:], actions=g['stream/act'][half:],     # ... and the packed path
                     rewards=g['stream/rew'][half:], next_observations=g['stream/nobs'][half:],
                     terminals=g['stream/term'][half:], agent_infos=[{}] * (T - half), env_infos=[{}] * (T - half)))
    assert rb.num_steps_can_sample() == T
    for s in range(n_steps):
        np.random.seed(1000 + s)
        batch = rb.random_batch(B)
        batch['buffer'] = rb
        if s == 0:      # first call: numpy dict, bit-exact float32 of the reference's batch
            idx = g['indices'][0]
            assert np.array_equal(batch['observations'].astype(np.float32), g['stream/obs'][idx].astype(np.float32))
            assert np.array_equal(batch['rewards'].astype(np.float32), g['stream/rew'][idx].astype(np.float32))
        eps = synth_eps(2, B, A, seed=300 + s)
        tr.inject_noise(eps[0], eps[1])
        tr._need_to_update_eval_statistics = True
        tr.train(batch)
        check_diag(tr.eval_statistics, g['diag'][s], 2e-5)
    for n in NETS:
        ours = net_cpu(getattr(tr, n))
        for k in ours:
            # digests are coarse (a sum over the tensor): 2e-5 of the tensor's L2 mass after 20 steps; the element-wise
            # check on the stored full tensors below is the sharp one
            gu.assert_digest_close(ours[k], g['final/%s/%s' % (n, k)], 2e-5, n + '/' + k)
        assert max_abs(ours['last_fc.weight'], g['final_full/%s/last_fc.weight' % n]) <= 5e-6
        assert max_abs(ours['fc1.bias'], g['final_full/%s/fc1.bias' % n]) <= 2e-5      # values ~0.1: 2e-4 relative, << 2*lr
    assert max_abs(tr.log_alpha.cpu(), g['final/log_alpha']) <= 2e-6
