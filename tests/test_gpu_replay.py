"""GPU-resident replay buffer against the oracle / the reference's golden batches: bit-exact
gathered batches for the same numpy index stream."""
import numpy as np
import pytest
import torch

from oracle import oac_oracle as orc
from tests import golden_util as gu
from tests.gpu_util import Box

pytestmark = pytest.mark.gpu


def test_replay_counts_golden_bit_exact():
    from oac_explore_b200.replay_buffer import ReplayBufferCount
    g = gu.load("replay_counts.npz")
    O, A, N, T, B = [int(v) for v in g['meta']]
    rb = ReplayBufferCount(N, Box(O), Box(A))
    nb = 0
    for t in range(T):
        rb.add_sample(g['stream/obs'][t], g['stream/act'][t], g['stream/rew'][t], g['stream/nobs'][t],
                      g['stream/term'][t], env_info={})
        if nb < int(g['n_batches']) and t == int(g['batch%d/t' % nb]):
            np.random.seed(t)
            b = rb.random_batch(B)
            for k, v in b.items():
                ref = g['batch%d/%s' % (nb, k)]
                assert v.dtype == ref.dtype, k
                # the device store is float32(store); compare as the trainer sees it (utils/core.py:45)
                assert np.array_equal(v.astype(np.float32), ref.astype(np.float32)), k
            nb += 1
    assert nb == int(g['n_batches'])
    assert [rb._top, rb._size] == list(g['final/top_size'])
    assert np.array_equal(rb._counts, g['final/counts'])
    ss = rb.get_snapshot()
    for k in ('_observations', '_next_obs', '_actions', '_rewards', '_terminals', '_top', '_size', '_counts'):
        assert k in ss
    assert ss['_terminals'].dtype == np.uint8


@pytest.mark.parametrize("O,A,N,B", [(376, 17, 20000, 256), (1, 1, 1000, 256), (5, 2, 300, 7)])
def test_replay_vs_oracle_and_fast_path(O, A, N, B):
    from oac_explore_b200.replay_buffer import ReplayBuffer
    rng = np.random.RandomState(0)
    rb = ReplayBuffer(N, Box(O), Box(A))
    ob = orc.ReplayBuffer(N, O, A)
    n_add = N + N // 3           # wrap the ring
    obs = rng.randn(n_add, O); act = rng.rand(n_add, A) * 2 - 1
    rew = rng.randn(n_add); nobs = rng.randn(n_add, O); term = rng.rand(n_add) < 0.05
    for t in range(n_add):
        rb.add_sample(obs[t], act[t], rew[t], nobs[t], term[t], env_info={})
        ob.add_sample(obs[t], act[t], rew[t], nobs[t], term[t])
    for s in range(3):
        np.random.seed(s)
        b1 = rb.random_batch(B)
        np.random.seed(s)
        b2 = ob.random_batch(B)
        for k in b2:
            assert b1[k].dtype == b2[k].dtype
            assert np.array_equal(b1[k].astype(np.float32), b2[k].astype(np.float32)), k
    # dense device gather == float32 cast of the oracle's batch, bit for bit
    np.random.seed(11)
    idx = np.random.randint(0, rb._size, B)
    d = rb.gather_dense(rb._upload_indices(idx), B)
    ref = orc.np_to_torch_batch(ob.gather(idx))
    for k in ref:
        assert torch.equal(d[k].cpu(), ref[k]), k
    assert rb.num_steps_can_sample() == ob.num_steps_can_sample() == N
    assert np.array_equal(rb.get_dataset().astype(np.float32), ob._observations[:N].astype(np.float32))


def test_replay_resident_gather_feeds_trainer():
    """random_batch gathers straight into an attached trainer's batch rows (no host round trip)."""
    from oac_explore_b200.replay_buffer import ReplayBuffer
    from tests.test_gpu_sac import make_trainer
    O, A, N, B, H = 11, 3, 500, 32, 32
    rng = np.random.RandomState(1)
    rb = ReplayBuffer(N, Box(O), Box(A))
    ob = orc.ReplayBuffer(N, O, A)
    for t in range(N):
        s = (rng.randn(O), rng.rand(A), rng.randn(), rng.randn(O), rng.rand() < 0.1)
        rb.add_sample(*s, env_info={})
        ob.add_sample(*s)
    tr = make_trainer(O, A, H)
    rb.attach(tr)
    np.random.seed(5)
    b = rb.random_batch(B)
    np.random.seed(5)
    ref = orc.np_to_torch_batch(ob.random_batch(B))
    assert b['_oac_resident'] is tr._engine
    for k in ref:
        assert torch.equal(b[k].cpu(), ref[k]), k
    e = tr._engine
    assert torch.equal(e.x_block(1)[:, :O].cpu(), ref['observations'])
    assert torch.equal(e.x_block(3)[:, :O].cpu(), ref['next_observations'])
    b['buffer'] = rb
    tr.train(b)          # consumes the resident batch
    torch.cuda.synchronize()
    assert tr._n_train_steps_total == 1


def test_full_size_roundtrip_properties():
    """BASELINE size (1M x Humanoid): gather is a permutation-invariant copy -- checksum of the
    gathered rows equals the checksum of the same rows read directly."""
    from oac_explore_b200.replay_buffer import ReplayBuffer
    O, A, N, B = 376, 17, 1000000, 256
    rb = ReplayBuffer(N, Box(O), Box(A))
    g = torch.Generator(device='cuda').manual_seed(0)
    rb._observations.normal_(generator=g); rb._next_obs.normal_(generator=g)
    rb._actions.uniform_(-1, 1, generator=g); rb._rewards.normal_(generator=g)
    rb._size, rb._top = N, 0
    np.random.seed(0)
    idx = np.random.randint(0, N, B)
    d = rb.gather_dense(rb._upload_indices(idx), B)
    it = torch.from_numpy(idx).cuda()
    assert torch.equal(d['observations'], rb._observations[it])
    assert torch.equal(d['next_observations'], rb._next_obs[it])
    assert torch.equal(d['actions'], rb._actions[it])
    assert torch.equal(d['rewards'], rb._rewards[it])


def test_pinned_index_ring_without_host_synchronisation():
    """The gather kernel reads its indices from a ring of pinned host slots while the host runs ahead: 2 600 batches
    (more than two turns of the 1 024-slot ring) are issued without any synchronisation in between and every one of
    them must still be the batch of ITS index draw."""
    from oac_explore_b200.replay_buffer import ReplayBuffer
    from tests.test_gpu_sac import make_trainer
    O, A, N, B, H = 7, 2, 300, 16, 32
    rb = ReplayBuffer(N, Box(O), Box(A))
    g = torch.Generator(device='cuda').manual_seed(3)
    rb._observations.normal_(generator=g)
    rb._size = N
    tr = make_trainer(O, A, H)
    tr._ensure_engine(B)
    rb.attach(tr)
    n_batches = 2600
    keep = torch.empty((n_batches, B, O), device='cuda')
    np.random.seed(11)
    for i in range(n_batches):
        rb.random_batch(B)
        keep[i].copy_(tr._engine.x_block(2)[:, :O])          # stream-ordered after this batch's gather
    torch.cuda.synchronize()
    np.random.seed(11)
    store = rb._observations.cpu()
    for i in range(n_batches):
        idx = torch.from_numpy(np.random.randint(0, N, B))
        assert torch.equal(keep[i].cpu(), store[idx]), i


# ---- round 2 ----
def _path(rng, T, O, A):
    return dict(observations=rng.randn(T, O), actions=rng.rand(T, A) * 2 - 1, rewards=rng.randn(T, 1),
                next_observations=rng.randn(T, O), terminals=(rng.rand(T, 1) < 0.1),
                agent_infos=[{}] * T, env_infos=[{}] * T)


@pytest.mark.parametrize("N,lens", [(100, [30, 50, 45, 7]), (9000, [4096, 4096, 300, 5000])])
def test_add_path_packed_equals_add_sample_loop(N, lens):
    """add_path packs a whole path with vectorised stores + one scatter launch (replay_buffer.py:57-82 loops per sample):
    same store, _top, _size and counts as the oracle's per-sample loop, across ring wraps, staging-buffer switches and
    paths longer than the staging buffer."""
    from oac_explore_b200.replay_buffer import ReplayBufferCount
    O, A = 6, 2
    rng = np.random.RandomState(3)
    rb = ReplayBufferCount(N, Box(O), Box(A))
    ob = orc.ReplayBufferCount(N, O, A)
    for i, T in enumerate(lens):
        p = _path(rng, T, O, A)
        rb.add_paths([p])
        for t in range(T):
            ob.add_sample(p['observations'][t], p['actions'][t], p['rewards'][t], p['next_observations'][t], p['terminals'][t])
        if i == 1:          # sample in between so some counts are non-zero before the next overwrite zeroes them
            np.random.seed(1); rb.random_batch(16)
            np.random.seed(1); ob.random_batch(16)
    assert (rb._top, rb._size) == (ob._top, ob._size)
    ss = rb.get_snapshot()
    n = ob._size
    assert np.array_equal(ss['_observations'][:n], ob._observations[:n].astype(np.float32))
    assert np.array_equal(ss['_next_obs'][:n], ob._next_obs[:n].astype(np.float32))
    assert np.array_equal(ss['_actions'][:n], ob._actions[:n].astype(np.float32))
    assert np.array_equal(ss['_rewards'][:n], ob._rewards[:n].astype(np.float32))
    assert np.array_equal(ss['_terminals'][:n], ob._terminals[:n])
    assert np.array_equal(rb._counts[:n], ob._counts[:n])


@pytest.mark.parametrize("B", [256, 2048, 7000])
def test_counts_bump_once_per_distinct_index(B):
    """numpy fancy ``counts[idx] += 1`` increments a duplicated index ONCE (replay_buffer.py:195): heavy duplication
    (B draws from 50 slots), the shared-memory staging kernel (B <= 6144) and the plain-scan fallback (B = 7000)."""
    from oac_explore_b200.replay_buffer import ReplayBufferCount
    O, A, N = 3, 1, 50
    rb = ReplayBufferCount(N, Box(O), Box(A))
    ob = orc.ReplayBufferCount(N, O, A)
    rng = np.random.RandomState(0)
    for t in range(N):
        s = (rng.randn(O), rng.rand(A), rng.randn(), rng.randn(O), False)
        rb.add_sample(*s, env_info={}); ob.add_sample(*s)
    for rep in range(3):
        np.random.seed(rep); b1 = rb.random_batch(B)
        np.random.seed(rep); b2 = ob.random_batch(B)
        assert np.array_equal(b1['counts'], b2['counts'])
    assert np.array_equal(rb._counts, ob._counts) and ob._counts.max() == 3


def test_priority_sample_matches_reference_draw():
    """replay_buffer.py:181-185: p ~ 1 / (count + 1), np.random.choice on the global stream."""
    from oac_explore_b200.replay_buffer import ReplayBufferCount
    O, A, N, B = 3, 1, 40, 8
    rb = ReplayBufferCount(N, Box(O), Box(A), priority_sample=True)
    ob = orc.ReplayBufferCount(N, O, A)
    rng = np.random.RandomState(0)
    for t in range(N):
        s = (rng.randn(O), rng.rand(A), rng.randn(), rng.randn(O), False)
        rb.add_sample(*s, env_info={}); ob.add_sample(*s)
    for rep in range(4):
        np.random.seed(rep)
        b1 = rb.random_batch(B)
        np.random.seed(rep)
        probs = 1 / (ob._counts[:ob._size] + 1)
        probs /= probs.sum()
        idx = np.random.choice(np.arange(ob._size), size=B, p=probs[:, 0])
        b2 = ob.gather(idx)
        for k in b2:
            assert np.array_equal(b1[k].astype(np.float32), b2[k].astype(np.float32)), (rep, k)
    assert np.array_equal(rb._counts, ob._counts)


def test_random_batch_return_copies_flag():
    """With a trainer attached random_batch returns views of the trainer's batch rows (overwritten by the next call);
    ``return_copies`` gives private tensors for callers that keep batches (rl_algorithm.py:163-165)."""
    from oac_explore_b200.replay_buffer import ReplayBuffer
    from tests.test_gpu_sac import make_trainer
    O, A, N, B, H = 5, 2, 200, 16, 32
    rb = ReplayBuffer(N, Box(O), Box(A))
    g = torch.Generator(device='cuda').manual_seed(1)
    rb._observations.normal_(generator=g)
    rb._size = N
    tr = make_trainer(O, A, H)
    rb.attach(tr)
    np.random.seed(0); v1 = rb.random_batch(B)['observations']
    keep = v1.clone()
    np.random.seed(1); rb.random_batch(B)
    assert not torch.equal(v1, keep)                 # a view: the second batch replaced it
    rb.return_copies = True
    np.random.seed(0); c1 = rb.random_batch(B)
    assert '_oac_resident' not in c1
    np.random.seed(1); rb.random_batch(B)
    assert torch.equal(c1['observations'], keep)     # a private copy: unchanged
    c1['buffer'] = rb
    tr.train(c1)                                     # goes through the normal upload path
    assert tr._n_train_steps_total == 1
