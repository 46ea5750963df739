"""Seeds batched in one engine (grid.z = seeds) == the same seeds trained one by one, bit for bit."""
import numpy as np
import pytest
import torch

from tests.util import synth_batch, synth_eps
from tests.gpu_util import Box, net_cpu
from tests.test_gpu_sac import make_trainer, NETS

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("gemm_path", [0, 1, 2])
def test_group_equals_singles(gemm_path):
    from oac_explore_b200.seed_group import SACSeedGroup
    O, A, B, H = 376, 17, 256, 256
    ids = [3, 11, 4]
    grp = SACSeedGroup(ids, O, A, hidden=H, batch=B, gemm_path=gemm_path)
    singles = []
    for sid in ids:
        torch.manual_seed(sid)
        singles.append(make_trainer(O, A, H, gemm_path=gemm_path))
    for step in range(2):
        for slot, sid in enumerate(ids):
            batch = synth_batch(B, O, A, seed=1000 * sid + step)
            eps = synth_eps(2, B, A, seed=77 * sid + step)
            grp.load_batch(slot, batch)
            grp.inject_noise(slot, eps[0], eps[1])
            singles[slot].inject_noise(eps[0], eps[1])
            singles[slot].train_from_torch({k: v.cuda() for k, v in batch.items()})
        grp.step(external_eps=True)
    torch.cuda.synchronize()
    for slot in range(len(ids)):
        for n in NETS:
            a, b = net_cpu(grp.nets[slot][n]), net_cpu(getattr(singles[slot], n))
            for k in a:
                assert torch.equal(a[k], b[k]), (slot, n, k)
    st = grp.stats().cpu()
    assert st.shape == (3, 32) and torch.isfinite(st).all()


def test_group_gather_shared_store():
    """One gather launch feeds every seed of the group from a shared store with per-seed index streams."""
    from oac_explore_b200.seed_group import SACSeedGroup
    from oac_explore_b200.replay_buffer import ReplayBuffer
    O, A, B, H, N = 11, 3, 32, 32, 400
    rb = ReplayBuffer(N, Box(O), Box(A))
    g = torch.Generator(device='cuda').manual_seed(0)
    rb._observations.normal_(generator=g); rb._next_obs.normal_(generator=g)
    rb._actions.uniform_(-1, 1, generator=g); rb._rewards.normal_(generator=g)
    rb._size = N
    grp = SACSeedGroup([0, 1, 2, 3], O, A, hidden=H, batch=B, gemm_path=0)
    idx = np.random.RandomState(0).randint(0, N, (4, B))
    grp.gather(rb, idx)
    torch.cuda.synchronize()
    for s in range(4):
        it = torch.from_numpy(idx[s]).cuda()
        assert torch.equal(grp.engine.x_block(2, seed=s)[:, :O], rb._observations[it])
        assert torch.equal(grp.engine.x_block(1, seed=s)[:, :O], rb._observations[it])
        assert torch.equal(grp.engine.x_block(2, seed=s)[:, O:O + A], rb._actions[it])
        assert torch.equal(grp.engine.x_block(3, seed=s)[:, :O], rb._next_obs[it])
    grp.step()
    torch.cuda.synchronize()
    assert torch.isfinite(grp.stats()).all()


@pytest.mark.parametrize("O,A", [(376, 17), (44, 6), (11, 3)])
def test_group_gather_many_rows_bit_exact(O, A):
    """>= 2048 sampled rows per launch take the warp-per-row gather kernel (16-byte pieces, rows of whole float4); other
    shapes fall back to one CTA per row.  Either way every X block / IO slot is a bit-exact copy of the store rows."""
    from oac_explore_b200.seed_group import SACSeedGroup
    from oac_explore_b200.replay_buffer import ReplayBuffer
    B, H, N, S = 256, 32, 5000, 9
    rb = ReplayBuffer(N, Box(O), Box(A))
    g = torch.Generator(device='cuda').manual_seed(1)
    rb._observations.normal_(generator=g); rb._next_obs.normal_(generator=g)
    rb._actions.uniform_(-1, 1, generator=g); rb._rewards.normal_(generator=g)
    rb._terminals.copy_((torch.rand(rb._terminals.shape, generator=g, device='cuda') < 0.1).float())
    rb._size = N
    grp = SACSeedGroup(list(range(S)), O, A, hidden=H, batch=B, gemm_path=0)
    idx = np.random.RandomState(3).randint(0, N, (S, B))
    grp.gather(rb, idx)
    torch.cuda.synchronize()
    e = grp.engine
    for s in range(S):
        it = torch.from_numpy(idx[s]).cuda()
        assert torch.equal(e.x_block(2, seed=s)[:, :O], rb._observations[it])
        assert torch.equal(e.x_block(1, seed=s)[:, :O], rb._observations[it])
        assert torch.equal(e.x_block(2, seed=s)[:, O:O + A], rb._actions[it])
        assert torch.equal(e.x_block(3, seed=s)[:, :O], rb._next_obs[it])
        assert torch.equal(e.io_view(e.lay.off_rewards, (B,), seed=s), rb._rewards[it].reshape(-1))
        assert torch.equal(e.io_view(e.lay.off_terminals, (B,), seed=s), rb._terminals[it].reshape(-1))


@pytest.mark.parametrize("share", [True, False])
def test_particle_and_gaussian_groups_equal_singles(share):
    """P-OAC and G-OAC seeds batched in one engine (ParticleSeedGroup / GaussianSeedGroup) == the same seeds trained one by
    one through ParticleTrainer / GaussianTrainer, bit for bit (fp32 path), including the per-seed statistics vectors."""
    from oac_explore_b200.seed_group import ParticleSeedGroup, GaussianSeedGroup
    from tests.test_gpu_poac_goac import make_poac, make_goac
    O, A, B, H, P = 23, 5, 64, 64, 4
    ids = [2, 9, 5]
    kw = dict(policy_lr=3e-4, qf_lr=3e-4, soft_target_tau=5e-3, q_min=0.0, q_max=500.0)
    pg = ParticleSeedGroup(ids, O, A, hidden=H, batch=B, n_estimators=P, share_layers=share, **kw)
    gg = GaussianSeedGroup(ids, O, A, hidden=H, batch=B, share_layers=share, std_lr=3e-5, **kw)
    ps, gs = [], []
    for sid in ids:
        torch.manual_seed(sid)
        ps.append(make_poac(O, A, H, P, share, False))
        torch.manual_seed(sid)
        gs.append(make_goac(O, A, H, share, False))
    for step in range(2):
        for slot, sid in enumerate(ids):
            batch = synth_batch(B, O, A, seed=100 * sid + step)
            eps = synth_eps(2, B, A, seed=7 * sid + step)
            pg.load_batch(slot, batch); pg.inject_noise(slot, eps[0], eps[1])
            gg.load_batch(slot, batch)
            ps[slot].inject_noise(eps_obs=eps[0], eps_next=eps[1])
            ps[slot].train_from_torch({k: v.cuda() for k, v in batch.items()})
            gs[slot].train_from_torch({k: v.cuda() for k, v in batch.items()})
        pg.step(external_eps=True)
        gg.step()
    torch.cuda.synchronize()
    for slot in range(len(ids)):
        a = pg.nets[slot]
        pairs = [(a['policy'], ps[slot].policy)] + list(zip(a['qfs'], ps[slot].qfs)) + list(zip(a['tfs'], ps[slot].tfs))
        b = gg.nets[slot]
        pairs += [(b['policy'], gs[slot].policy), (b['target_policy'], gs[slot].target_policy)]
        pairs += list(zip(b['qfs'], gs[slot].qfs)) + list(zip(b['tfs'], gs[slot].tfs))
        for x, y in pairs:
            sx, sy = net_cpu(x), net_cpu(y)
            for k in sx:
                assert torch.equal(sx[k], sy[k]), (slot, k)
        assert torch.equal(pg.stats()[slot].cpu(), torch.from_numpy(ps[slot]._engine.stats_host()[0]))
        assert torch.equal(gg.stats()[slot].cpu(), torch.from_numpy(gs[slot]._engine.stats_host()[0]))
    assert pg.stats().shape == (3, 11 + 9 * P) and gg.stats().shape == (3, 29)
