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
