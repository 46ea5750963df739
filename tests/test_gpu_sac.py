"""CUDA SACTrainer (fused sm_100a step, through the C ABI) against the oracle and the
reference's golden vectors.  fp32 path tolerance: norm-wise rel <= 1e-5 on values,
losses and gradients-driven weight updates."""
import numpy as np
import pytest
import torch

from oracle import oac_oracle as orc
from tests.util import synth_batch, synth_eps, rel_err, max_abs
from tests import golden_util as gu
from tests.gpu_util import Box, producers, load_net, net_cpu

pytestmark = pytest.mark.gpu

NETS = ['policy', 'qf1', 'qf2', 'target_qf1', 'target_qf2']


def make_trainer(O, A, H, **kw):
    from oac_explore_b200.trainer import SACTrainer
    pp, qp = producers(O, A, H)
    args = dict(policy_lr=3e-4, qf_lr=3e-4, soft_target_tau=5e-3, use_automatic_entropy_tuning=True)
    args.update(kw)
    return SACTrainer(pp, qp, action_space=Box(A), **args)


def run_pair(tr, st, O, A, B, n_steps, mode="A", seed0=10):
    outs = []
    for s in range(n_steps):
        batch = synth_batch(B, O, A, seed=seed0 + s)
        eps = synth_eps(2, B, A, seed=seed0 * 10 + s)
        out = orc.sac_step(st, batch, eps[0], eps[1], mode=mode)
        tr.inject_noise(eps[0], eps[1])
        tr._need_to_update_eval_statistics = True
        tr.train_from_torch({k: v.cuda() for k, v in batch.items()})
        outs.append(out)
    torch.cuda.synchronize()
    return outs


def test_same_seed_same_init():
    """torch.manual_seed(s) gives the reference's initial weights (same RNG consumption)."""
    O, A, H = 11, 3, 32
    torch.manual_seed(7)
    tr = make_trainer(O, A, H)
    torch.manual_seed(7)
    st = orc.SACState(O, A, hidden=(H, H))
    for n in NETS:
        ours = net_cpu(getattr(tr, n))
        for k, v in getattr(st, n).items():
            assert torch.equal(ours[k], v), (n, k)


def test_sac_small_golden():
    g = gu.load("sac_small.npz")
    O, A, B, n_steps, seed = [int(v) for v in g['meta'][:5]]
    H = int(g['meta'][5])
    tr = make_trainer(O, A, H)
    for n in NETS:
        load_net(getattr(tr, n), gu.net_from(g, 'init/' + n))
    for s in range(n_steps):
        batch = synth_batch(B, O, A, seed=10 + s)
        eps = synth_eps(2, B, A, seed=100 + s)
        tr.inject_noise(eps[0], eps[1])
        tr._need_to_update_eval_statistics = True
        tr.train_from_torch({k: v.cuda() for k, v in batch.items()})
        es, d = tr.eval_statistics, g['diag'][s]
        assert abs(es['QF1 Loss'] - d[0]) <= 1e-5 * abs(d[0])
        assert abs(es['QF2 Loss'] - d[1]) <= 1e-5 * abs(d[1])
        assert abs(es['Policy Loss'] - d[2]) <= 1e-5 * abs(d[2]) + 1e-6
        assert abs(es['Alpha'] - d[3]) <= 1e-6
        assert abs(es['Log Pis Mean'] - d[4]) <= 1e-5
        assert abs(es['Q Targets Mean'] - d[5]) <= 1e-5
    for n in NETS:
        ours = net_cpu(getattr(tr, n))
        for k, v in gu.net_from(g, 'final/' + n).items():
            # Adam's first steps move each weight by ~lr*sign(g): allow 2*lr on sign-ambiguous elements
            assert rel_err(ours[k], v) <= 1e-5 or max_abs(ours[k], v) <= 2e-6, (n, k, rel_err(ours[k], v))
    assert max_abs(tr.log_alpha.cpu(), g['final/log_alpha']) <= 1e-6


@pytest.mark.parametrize("O,A,B,H", [(376, 17, 256, 256), (1, 1, 256, 256), (11, 3, 32, 32), (23, 6, 100, 64)])
@pytest.mark.parametrize("mode", ["A", "B"])
def test_sac_vs_oracle(O, A, B, H, mode):
    if mode == "B" and O != 376:
        pytest.skip("mode B checked at Humanoid shapes only")
    torch.manual_seed(0)
    tr = make_trainer(O, A, H, stale_graph_mode=mode)
    torch.manual_seed(0)
    st = orc.SACState(O, A, hidden=(H, H))
    outs = run_pair(tr, st, O, A, B, 3, mode=mode)
    es, out = tr.eval_statistics, outs[-1]
    assert abs(es['QF1 Loss'] - float(out['qf1_loss'])) <= 1e-5 * abs(float(out['qf1_loss']))
    assert abs(es['QF2 Loss'] - float(out['qf2_loss'])) <= 1e-5 * abs(float(out['qf2_loss']))
    assert abs(es['Alpha'] - float(out['alpha'])) <= 1e-6
    e = tr._engine
    q_pred = e.io_view(e.lay.off_q_pred, (B, 2)).cpu()
    assert rel_err(q_pred[:, 0], out['q1_pred'][:, 0]) <= 1e-5
    assert rel_err(e.io_view(e.lay.off_q_target, (B, 2)).cpu()[:, 0], out['q_target'][:, 0]) <= 1e-5
    assert rel_err(e.io_view(e.lay.off_log_pi, (3 * B,)).cpu()[:B], out['log_pi'][:, 0]) <= 1e-5
    for n in NETS:
        ours = net_cpu(getattr(tr, n))
        for k, v in getattr(st, n).items():
            r, m = rel_err(ours[k], v), max_abs(ours[k], v)
            assert r <= 1e-5 or m <= 6e-6, (n, k, r, m)


def test_sac_gradients_one_step():
    """After ONE Adam step from zero moments m = (1-b1) g, so the first moment IS the gradient:
    compare it (norm-wise, <= 1e-5... a few 1e-5 for the policy, whose fp32 autograd noise in the
    reference itself is ~1e-5) with the oracle's gradients."""
    O, A, B, H = 376, 17, 256, 256
    torch.manual_seed(3)
    tr = make_trainer(O, A, H)
    torch.manual_seed(3)
    st = orc.SACState(O, A, hidden=(H, H))
    out = run_pair(tr, st, O, A, B, 1)[0]
    e = tr._engine
    for idx, gname in ((0, 'grad_policy'), (1, 'grad_qf1'), (2, 'grad_qf2')):
        m = e.net_views(idx, arena=e.adam_m)
        for k, gref in out[gname].items():
            got = m[k].cpu() / 0.1
            tol = 5e-5 if idx == 0 else 1e-5
            assert rel_err(got, gref) <= tol, (gname, k, rel_err(got, gref))


def test_sac_fp64_arbitration():
    """Both the CUDA fp32 path and the fp32 oracle must sit within fp32 noise of the fp64 oracle."""
    O, A, B, H = 376, 17, 256, 256
    torch.manual_seed(5)
    tr = make_trainer(O, A, H)
    torch.manual_seed(5)
    st32 = orc.SACState(O, A, hidden=(H, H))
    torch.manual_seed(5)
    st64 = orc.SACState(O, A, hidden=(H, H), dtype=torch.float64)
    batch = synth_batch(B, O, A, seed=42)
    eps = synth_eps(2, B, A, seed=43)
    o32 = orc.sac_step(st32, batch, eps[0], eps[1])
    o64 = orc.sac_step(st64, {k: v.double() for k, v in batch.items()}, eps[0].double(), eps[1].double())
    tr.inject_noise(eps[0], eps[1])
    tr.train_from_torch({k: v.cuda() for k, v in batch.items()})
    m = tr._engine.net_views(0, arena=tr._engine.adam_m)
    for k, g64 in o64['grad_policy'].items():
        ours = rel_err(m[k].cpu() / 0.1, g64)
        ref32 = rel_err(o32['grad_policy'][k], g64)
        assert ours <= max(3 * ref32, 2e-5), (k, ours, ref32)


def test_auto_alpha_off_and_polyak_kat():
    O, A, B, H = 11, 3, 32, 32
    torch.manual_seed(1)
    tr = make_trainer(O, A, H, use_automatic_entropy_tuning=False, soft_target_tau=1.0)
    torch.manual_seed(1)
    st = orc.SACState(O, A, hidden=(H, H), use_automatic_entropy_tuning=False, soft_target_tau=1.0)
    run_pair(tr, st, O, A, B, 2)
    # tau = 1: the targets are copies of the online critics (KAT iv)
    for a, b in (('qf1', 'target_qf1'), ('qf2', 'target_qf2')):
        x, y = net_cpu(getattr(tr, a)), net_cpu(getattr(tr, b))
        for k in x:
            assert torch.equal(x[k], y[k])
    for n in NETS:
        ours = net_cpu(getattr(tr, n))
        for k, v in getattr(st, n).items():
            assert rel_err(ours[k], v) <= 1e-5 or max_abs(ours[k], v) <= 6e-6


def test_adam_first_step_kat():
    """KAT iii: the first Adam step moves every weight with a non-zero gradient by ~lr."""
    O, A, B, H = 11, 3, 32, 32
    torch.manual_seed(2)
    tr = make_trainer(O, A, H)
    before = net_cpu(tr.qf1)['fc1.weight'].clone()
    batch = synth_batch(B, O, A, seed=1)
    tr.train_from_torch({k: v.cuda() for k, v in batch.items()})
    after = net_cpu(tr.qf1)['fc1.weight']
    m = tr._engine.net_views(1, arena=tr._engine.adam_m)['fc1.weight'].cpu()
    moved = (after - before).abs()
    nz = m.abs() > 1e-6     # |g| > 1e-5 >> adam eps: update = lr * g / (|g| + 1e-8)
    assert torch.allclose(moved[nz], torch.full_like(moved[nz], 3e-4), rtol=2e-2)
    assert torch.all(moved[m == 0] == 0)


def test_device_noise_statistics():
    """Without injected eps the step draws N(0,1) on the device (Philox): check moments."""
    O, A, B, H = 11, 3, 256, 32
    tr = make_trainer(O, A, H, rng_seed=123)
    batch = synth_batch(B, O, A, seed=1)
    vals = []
    for _ in range(8):
        tr.train_from_torch({k: v.cuda() for k, v in batch.items()})
        # save slot 3 of the policy-head save buffer is private; recover eps from z = mean + std*eps is not
        # exposed, so use log_pi finiteness + the action range as the observable
        lp = tr._engine.io_view(tr._engine.lay.off_log_pi, (3 * B,)).cpu()[:2 * B]
        assert torch.isfinite(lp).all()
        vals.append(tr._engine.x_block(1)[:, O:O + A].cpu().clone())
    acts = torch.stack(vals)
    assert acts.abs().max() <= 1.0
    assert acts.std() > 0.3 and abs(float(acts.mean())) < 0.1
    assert not torch.equal(vals[0], vals[1])


def test_snapshot_roundtrip_and_batch_resize():
    O, A, B, H = 11, 3, 32, 32
    torch.manual_seed(4)
    tr = make_trainer(O, A, H)
    torch.manual_seed(4)
    st = orc.SACState(O, A, hidden=(H, H))
    run_pair(tr, st, O, A, B, 2)
    ss = tr.get_snapshot()
    for key in ('policy_state_dict', 'policy_optim_state_dict', 'qf1_state_dict', 'qf1_optim_state_dict',
                'target_qf1_state_dict', 'qf2_state_dict', 'qf2_optim_state_dict', 'target_qf2_state_dict',
                'eval_statistics', '_n_train_steps_total', '_need_to_update_eval_statistics', 'log_alpha',
                'alpha_optim_state_dict'):
        assert key in ss
    assert list(ss['policy_state_dict'].keys()) == ['fc0.weight', 'fc0.bias', 'fc1.weight', 'fc1.bias',
                                                    'last_fc.weight', 'last_fc.bias',
                                                    'last_fc_log_std.weight', 'last_fc_log_std.bias']
    assert ss['qf1_optim_state_dict']['state'][0]['step'] == 2
    import copy
    ss = copy.deepcopy(ss)
    torch.manual_seed(99)
    tr2 = make_trainer(O, A, H)
    tr2.restore_from_snapshot(ss)
    # continue both (different batch size on purpose: the engine is rebuilt, state carried over)
    for t in (tr, tr2):
        batch = synth_batch(48, O, A, seed=77)
        eps = synth_eps(2, 48, A, seed=78)
        t.inject_noise(eps[0], eps[1])
        t.train_from_torch({k: v.cuda() for k, v in batch.items()})
    for n in NETS:
        x, y = net_cpu(getattr(tr, n)), net_cpu(getattr(tr2, n))
        for k in x:
            assert torch.equal(x[k], y[k]), (n, k)
    batch = synth_batch(48, O, A, seed=77)
    eps = synth_eps(2, 48, A, seed=78)
    orc.sac_step(st, batch, eps[0], eps[1])
    for n in NETS:
        ours = net_cpu(getattr(tr, n))
        for k, v in getattr(st, n).items():
            assert rel_err(ours[k], v) <= 1e-5 or max_abs(ours[k], v) <= 6e-6


@pytest.mark.parametrize("H", [256, 320])
def test_fused_forward_layers_match_separate_launches(monkeypatch, H):
    """gemm_fwd2_kernel runs layer 1 and layer 2 of every 32-row strip in one thread-block cluster with the arithmetic of the
    two gemm_sk_kernel launches it replaces: weights after three steps must be equal bit for bit (H = 256: clusters of 8;
    H = 320: 10 column tiles exceed the portable cluster size, so the fused stage falls back to its two launches)."""
    O, A, B = 376, 17, 256
    trainers = []
    for off in ("", "1"):
        if off:
            monkeypatch.setenv("OAC_NO_FWD2", off)
        torch.manual_seed(6)
        trainers.append(make_trainer(O, A, H))
    monkeypatch.delenv("OAC_NO_FWD2")
    if H == 256:
        assert trainers[0]._engine.launches_per_step == trainers[1]._engine.launches_per_step - 3
    for step in range(3):
        batch = synth_batch(B, O, A, seed=50 + step)
        eps = synth_eps(2, B, A, seed=500 + step)
        for tr in trainers:
            tr.inject_noise(eps[0], eps[1])
            tr.train_from_torch({k: v.cuda() for k, v in batch.items()})
    torch.cuda.synchronize()
    for n in NETS:
        a, b = net_cpu(getattr(trainers[0], n)), net_cpu(getattr(trainers[1], n))
        for k in a:
            assert torch.equal(a[k], b[k]), (n, k)


def test_step_scalars_ring_one_step_in_flight():
    """The step's host scalars go to two mapped slots chosen by step parity and stamped with the step number: the host may
    read step i's scalars (after ITS event) while step i + 1 is already in flight; values equal the blocking read."""
    O, A, H, B = 17, 6, 64, 64
    torch.manual_seed(3)
    tr = make_trainer(O, A, H)
    batches = [{k: v.cuda() for k, v in synth_batch(B, O, A, seed=40 + s).items()} for s in range(7)]
    # blocking pass on a twin trainer: the expected per-step alpha
    torch.manual_seed(3)
    tw = make_trainer(O, A, H)
    want = []
    for s, b in enumerate(batches):
        eps = synth_eps(2, B, A, seed=900 + s)
        tw.inject_noise(eps[0], eps[1])
        tw.train_from_torch(dict(b))
        torch.cuda.synchronize()
        sc = tw._engine.step_scalars()
        assert int(sc[0, 3]) == s + 1
        want.append((float(sc[0, 0]), float(sc[0, 1]), float(sc[0, 2])))
    tr._ensure_engine(B)
    eng = tr._engine
    stream = torch.cuda.current_stream()
    ev = [torch.cuda.Event(), torch.cuda.Event()]
    got = []
    for s, b in enumerate(batches):
        eps = synth_eps(2, B, A, seed=900 + s)
        tr.inject_noise(eps[0], eps[1])
        tr.train_from_torch(dict(b))
        ev[s & 1].record(stream)
        if s:
            ev[(s - 1) & 1].synchronize()
            sc = eng.step_scalars(eng.steps - 1)
            assert int(sc[0, 3]) == s
            got.append((float(sc[0, 0]), float(sc[0, 1]), float(sc[0, 2])))
    ev[(len(batches) - 1) & 1].synchronize()
    sc = eng.step_scalars()
    assert int(sc[0, 3]) == len(batches)
    got.append((float(sc[0, 0]), float(sc[0, 1]), float(sc[0, 2])))
    assert got == want
    # a restored step counter moves the host mirror with it
    eng.set_train_steps(100)
    tr.train_from_torch(dict(batches[0]))
    torch.cuda.synchronize()
    assert int(eng.step_scalars()[0, 3]) == 101 and eng.steps == 101
