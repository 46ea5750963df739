"""The tcgen05 kind::tf32 path against the TF32 arithmetic model of the oracle (``orc.tf32_mode``: both operands of every
forward, input-gradient and weight-gradient product rounded to tf32, fp32 accumulate), at the tolerance BASELINE.json's
north_star states for TF32: norm-wise rel <= 1e-3 on Q-values, losses, GRADIENTS and UPDATED WEIGHTS.

Round 1 compared the TF32 path with the fp32 oracle only and had to allow 3e-2 on gradients (a 3e-4 relative error on a
pre-activation flips the ReLU mask of the units that sit that close to zero: dozens per layer).  Against the model the
operand rounding is the same on both sides, and what is left is the tensor core's accumulation (the tcgen05 accumulator
truncates): pre-activations agree to ~4e-5 instead of 3e-4, which still flips the mask of the zero to two (sample, unit)
pairs per layer that sit within 4e-5 of zero.  One flipped unit of layer 2 changes dh1 of ITS SAMPLE in every unit, i.e.
every row of dW0 (measured: 2.4e-3 .. 7e-3 norm-wise, profiles/r02_tf32_flip_evidence.txt), with nothing wrong.  So the
gradient criterion is stated per sample (tests/util.py::per_sample_flip_split / rows_off):
  * per net, the first layer's gradient is split per sample (dW0 pinv(X) = dh1^T): at most MAX_FLIPS samples (of 256) may
    differ by more than 1e-3, and with those samples taken out of both sides it agrees norm-wise to <= 1e-3 (measured
    <= 3e-4);
  * a net WITHOUT such a sample (about a quarter of the nets; the 64-seed test requires it of at least a sixth) must agree norm-wise to
    <= 1e-3 on EVERY tensor; a net with some is bounded by what that many flips can do (3e-2);
  * values, losses and updated weights: norm-wise <= 1e-3 against the model AND against the fp32 oracle.
``test_tf32_model_differs_from_fp32_like_the_kernel`` shows the model reproduces the >3e-3 gradient gap to fp32 that the
round-1 tests tolerated, i.e. that gap IS tf32 rounding and not a kernel defect."""
import pytest
import torch

from oracle import oac_oracle as orc
from tests.util import synth_batch, synth_eps, rel_err, max_abs, per_sample_flip_split, rows_off
from tests.gpu_util import net_cpu
from tests.test_gpu_sac import make_trainer, NETS
from tests.test_gpu_poac_goac import make_poac, make_goac

pytestmark = pytest.mark.gpu

TOL = 1e-3          # north_star: "rel <= 1e-3 with TF32 tensor cores"
MAX_FLIPS = 6       # (sample, unit) pairs per net whose ReLU mask may differ (pre-activation within ~4e-5 of zero)
O, A, B, H = 376, 17, 256, 256


FLIP_TOL = 3e-2     # what a handful of flipped units can do to a gradient norm-wise (each one ~5e-3, measured)
LR = 3e-4


def check_weights(ours, ref, steps, what):
    """Updated weights: norm-wise <= 1e-3, or element-wise within the Adam floor.  Adam's first steps move every weight by
    ~lr * sign(g) whatever |g| is, so an element whose gradient is smaller than the 1e-3 relative tf32 noise changes sign
    between two equally valid evaluations and ends 2 * lr per step apart ("weights-after-K-steps
    with an atol ~ 2 lr floor"); on the log_std head, whose weights are themselves ~1e-3, three such elements are 1.6e-2
    norm-wise (measured, seed 55)."""
    r, m = rel_err(ours, ref), max_abs(ours, ref)
    assert r <= TOL or m <= 2 * LR * steps * 1.01, (what, r, m)


def check_net_grads(got, ref, X, what):
    """got / ref: state_dict-keyed gradients of one MLP; X: the [B, K] input rows of its first layer (fp32).
    The first layer's gradient is decomposed per sample (dW0 pinv(X) = dh1^T): every flipped ReLU unit upstream of this
    net's loss -- in the net itself or, for the policy, in the critics its loss runs through -- shows up as a sample whose
    dh1 differs.  No such sample: EVERY tensor of the net must agree norm-wise to 1e-3.  Some: the clean part of the
    first layer must, and the other tensors are bounded by what that many flips can do.  Returns the number of samples."""
    n_bad, cw, cb = per_sample_flip_split(got['fc0.weight'], ref['fc0.weight'], got['fc0.bias'], ref['fc0.bias'],
                                          orc.round_tf32(X))
    assert n_bad <= MAX_FLIPS and cw <= TOL and cb <= TOL, (what, 'fc0', n_bad, cw, cb)
    for k in got:
        r = rel_err(got[k], ref[k])
        assert r <= (TOL if n_bad == 0 else FLIP_TOL), (what, k, n_bad, r)
    return n_bad


def grads_of(e, idx, seed=0):
    """first-step gradients = exp_avg / (1 - beta1)"""
    return {k: v.cpu() / 0.1 for k, v in e.net_views(idx, seed=seed, arena=e.adam_m).items()}


def test_sac_single_seed_tf32_vs_model():
    """One seed on the TF32 path (few-seed regime: head layers stay fp32 in the glue kernels -> model mode "trunk")."""
    torch.manual_seed(0)
    tr = make_trainer(O, A, H, gemm_path=1)
    torch.manual_seed(0)
    st = orc.SACState(O, A, hidden=(H, H))          # tf32 model
    torch.manual_seed(0)
    st32 = orc.SACState(O, A, hidden=(H, H))        # fp32 oracle (values / losses)
    for s in range(3):
        batch = synth_batch(B, O, A, seed=10 + s)
        eps = synth_eps(2, B, A, seed=100 + s)
        with orc.tf32_mode("trunk"):
            out = orc.sac_step(st, batch, eps[0], eps[1])
        out32 = orc.sac_step(st32, batch, eps[0], eps[1])
        tr.inject_noise(eps[0], eps[1])
        tr._need_to_update_eval_statistics = True
        tr.train_from_torch({k: v.cuda() for k, v in batch.items()})
        e = tr._engine
        qp = e.io_view(e.lay.off_q_pred, (B, 2)).cpu()
        for ref, tol in ((out, TOL), (out32, TOL)):
            assert rel_err(qp[:, 0], ref['q1_pred'][:, 0]) <= tol
            assert rel_err(qp[:, 1], ref['q2_pred'][:, 0]) <= tol
            assert rel_err(e.io_view(e.lay.off_q_target, (B, 2)).cpu()[:, 0], ref['q_target'][:, 0]) <= tol
            assert rel_err(e.io_view(e.lay.off_log_pi, (3 * B,)).cpu()[:B], ref['log_pi'][:, 0]) <= tol
            assert abs(tr.eval_statistics['QF1 Loss'] - float(ref['qf1_loss'])) <= tol * abs(float(ref['qf1_loss']))
            assert abs(tr.eval_statistics['QF2 Loss'] - float(ref['qf2_loss'])) <= tol * abs(float(ref['qf2_loss']))
        if s == 0:
            xq = torch.cat([batch['observations'], batch['actions']], dim=1)
            check_net_grads(grads_of(e, 0), out['grad_policy'], batch['observations'], 'policy')
            check_net_grads(grads_of(e, 1), out['grad_qf1'], xq, 'qf1')
            check_net_grads(grads_of(e, 2), out['grad_qf2'], xq, 'qf2')
    for n in NETS:
        ours = net_cpu(getattr(tr, n))
        for k, v in getattr(st, n).items():
            check_weights(ours[k], v, 3, (n, k))


def test_tf32_model_differs_from_fp32_like_the_kernel():
    """The ~1e-2 gradient gap between the TF32 path and the fp32 oracle is tf32 rounding, not a kernel defect: the CPU
    model shows the same gap to fp32 (> 3e-3 on the first layer's gradient) while the kernel sits within 1e-3 of the model."""
    torch.manual_seed(0)
    tr = make_trainer(O, A, H, gemm_path=1)
    torch.manual_seed(0)
    st = orc.SACState(O, A, hidden=(H, H))
    torch.manual_seed(0)
    st32 = orc.SACState(O, A, hidden=(H, H))
    batch = synth_batch(B, O, A, seed=10)
    eps = synth_eps(2, B, A, seed=100)
    with orc.tf32_mode("trunk"):
        out = orc.sac_step(st, batch, eps[0], eps[1])
    out32 = orc.sac_step(st32, batch, eps[0], eps[1])
    tr.inject_noise(eps[0], eps[1])
    tr.train_from_torch({k: v.cuda() for k, v in batch.items()})
    g = grads_of(tr._engine, 1)
    xq = orc.round_tf32(torch.cat([batch['observations'], batch['actions']], dim=1))
    model_vs_fp32 = rel_err(out['grad_qf1']['fc0.weight'], out32['grad_qf1']['fc0.weight'])
    kernel_vs_fp32 = rel_err(g['fc0.weight'], out32['grad_qf1']['fc0.weight'])
    assert model_vs_fp32 > 3e-3 and kernel_vs_fp32 > 3e-3
    # against fp32 dozens of samples carry a flipped unit; against the model at most a handful
    bad32, _, _ = per_sample_flip_split(g['fc0.weight'], out32['grad_qf1']['fc0.weight'], g['fc0.bias'],
                                        out32['grad_qf1']['fc0.bias'], xq)
    bad, cw, cb = per_sample_flip_split(g['fc0.weight'], out['grad_qf1']['fc0.weight'], g['fc0.bias'],
                                        out['grad_qf1']['fc0.bias'], xq)
    assert bad <= MAX_FLIPS and cw <= TOL and bad32 >= 4 * max(bad, 1), (bad, bad32, cw)


def test_group_of_64_seeds_tf32_vs_oracle():
    """BASELINE config 5 as benchmarked: 64 seeds in one SACSeedGroup on the warp-specialised TMA + tcgen05 program
    (>= 12 gemm_ws stages).  A sample of 8 seeds is compared with the ORACLE (not with the repo's own fp32 singles):
    the tf32 model at <= 1e-3 on values, losses, gradients and weights, the fp32 oracle at <= 1e-3 on values / losses."""
    from oac_explore_b200.seed_group import SACSeedGroup
    S = 64
    ids = list(range(S))
    grp = SACSeedGroup(ids, O, A, hidden=H, batch=B, gemm_path=1)
    e = grp.engine
    assert e.ws_stages >= 12, e.ws_stages
    sample = [0, 7, 13, 21, 34, 42, 55, 63]
    states, states32 = {}, {}
    flips = []
    for sid in sample:
        torch.manual_seed(sid)
        states[sid] = orc.SACState(O, A, hidden=(H, H))
        torch.manual_seed(sid)
        states32[sid] = orc.SACState(O, A, hidden=(H, H))
    for step in range(2):
        outs, outs32 = {}, {}
        for slot, sid in enumerate(ids):
            batch = synth_batch(B, O, A, seed=1000 * sid + step)
            eps = synth_eps(2, B, A, seed=77 * sid + step)
            grp.load_batch(slot, batch)
            grp.inject_noise(slot, eps[0], eps[1])
            if sid in states:
                with orc.tf32_mode("many"):
                    outs[sid] = orc.sac_step(states[sid], batch, eps[0], eps[1])
                outs32[sid] = orc.sac_step(states32[sid], batch, eps[0], eps[1])
        grp.step(external_eps=True)
        torch.cuda.synchronize()
        stats = grp.stats().cpu()
        for sid in sample:
            slot = sid
            qp = e.io_view(e.lay.off_q_pred, (B, 2), seed=slot).cpu()
            qt = e.io_view(e.lay.off_q_target, (B, 2), seed=slot).cpu()[:, 0]
            lp = e.io_view(e.lay.off_log_pi, (3 * B,), seed=slot).cpu()[:B]
            for ref in (outs[sid], outs32[sid]):
                assert rel_err(qp[:, 0], ref['q1_pred'][:, 0]) <= TOL, (sid, step)
                assert rel_err(qp[:, 1], ref['q2_pred'][:, 0]) <= TOL, (sid, step)
                assert rel_err(qt, ref['q_target'][:, 0]) <= TOL, (sid, step)
                assert rel_err(lp, ref['log_pi'][:, 0]) <= TOL, (sid, step)
                assert abs(float(stats[slot, 2]) - float(ref['qf1_loss'])) <= TOL * abs(float(ref['qf1_loss']))
                assert abs(float(stats[slot, 3]) - float(ref['qf2_loss'])) <= TOL * abs(float(ref['qf2_loss']))
            if step == 0:
                batch = synth_batch(B, O, A, seed=1000 * sid + step)
                xq = torch.cat([batch['observations'], batch['actions']], dim=1)
                flips.append(check_net_grads(grads_of(e, 0, seed=slot), outs[sid]['grad_policy'], batch['observations'], (sid, 'policy')))
                flips.append(check_net_grads(grads_of(e, 1, seed=slot), outs[sid]['grad_qf1'], xq, (sid, 'qf1')))
                flips.append(check_net_grads(grads_of(e, 2, seed=slot), outs[sid]['grad_qf2'], xq, (sid, 'qf2')))
    # With ~4e-5 of accumulation noise on 2 x 65 536 pre-activations per net (+ the critics' policy-loss rows for the
    # policy) a net carries 1.3 flipped units on average, so about a quarter of the nets have none: the strict bound
    # (EVERY tensor <= 1e-3) must have been exercised by at least a sixth of the 24 nets (measured: 8)
    assert sum(1 for f in flips if f == 0) >= len(flips) // 6, flips
    for sid in sample:
        for n in NETS:
            ours = net_cpu(grp.nets[sid][n])
            for k, v in getattr(states[sid], n).items():
                check_weights(ours[k], v, 2, (sid, n, k))


@pytest.mark.parametrize("share", [True, False])
def test_poac_goac_tensor_path_vs_model(share):
    """P-OAC / G-OAC in the many-row regime (model mode "many": everything but the critic head's forward / input
    gradient, which the critic_head glue kernel computes in fp32, is a GEMM stage): gradients <= 1e-3."""
    o, a, b, h, P = 24, 4, 2048, 64, 5
    torch.manual_seed(1)
    tr = make_poac(o, a, h, P, share, False)
    tr.gemm_path = 1
    tr._make_engine(b)
    assert tr._engine.ws_stages >= 12
    torch.manual_seed(1)
    st = orc.ParticleState(o, a, hidden=(h, h), n_estimators=P, share_layers=share, q_min=0., q_max=500.)
    batch = synth_batch(b, o, a, seed=20)
    eps = synth_eps(2, b, a, seed=200)
    with orc.tf32_mode("many"):
        ref = orc.poac_step(st, batch, eps[0], eps[1])
    tr.inject_noise(eps_obs=eps[1], eps_next=eps[0])
    tr.train_from_torch({k: v.cuda() for k, v in batch.items()})
    e = tr._engine
    assert rel_err(e.io_view(e.lay.off_q_target, (b, P)).cpu().t(), ref['q_target'][:, :, 0]) <= TOL
    for k, got in grads_of(e, 0).items():
        assert rel_err(got, ref['grad_policy'][k]) <= TOL, ('poac policy', k, rel_err(got, ref['grad_policy'][k]))
    for i in range(len(st.qfs)):
        for k, got in grads_of(e, 1 + i).items():
            gref = ref['grad_qf'][i][k]
            assert rel_err(got, gref) <= TOL, ('poac qf', i, k, rel_err(got, gref))

    torch.manual_seed(2)
    tg = make_goac(o, a, h, share, False)
    tg.gemm_path = 1
    tg._make_engine(b)
    torch.manual_seed(2)
    sg = orc.GaussianState(o, a, hidden=(h, h), share_layers=share, q_min=0., q_max=500.)
    batch = synth_batch(b, o, a, seed=30)
    with orc.tf32_mode("many"):
        og = orc.goac_step(sg, batch)
    tg.train_from_torch({k: v.cuda() for k, v in batch.items()})
    e = tg._engine
    for idx, gname in ((0, 'grad_policy'), (1, 'grad_target_policy'), (2, 'grad_q')):
        for k, got in grads_of(e, idx).items():
            if og[gname][k] is None:
                continue
            assert rel_err(got, og[gname][k]) <= TOL, (gname, k, rel_err(got, og[gname][k]))


def test_strip_fused_forward_chain_vs_model_and_unfused(monkeypatch):
    """Small seed groups run the forward layers of every 128-row strip as ONE launch (gemm_chain.cuh): h1 / h2 go from
    layer to layer through tensor memory (accumulator converted in place -> tcgen05.mma with A in TMEM).  8 seeds: the
    chained program against the oracle's tf32 model (mode "chain": hidden activations are stored tf32-rounded) and against
    the unfused program (OAC_NO_CHAIN=1) on the same inputs."""
    from oac_explore_b200.seed_group import SACSeedGroup
    S = 8
    ids = list(range(S))
    grp = SACSeedGroup(ids, O, A, hidden=H, batch=B, gemm_path=1)
    monkeypatch.setenv("OAC_NO_CHAIN", "1")
    ref_grp = SACSeedGroup(ids, O, A, hidden=H, batch=B, gemm_path=1)
    monkeypatch.delenv("OAC_NO_CHAIN")
    # 4 launches fewer: policy l1 > l2 > l3, critic l1 > l2 and (backward) policy_dh2 > policy_dh1 are one launch each
    assert grp.engine.launches_per_step == ref_grp.engine.launches_per_step - 4, (grp.engine.launches_per_step, ref_grp.engine.launches_per_step)
    states, outs = {}, {}
    for sid in (0, 5):
        torch.manual_seed(sid)
        states[sid] = orc.SACState(O, A, hidden=(H, H))
    for step in range(2):
        for slot, sid in enumerate(ids):
            batch = synth_batch(B, O, A, seed=3000 * sid + step)
            eps = synth_eps(2, B, A, seed=91 * sid + step)
            for g_ in (grp, ref_grp):
                g_.load_batch(slot, batch)
                g_.inject_noise(slot, eps[0], eps[1])
            if sid in states:
                with orc.tf32_mode("chain"):
                    outs[sid] = orc.sac_step(states[sid], batch, eps[0], eps[1])
        grp.step(external_eps=True)
        ref_grp.step(external_eps=True)
        torch.cuda.synchronize()
        e, er = grp.engine, ref_grp.engine
        for slot in range(S):
            for off, shape in ((e.lay.off_q_pred, (B, 2)), (e.lay.off_q_target, (B, 2)), (e.lay.off_log_pi, (3 * B,))):
                assert rel_err(e.io_view(off, shape, seed=slot).cpu(), er.io_view(off, shape, seed=slot).cpu()) <= TOL
        for sid, o in outs.items():
            qp = e.io_view(e.lay.off_q_pred, (B, 2), seed=sid).cpu()
            assert rel_err(qp[:, 0], o['q1_pred'][:, 0]) <= TOL and rel_err(qp[:, 1], o['q2_pred'][:, 0]) <= TOL
            assert rel_err(e.io_view(e.lay.off_q_target, (B, 2), seed=sid).cpu()[:, 0], o['q_target'][:, 0]) <= TOL
            if step == 0:
                batch = synth_batch(B, O, A, seed=3000 * sid)
                xq = torch.cat([batch['observations'], batch['actions']], dim=1)
                check_net_grads(grads_of(e, 0, seed=sid), o['grad_policy'], batch['observations'], (sid, 'policy'))
                check_net_grads(grads_of(e, 1, seed=sid), o['grad_qf1'], xq, (sid, 'qf1'))
                check_net_grads(grads_of(e, 2, seed=sid), o['grad_qf2'], xq, (sid, 'qf2'))
    for sid in states:
        for n in NETS:
            ours = net_cpu(grp.nets[sid][n])
            for k, v in getattr(states[sid], n).items():
                check_weights(ours[k], v, 2, (sid, n, k))
