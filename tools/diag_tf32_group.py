"""Diagnostic: gradients of a TF32 seed group against the oracle's tf32 regimes (prints norm-wise errors)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import oac_oracle as orc
from tests.util import synth_batch, synth_eps, rel_err
from oac_explore_b200.seed_group import SACSeedGroup

O, A, B, H = 376, 17, 256, 256
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
grp = SACSeedGroup(list(range(S)), O, A, hidden=H, batch=B, gemm_path=1)
e = grp.engine
print("ws_stages", e.ws_stages, "launches", e.launches_per_step)
outs = {}
for slot in range(S):
    batch = synth_batch(B, O, A, seed=1000 * slot)
    eps = synth_eps(2, B, A, seed=77 * slot)
    grp.load_batch(slot, batch)
    grp.inject_noise(slot, eps[0], eps[1])
    if slot == 0:
        for mode in (None, "trunk", "many", "all"):
            torch.manual_seed(0)
            st = orc.SACState(O, A, hidden=(H, H))
            with orc.tf32_mode(mode):
                outs[mode] = orc.sac_step(st, batch, eps[0], eps[1])
grp.step(external_eps=True)
torch.cuda.synchronize()
for idx, gname in ((0, 'grad_policy'), (1, 'grad_qf1'), (2, 'grad_qf2')):
    m = e.net_views(idx, seed=0, arena=e.adam_m)
    for k in m:
        got = m[k].cpu() / 0.1
        print("%-12s %-24s" % (gname, k), "  ".join("%s %.2e" % (mode, rel_err(got, outs[mode][gname][k])) for mode in outs))
qp = e.io_view(e.lay.off_q_pred, (B, 2), seed=0).cpu()
qn = e.io_view(e.lay.off_q_new, (B, 2), seed=0).cpu()
for mode in outs:
    o = outs[mode]
    print(mode, "q1_pred %.2e  q_target %.2e log_pi %.2e" % (
        rel_err(qp[:, 0], o['q1_pred'][:, 0]),
        rel_err(e.io_view(e.lay.off_q_target, (B, 2), seed=0).cpu()[:, 0], o['q_target'][:, 0]),
        rel_err(e.io_view(e.lay.off_log_pi, (3 * B,), seed=0).cpu()[:B], o['log_pi'][:, 0])),
        "q_new(min) %.2e" % rel_err(torch.minimum(qn[:, 0], qn[:, 1]), o['q_new'][:, 0]),
        "grad_action vs", )

# row-wise picture: a flipped ReLU unit (sample b, unit n) of layer 1 changes ROW n of dW0 only
import numpy as np
for idx, gname in ((0, 'grad_policy'), (1, 'grad_qf1'), (2, 'grad_qf2')):
    m = e.net_views(idx, seed=0, arena=e.adam_m)
    for k in ('fc0.weight', 'fc1.weight'):
        got = (m[k].cpu() / 0.1).double()
        ref = outs['many'][gname][k].double()
        rows = ((got - ref).norm(dim=1) / ref.norm(dim=1).clamp_min(1e-30)).numpy()
        print("%-12s %-12s rows>1e-3: %3d  median %.1e  p90 %.1e  max %.1e   norm-wise without the rows>1e-3: %.1e" % (
            gname, k, int((rows > 1e-3).sum()), np.median(rows), np.quantile(rows, 0.9), rows.max(),
            float((got - ref)[rows <= 1e-3].norm() / ref[rows <= 1e-3].norm())))
