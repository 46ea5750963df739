#!/usr/bin/env python
"""Per-stage times of the seed-group step program (stage by stage, CUDA events) next to the graph-replayed step time.

    python tools/stage_profile.py [--seeds 8 16 64] [--gemm-path tf32] [--steps 50]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--seeds", type=int, nargs="+", default=[8, 64])
ap.add_argument("--gemm-path", default="tf32")
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--replay", type=int, default=200000)
args = ap.parse_args()

import bench
from oac_explore_b200.replay_buffer import ReplayBuffer
from oac_explore_b200.seed_group import SACSeedGroup

dev = torch.device("cuda", 0)
rb = ReplayBuffer(args.replay, bench.Box(bench.O), bench.Box(bench.A))
g = torch.Generator(device=dev).manual_seed(0)
rb._observations.normal_(generator=g); rb._next_obs.normal_(generator=g)
rb._actions.uniform_(-1, 1, generator=g); rb._rewards.normal_(generator=g)
rb._size = args.replay
np.random.seed(0)
for S in args.seeds:
    grp = SACSeedGroup(list(range(S)), bench.O, bench.A, hidden=bench.H, batch=bench.B,
                       gemm_path=bench.GEMM_PATHS[args.gemm_path], **bench.HP)
    idx = torch.from_numpy(np.random.randint(0, args.replay, (args.steps + 5, S, bench.B))).to(dev)
    for i in range(5):
        grp.gather(rb, idx[i]); grp.step()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for i in range(5, 5 + args.steps):
        grp.gather(rb, idx[i]); grp.step()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / args.steps
    scratch = SACSeedGroup(list(range(S)), bench.O, bench.A, hidden=bench.H, batch=bench.B,
                           gemm_path=bench.GEMM_PATHS[args.gemm_path], **bench.HP)
    rb.gather_into(scratch.engine, idx[0], bench.B, n_seeds=S)
    prof = scratch.engine.profile(iters=10)
    tot = sum(p[1] for p in prof)
    print("== %d seeds, %s: step %.4f ms (%.0f seed-updates/s), %d launches, stage sum %.4f ms"
          % (S, args.gemm_path, ms, S / ms * 1e3, grp.engine.launches_per_step + 1, tot))
    for i, p in enumerate(prof):
        print("   %-32s %8.4f ms %5.1f%%  %s" % (p[0], p[1], 100 * p[1] / tot, ("%.0f TFLOP/s" % (p[3] * S / p[1] / 1e9)) if p[2] else ""))
    del grp, scratch
