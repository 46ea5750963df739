#!/usr/bin/env python
"""Small driver for ncu: N update steps (gather + fused step) at Humanoid shapes.

    python tools/profile_step.py [--steps 5] [--algo sac] [--seeds 1] [--no-graph]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--algo", default="sac")
ap.add_argument("--no-graph", action="store_true")
ap.add_argument("--replay", type=int, default=200000)
ap.add_argument("--gemm-path", default="fp32")
ap.add_argument("--seeds", type=int, default=1)
args = ap.parse_args()
if args.no_graph:
    os.environ["OAC_NO_GRAPH"] = "1"

import bench
from oac_explore_b200.replay_buffer import ReplayBuffer

bench.N_REPLAY = args.replay
GP = bench.GEMM_PATHS[args.gemm_path]
dev = torch.device("cuda", 0)
rb = ReplayBuffer(args.replay, bench.Box(bench.O), bench.Box(bench.A))
g = torch.Generator(device=dev).manual_seed(0)
rb._observations.normal_(generator=g); rb._next_obs.normal_(generator=g)
rb._actions.uniform_(-1, 1, generator=g); rb._rewards.normal_(generator=g)
rb._size = args.replay
np.random.seed(0)
S = args.seeds
idx = torch.from_numpy(np.random.randint(0, args.replay, (args.steps, S, bench.B))).to(dev)
if S == 1:
    tr = bench.build_trainer(args.algo, 0, GP)
    rb.attach(tr)
    engine = tr._engine
else:
    from oac_explore_b200.seed_group import SACSeedGroup
    grp = SACSeedGroup(list(range(S)), bench.O, bench.A, hidden=bench.H, batch=bench.B, gemm_path=GP, **bench.HP)
    engine = grp.engine
torch.cuda.synchronize()
for i in range(args.steps):
    rb.gather_into(engine, idx[i], bench.B, n_seeds=S)
    engine.step()
torch.cuda.synchronize()
print("done", args.steps, "steps;", engine.launches_per_step, "launches/step; seeds", S)
