#!/usr/bin/env python
"""Exploration-step latency / throughput: get_optimistic_exploration_action on Humanoid
shapes, end to end (host observation -> pinned -> kernel -> host action), and explore_batch for n observations;
the oracle port of the reference's CPU path is timed beside it.

    python tools/bench_explore.py [--n 1 16 256]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import bench
from oac_explore_b200.optimistic_exploration import get_optimistic_exploration_action, explore_batch

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, nargs="*", default=[1, 16, 256])
ap.add_argument("--iters", type=int, default=2000)
args = ap.parse_args()

tr = bench.build_trainer("sac", 0)
hp = dict(beta_UB=4.66, delta=23.53, share_layers=False)
rng = np.random.RandomState(0)
out = {}
ob = rng.randn(bench.O)
for _ in range(50):
    get_optimistic_exploration_action(ob, policy=tr.policy, qfs=tr.qfs, hyper_params=hp)
t0 = time.perf_counter()
for _ in range(args.iters):
    get_optimistic_exploration_action(ob, policy=tr.policy, qfs=tr.qfs, hyper_params=hp)
dt = time.perf_counter() - t0
out["single_call_us"] = 1e6 * dt / args.iters
out["single_call_actions_per_s"] = args.iters / dt
for n in args.n:
    obs = rng.randn(n, bench.O)
    for _ in range(20):
        explore_batch(obs, tr.policy, tr.qfs, hp)
    it = max(50, args.iters // 4)
    t0 = time.perf_counter()
    for _ in range(it):
        explore_batch(obs, tr.policy, tr.qfs, hp)
    dt = time.perf_counter() - t0
    out["batch_%d_us_per_call" % n] = 1e6 * dt / it
    out["batch_%d_actions_per_s" % n] = n * it / dt
# policy.get_action (evaluation rollouts / warm-up collection, path_collector.py:216-220)
from oac_explore_b200.networks import MakeDeterministic
for name, pol in (("get_action_stochastic", tr.policy), ("get_action_deterministic", MakeDeterministic(tr.policy))):
    for _ in range(50):
        pol.get_action(ob)
    t0 = time.perf_counter()
    for _ in range(args.iters):
        pol.get_action(ob)
    out[name + "_us"] = 1e6 * (time.perf_counter() - t0) / args.iters
# reference CPU path (oracle port), one thread (launcher_util.py:90) and all threads
from oracle import oac_oracle as orc
torch.manual_seed(0)
st = orc.SACState(bench.O, bench.A, hidden=(bench.H, bench.H))
obt = torch.from_numpy(ob).float()
for threads in (1, os.cpu_count() or 1):
    torch.set_num_threads(threads)
    for _ in range(20):
        orc.explore(obt, st.policy, [st.qf1, st.qf2], 4.66, 23.53)
    t0 = time.perf_counter()
    for _ in range(300):
        orc.explore(obt, st.policy, [st.qf1, st.qf2], 4.66, 23.53)
    dt = time.perf_counter() - t0
    out["cpu_port_%d_threads_actions_per_s" % threads] = 300 / dt
print(json.dumps(out))
