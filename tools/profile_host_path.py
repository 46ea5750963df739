"""cProfile of the end-to-end loop of bench.py's headline (random_batch + train + sync + host read): where the host time goes."""
import cProfile, pstats, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from oac_explore_b200.replay_buffer import ReplayBuffer
dev = torch.device("cuda", 0)
rb = ReplayBuffer(200000, bench.Box(bench.O), bench.Box(bench.A))
rb._observations.normal_(); rb._next_obs.normal_(); rb._actions.uniform_(-1, 1); rb._rewards.normal_(); rb._size = 200000
bench.N_REPLAY = 200000
w = bench._Single("sac", 0, rb, 0)
stream = torch.cuda.current_stream()
hs = w.engine.host_scalars
for _ in range(200):
    w.api_step(); stream.synchronize()
def loop(n):
    acc = 0.0
    for _ in range(n):
        w.api_step(); stream.synchronize(); acc += float(hs[(w.engine.steps - 1) & 1, 0, 0])
    return acc
t0 = time.perf_counter(); loop(3000); dt = (time.perf_counter() - t0) / 3000
print("e2e (blocking) %.1f us/step" % (dt * 1e6))
stream.synchronize()
t0 = time.perf_counter()
for _ in range(400):
    w.api_step()
dt = (time.perf_counter() - t0) / 400
stream.synchronize()
print("host enqueue only (no wait, 400 steps run ahead) %.1f us/step" % (dt * 1e6))
pr = cProfile.Profile(); pr.enable(); loop(3000); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
