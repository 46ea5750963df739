"""Single-seed OAC step at Humanoid shapes: graph-replayed step time and the stage-by-stage profile (measurement aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from oac_explore_b200.replay_buffer import ReplayBuffer
dev = torch.device("cuda", 0)
rb = ReplayBuffer(200000, bench.Box(bench.O), bench.Box(bench.A))
rb._observations.normal_(); rb._next_obs.normal_(); rb._actions.uniform_(-1, 1); rb._rewards.normal_(); rb._size = 200000
bench.N_REPLAY = 200000
w = bench._Single("sac", 0, rb, 0)
idx = torch.from_numpy(np.random.randint(0, 200000, (2100, 1, bench.B))).to(dev)
for i in range(100): w.device_step(idx[i])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for i in range(100, 2100): w.device_step(idx[i])
e1.record(); torch.cuda.synchronize()
print("step %.2f us, %d launches" % (e0.elapsed_time(e1) / 2000 * 1e3, w.engine.launches_per_step))
prof = w.engine.profile(iters=50)
print(" ".join("%s=%.1f" % (p[0].split('+')[0], p[1] * 1e3) for p in prof))
