import cProfile, pstats, sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from oac_explore_b200.replay_buffer import ReplayBuffer
bench.GEMM_PATH = 0
dev = torch.device("cuda", 0)
rb = ReplayBuffer(200000, bench.Box(bench.O), bench.Box(bench.A))
rb._observations.normal_(); rb._next_obs.normal_(); rb._size = 200000
tr = bench.build_trainer("sac", 0)
rb.attach(tr)
sc_host = torch.zeros((1, 16)).pin_memory()
stream = torch.cuda.current_stream()
def step():
    batch = rb.random_batch(256)
    batch['buffer'] = rb
    tr.train(batch)
    sc_host.copy_(tr._engine.scalars().view(1, 16), non_blocking=True)
    stream.synchronize()
for _ in range(200): step()
pr = cProfile.Profile()
pr.enable()
for _ in range(3000): step()
pr.disable()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(18)
