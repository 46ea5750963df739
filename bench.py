#!/usr/bin/env python
"""bench.py -- OAC gradient-updates/s on Humanoid shapes (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--seeds-per-gpu S] [--algo sac|poac|goac]

A "step" is one pass of the hot path over one batch: ReplayBuffer.random_batch(256) from the
1M-transition GPU-resident store + one train_from_torch (SAC/OAC update), per seed.

  value : whole-job seed-updates/s with the inputs (the index stream) already in HBM when the
          timed region starts: K x (gather kernel + step graph) between two CUDA events.
  e2e   : the same metric through the public, reference-facing API with HOST inputs:
          replay_buffer.random_batch(B) (np.random indices -> pinned -> H2D) + trainer.train(batch)
          + a D2H read of the step's scalars, every step, inside the timed region.
  N > 1 : one process per GPU (torchrun), independent seeds on each GPU (the reference's
          `seed % n_gpus` rule, main.py:575-576), no data-path collective; NCCL only gathers the
          per-seed statistics after the timed region.  scaling = weak.
  --impl reference : the reference's CPU path (oracle port of its PyTorch code; /root/reference
          does not exist on the GPU box) on all host threads, same config and metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

O, A, H, B, N_REPLAY = 376, 17, 256, 256, 1000000
# algorithmic work per update: necessary GEMM MACs only
FLOP_PER_UPDATE = {"sac": 2 * 256 * 2188800, "poac": 2 * 256 * 1401088, "goac": 2 * 256 * 2023680}
GATHER_BYTES = 2 * B * (2 * O + A + 2) * 4      # algorithmic bytes per batch: 789 504 read + the same written
HP = dict(policy_lr=3e-4, qf_lr=3e-4, soft_target_tau=5e-3, discount=0.99, reward_scale=1.0)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


def ncu_traffic(seeds, gemm_path):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant GEMM kernel per launch, from the committed
    `ncu --set full` summary of the same configuration (profiles/traffic.json, written by tools/summarize_profiles.py);
    None when that configuration was not captured."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.isfile(p):
        return None
    try:
        return json.load(open(p)).get("%d:%s" % (seeds, gemm_path))
    except Exception:
        return None


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace('.', '').isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 6:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class Box(object):
    def __init__(self, dim):
        self.low = np.full((dim,), -1.0, dtype=np.float32)
        self.high = np.full((dim,), 1.0, dtype=np.float32)
        self.shape = (dim,)


def synthetic_store_numpy(n, rng):
    return dict(obs=rng.standard_normal((n, O), dtype=np.float32), next_obs=rng.standard_normal((n, O), dtype=np.float32),
                actions=rng.uniform(-1, 1, (n, A)).astype(np.float32), rewards=rng.standard_normal((n, 1), dtype=np.float32),
                terminals=(rng.random((n, 1)) < 0.01).astype(np.float32))


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's PyTorch CPU path
# ------------------------------------------------------------------------------------------
def cpu_reference_rate(steps, warmup, threads, n_store=50000):
    """random_batch -> np_to_pytorch_batch -> train_from_torch (Mode A) on the host CPU.
    The store is a bounded sample (n_store rows of the 1M synthetic store; row reads are
    random either way) so the run stays within seconds."""
    from oracle import oac_oracle as orc
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    st = orc.SACState(O, A, hidden=(H, H), **{k: HP[k] for k in ("policy_lr", "qf_lr", "soft_target_tau", "discount", "reward_scale")})
    rng = np.random.default_rng(0)
    s = synthetic_store_numpy(n_store, rng)
    rb = orc.ReplayBuffer(n_store, O, A)
    rb._observations[:] = s["obs"]; rb._next_obs[:] = s["next_obs"]; rb._actions[:] = s["actions"]
    rb._rewards[:] = s["rewards"]; rb._terminals[:] = s["terminals"].astype(np.uint8)
    rb._size = n_store
    np.random.seed(0)

    def one():
        batch = orc.np_to_torch_batch(rb.random_batch(B))
        orc.sac_step(st, batch, None, None)
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return steps / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    rate, dt = cpu_reference_rate(args.steps, args.warmup, threads)
    line = {"metric": "OAC grad-updates/sec (Humanoid shapes, B=256)", "value": rate, "unit": "updates/s",
            "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "OAC Humanoid-v2 shapes (obs 376, act 17), batch 256, twin-Q 2x256, 1 seed, CPU",
                       "replay": "50k-row sample of the synthetic 1M store"},
            "cpu_baseline": {"value": rate, "unit": "updates/s", "cores": threads, "kind": "port",
                             "sample": "%d updates of the oracle port (reference PyTorch CPU path, mode A)" % args.steps},
            "e2e": {"value": rate, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
GEMM_PATHS = {"fp32": 0, "tf32": 1, "tf32x3": 2}
GEMM_PATH = 2


def build_trainer(algo, seed):
    from oac_explore_b200.networks import get_policy_producer, get_q_producer
    torch.manual_seed(seed)
    if algo == "sac":
        from oac_explore_b200.trainer import SACTrainer
        pp, qp = get_policy_producer(O, A, [H, H]), get_q_producer(O, A, [H, H])
        return SACTrainer(pp, qp, action_space=Box(A), use_automatic_entropy_tuning=True, rng_seed=seed,
                          target_update_period=1, gemm_path=GEMM_PATH, **HP)
    if algo == "poac":
        from oac_explore_b200.particle_trainer_oac import ParticleTrainer
        pp, qp = get_policy_producer(O, A, [H, H]), get_q_producer(O, A, [H, H], output_size=10)
        return ParticleTrainer(pp, qp, n_estimators=10, action_space=Box(A), share_layers=True, deterministic=False,
                               delta=0.95, q_min=0.0, q_max=500.0, rng_seed=seed, gemm_path=GEMM_PATH, **HP)
    from oac_explore_b200.gaussian_trainer import GaussianTrainer
    pp, qp = get_policy_producer(O, A, [H, H]), get_q_producer(O, A, [H, H], output_size=2)
    return GaussianTrainer(pp, qp, action_space=Box(A), share_layers=True, delta=0.95, q_min=0.0, q_max=500.0,
                           gemm_path=GEMM_PATH, **HP)


class _Single(object):
    """Adapter: one reference-style trainer + buffer (the public API of BASELINE config 2)."""

    def __init__(self, args, rank, rb):
        self.tr = build_trainer(args.algo, seed=rank)
        self.rb = rb
        rb.attach(self.tr)
        self.engine = self.tr._engine
        self.S = 1

    def device_step(self, idx_dev):              # idx_dev [1, B] on the device
        self.rb.gather_into(self.engine, idx_dev, B)
        self.engine.step()

    def api_step(self):
        batch = self.rb.random_batch(B)          # host np.random indices -> pinned -> H2D -> gather kernel
        batch['buffer'] = self.rb
        self.tr.train(batch)                     # fused step (CUDA graph)

    def scalars(self):
        if getattr(self, '_sc', None) is None or self._sc_engine is not self.tr._engine:
            self._sc_engine = self.tr._engine
            self._sc = self._sc_engine.scalars().view(1, 16)
        return self._sc

    api = "ReplayBuffer.random_batch(256) + SACTrainer.train(batch) + stream sync + host read of the step scalars, every step"


class _Group(object):
    """Adapter: S independent seeds batched in one engine (BASELINE config 5)."""

    def __init__(self, args, rank, world, rb):
        from oac_explore_b200.seed_group import SACSeedGroup
        S = args.seeds_per_gpu
        ids = [rank + world * i for i in range(S)]               # seed % n_gpus == rank (main.py:575-576)
        self.grp = SACSeedGroup(ids, O, A, hidden=H, batch=B, gemm_path=GEMM_PATH, rng_seed=rank, **HP)
        self.rb, self.engine, self.S = rb, self.grp.engine, S

    def device_step(self, idx_dev):              # idx_dev [S, B]
        self.grp.gather(self.rb, idx_dev)
        self.grp.step()

    def api_step(self):
        self.grp.gather(self.rb, np.random.randint(0, N_REPLAY, (self.S, B)))
        self.grp.step()

    def scalars(self):
        return self.engine.io[:, self.engine.lay.off_scalars:self.engine.lay.off_scalars + 16]

    api = "SACSeedGroup.gather(replay, host indices [S,256]) + .step() + stream sync + host read of the per-seed scalars, every step"


def batched_brief(rb, dev, pk, S=64, steps=20):
    from oac_explore_b200.seed_group import SACSeedGroup
    grp = SACSeedGroup(list(range(S)), O, A, hidden=H, batch=B, gemm_path=GEMM_PATHS["tf32"], **HP)
    idx = torch.from_numpy(np.random.randint(0, N_REPLAY, (steps + 3, S, B))).to(dev)
    for i in range(3):
        grp.gather(rb, idx[i]); grp.step()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for i in range(3, 3 + steps):
        grp.gather(rb, idx[i]); grp.step()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    prof = grp.engine.profile(iters=5)
    gemm_ms = sum(p[1] for p in prof if p[2])
    flops = FLOP_PER_UPDATE["sac"] * S
    tf32_peak = pk["bf16"] / 2.0
    return {"seeds_per_gpu": S, "gemm_path": "tf32", "value": S / (ms * 1e-3), "unit": "seed-updates/s", "ms_per_step": ms,
            "launches_per_step": grp.engine.launches_per_step + 1, "tma_tcgen05_stages": grp.engine.ws_stages,
            "roofline": {"bound": "tensor", "kernel": "gemm_ws_kernel (TMA + tcgen05 kind::tf32), all GEMM stages of one step",
                         "achieved": flops / (gemm_ms * 1e-3) / 1e12, "peak": tf32_peak, "unit": "TFLOP/s",
                         "frac": flops / (gemm_ms * 1e-3) / 1e12 / tf32_peak, "whole_step_frac": flops / (ms * 1e-3) / 1e12 / tf32_peak,
                         "traffic": ncu_traffic(S, "tf32")}}


def run_ours(args):
    import torch.distributed as dist
    from oac_explore_b200.replay_buffer import ReplayBuffer
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    # 1M-transition synthetic store, generated on the device (obs/next N(0,1), actions U(-1,1),
    # rewards N(0,1), terminals Bernoulli(0.01)); the store is setup, not step input.  The seeds of a
    # GPU read one shared store with independent index streams (seeds only read it).
    rb = ReplayBuffer(N_REPLAY, Box(O), Box(A))
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    rb._observations.normal_(generator=g); rb._next_obs.normal_(generator=g)
    rb._actions.uniform_(-1, 1, generator=g); rb._rewards.normal_(generator=g)
    rb._terminals.copy_((torch.rand(N_REPLAY, 1, device=dev, generator=g) < 0.01).float())
    rb._size, rb._top = N_REPLAY, 0
    S = args.seeds_per_gpu
    w = _Single(args, rank, rb) if S == 1 else _Group(args, rank, world, rb)
    e = w.engine
    K, W = args.steps, args.warmup
    np.random.seed(rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput ----------------
    idx_all = torch.from_numpy(np.random.randint(0, N_REPLAY, (W + K, S, B))).to(dev)
    stream = torch.cuda.current_stream()
    for i in range(W):
        w.device_step(idx_all[i])
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for i in range(W, W + K):
        w.device_step(idx_all[i])
    ev1.record(stream)
    barrier()
    ms_dev = ev0.elapsed_time(ev1)
    clk = clocks.stop()

    # ---------------- end to end through the public API ----------------
    # every step: host index draw -> pinned ring (read by the gather kernel over PCIe) -> update -> the step's scalars
    # (alpha, alpha loss, mean log pi per seed) land in the engine's mapped pinned host tensor -> synchronise -> read
    host_sc = e.host_scalars
    acc = 0.0
    for _ in range(W):
        w.api_step(); stream.synchronize()
    barrier()
    ev0.record(stream)
    t0 = time.perf_counter()
    for _ in range(K):
        w.api_step()
        stream.synchronize()
        acc += float(host_sc[0, 0])          # the host consumes the result of THIS step before the next one starts
    ev1.record(stream)
    barrier()
    ms_e2e = max(ev0.elapsed_time(ev1), 1000.0 * (time.perf_counter() - t0))
    assert acc == acc and (args.algo == "goac" or acc > 0.0), "the step's scalars never reached the host"

    # max over ranks
    t = torch.tensor([ms_dev, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # the only collective of the design: per-seed statistics gathered after the timed region
        from oac_explore_b200.seed_group import allgather_stats
        ids = [rank + world * i for i in range(S)]
        allgather_stats(w.scalars().view(S, 16).clone(), ids, S * world)
    ms_dev, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        # ---------------- roofline of the dominant kernel + cpu baseline (rank 0, N=1 only) --------
        pk = peaks()
        roof, cpu, roof_b = None, None, None
        if world == 1:
            if S == 1:
                scratch = build_trainer(args.algo, seed=99)
                scratch._ensure_engine(B)
                se = scratch._engine
            else:
                from oac_explore_b200.seed_group import SACSeedGroup
                scratch = SACSeedGroup(list(range(S)), O, A, hidden=H, batch=B, gemm_path=GEMM_PATH, **HP)
                se = scratch.engine
            rb.gather_into(se, idx_all[0], B, n_seeds=S)
            prof = se.profile(iters=50 if S == 1 else 10)
            gemm_ms = sum(p[1] for p in prof if p[2])
            gemm_flops = sum(p[3] for p in prof if p[2]) * S
            all_ms = sum(p[1] for p in prof)
            tf32_peak = pk["bf16"] / 2.0          # kind::tf32 runs at half the bf16 rate
            achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12
            kname = {0: "gemm_sk_kernel / gemm_fwd2_kernel (fp32 FFMA, 4-way split-K 32x32 tiles, TMA-staged operands)",
                     1: "gemm_ws_kernel (TMA + tcgen05 kind::tf32, warp-specialised, TMEM accumulators)",
                     2: "gemm_tc_kernel (tcgen05 kind::tf32, 3xTF32 split, TMEM accumulators)"}[GEMM_PATH]
            roof = {"bound": "tensor", "kernel": kname + ", all GEMM stages of one step",
                    "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s", "frac": achieved / tf32_peak,
                    "peak_source": "%s bf16 %.1f TFLOP/s / 2 (tf32 rate)" % (pk["source"], pk["bf16"]),
                    "traffic": ncu_traffic(S, args.gemm_path), "share_of_step_kernel_time": gemm_ms / all_ms,
                    "algorithmic_flops_per_step": FLOP_PER_UPDATE[args.algo] * S, "executed_flops_per_step": gemm_flops,
                    "note": ("fp32 FFMA kernel: the tensor pipe is not used on this path (reference-matching numerics); against "
                             "the fp32 FMA peak of 74.4 TFLOP/s (148 SMs x 128 FMA/clk x 1.965 GHz) the fraction is %.3f.  A single "
                             "seed is a latency-bound chain of ~13 dependent stages, not a throughput problem"
                             % (achieved / 74.4)) if GEMM_PATH == 0 else "",
                    "stages_ms": {p[0] + "#%d" % i: round(p[1], 5) for i, p in enumerate(prof)}}
            # replay gather: HBM roofline, timed alone: R launches captured into one CUDA graph (at S = 1 the Python /
            # ctypes call costs more than the kernel), different index rows per launch, CUDA events around the replay
            R = 100
            for _ in range(3):
                rb.gather_into(e, idx_all[0], B, n_seeds=S)
            torch.cuda.synchronize()
            gg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gg):
                for i in range(R):
                    rb.gather_into(e, idx_all[i % (W + K)], B, n_seeds=S)
            gg.replay(); torch.cuda.synchronize()
            ev0.record(); gg.replay(); ev1.record(); torch.cuda.synchronize()
            g_ms = ev0.elapsed_time(ev1) / R
            roof["replay_gather"] = {"bound": "hbm", "achieved": S * GATHER_BYTES / (g_ms * 1e-3) / 1e9, "peak": pk["hbm"],
                                     "unit": "GB/s", "frac": S * GATHER_BYTES / (g_ms * 1e-3) / 1e9 / pk["hbm"],
                                     "us_per_launch": g_ms * 1e3, "algorithmic_bytes": S * GATHER_BYTES,
                                     "note": "bytes = rows read + batch rows written; back-to-back launches inside one CUDA graph"}
            if S > 1:
                # config 5 is HBM-bound before it is tensor-bound: algorithmic state traffic per seed-update
                # = read 838 695 W + r/w 1 009 738 m,v + write 504 869 + 333 826 W floats = 14.8 MB
                state_bytes = 4.0 * (838695 + 2 * 1009738 + 504869 + 333826)
                hb = S * state_bytes / (ms_dev / K * 1e-3) / 1e9
                roof["hbm_state_traffic"] = {"bound": "hbm", "achieved": hb, "peak": pk["hbm"], "unit": "GB/s", "frac": hb / pk["hbm"],
                                             "algorithmic_bytes_per_seed_update": state_bytes,
                                             "note": "whole step: weights + Adam state only; activations (~28 MB per seed-update at "
                                                     "64 seeds) also stream through HBM"}
            else:
                # the batched-seed configuration (BASELINE config 5) in brief, so that the default run also shows the
                # tensor-core path: 64 seeds, TMA + tcgen05 kind::tf32, device-resident timing
                del scratch, se
                roof_b = batched_brief(rb, dev, pk)
            threads = os.cpu_count() or 1
            n_cpu = 150
            rate_all, _ = cpu_reference_rate(n_cpu, 5, threads)
            rate_1, _ = cpu_reference_rate(60, 3, 1)
            cpu = {"value": rate_all, "unit": "updates/s", "cores": threads, "kind": "port",
                   "sample": "%d updates of the oracle port of the reference's PyTorch CPU path (mode A), "
                             "50k-row store sample" % n_cpu,
                   "single_thread_value": rate_1}
        n_seeds = world * S
        workload = ("OAC Humanoid-v2 shapes (obs 376, act 17) synthetic 1M replay, batch 256, twin-Q 2x256, "
                    "%d seed%s per B200" % (S, "" if S == 1 else "s batched as grouped GEMMs"))
        line = {"metric": "OAC grad-updates/sec (Humanoid shapes, B=256)" if S == 1 else
                          "OAC seed-updates/sec (Humanoid shapes, B=256, %d seeds per GPU)" % S,
                "value": n_seeds * K / (ms_dev * 1e-3),
                "unit": "updates/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_dev / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload if args.algo == "sac" else args.algo,
                           "algo": args.algo, "seeds_per_gpu": S, "gemm_path": args.gemm_path, "stale_graph_mode": "A",
                           "l2": "inputs are random rows of a 3.1 GB replay store (>> 126 MB L2); the 3.4 MB of "
                                 "weights per seed stay cache-resident as in the real training loop; no explicit flush",
                           "cuda_graph": True},
                "clocks": clk,
                "e2e": {"value": n_seeds * K / (ms_e2e * 1e-3), "unit": "updates/s", "h2d_bytes_per_step": S * B * 8,
                        "d2h_bytes_per_step": S * 12, "ms_per_step": ms_e2e / K, "api": w.api,
                        "transfers": "indices: pinned host ring read by the gather kernel (single seed) / pinned -> H2D copy "
                                     "(seed group); result: 3 scalars per seed stored by the step into mapped pinned host memory"},
                "gpu_launches": (e.launches_per_step + 1) * K * 2,
                "launches_per_step": e.launches_per_step + 1}
        if roof is not None:
            line["roofline"] = roof
        if roof_b is not None:
            line["batched_seeds"] = roof_b
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--algo", default="sac", choices=["sac", "poac", "goac"])
    ap.add_argument("--seeds-per-gpu", type=int, default=1,
                    help="independent OAC seeds batched per GPU (BASELINE config 5: 64 total)")
    ap.add_argument("--gemm-path", default="auto", choices=["auto"] + list(GEMM_PATHS),
                    help="fp32: SIMT FFMA (reference-matching numerics, <=1e-5); tf32: TMA + tcgen05 kind::tf32 (<=1e-3); "
                         "tf32x3: tcgen05 3xTF32 (fp32-grade accuracy).  auto: fp32 for one seed per GPU (the step is "
                         "latency-bound there and the FFMA path is the fastest at reference numerics), tf32 for batched seeds")
    args = ap.parse_args()
    global GEMM_PATH
    if args.gemm_path == "auto":
        args.gemm_path = "fp32" if args.seeds_per_gpu == 1 else "tf32"
    GEMM_PATH = GEMM_PATHS[args.gemm_path]
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        if args.steps > 400:
            args.steps = 400      # bounded CPU sample (~15 ms per update)
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
