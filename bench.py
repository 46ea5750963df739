#!/usr/bin/env python
"""bench.py -- OAC gradient-updates/s on Humanoid shapes (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--seeds-per-gpu S] [--algo sac|poac|goac] [--total-seeds 64]

A "step" is one pass of the hot path over one batch: ReplayBuffer.random_batch(256) from the
1M-transition GPU-resident store + one train_from_torch (SAC/OAC update), per seed.

  value : whole-job updates/s of the headline workload (BASELINE config 2: ONE seed per GPU, fp32 path with
          reference-matching numerics) with the inputs (the index stream) already in HBM when the timed region starts:
          K x (gather kernel + step graph) between two CUDA events, max over ranks.
  e2e   : the same metric through the public, reference-facing API with HOST inputs:
          replay_buffer.random_batch(B) (np.random indices -> pinned host memory -> async H2D on a copy stream) +
          trainer.train(batch) + a host read of the step's scalars, every step, inside the timed region.
  batched_seeds : BASELINE config 5 at EVERY N: 64 independent seeds in total, 64 / N per GPU (seed s on GPU s % N, the
          reference's rule, main.py:575-576), batched as grouped GEMMs on the TMA + tcgen05 kind::tf32 path: value
          (seed-updates/s over all GPUs, max-over-ranks timing), e2e, tensor and HBM rooflines.  Total work is fixed as N
          grows (strong scaling of the 64-seed job); the headline `value` above is weak scaling (one seed per GPU).
  N > 1 : one process per GPU (torchrun), independent seeds on each GPU, no data-path collective; NCCL only all-gathers
          the per-seed statistics vectors after the timed regions.  Every rank pins itself to its own share of the host cores.
  N = 1 : additionally `variants` (P-OAC, G-OAC updates/s; exploration latency at 1 / 16 / 256 observations per call), the
          rooflines of the dominant kernels and `cpu_baseline`.
  --impl reference : the reference's OWN code (oracle/_ref: its hot-path modules byte-compiled by oracle/build_ref.py,
          imported with the gym / matplotlib / gtimer stand-ins and the torch-1.4 optimizer patch) running
          random_batch -> np_to_pytorch_batch -> train_from_torch on the host CPU, full 1M-row float64 store,
          >= 5 warm-up + 200 timed updates, median of 3 repeats, on all host threads (the
          single-thread rate -- the reference's own setting, launcher_util.py:90 -- is reported next to it).
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
# algorithmic state traffic per seed-update when the state is not on-chip: read 838 695 W, read+write
# 1 009 738 m and v, write 504 869 + 333 826 W floats
STATE_BYTES = 4.0 * (838695 + 2 * 1009738 + 504869 + 333826)
HP = dict(policy_lr=3e-4, qf_lr=3e-4, soft_target_tau=5e-3, discount=0.99, reward_scale=1.0)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant GEMM kernel per launch, from the committed
    `ncu --set full` summary of the same configuration (profiles/traffic.json); None when it was not captured."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.isfile(p):
        return None
    try:
        return json.load(open(p)).get(key)
    except Exception:
        return None


class ClockSampler(object):
    """SM clock and throttle reasons while the timed regions run (B200_PROFILING.md recipe).  NVML is polled from a
    thread every few milliseconds (an `nvidia-smi -lms 100` process sees nothing of a 2 ms timed region); nvidia-smi is
    the fallback.  `mark_load(True/False)` brackets the timed regions so the median is taken under load."""
    REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40), ("sw_power_cap", 0x4))

    def __init__(self, index, period=0.004):
        self.index, self.period = index, period
        self.rows, self.loaded, self._stop, self.thread = [], False, False, None
        self.how = None

    def _poll_nvml(self):
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop:
            try:
                self.rows.append((self.loaded, float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), float(mx),
                                  int(get_reasons(h))))
            except Exception:
                pass
            time.sleep(self.period)

    def _poll_smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active"
        while not self._stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.rows.append((self.loaded, float(out[0]), float(out[1]), int(out[2].strip(), 16)))
            except Exception:
                time.sleep(0.05)

    def start(self):
        try:
            import pynvml  # noqa: F401
            target, self.how = self._poll_nvml, "nvml poll every %d ms" % int(self.period * 1000)
        except Exception:
            target, self.how = self._poll_smi, "nvidia-smi query loop"
        self.thread = threading.Thread(target=target, daemon=True)
        self.thread.start()

    def mark_load(self, on):
        self.loaded = on

    def stop(self):
        self._stop = True
        if self.thread is not None:
            self.thread.join(timeout=2)
        rows = [r for r in self.rows if r[0]] or self.rows
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples"], "samples": 0}
        bits = 0
        for r in rows:
            bits |= r[3]
        return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": max(r[2] for r in rows),
                "reasons": sorted(n for n, b in self.REASONS if bits & b), "samples": len(rows), "how": self.how,
                "window": "samples taken while the timed loops (or, when those are shorter than the sampling period, an untimed "
                          "continuation of the same loop) were running"}


class Box(object):
    def __init__(self, dim, low=-1.0, high=1.0):
        self.low = np.full((dim,), low, dtype=np.float32)
        self.high = np.full((dim,), high, dtype=np.float32)
        self.shape = (dim,)


def pin_rank_to_cores(local, local_world):
    """One NUMA-agnostic slice of the host cores per rank: 8 unpinned processes sharing the same cores showed up as a 4 %
    max-over-ranks penalty on the replica line of round 1."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = len(cores) // max(local_world, 1)
        if local_world > 1 and per >= 1:
            os.sched_setaffinity(0, cores[local * per:(local + 1) * per])
            torch.set_num_threads(max(1, min(per, 8)))
            return per
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own code (oracle/_ref), else the oracle port
# ------------------------------------------------------------------------------------------
def fill_reference_store(rb, n):
    """numpy default_rng(0): obs/next_obs ~ N(0,1), actions ~ U(-1,1), rewards ~ N(0,1),
    terminals ~ Bernoulli(0.01); float64 / uint8 like the reference's store."""
    rng = np.random.default_rng(0)
    rng.standard_normal(out=rb._observations)
    rng.standard_normal(out=rb._next_obs)
    rng.random(out=rb._actions)
    rb._actions *= 2.0
    rb._actions -= 1.0
    rng.standard_normal(out=rb._rewards)
    rb._terminals[:] = (rng.random((n, 1)) < 0.01)
    rb._size, rb._top = n, 0


def cpu_reference_rates(timed=200, warmup=5, repeats=3, n_store=N_REPLAY, budget_s=150.0):
    """random_batch -> np_to_pytorch_batch -> train_from_torch on the host CPU.  Returns a dict with the all-threads and
    single-thread rates (median of `repeats` x `timed` updates each) and what ran."""
    from oracle import ref_import as ri
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    use_ref = ri.import_root() is not None
    torch.manual_seed(0)
    if use_ref:
        ref = ri.load_reference()
        ob_space, ac_space = ri.make_spaces(O, A)
        pp, qp = ri.make_producers(O, A, hidden=(H, H))
        tr = ref.trainer.SACTrainer(pp, qp, action_space=ac_space, use_automatic_entropy_tuning=True, **HP)
        ri.mode_a(tr)                                            # torch-1.4 semantics
        rb = ref.replay_buffer.ReplayBuffer(n_store, ob_space, ac_space)
        fill_reference_store(rb, n_store)

        def one():
            batch = rb.random_batch(B)                           # rl_algorithm.py:161
            batch['buffer'] = rb                                 # :166
            tr.train(batch)                                      # :167 -> np_to_pytorch_batch -> train_from_torch
        kind = "reference"
        what = ("the reference's own ReplayBuffer.random_batch(256) + SACTrainer.train (np_to_pytorch_batch + "
                "train_from_torch), imported from %s" % ("its sources" if ref.root == ri.REFERENCE_ROOT else
                                                           "oracle/_ref (byte-compiled from its sources)"))
    else:
        from oracle import oac_oracle as orc
        st = orc.SACState(O, A, hidden=(H, H), **HP)
        rb = orc.ReplayBuffer(n_store, O, A)
        fill_reference_store(rb, n_store)

        def one():
            orc.sac_step(st, orc.np_to_torch_batch(rb.random_batch(B)), None, None)
        kind = "port"
        what = "the oracle port of the reference's PyTorch CPU path (oracle/_ref not built)"
    np.random.seed(0)

    def measure(nthreads, deadline):
        torch.set_num_threads(nthreads)
        for _ in range(warmup):
            one()
        rates = []
        for _ in range(repeats):
            t0 = time.perf_counter()
            for _ in range(timed):
                one()
            rates.append(timed / (time.perf_counter() - t0))
            if time.perf_counter() > deadline:
                break
        return float(np.median(rates)), len(rates)

    t_start = time.perf_counter()
    rate_all, n_all = measure(threads, t_start + budget_s * 0.5)
    rate_1, n_1 = measure(1, t_start + budget_s)
    torch.set_num_threads(threads)
    return dict(all=rate_all, single=rate_1, threads=threads, kind=kind, what=what, timed=timed, warmup=warmup,
                repeats=(n_all, n_1), n_store=n_store)


def cpu_baseline_record(r):
    """`value` is the BEST of the two thread settings (all host threads / one thread): the CPU arm gets whichever is faster
    on this box, so the GPU / CPU ratio is the conservative one."""
    best_all = r["all"] >= r["single"]
    return {"value": max(r["all"], r["single"]), "unit": "updates/s", "cores": r["threads"] if best_all else 1,
            "kind": r["kind"],
            "sample": "%s; 1M-row float64 store (6.2 GB), %d warm-up + %d timed updates, median of %d repeats; best of %d threads "
                      "(%.1f /s) and 1 thread (%.1f /s)"
                      % (r["what"], r["warmup"], r["timed"], r["repeats"][0], r["threads"], r["all"], r["single"]),
            "all_threads_value": r["all"], "all_threads": r["threads"], "single_thread_value": r["single"],
            "single_thread_note": "torch.set_num_threads(1) is the reference's own setting (launcher_util.py:90)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return                        # rank 0 alone runs and prints the CPU arm
    r = cpu_reference_rates()
    rate = max(r["all"], r["single"])
    line = {"metric": "OAC grad-updates/sec (Humanoid shapes, B=256)", "value": rate, "unit": "updates/s",
            "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 / rate, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "OAC Humanoid-v2 shapes (obs 376, act 17) synthetic 1M replay, batch 256, twin-Q 2x256, "
                                   "1 seed, host CPU",
                       "protocol": "BASELINE.md section 3: %d warm-up + %d timed updates, median of %d repeats, whatever "
                                   "--steps / --warmup say (a bounded sample: the run ends within a minute)"
                                   % (r["warmup"], r["timed"], r["repeats"][0]),
                       "processes": 1,
                       "note": "ONE CPU process whatever --gpus says: at N > 1 the repo arm's value is N seeds on N GPUs"},
            "cpu_baseline": cpu_baseline_record(r),
            "e2e": {"value": rate, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
GEMM_PATHS = {"fp32": 0, "tf32": 1, "tf32x3": 2}


def build_trainer(algo, seed, gemm_path):
    from oac_explore_b200.networks import get_policy_producer, get_q_producer
    torch.manual_seed(seed)
    if algo == "sac":
        from oac_explore_b200.trainer import SACTrainer
        pp, qp = get_policy_producer(O, A, [H, H]), get_q_producer(O, A, [H, H])
        return SACTrainer(pp, qp, action_space=Box(A), use_automatic_entropy_tuning=True, rng_seed=seed,
                          target_update_period=1, gemm_path=gemm_path, **HP)
    if algo == "poac":
        from oac_explore_b200.particle_trainer_oac import ParticleTrainer
        pp, qp = get_policy_producer(O, A, [H, H]), get_q_producer(O, A, [H, H], output_size=10)
        return ParticleTrainer(pp, qp, n_estimators=10, action_space=Box(A), share_layers=True, deterministic=False,
                               delta=0.95, q_min=0.0, q_max=500.0, rng_seed=seed, gemm_path=gemm_path, **HP)
    from oac_explore_b200.gaussian_trainer import GaussianTrainer
    pp, qp = get_policy_producer(O, A, [H, H]), get_q_producer(O, A, [H, H], output_size=2)
    return GaussianTrainer(pp, qp, action_space=Box(A), share_layers=True, delta=0.95, q_min=0.0, q_max=500.0,
                           gemm_path=gemm_path, **HP)


class _Single(object):
    """Adapter: one reference-style trainer + buffer (the public API of BASELINE config 2)."""

    def __init__(self, algo, seed, rb, gemm_path):
        self.tr = build_trainer(algo, seed, gemm_path)
        self.rb = rb
        rb.attach(self.tr)
        self.tr._ensure_engine(B)
        self.engine = self.tr._engine
        self.S = 1

    def device_step(self, idx_dev):              # idx_dev [1, B] on the device
        self.rb.gather_into(self.engine, idx_dev, B)
        self.engine.step()

    def api_step(self):
        batch = self.rb.random_batch(B)          # host np.random indices -> pinned ring -> device ring (copy stream) -> gather
        batch['buffer'] = self.rb
        self.tr.train(batch)                     # fused step (CUDA graph)

    api = "ReplayBuffer.random_batch(256) + SACTrainer.train(batch) + host read of the step's scalars, every step"


class _Group(object):
    """Adapter: S independent seeds batched in one engine (BASELINE config 5)."""

    def __init__(self, seed_ids, rb, gemm_path):
        from oac_explore_b200.seed_group import SACSeedGroup
        self.grp = SACSeedGroup(seed_ids, O, A, hidden=H, batch=B, gemm_path=gemm_path, **HP)
        self.rb, self.engine, self.S = rb, self.grp.engine, len(seed_ids)

    def device_step(self, idx_dev):              # idx_dev [S, B]
        self.grp.gather(self.rb, idx_dev)
        self.grp.step()

    def api_step(self):
        self.grp.gather(self.rb, np.random.randint(0, N_REPLAY, (self.S, B)))
        self.grp.step()

    api = "SACSeedGroup.gather(replay, host indices [S,256]) + .step() + host read of the per-seed scalars, every step"


def timed_loops(w, K, W, dev, barrier, clocks=None, check_scalar=True):
    """(ms for K device-resident steps, ms for K end-to-end steps) of one adapter, CUDA events on the launching stream
    bracketed by barrier + synchronize."""
    S = w.S
    idx_all = torch.from_numpy(np.random.randint(0, N_REPLAY, (W + K, S, B))).to(dev)
    stream = torch.cuda.current_stream()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(W):
        w.device_step(idx_all[i])
    barrier()
    if clocks:
        clocks.mark_load(True)
    ev0.record(stream)
    for i in range(W, W + K):
        w.device_step(idx_all[i])
    ev1.record(stream)
    barrier()
    ms_dev = ev0.elapsed_time(ev1)
    # end to end: every step, host index draw -> pinned -> update -> the step's scalars (alpha, alpha loss, mean log pi per
    # seed, the step's number) land in the engine's mapped pinned host tensor -> the host reads them.
    #   blocking : synchronise after every step and read its scalars before the next one is issued;
    #   pipelined: the reference's training loop (rl_algorithm.py:159-165: random_batch -> trainer.train, nothing waits on
    #              a step) with the per-step host read kept: while step i runs, the host draws and enqueues step i + 1,
    #              then waits for step i's event and reads ITS scalars (two mapped slots, stamped with the step number).
    #              Every step's inputs still go host -> device and every step's result is still read inside the region.
    eng = w.engine
    acc = 0.0
    for _ in range(W):
        w.api_step(); stream.synchronize()
    barrier()
    ev0.record(stream)
    t0 = time.perf_counter()
    for _ in range(K):
        w.api_step()
        stream.synchronize()
        acc += float(eng.step_scalars()[0, 0])
    ev1.record(stream)
    barrier()
    ms_block = max(ev0.elapsed_time(ev1), 1000.0 * (time.perf_counter() - t0))
    done = [torch.cuda.Event(), torch.cuda.Event()]
    stamps_ok = True
    ev0.record(stream)
    t0 = time.perf_counter()
    for i in range(K):
        w.api_step()
        done[i & 1].record(stream)
        if i:
            done[(i - 1) & 1].synchronize()
            sc = eng.step_scalars(eng.steps - 1)
            acc += float(sc[0, 0])
            stamps_ok &= int(sc[0, 3]) == eng.steps - 1
    done[(K - 1) & 1].synchronize()
    sc = eng.step_scalars(eng.steps)
    acc += float(sc[0, 0])
    stamps_ok &= int(sc[0, 3]) == eng.steps
    ev1.record(stream)
    barrier()
    ms_e2e = max(ev0.elapsed_time(ev1), 1000.0 * (time.perf_counter() - t0))
    if clocks:
        clocks.mark_load(False)
    assert stamps_ok, "a pipelined read saw another step's scalars"
    if check_scalar:
        assert acc == acc and acc > 0.0, "the step's scalars never reached the host"
    return ms_dev, ms_e2e, ms_block, idx_all


def stage_profile(engine, iters):
    prof = engine.profile(iters=iters)
    gemm_ms = sum(p[1] for p in prof if p[2])
    gemm_flops = sum(p[3] for p in prof if p[2]) * engine.cfg.n_seeds
    return prof, gemm_ms, gemm_flops, sum(p[1] for p in prof)


def variants_brief(rb, dev, steps=20, warm=5):
    """N = 1 only: the other configurations of BASELINE.json in brief (device-resident timing, CUDA events)."""
    from oac_explore_b200.optimistic_exploration import explore_batch
    out = {}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    idx = torch.from_numpy(np.random.randint(0, N_REPLAY, (warm + steps, 1, B))).to(dev)
    keep = None
    for algo, name in (("poac", "P-OAC particle_trainer_oac, 10 Q-particles on a shared trunk (config 3)"),
                       ("goac", "G-OAC gaussian_trainer, shared mean/std critic (config 4)")):
        w = _Single(algo, 7, rb, GEMM_PATHS["fp32"])
        for i in range(warm):
            w.device_step(idx[i])
        torch.cuda.synchronize()
        ev0.record()
        for i in range(warm, warm + steps):
            w.device_step(idx[i])
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / steps
        out[algo] = {"workload": name, "value": 1000.0 / ms, "unit": "updates/s", "ms_per_step": ms, "steps": steps,
                     "gemm_path": "fp32", "launches_per_step": w.engine.launches_per_step + 1,
                     "algorithmic_flops_per_update": FLOP_PER_UPDATE[algo]}
        keep = w if algo == "poac" else keep
    # exploration: host observation in, host action out, per call (path_collector.py:214-232 calls it per env step)
    w = _Single("sac", 3, rb, GEMM_PATHS["fp32"])
    hp = dict(beta_UB=4.66, delta=23.53, share_layers=False)
    rng = np.random.RandomState(0)
    ex = {}
    for n in (1, 16, 256):
        obs = rng.randn(n, O)
        for _ in range(5):
            explore_batch(obs, w.tr.policy, w.tr.qfs, hp)
        t0 = time.perf_counter()
        for _ in range(steps):
            explore_batch(obs, w.tr.policy, w.tr.qfs, hp)
        dt = (time.perf_counter() - t0) / steps
        ex["%d_obs" % n] = {"us_per_call": dt * 1e6, "actions_per_s": n / dt}
    out["explore"] = {"workload": "get_optimistic_exploration_action (twin-Q, beta_UB=4.66, delta=23.53), host observation in -> "
                                  "host action out, wall clock per call", "calls": steps, **ex}
    rb.attach(None)
    del w
    # per-seed batched exploration: ONE launch serves the current observation of every seed of a 64-seed group
    # with that seed's own policy and critics; host observations in -> host actions out
    from oac_explore_b200.seed_group import SACSeedGroup
    S = 64
    grp = SACSeedGroup(list(range(S)), O, A, hidden=H, batch=B, gemm_path=GEMM_PATHS["tf32"], **HP)
    exg = grp.explorer(hp)
    obs = rng.randn(S, O)
    for _ in range(5):
        exg.actions(obs)
    t0 = time.perf_counter()
    for _ in range(steps):
        exg.actions(obs)
    dt = (time.perf_counter() - t0) / steps
    # double buffered: the second half of the environments is submitted before the first half is collected
    half = S // 2
    sl = np.arange(S)
    for _ in range(3):
        a_ = exg.submit(obs[:half], slots=sl[:half]); b_ = exg.submit(obs[half:], slots=sl[half:]); exg.collect(a_); exg.collect(b_)
    t0 = time.perf_counter()
    for _ in range(steps):
        a_ = exg.submit(obs[:half], slots=sl[:half]); b_ = exg.submit(obs[half:], slots=sl[half:]); exg.collect(a_); exg.collect(b_)
    dt2 = (time.perf_counter() - t0) / steps
    out["explore_group"] = {"workload": "GroupExplorer: one oac_explore launch for the 64 seeds of a SACSeedGroup, every observation "
                                        "evaluated with its own seed's policy and critics (path_collector.py:214-232 vectorised over seeds)",
                            "seeds": S, "us_per_call": dt * 1e6, "actions_per_s": S / dt,
                            "double_buffered_us_per_64": dt2 * 1e6, "calls": steps}
    del grp, exg
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch.distributed as dist
    from oac_explore_b200.replay_buffer import ReplayBuffer
    from oac_explore_b200.seed_group import allgather_stats
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    cores = pin_rank_to_cores(local, local_world)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    pk = peaks()
    tf32_peak = pk["bf16"] / 2.0              # kind::tf32 runs at half the bf16 rate

    # 1M-transition synthetic store, generated on the device (obs/next N(0,1), actions U(-1,1), rewards N(0,1), terminals
    # Bernoulli(0.01)); the store is setup, not step input.  The seeds of a GPU read one shared store with independent
    # index streams (seeds only read it).
    rb = ReplayBuffer(N_REPLAY, Box(O), Box(A))
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    rb._observations.normal_(generator=g); rb._next_obs.normal_(generator=g)
    rb._actions.uniform_(-1, 1, generator=g); rb._rewards.normal_(generator=g)
    rb._terminals.copy_((torch.rand(N_REPLAY, 1, device=dev, generator=g) < 0.01).float())
    rb._size, rb._top = N_REPLAY, 0
    K, W = args.steps, args.warmup
    np.random.seed(rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(vals):
        t = torch.tensor(vals, device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    clocks = ClockSampler(local)
    clocks.start()
    # ---------------- headline: S seeds per GPU (default 1: BASELINE config 2, fp32 path) ----------------
    S = args.seeds_per_gpu
    gp = GEMM_PATHS[args.gemm_path]
    if S == 1:
        w = _Single(args.algo, rank, rb, gp)
    else:
        w = _Group([rank + world * i for i in range(S)], rb, gp)
    e = w.engine
    ms_dev, ms_e2e, ms_block, idx_all = timed_loops(w, K, W, dev, barrier, clocks, check_scalar=args.algo != "goac")
    ms_dev, ms_e2e, ms_block = max_over_ranks([ms_dev, ms_e2e, ms_block])
    launches = (e.launches_per_step + 1) * K * 3

    # ---------------- BASELINE config 5 at every N: 64 seeds in total, 64 / N per GPU, TF32 ----------------
    batched = None
    if args.total_seeds > 0 and S == 1 and args.algo == "sac":
        total = args.total_seeds
        ids = [s for s in range(total) if s % world == rank]          # main.py:575-576: gpu = seed % n_gpus
        rb.attach(None)
        wg = _Group(ids, rb, GEMM_PATHS["tf32"])
        Kb = max(10, min(K, 200))
        b_dev, b_e2e, b_block, idx_b = timed_loops(wg, Kb, max(3, min(W, 10)), dev, barrier, clocks)
        b_dev, b_e2e, b_block = max_over_ranks([b_dev, b_e2e, b_block])
        launches += (wg.engine.launches_per_step + 1) * Kb * 3
        # the only collective of the design: every seed's statistics vector, gathered after the timed regions
        stats = allgather_stats(wg.grp.stats(), ids, total)
        assert stats.shape[0] == total and bool(torch.isfinite(stats).all())
        Sg = len(ids)
        ms_b = b_dev / Kb
        batched = {"workload": "BASELINE config 5: %d independent OAC seeds in total, batched as grouped GEMMs, %d per B200 "
                               "(seed s on GPU s %% %d)" % (total, Sg, world),
                   "total_seeds": total, "seeds_per_gpu": Sg, "n_gpus": world, "gemm_path": "tf32", "scaling": "strong",
                   "value": total * Kb / (b_dev * 1e-3), "unit": "seed-updates/s", "ms_per_step": ms_b, "steps": Kb,
                   "e2e": {"value": total * Kb / (b_e2e * 1e-3), "unit": "seed-updates/s", "ms_per_step": b_e2e / Kb,
                           "h2d_bytes_per_step": Sg * B * 8, "d2h_bytes_per_step": Sg * 16, "api": wg.api,
                           "blocking_value": total * Kb / (b_block * 1e-3), "blocking_ms_per_step": b_block / Kb},
                   "launches_per_step": wg.engine.launches_per_step + 1, "tma_tcgen05_stages": wg.engine.ws_stages,
                   "stats_allgather": {"seeds": total, "floats_per_seed": int(stats.shape[1]),
                                       "collective": "nccl all_gather" if world > 1 else "none (one rank)"},
                   "roofline_tensor": {"bound": "tensor", "achieved": FLOP_PER_UPDATE["sac"] * Sg / (ms_b * 1e-3) / 1e12,
                                       "peak": tf32_peak, "unit": "TFLOP/s",
                                       "frac": FLOP_PER_UPDATE["sac"] * Sg / (ms_b * 1e-3) / 1e12 / tf32_peak,
                                       "note": "whole step, rank 0's GPU: algorithmic FLOPs of its seeds / step time"},
                   "roofline_hbm": {"bound": "hbm", "achieved": Sg * STATE_BYTES / (ms_b * 1e-3) / 1e9, "peak": pk["hbm"],
                                    "unit": "GB/s", "frac": Sg * STATE_BYTES / (ms_b * 1e-3) / 1e9 / pk["hbm"],
                                    "algorithmic_bytes_per_seed_update": STATE_BYTES,
                                    "note": "whole step, rank 0's GPU: weights + Adam state only (the mandatory traffic when "
                                            "the state is not on-chip: 76 FLOP/B, below the TF32 ridge -> HBM-bound, SURVEY 8d)"}}
        if rank == 0 and world == 1:
            # stage-by-stage profile of the same program (scratch group: the profile loop drifts the state)
            scratch = _Group(ids, rb, GEMM_PATHS["tf32"])
            rb.gather_into(scratch.engine, idx_b[0], B, n_seeds=Sg)
            prof, gemm_ms, gemm_flops, all_ms = stage_profile(scratch.engine, 10)
            batched["roofline"] = {"bound": "tensor",
                                   "kernel": "gemm_ws_kernel (TMA + tcgen05 kind::tf32, warp-specialised persistent), all GEMM "
                                             "stages of one step",
                                   "achieved": gemm_flops / (gemm_ms * 1e-3) / 1e12, "peak": tf32_peak, "unit": "TFLOP/s",
                                   "frac": gemm_flops / (gemm_ms * 1e-3) / 1e12 / tf32_peak,
                                   "share_of_step_kernel_time": gemm_ms / all_ms, "traffic": ncu_traffic("%d:tf32" % Sg),
                                   "stages_ms": {p[0] + "#%d" % i: round(p[1], 5) for i, p in enumerate(prof)}}
            del scratch
        del wg
        torch.cuda.empty_cache()
        if world > 1:
            # the same program with the per-GPU load held at `total` seeds (weak scaling: N x total seeds in all): what the N
            # GPUs deliver when every one of them has a full group -- the strong-scaling curve above is bounded by the
            # per-stage floors of small groups, not by anything between the GPUs (there is no data-path collective)
            idw = [rank * total + i for i in range(total)]
            ww = _Group(idw, rb, GEMM_PATHS["tf32"])
            Kw = max(10, min(K, 100))
            w_dev, w_e2e, w_block, _ = timed_loops(ww, Kw, 3, dev, barrier, clocks)
            w_dev, w_e2e, w_block = max_over_ranks([w_dev, w_e2e, w_block])
            launches += (ww.engine.launches_per_step + 1) * Kw * 3
            batched["weak"] = {"workload": "%d seeds on EVERY GPU (%d in all), same program" % (total, total * world),
                               "scaling": "weak", "seeds_per_gpu": total, "total_seeds": total * world,
                               "value": total * world * Kw / (w_dev * 1e-3), "unit": "seed-updates/s",
                               "ms_per_step": w_dev / Kw, "steps": Kw,
                               "e2e": {"value": total * world * Kw / (w_e2e * 1e-3), "unit": "seed-updates/s",
                                       "blocking_value": total * world * Kw / (w_block * 1e-3)}}
            del ww
            torch.cuda.empty_cache()

    # keep the GPU busy with the headline loop while the clock sampler collects (the timed regions may be shorter than
    # its period); untimed
    if rank == 0:
        t_end = time.perf_counter() + 0.25
        clocks.mark_load(True)
        i = 0
        while time.perf_counter() < t_end:
            w.device_step(idx_all[i % (W + K)]); i += 1
            if i % 50 == 0:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        clocks.mark_load(False)
    clk = clocks.stop()

    if rank == 0:
        roof, cpu, variants = None, None, None
        if world == 1:
            # ---------------- rooflines of the dominant kernels (N = 1) ----------------
            if S == 1:
                scratch = build_trainer(args.algo, 99, gp)
                scratch._ensure_engine(B)
                se = scratch._engine
            else:
                scratch = _Group(list(range(S)), rb, gp)
                se = scratch.engine
            rb.gather_into(se, idx_all[0], B, n_seeds=S)
            prof, gemm_ms, gemm_flops, all_ms = stage_profile(se, 50 if S == 1 else 10)
            achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12
            kname = {0: "gemm_sk_kernel / gemm_fwd2_kernel (fp32 FFMA, 4-way split-K 32x32 tiles, TMA-staged operands)",
                     1: "gemm_ws_kernel (TMA + tcgen05 kind::tf32, warp-specialised, TMEM accumulators)",
                     2: "gemm_tc_kernel (tcgen05 kind::tf32, 3xTF32 split, TMEM accumulators)"}[gp]
            roof = {"bound": "tensor", "kernel": kname + ", all GEMM stages of one step",
                    "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s", "frac": achieved / tf32_peak,
                    "peak_source": "%s bf16 %.1f TFLOP/s / 2 (tf32 rate)" % (pk["source"], pk["bf16"]),
                    "traffic": ncu_traffic("%d:%s" % (S, args.gemm_path)), "share_of_step_kernel_time": gemm_ms / all_ms,
                    "algorithmic_flops_per_step": FLOP_PER_UPDATE[args.algo] * S, "executed_flops_per_step": gemm_flops,
                    "note": ("fp32 FFMA kernel: the tensor pipe is not used on this path (reference-matching numerics); against "
                             "the fp32 FMA peak of 74.4 TFLOP/s (148 SMs x 128 FMA/clk x 1.965 GHz) the fraction is %.3f.  A single "
                             "seed is a latency-bound chain of dependent stages, not a throughput problem: the figure that "
                             "matters is us per update" % (achieved / 74.4)) if gp == 0 else "",
                    "stages_ms": {p[0] + "#%d" % i: round(p[1], 5) for i, p in enumerate(prof)}}
            del scratch, se
            # replay gather: HBM roofline, timed alone: R launches captured into one CUDA graph (at S = 1 the Python /
            # ctypes call costs more than the kernel), different index rows per launch, CUDA events around the replay
            R = 100
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
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
            del gg
            if args.algo == "sac" and S == 1:
                variants = variants_brief(rb, dev)
            cpu = cpu_baseline_record(cpu_reference_rates())
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
                           "cuda_graph": True, "host_cores_per_rank": cores},
                "clocks": clk,
                "e2e": {"value": n_seeds * K / (ms_e2e * 1e-3), "unit": "updates/s", "h2d_bytes_per_step": S * B * 8,
                        "d2h_bytes_per_step": S * 16, "ms_per_step": ms_e2e / K, "api": w.api,
                        "blocking_value": n_seeds * K / (ms_block * 1e-3), "blocking_ms_per_step": ms_block / K,
                        "pipelining": "value: the host enqueues step i+1 (index draw, gather, update) while step i runs, then "
                                      "waits for step i's event and reads its scalars (two mapped slots stamped with the step "
                                      "number, checked) -- the reference's loop never blocks on a step (rl_algorithm.py:159-165); "
                                      "blocking_value: stream synchronise + read after EVERY step before the next is issued",
                        "transfers": "indices: pinned host ring -> asynchronous H2D copy on a copy stream (overlaps the previous step); "
                                     "result: 4 scalars per seed stored by the step into mapped pinned host memory"},
                "gpu_launches": launches,
                "launches_per_step": e.launches_per_step + 1}
        if roof is not None:
            line["roofline"] = roof
        if batched is not None:
            line["batched_seeds"] = batched
        if variants is not None:
            line["variants"] = variants
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--algo", default="sac", choices=["sac", "poac", "goac"])
    ap.add_argument("--seeds-per-gpu", type=int, default=1,
                    help="headline workload: independent OAC seeds batched per GPU (1 = BASELINE config 2)")
    ap.add_argument("--total-seeds", type=int, default=64,
                    help="BASELINE config 5, measured next to the headline at every N: this many seeds in total, split "
                         "seed %% N over the GPUs (0: skip)")
    ap.add_argument("--gemm-path", default="auto", choices=["auto"] + list(GEMM_PATHS),
                    help="fp32: SIMT FFMA (reference-matching numerics, <=1e-5); tf32: TMA + tcgen05 kind::tf32 (<=1e-3); "
                         "tf32x3: tcgen05 3xTF32 (fp32-grade accuracy).  auto: fp32 for one seed per GPU (the step is "
                         "latency-bound there and the FFMA path is the fastest at reference numerics), tf32 for batched seeds")
    args = ap.parse_args()
    if args.gemm_path == "auto":
        args.gemm_path = "fp32" if args.seeds_per_gpu == 1 else "tf32"
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
