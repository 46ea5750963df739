"""Independent OAC seeds batched in one handle (BASELINE.json config 5) and sharded over GPUs.

The reference's only parallelism is one OS process per seed, pinned to GPU ``seed % n_gpus``
(main.py:575-576, reproduce_*.sh).  Here the seeds of one GPU share ONE engine: every stage of the
fused step is launched once with ``grid.z = n_seeds`` (grouped GEMMs over seeds x networks), each
seed keeping its own weights, Adam state, entropy temperature, step counters and noise stream.
Seeds never exchange data on the step path; ``allgather_stats`` (NCCL over NVLink in production,
gloo in the CPU tests) only collects the per-seed statistics vector for reporting.
"""
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from .engine import Engine, make_config
from .networks import get_policy_producer, get_q_producer

from .trainer import SACTrainer as _SACTrainer

# the per-seed statistics vector = SACTrainer's eval_statistics keys (trainer/trainer.py:243-279), reduced on the device
STAT_NAMES = tuple(_SACTrainer.STAT_KEYS)
N_STATS = len(STAT_NAMES)


def partition_seeds(n_seeds_total, rank, world_size):
    """Seed ids owned by ``rank``: the reference's ``seed % n_gpus`` placement (main.py:575-576)."""
    return [s for s in range(n_seeds_total) if s % world_size == rank]


class SACSeedGroup(object):
    """S independent SACTrainer instances (trainer/trainer.py:14-97) in one engine."""

    def __init__(self, seed_ids, obs_dim, act_dim, hidden=256, batch=256, gemm_path=_lib.GEMM_TF32,
                 discount=0.99, reward_scale=1.0, policy_lr=3e-4, qf_lr=3e-4, soft_target_tau=5e-3,
                 target_update_period=1, use_automatic_entropy_tuning=True, target_entropy=None,
                 stale_graph_mode="A", rng_seed=0):
        self.seed_ids = list(seed_ids)
        S = len(self.seed_ids)
        if S < 1:
            raise ValueError("empty seed group")
        cfg = make_config(_lib.ALGO_SAC, obs_dim, act_dim, hidden, batch, n_seeds=S,
                          auto_alpha=use_automatic_entropy_tuning, stale_graph_mode=stale_graph_mode,
                          target_update_period=target_update_period, gemm_path=gemm_path, discount=discount,
                          reward_scale=reward_scale, soft_target_tau=soft_target_tau, policy_lr=policy_lr,
                          qf_lr=qf_lr, target_entropy=target_entropy, rng_seed=rng_seed)
        self.engine = Engine(cfg)
        self.O, self.A, self.B = obs_dim, act_dim, batch
        pp = get_policy_producer(obs_dim, act_dim, [hidden, hidden])
        qp = get_q_producer(obs_dim, act_dim, [hidden, hidden])
        self.nets = []          # per seed: OrderedDict(policy, qf1, qf2, target_qf1, target_qf2)
        order = (("policy", 0), ("qf1", 1), ("qf2", 2), ("target_qf1", 4), ("target_qf2", 5))
        for s, sid in enumerate(self.seed_ids):
            torch.manual_seed(sid)                      # main.py:115: torch.manual_seed(seed) before construction
            objs = OrderedDict()
            for name, idx in order:                     # construction order of trainer/trainer.py:58-71
                net = pp() if name == "policy" else qp()
                net._bind(self.engine.net_views(idx, seed=s), self.engine.params[s], self.engine.net_layout(idx))
                objs[name] = net
            self.nets.append(objs)
        self._n_train_steps_total = 0
        # the step's Philox noise is keyed by the REAL seed id, not the slot: seeds keep their stream wherever they are
        # placed, and ranks holding different seeds never share one
        ids = torch.tensor(self.seed_ids, dtype=torch.int64)
        self.engine.counters[:, _lib.CNT_RNG_LO] = (ids & 0x7fffffff).to(torch.int32).to(self.engine.device)
        self.engine.counters[:, _lib.CNT_RNG_HI] = (ids >> 31).to(torch.int32).to(self.engine.device)
        # index upload: the host runs ahead of the step stream (a step takes longer than drawing the next indices), so the pinned
        # staging buffers AND their device copies form a ring; the host -> device copy of step i + 1 runs on a copy stream of its
        # own while step i computes (on the step stream it sat between two steps: ~10 us per step of a small group), and a ring
        # slot is only rewritten after the gather that read its device copy has completed
        self._idx_ring = [torch.zeros((S, batch), dtype=torch.int64).pin_memory() for _ in range(self._IDX_RING)]
        self._idx_dev_ring = [torch.zeros((S, batch), dtype=torch.int64, device=self.engine.device) for _ in range(self._IDX_RING)]
        self._idx_copied = [torch.cuda.Event() for _ in range(self._IDX_RING)]
        self._idx_consumed = [None] * self._IDX_RING
        self._idx_slot = 0
        self._copy_stream = torch.cuda.Stream(device=self.engine.device)

    _IDX_RING = 4

    @property
    def n_seeds(self):
        return len(self.seed_ids)

    def gather(self, replay, indices):
        """indices: [S, B] int64 (host numpy or device tensor) -> one gather launch for all seeds."""
        k = None
        if isinstance(indices, np.ndarray):
            k = self._idx_slot
            self._idx_slot = (k + 1) % self._IDX_RING
            if self._idx_consumed[k] is not None:
                self._idx_consumed[k].synchronize()         # the gather that read this slot's device copy is done
            host = self._idx_ring[k]
            host.numpy()[...] = indices
            main = torch.cuda.current_stream()
            if main.query():
                indices = host          # idle stream (a caller that synchronises every step): nothing to overlap a copy with --
                                        # the gather reads the pinned slot in place (unified addressing)
            else:
                with torch.cuda.stream(self._copy_stream):
                    self._idx_dev_ring[k].copy_(host, non_blocking=True)
                    self._idx_copied[k].record(self._copy_stream)
                main.wait_event(self._idx_copied[k])
                indices = self._idx_dev_ring[k]
        replay.gather_into(self.engine, indices, self.B, seed=0, n_seeds=self.n_seeds)
        if k is not None:
            ev = self._idx_consumed[k] or torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self._idx_consumed[k] = ev

    def load_batch(self, seed_slot, batch):
        f = lambda t: t.to(self.engine.device, torch.float32)
        self.engine.load_batch(f(batch['observations']), f(batch['actions']), f(batch['rewards']),
                               f(batch['terminals']), f(batch['next_observations']), seed=seed_slot)

    def inject_noise(self, seed_slot, eps_obs, eps_next):
        self.engine.set_eps(eps_obs.to(self.engine.device), eps_next.to(self.engine.device), seed=seed_slot)

    def step(self, external_eps=False):
        self.engine.step(external_eps=external_eps)
        self._n_train_steps_total += 1

    def stats(self):
        """[S, N_STATS] fp32 device tensor (STAT_NAMES: the reference's ``eval_statistics`` keys) of the last step,
        reduced by one kernel (``oac_trainer_stats``) -- the all-gather payload."""
        return self.engine.stats_device()

    def explorer(self, hyper_params, **kw):
        """One-launch exploration for all seeds' current observations (``optimistic_exploration.GroupExplorer``)."""
        from .optimistic_exploration import GroupExplorer
        return GroupExplorer(self, hyper_params, **kw)


class _AlgoSeedGroup(object):
    """Shared plumbing of the P-OAC / G-OAC seed groups: S independent trainers of one algorithm in one engine (every
    stage launched once over seeds x networks), per-seed weights / Adam state / counters / noise keys.  Subclasses build
    the per-seed network objects in the reference's construction (= torch RNG) order."""

    def __init__(self, cfg, seed_ids, obs_dim, act_dim, batch):
        self.seed_ids = list(seed_ids)
        if len(self.seed_ids) < 1:
            raise ValueError("empty seed group")
        self.engine = Engine(cfg)
        self.O, self.A, self.B = obs_dim, act_dim, batch
        ids = torch.tensor(self.seed_ids, dtype=torch.int64)
        self.engine.counters[:, _lib.CNT_RNG_LO] = (ids & 0x7fffffff).to(torch.int32).to(self.engine.device)
        self.engine.counters[:, _lib.CNT_RNG_HI] = (ids >> 31).to(torch.int32).to(self.engine.device)
        self.nets = []
        self._n_train_steps_total = 0
        self._uses_counts = bool(cfg.counts)
        self._with_tp = cfg.algo == _lib.ALGO_GOAC

    n_seeds = property(lambda self: len(self.seed_ids))

    def _bind(self, net, idx, s):
        net._bind(self.engine.net_views(idx, seed=s), self.engine.params[s], self.engine.net_layout(idx))
        return net

    def load_batch(self, seed_slot, batch):
        f = lambda t: t.to(self.engine.device, torch.float32)
        self.engine.load_batch(f(batch['observations']), f(batch['actions']), f(batch['rewards']), f(batch['terminals']),
                               f(batch['next_observations']),
                               f(batch['counts']) if (self._uses_counts and 'counts' in batch) else None,
                               seed=seed_slot, with_tp=self._with_tp)

    def inject_noise(self, seed_slot, eps_obs, eps_next):
        self.engine.set_eps(eps_obs.to(self.engine.device), eps_next.to(self.engine.device), seed=seed_slot)

    def step(self, external_eps=False):
        self.engine.step(external_eps=external_eps)
        self._n_train_steps_total += 1

    def stats(self):
        """[S, n_stats] device tensor: every seed's ``eval_statistics`` vector (key order: include/oac_b200.h)."""
        return self.engine.stats_device()


class ParticleSeedGroup(_AlgoSeedGroup):
    """S independent ``ParticleTrainer`` (P-OAC) instances in one engine (trainer/particle_trainer_oac.py:13-113; BASELINE
    config 3 batched over seeds).  ``nets[s]`` = dict(policy, qfs=[...], tfs=[...])."""

    def __init__(self, seed_ids, obs_dim, act_dim, hidden=256, batch=256, n_estimators=10, share_layers=True,
                 gemm_path=_lib.GEMM_FP32, discount=0.99, reward_scale=1.0, policy_lr=3e-4, qf_lr=3e-4, soft_target_tau=5e-3,
                 target_update_period=1, use_automatic_entropy_tuning=True, target_entropy=None, deterministic=False,
                 q_min=0.0, q_max=100.0, counts=False, train_bias=True, std_soft_update=False, std_soft_update_prob=0.0,
                 rng_seed=0):
        S = len(list(seed_ids))
        cfg = make_config(_lib.ALGO_POAC, obs_dim, act_dim, hidden, batch, n_seeds=S, n_particles=n_estimators,
                          share_layers=share_layers, deterministic=deterministic, auto_alpha=use_automatic_entropy_tuning,
                          counts=counts, train_bias=train_bias, target_update_period=target_update_period, gemm_path=gemm_path,
                          discount=discount, reward_scale=reward_scale, soft_target_tau=soft_target_tau, policy_lr=policy_lr,
                          qf_lr=qf_lr, target_entropy=target_entropy, rng_seed=rng_seed, std_soft_update=std_soft_update,
                          std_soft_update_prob=std_soft_update_prob)
        super().__init__(cfg, seed_ids, obs_dim, act_dim, batch)
        n = 1 if share_layers else n_estimators
        pp = get_policy_producer(obs_dim, act_dim, [hidden, hidden])
        qp = get_q_producer(obs_dim, act_dim, [hidden, hidden], output_size=n_estimators if share_layers else 1)
        init = np.linspace(q_min, q_max, n_estimators)                       # :75
        for s, sid in enumerate(self.seed_ids):
            torch.manual_seed(sid)
            policy = self._bind(pp(), 0, s)                                  # SACTrainer.__init__ part (:46-58): policy + 4 dropped critics
            for _ in range(4):
                qp()
            qfs, tfs = [], []
            for i in range(n):                                               # :99-113: (qf, tf) pairs
                b = init if share_layers else init[i]
                qfs.append(self._bind(qp(bias=b, train_bias=train_bias), 1 + i, s))
                tfs.append(self._bind(qp(bias=b, train_bias=train_bias), 2 + n + i, s))
            self.nets.append(dict(policy=policy, qfs=qfs, tfs=tfs))


class GaussianSeedGroup(_AlgoSeedGroup):
    """S independent ``GaussianTrainer`` (G-OAC) instances in one engine (trainer/gaussian_trainer.py:14-160; BASELINE
    config 4 batched over seeds).  ``nets[s]`` = dict(policy, target_policy, qfs=[q(, std)], tfs=[q_target(, std_target)])."""

    def __init__(self, seed_ids, obs_dim, act_dim, hidden=256, batch=256, share_layers=True, gemm_path=_lib.GEMM_FP32,
                 discount=0.99, reward_scale=1.0, policy_lr=3e-4, qf_lr=3e-4, std_lr=3e-5, soft_target_tau=5e-3,
                 target_update_period=1, delta=0.95, q_min=0.0, q_max=100.0, counts=False, train_bias=True, rng_seed=0):
        from .gaussian_trainer import norm_ppf
        S = len(list(seed_ids))
        mean, std = (q_max + q_min) / 2, (q_max - q_min) / np.sqrt(12)       # :70-72
        cfg = make_config(_lib.ALGO_GOAC, obs_dim, act_dim, hidden, batch, n_seeds=S, share_layers=share_layers,
                          deterministic=True, auto_alpha=False, counts=counts, train_bias=train_bias,
                          target_update_period=target_update_period, gemm_path=gemm_path, discount=discount,
                          reward_scale=reward_scale, soft_target_tau=soft_target_tau, policy_lr=policy_lr, qf_lr=qf_lr,
                          std_lr=std_lr, standard_bound=norm_ppf(delta), std_init=float(std), rng_seed=rng_seed)
        super().__init__(cfg, seed_ids, obs_dim, act_dim, batch)
        pp = get_policy_producer(obs_dim, act_dim, [hidden, hidden])
        qp = get_q_producer(obs_dim, act_dim, [hidden, hidden], output_size=2 if share_layers else 1)
        n = 1 if share_layers else 2
        for s, sid in enumerate(self.seed_ids):
            torch.manual_seed(sid)
            policy = self._bind(pp(), 0, s)
            for _ in range(4):
                qp()
            if share_layers:                                                 # :92-99
                b = np.array([mean, np.log(std)])
                q = self._bind(qp(bias=b, positive=[False, True], train_bias=train_bias), 2, s)
                qt = self._bind(qp(bias=b, positive=[False, True], train_bias=train_bias), 3 + n, s)
                qfs, tfs = [q], [qt]
            else:                                                            # :100-112
                q = self._bind(qp(bias=mean), 2, s)
                qt = self._bind(qp(bias=mean), 3 + n, s)
                sd = self._bind(qp(bias=np.log(std), positive=True, train_bias=train_bias), 3, s)
                sdt = self._bind(qp(bias=np.log(std), positive=True, train_bias=train_bias), 4 + n, s)
                qfs, tfs = [q, sd], [qt, sdt]
            target_policy = self._bind(pp(), 1, s)                           # :143
            self.nets.append(dict(policy=policy, target_policy=target_policy, qfs=qfs, tfs=tfs))


def allgather_stats(local_stats, seed_ids, n_seeds_total, group=None):
    """Collects every rank's [S_local, n] statistics into one [n_seeds_total, n] tensor ordered by seed id.
    The only collective of the design; runs once per reporting interval, never inside a step."""
    import torch.distributed as dist
    n = local_stats.shape[1]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        out = torch.zeros((n_seeds_total, n), dtype=local_stats.dtype, device=local_stats.device)
        out[torch.as_tensor(list(seed_ids), device=local_stats.device)] = local_stats
        return out
    world = dist.get_world_size(group)
    per = (n_seeds_total + world - 1) // world
    pad = torch.zeros((per, n + 1), dtype=local_stats.dtype, device=local_stats.device)
    pad[:, n] = -1
    pad[:len(seed_ids), :n] = local_stats
    pad[:len(seed_ids), n] = torch.as_tensor(list(seed_ids), dtype=local_stats.dtype, device=local_stats.device)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    out = torch.zeros((n_seeds_total, n), dtype=local_stats.dtype, device=local_stats.device)
    for b in bufs:
        ids = b[:, n]
        keep = ids >= 0
        out[ids[keep].long()] = b[keep, :n]
    return out
