"""ParticleTrainer (P-OAC with OAC exploration): drop-in for the reference's
``trainer/particle_trainer_oac.py`` (imported as ``ParticleTrainerOAC`` at main.py:20),
backed by the fused sm_100a step.  Covers the configuration main.py wires for ``--alg p-oac``
with optimistic exploration: P Q-particles as P heads of one trunk (``share_layers``) or P
separate nets, sorted-particle regression targets, optional ``counts`` re-centering, policy
loss on the lowest particle (:169-324)."""
from collections import OrderedDict

import numpy as np
import torch
import torch.optim as optim

from . import _lib
from .networks import from_numpy
from .trainer import _EngineTrainer, _AdamHandle, create_stats_ordered_dict, get_numpy


class ParticleTrainer(_EngineTrainer):
    ALGO = _lib.ALGO_POAC

    def __init__(self, policy_producer, q_producer, n_estimators=2, action_space=None, discount=0.99,
                 reward_scale=1.0, delta=0.95, policy_lr=1e-3, qf_lr=1e-3, optimizer_class=optim.Adam,
                 soft_target_tau=1e-2, target_update_period=1, use_automatic_entropy_tuning=True,
                 target_entropy=None, deterministic=True, q_min=0, q_max=100, ensemble=False, n_policies=1,
                 share_layers=False, r_mellow_max=1., b_mellow_max=None, mellow_max=False, counts=False,
                 mean_update=False, global_opt=False, std_soft_update=False, std_soft_update_prob=0.,
                 train_bias=True, lb=0.1, rng_seed=None, gemm_path=_lib.GEMM_FP32):
        self.gemm_path = gemm_path
        if optimizer_class is not optim.Adam:
            raise NotImplementedError("the fused step implements torch.optim.Adam")
        if ensemble or global_opt or mean_update or mellow_max:
            raise NotImplementedError("ensemble / global_opt / mean_update variants are "
                                      "outside the P-OAC hot path (SURVEY.md section 8f rank 4)")
        assert not counts or not std_soft_update                     # :97
        self.std_soft_update, self.std_soft_update_prob = std_soft_update, std_soft_update_prob
        self.use_automatic_entropy_tuning = use_automatic_entropy_tuning
        self.target_entropy = None
        if use_automatic_entropy_tuning:
            self.target_entropy = target_entropy if target_entropy else -np.prod(action_space.shape).item()
        self.soft_target_tau, self.target_update_period = soft_target_tau, target_update_period
        self.deterministic = deterministic
        self.discount, self.reward_scale = discount, reward_scale
        self.policy_lr, self.qf_lr = policy_lr, qf_lr
        self._rng_seed = int(torch.initial_seed() & 0x7fffffffffffffff) if rng_seed is None else rng_seed
        # SACTrainer.__init__ part (trainer/trainer.py:58-71): policy + four critics that are dropped
        self.policy = policy_producer()
        for _ in range(4):
            q_producer()
        quantiles = [i * 1. / (n_estimators - 1) for i in range(n_estimators)]     # :60-74
        self.delta_index = next(p for p in range(n_estimators) if quantiles[p] >= delta)
        self.lb_index = next(p for p in range(n_estimators) if quantiles[p] >= lb)
        initial_values = np.linspace(q_min, q_max, n_estimators)
        self.share_layers, self.num_particles = share_layers, n_estimators
        self.n_estimators = 1 if share_layers else n_estimators
        self.q_min, self.q_max, self.delta = q_min, q_max, delta
        self.counts, self.train_bias = counts, train_bias
        self.action_space = action_space
        self.qfs, self.tfs = [], []
        for i in range(self.n_estimators):                                          # :99-113
            b = initial_values if share_layers else initial_values[i]
            self.qfs.append(q_producer(bias=b, train_bias=train_bias))
            self.tfs.append(q_producer(bias=b, train_bias=train_bias))
        if self.qfs[0].output_size != (n_estimators if share_layers else 1):
            raise ValueError("q_producer output_size does not match n_estimators / share_layers")
        self._O, self._H, self._A = self.policy.input_size, self.policy.hidden, self.policy.action_dim
        self._uses_counts = counts
        n = self.n_estimators
        self._log_alpha_index = 1 + n
        self._engine = None
        self._make_engine(self.DEFAULT_BATCH)
        self.policy_optimizer = _AdamHandle(self, 0, 0, policy_lr)
        if use_automatic_entropy_tuning:
            self.alpha_optimizer = _AdamHandle(self, 1 + n, 1, policy_lr)
        self.qf_optimizers = [_AdamHandle(self, 1 + i, 2 + i, qf_lr) for i in range(n)]
        self.eval_statistics = OrderedDict()
        self._n_train_steps_total = 0
        self._need_to_update_eval_statistics = True

    def _engine_kwargs(self):
        return dict(n_particles=self.num_particles, share_layers=self.share_layers,
                    deterministic=self.deterministic, auto_alpha=self.use_automatic_entropy_tuning,
                    counts=self.counts, train_bias=self.train_bias,
                    target_update_period=self.target_update_period, discount=self.discount,
                    reward_scale=self.reward_scale, soft_target_tau=self.soft_target_tau,
                    policy_lr=self.policy_lr, qf_lr=self.qf_lr, target_entropy=self.target_entropy,
                    rng_seed=self._rng_seed, gemm_path=self.gemm_path, std_soft_update=self.std_soft_update,
                    std_soft_update_prob=self.std_soft_update_prob)

    def _net_objects(self):
        # layout order: policy, qf[0..n), log_alpha | tf[0..n)
        n = self.n_estimators
        return ([(self.policy, 0)] + [(q, 1 + i) for i, q in enumerate(self.qfs)] +
                [(t, 2 + n + i) for i, t in enumerate(self.tfs)])

    def predict(self, obs, action, all_particles=False, upper_bound=True, beta_UB=None):
        """:147-167."""
        if not isinstance(obs, torch.Tensor):
            obs, action = from_numpy(np.array(obs)), from_numpy(np.array(action))
        qs = torch.stack([q(obs, action) for q in self.qfs], dim=0)
        if self.share_layers:
            qs = qs.permute(2, 1, 0)
        sorted_qs = torch.sort(qs, dim=0)[0]
        out = sorted_qs[self.delta_index] if upper_bound else torch.mean(qs, dim=0)
        if all_particles:
            return sorted_qs, out
        return out

    def _stat_keys(self):
        keys = ['QF mean', 'QF std']
        for i in range(self.num_particles):
            keys.append('QF' + str(i) + ' Loss')
            keys += ['Q' + str(i) + 'Predictions' + s for s in (' Mean', ' Std', ' Max', ' Min')]
            keys += ['Q' + str(i) + 'Targets' + s for s in (' Mean', ' Std', ' Max', ' Min')]
        keys.append('Policy Loss')
        keys += [n + s for n in ('Policy mu', 'Policy log std') for s in (' Mean', ' Std', ' Max', ' Min')]
        return keys

    def _update_eval_statistics(self):
        """Keys of :332-362 that derive from the step's tensors, reduced on the device (``oac_trainer_stats``)."""
        vec = self._engine.stats_host()[0]
        st = self.eval_statistics
        for k, v in zip(self._stat_keys(), vec):
            st[k] = v

    @property
    def networks(self):
        return [self.policy] + self.qfs + self.tfs

    def get_snapshot(self):
        """:374-400 (same keys)."""
        data = dict(policy_state_dict=self.policy.state_dict(),
                    policy_optim_state_dict=self.policy_optimizer.state_dict(),
                    eval_statistics=self.eval_statistics, _n_train_steps_total=self._n_train_steps_total,
                    _need_to_update_eval_statistics=self._need_to_update_eval_statistics)
        if self.use_automatic_entropy_tuning:
            data['log_alpha'] = self.log_alpha
            data['alpha_optim_state_dict'] = self.alpha_optimizer.state_dict()
        data["qfs_state_dicts"] = [q.state_dict() for q in self.qfs]
        data["qfs_optims_state_dicts"] = [o.state_dict() for o in self.qf_optimizers]
        data["target_qfs_state_dicts"] = [t.state_dict() for t in self.tfs]
        return data

    def restore_from_snapshot(self, ss):
        self.policy.load_state_dict(ss['policy_state_dict'])
        self.policy_optimizer.load_state_dict(ss['policy_optim_state_dict'])
        for i in range(len(self.qfs)):
            self.qfs[i].load_state_dict(ss['qfs_state_dicts'][i])
            self.qf_optimizers[i].load_state_dict(ss['qfs_optims_state_dicts'][i])
            self.tfs[i].load_state_dict(ss['target_qfs_state_dicts'][i])
        if self.use_automatic_entropy_tuning and 'log_alpha' in ss:
            self.log_alpha.copy_(torch.as_tensor(ss['log_alpha']).to(self.log_alpha.device).reshape(1))
            self.alpha_optimizer.load_state_dict(ss['alpha_optim_state_dict'])
        self.eval_statistics = ss['eval_statistics']
        self._n_train_steps_total = ss['_n_train_steps_total']
        self._engine.set_train_steps(self._n_train_steps_total)
        self._need_to_update_eval_statistics = ss['_need_to_update_eval_statistics']


ParticleTrainerOAC = ParticleTrainer
