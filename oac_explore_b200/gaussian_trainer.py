"""GaussianTrainer (G-OAC): drop-in for the reference's ``trainer/gaussian_trainer.py``, backed by
the fused sm_100a step.  Covers what main.py wires for ``--alg g-oac``: deterministic policy
(``a = tanh(mean)``), a mean critic and a std critic (two heads of one trunk with
``share_layers``, else two nets with their own ``std_lr``), clamp of the std target to
``[0, (q_max-q_min)/sqrt(12)]``, optional ``counts``, the upper-bound policy loss
``-(Q + Phi^-1(delta) * sigma)`` and the extra target policy trained on the mean (:177-388)."""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.optim as optim

from . import _lib
from .networks import from_numpy
from .trainer import _EngineTrainer, _AdamHandle, create_stats_ordered_dict, get_numpy


def norm_ppf(p):
    """scipy.stats.norm.ppf(p) (gaussian_trainer.py:68) without the scipy import."""
    return float(math.sqrt(2.0) * torch.erfinv(torch.tensor(2.0 * p - 1.0, dtype=torch.float64)))


class GaussianTrainer(_EngineTrainer):
    ALGO = _lib.ALGO_GOAC

    def __init__(self, policy_producer, q_producer, n_estimators=2, action_space=None, discount=0.99,
                 reward_scale=1.0, delta=0.95, policy_lr=1e-3, qf_lr=3e-4, std_lr=3e-5,
                 optimizer_class=optim.Adam, soft_target_tau=1e-2, target_update_period=1,
                 use_automatic_entropy_tuning=False, target_entropy=None, deterministic=True, q_min=0,
                 q_max=100, pac=False, ensemble=False, n_policies=1, share_layers=False, r_mellow_max=1.,
                 b_mellow_max=None, mellow_max=False, counts=False, mean_update=False, global_opt=False,
                 std_soft_update=False, std_soft_update_prob=0., train_bias=True, use_target_policy=False,
                 rescale_targets_around_mean=False, rng_seed=None, gemm_path=_lib.GEMM_FP32):
        self.gemm_path = gemm_path
        if optimizer_class is not optim.Adam:
            raise NotImplementedError("the fused step implements torch.optim.Adam")
        if ensemble or global_opt or std_soft_update or mean_update or use_target_policy or not deterministic:
            raise NotImplementedError("only the g-oac configuration main.py wires (deterministic policy, no "
                                      "ensemble / global_opt / std_soft_update / mean_update) is on the hot path")
        self.use_automatic_entropy_tuning = use_automatic_entropy_tuning
        self.target_entropy = None
        if use_automatic_entropy_tuning:
            self.target_entropy = target_entropy if target_entropy else -np.prod(action_space.shape).item()
        self.soft_target_tau, self.target_update_period = soft_target_tau, target_update_period
        self.deterministic = deterministic
        self.discount, self.reward_scale = discount, reward_scale
        self.policy_lr, self.qf_lr, self.std_lr = policy_lr, qf_lr, std_lr
        self._rng_seed = 0 if rng_seed is None else rng_seed
        self.policy = policy_producer()                       # SACTrainer.__init__ (trainer/trainer.py:58-71)
        for _ in range(4):
            q_producer()
        self.action_space = action_space
        self.q_min, self.q_max = q_min, q_max
        self.standard_bound = norm_ppf(delta)                 # :68
        self.share_layers = share_layers
        mean = (q_max + q_min) / 2
        std = (q_max - q_min) / np.sqrt(12)
        log_std = np.log(std)
        self.delta, self.std_init, self.n_estimators = delta, std, n_estimators
        self.counts, self.train_bias = counts, train_bias
        if share_layers:                                       # :92-99
            self.q = q_producer(bias=np.array([mean, log_std]), positive=[False, True], train_bias=train_bias)
            self.q_target = q_producer(bias=np.array([mean, log_std]), positive=[False, True], train_bias=train_bias)
            self.qfs, self.tfs = [self.q], [self.q_target]
            if self.q.output_size != 2:
                raise ValueError("share_layers needs a q_producer with output_size=2")
        else:                                                  # :100-112
            self.q, self.q_target = q_producer(bias=mean), q_producer(bias=mean)
            self.std = q_producer(bias=log_std, positive=True, train_bias=train_bias)
            self.std_target = q_producer(bias=log_std, positive=True, train_bias=train_bias)
            self.qfs, self.tfs = [self.q, self.std], [self.q_target, self.std_target]
        self.target_policy = policy_producer()                # :143
        self._O, self._H, self._A = self.policy.input_size, self.policy.hidden, self.policy.action_dim
        self._uses_counts = counts
        n = len(self.qfs)
        self._log_alpha_index = 2 + n
        self._engine = None
        self._make_engine(self.DEFAULT_BATCH)
        self.policy_optimizer = _AdamHandle(self, 0, 0, policy_lr)
        self.target_policy_optimizer = _AdamHandle(self, 1, 1, policy_lr)
        self.q_optimizer = _AdamHandle(self, 2, 2, qf_lr)
        if not share_layers:
            self.std_optimizer = _AdamHandle(self, 3, 3, std_lr)
        self.alpha_optimizer = _AdamHandle(self, 2 + n, 15, policy_lr)   # created, never stepped (:456-457)
        self.eval_statistics = OrderedDict()
        self._n_train_steps_total = 0
        self._need_to_update_eval_statistics = True

    def _engine_kwargs(self):
        return dict(share_layers=self.share_layers, deterministic=True, auto_alpha=False, counts=self.counts,
                    train_bias=self.train_bias, target_update_period=self.target_update_period,
                    discount=self.discount, reward_scale=self.reward_scale,
                    soft_target_tau=self.soft_target_tau, policy_lr=self.policy_lr, qf_lr=self.qf_lr,
                    std_lr=self.std_lr, standard_bound=self.standard_bound, std_init=float(self.std_init),
                    rng_seed=self._rng_seed, gemm_path=self.gemm_path)

    def _net_objects(self):
        # layout order: policy, target_policy, q, [std], log_alpha | q_target, [std_target]
        n = len(self.qfs)
        return ([(self.policy, 0), (self.target_policy, 1)] + [(q, 2 + i) for i, q in enumerate(self.qfs)] +
                [(t, 3 + n + i) for i, t in enumerate(self.tfs)])

    def predict(self, obs, action, std=True):
        """:146-159."""
        obs, action = from_numpy(np.array(obs)), (action if isinstance(action, torch.Tensor) else from_numpy(np.array(action)))
        qs = self.q(obs, action)
        if self.share_layers:
            stds, qs = qs[:, 1].unsqueeze(-1), qs[:, 0].unsqueeze(-1)
        else:
            stds = self.std(obs, action)
        upper_bound = qs + self.standard_bound * stds
        if std:
            return [qs, stds], upper_bound
        return upper_bound

    STAT_KEYS = (['QF mean', 'QF std', 'QF Loss'] +
                 [n + s for n in ('Q Predictions', 'Q Target') for s in (' Mean', ' Std', ' Max', ' Min')] + ['STD Loss'] +
                 [n + s for n in ('Q STD Predictions', 'Q STD Target') for s in (' Mean', ' Std', ' Max', ' Min')] +
                 ['Policy Loss'] +
                 [n + s for n in ('Policy mu', 'Policy log std') for s in (' Mean', ' Std', ' Max', ' Min')])

    def _update_eval_statistics(self):
        """Keys of :397-436, reduced on the device (``oac_trainer_stats``); ``Policy mu`` / ``Policy log std`` are the
        target policy's outputs like in the reference (:361-363)."""
        vec = self._engine.stats_host()[0]
        st = self.eval_statistics
        for k, v in zip(self.STAT_KEYS, vec):
            st[k] = v

    @property
    def networks(self):
        return [self.policy] + self.qfs + self.tfs + [self.target_policy]

    def get_snapshot(self):
        """:450-481 (same keys)."""
        data = dict(policy_state_dict=self.policy.state_dict(),
                    policy_optim_state_dict=self.policy_optimizer.state_dict(),
                    log_alpha=self.log_alpha, alpha_optim_state_dict=self.alpha_optimizer.state_dict(),
                    eval_statistics=self.eval_statistics, _n_train_steps_total=self._n_train_steps_total,
                    _need_to_update_eval_statistics=self._need_to_update_eval_statistics)
        opts = [self.q_optimizer] + ([] if self.share_layers else [self.std_optimizer])
        data["qfs_state_dicts"] = [q.state_dict() for q in self.qfs]
        data["qfs_optims_state_dicts"] = [o.state_dict() for o in opts]
        data["target_qfs_state_dicts"] = [t.state_dict() for t in self.tfs]
        data["target_policy_state_dict"] = self.target_policy.state_dict()
        data["target_policy_opt_state_dict"] = self.target_policy_optimizer.state_dict()
        return data

    def restore_from_snapshot(self, ss):
        """:483-514."""
        self.policy.load_state_dict(ss['policy_state_dict'])
        self.policy_optimizer.load_state_dict(ss['policy_optim_state_dict'])
        opts = [self.q_optimizer] + ([] if self.share_layers else [self.std_optimizer])
        for i in range(len(self.qfs)):
            self.qfs[i].load_state_dict(ss['qfs_state_dicts'][i])
            opts[i].load_state_dict(ss['qfs_optims_state_dicts'][i])
            self.tfs[i].load_state_dict(ss['target_qfs_state_dicts'][i])
        self.log_alpha.copy_(torch.as_tensor(ss['log_alpha']).to(self.log_alpha.device).reshape(1))
        self.eval_statistics = ss['eval_statistics']
        self._n_train_steps_total = ss['_n_train_steps_total']
        self._engine.set_train_steps(self._n_train_steps_total)
        self._need_to_update_eval_statistics = ss['_need_to_update_eval_statistics']
        self.target_policy.load_state_dict(ss["target_policy_state_dict"])
        self.target_policy_optimizer.load_state_dict(ss["target_policy_opt_state_dict"])
