"""SACTrainer: drop-in for the reference's ``trainer/trainer.py`` (the live OAC/SAC trainer,
``main.py:14``), backed by the fused sm_100a step (``oac_trainer_step``).

Same constructor, attributes (``policy``, ``qf1`` .. ``target_qf2``, ``qfs``, ``tfs``,
``log_alpha``, ``networks``, ``deterministic``) and methods (``train``,
``train_from_torch``, ``predict``, ``get_diagnostics``, ``end_epoch``, ``get_snapshot``,
``restore_from_snapshot``) as trainer/trainer.py:14-369.  Numerics follow the reference's
torch-1.4 behaviour ("mode A") unless ``stale_graph_mode='B'``.
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.optim as optim

from . import _lib
from .engine import Engine, make_config
from .networks import from_numpy, _device


def create_stats_ordered_dict(name, data):
    """utils/eval_util.py:69-113 for ndarray data (the trainer's use)."""
    return OrderedDict([(name + ' Mean', np.mean(data)), (name + ' Std', np.std(data)),
                        (name + ' Max', np.max(data)), (name + ' Min', np.min(data))])


def get_numpy(t):
    return t.to('cpu').detach().numpy()


def np_to_pytorch_batch(np_batch):
    """utils/core.py:40-61: float32 device tensors; bool -> int; object arrays dropped."""
    out = {}
    for k, v in np_batch.items():
        if isinstance(v, torch.Tensor):
            out[k] = v
            continue
        if not isinstance(v, np.ndarray) or v.dtype == np.dtype('O'):
            continue
        if v.dtype == np.bool_:
            v = v.astype(int)
        out[k] = from_numpy(v)
    return out


class _AdamHandle(object):
    """``torch.optim.Adam``-shaped view (state_dict / load_state_dict) of the fused Adam state of
    one optimizer: the moments live in the engine's Adam arenas, the step count in its counters."""

    def __init__(self, trainer, net_index, counter, lr):
        self._t, self._net, self._counter, self.lr = trainer, net_index, counter, lr

    def _views(self):
        e = self._t._engine
        return (e.net_views(self._net), e.net_views(self._net, arena=e.adam_m),
                e.net_views(self._net, arena=e.adam_v))

    def state_dict(self):
        p, m, v = self._views()
        step = int(self._t._engine.counters[0, _lib.CNT_OPT0 + self._counter].item())
        state = {}
        for i, k in enumerate(p.keys()):
            if step > 0:
                state[i] = {'step': step, 'exp_avg': m[k].detach().clone(), 'exp_avg_sq': v[k].detach().clone()}
        group = dict(lr=self.lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False,
                     params=list(range(len(p))))
        return {'state': state, 'param_groups': [group]}

    def load_state_dict(self, sd):
        p, m, v = self._views()
        step = 0
        for i, k in enumerate(p.keys()):
            st = sd['state'].get(i)
            if st is None:
                m[k].zero_(), v[k].zero_()
                continue
            m[k].copy_(torch.as_tensor(st['exp_avg']).to(m[k].device).reshape(m[k].shape))
            v[k].copy_(torch.as_tensor(st['exp_avg_sq']).to(v[k].device).reshape(v[k].shape))
            step = max(step, int(st['step']))
        self._t._engine.counters[0, _lib.CNT_OPT0 + self._counter] = step

    def zero_grad(self):
        pass


class _EngineTrainer(object):
    """Shared plumbing of the three trainers: engine (re)creation, batch upload, noise injection."""
    ALGO = _lib.ALGO_SAC
    DEFAULT_BATCH = 256

    def _engine_kwargs(self):
        raise NotImplementedError

    def _net_objects(self):
        """[(net object, layout index)] in arena order."""
        raise NotImplementedError

    def _make_engine(self, batch):
        kw = self._engine_kwargs()
        cfg = make_config(self.ALGO, self._O, self._A, self._H, batch, **kw)
        new = Engine(cfg)
        old = getattr(self, '_engine', None)
        for net, idx in self._net_objects():
            net._bind(new.net_views(idx), new.params[0], new.net_layout(idx))
        if old is not None:
            new.adam_m.copy_(old.adam_m), new.adam_v.copy_(old.adam_v), new.counters.copy_(old.counters)
            new.steps = old.steps
            la_old, la_new = old.net_views(self._log_alpha_index), new.net_views(self._log_alpha_index)
            la_new['log_alpha'].copy_(la_old['log_alpha'])
        self._engine = new
        self.log_alpha = new.net_views(self._log_alpha_index)['log_alpha']
        buf = getattr(self, '_attached_buffer', None)
        if buf is not None:
            buf.attach(self)
        self._pending_eps = False

    def _ensure_engine(self, batch):
        if self._engine.B != batch:
            self._make_engine(batch)

    # ---- reference API -------------------------------------------------------------
    def train(self, np_batch):
        """trainer/trainer.py:99-103."""
        buffer = np_batch.pop('buffer', None)
        if buffer is not None and hasattr(buffer, 'attach') and getattr(self, '_attached_buffer', None) is not buffer:
            # from the next random_batch on, the buffer gathers straight into our IO slice
            self._attached_buffer = buffer
            buffer.attach(self)
        if np_batch.get('_oac_resident') is self._engine:
            np_batch['buffer'] = buffer
            self._step()
            return
        batch = np_to_pytorch_batch({k: v for k, v in np_batch.items() if not k.startswith('_oac')})
        batch['buffer'] = buffer
        self.train_from_torch(batch)

    def train_from_torch(self, batch):
        obs = batch['observations']
        self._ensure_engine(obs.shape[0])
        dev = self._engine.device
        f = lambda t: t.to(dev, torch.float32)
        self._engine.load_batch(f(obs), f(batch['actions']), f(batch['rewards']), f(batch['terminals']),
                                f(batch['next_observations']),
                                f(batch['counts']) if ('counts' in batch and self._uses_counts) else None,
                                with_tp=self.ALGO == _lib.ALGO_GOAC)
        self._step()

    def inject_noise(self, eps_obs=None, eps_next=None):
        """Parity hook: the next step reads its N(0,1) draws (TanhNormal.rsample,
        trainer/policies.py:179-187) from these [B,A] tensors instead of the device Philox stream."""
        b = (eps_obs if eps_obs is not None else eps_next).shape[0]
        self._ensure_engine(b)
        dev = self._engine.device
        self._engine.set_eps(None if eps_obs is None else eps_obs.to(dev, torch.float32),
                             None if eps_next is None else eps_next.to(dev, torch.float32))
        self._pending_eps = True

    def _step(self):
        self._engine.step(external_eps=self._pending_eps)
        self._pending_eps = False
        if self._need_to_update_eval_statistics:
            self._need_to_update_eval_statistics = False
            self._update_eval_statistics()
        self._n_train_steps_total += 1

    def get_diagnostics(self):
        return self.eval_statistics

    def end_epoch(self, epoch):
        self._need_to_update_eval_statistics = True

    # ---- io helpers ------------------------------------------------------------------
    def _io(self, name, shape):
        return self._engine.io_view(getattr(self._engine.lay, name), shape)


class SACTrainer(_EngineTrainer):
    ALGO = _lib.ALGO_SAC

    def __init__(self, policy_producer, q_producer, action_space=None, discount=0.99, reward_scale=1.0,
                 policy_lr=1e-3, qf_lr=1e-3, optimizer_class=optim.Adam, soft_target_tau=1e-2,
                 target_update_period=1, use_automatic_entropy_tuning=True, target_entropy=None,
                 deterministic=False, stale_graph_mode="A", rng_seed=None, gemm_path=_lib.GEMM_FP32):
        if optimizer_class is not optim.Adam:
            raise NotImplementedError("the fused step implements torch.optim.Adam (the reference's only choice)")
        self.use_automatic_entropy_tuning = use_automatic_entropy_tuning
        self.target_entropy = None
        if self.use_automatic_entropy_tuning:
            self.target_entropy = target_entropy if target_entropy else -np.prod(action_space.shape).item()
        self.soft_target_tau, self.target_update_period = soft_target_tau, target_update_period
        self.deterministic = deterministic
        self.discount, self.reward_scale = discount, reward_scale
        self.policy_lr, self.qf_lr = policy_lr, qf_lr
        self.stale_graph_mode, self.gemm_path = stale_graph_mode, gemm_path
        self._rng_seed = int(torch.initial_seed() & 0x7fffffffffffffff) if rng_seed is None else rng_seed
        # construction order = RNG order of trainer/trainer.py:58-71
        self.policy = policy_producer()
        self.qf1, self.qf2 = q_producer(), q_producer()
        self.target_qf1, self.target_qf2 = q_producer(), q_producer()
        self._O, self._H = self.policy.input_size, self.policy.hidden
        self._A = self.policy.action_dim
        self._uses_counts = False
        self._log_alpha_index = 3
        self._engine = None
        self._make_engine(self.DEFAULT_BATCH)
        if self.use_automatic_entropy_tuning:
            self.alpha_optimizer = _AdamHandle(self, 3, 3, policy_lr)
        self.policy_optimizer = _AdamHandle(self, 0, 0, policy_lr)
        self.qf1_optimizer = _AdamHandle(self, 1, 1, qf_lr)
        self.qf2_optimizer = _AdamHandle(self, 2, 2, qf_lr)
        self.eval_statistics = OrderedDict()
        self._n_train_steps_total = 0
        self._need_to_update_eval_statistics = True
        self.qfs = [self.qf1, self.qf2]
        self.tfs = [self.target_qf1, self.target_qf2]

    def _engine_kwargs(self):
        return dict(deterministic=self.deterministic, auto_alpha=self.use_automatic_entropy_tuning,
                    stale_graph_mode=self.stale_graph_mode, target_update_period=self.target_update_period,
                    gemm_path=self.gemm_path, discount=self.discount, reward_scale=self.reward_scale,
                    soft_target_tau=self.soft_target_tau, policy_lr=self.policy_lr, qf_lr=self.qf_lr,
                    target_entropy=self.target_entropy, rng_seed=self._rng_seed)

    def _net_objects(self):
        # layout order: policy, qf1, qf2, log_alpha | target_qf1, target_qf2
        return [(self.policy, 0), (self.qf1, 1), (self.qf2, 2), (self.target_qf1, 4), (self.target_qf2, 5)]

    def predict(self, obs, action, upper_bound=True, beta_UB=4.46, both_values=False):
        """trainer/trainer.py:105-123."""
        if isinstance(obs, np.ndarray):
            obs, action = from_numpy(obs), from_numpy(action)
        if obs.dim() == 1:
            obs, action = obs[None], action[None]
        Q1, Q2 = self.qfs[0](obs, action), self.qfs[1](obs, action)
        mu_Q = (Q1 + Q2) / 2.0
        sigma_Q = torch.abs(Q1 - Q2) / 2.0
        if both_values:
            return mu_Q, sigma_Q
        if not upper_bound:
            return mu_Q
        return mu_Q + beta_UB * sigma_Q

    STAT_KEYS = (['QF mean', 'QF std', 'QF1 Loss', 'QF2 Loss', 'Q Loss', 'Policy Loss'] +
                 [n + s for n in ('Q1 Predictions', 'Q2 Predictions', 'Q Targets', 'Log Pis', 'Policy mu', 'Policy log std')
                  for s in (' Mean', ' Std', ' Max', ' Min')] + ['Alpha', 'Alpha Loss'])

    def _update_eval_statistics(self):
        """Same keys, order and definitions as trainer/trainer.py:230-279, reduced on the device by ONE kernel
        (``oac_trainer_stats``) straight into mapped host memory instead of seven device-to-host copies + numpy."""
        vec = self._engine.stats_host()[0]
        st = self.eval_statistics
        for k, v in zip(self.STAT_KEYS, vec):
            if k.startswith('Alpha') and not self.use_automatic_entropy_tuning:
                continue
            st[k] = v

    @property
    def networks(self):
        return [self.policy, self.qf1, self.qf2, self.target_qf1, self.target_qf2]

    def get_snapshot(self):
        """trainer/trainer.py:299-317 (same keys)."""
        snapshot = dict(
            policy_state_dict=self.policy.state_dict(),
            policy_optim_state_dict=self.policy_optimizer.state_dict(),
            qf1_state_dict=self.qf1.state_dict(), qf1_optim_state_dict=self.qf1_optimizer.state_dict(),
            target_qf1_state_dict=self.target_qf1.state_dict(),
            qf2_state_dict=self.qf2.state_dict(), qf2_optim_state_dict=self.qf2_optimizer.state_dict(),
            target_qf2_state_dict=self.target_qf2.state_dict(),
            eval_statistics=self.eval_statistics, _n_train_steps_total=self._n_train_steps_total,
            _need_to_update_eval_statistics=self._need_to_update_eval_statistics)
        if self.use_automatic_entropy_tuning:
            snapshot['log_alpha'] = self.log_alpha
            snapshot['alpha_optim_state_dict'] = self.alpha_optimizer.state_dict()
        return snapshot

    def restore_from_snapshot(self, ss):
        """trainer/trainer.py:334-369."""
        self.policy.load_state_dict(ss['policy_state_dict'])
        self.policy_optimizer.load_state_dict(ss['policy_optim_state_dict'])
        self.qf1.load_state_dict(ss['qf1_state_dict'])
        self.qf1_optimizer.load_state_dict(ss['qf1_optim_state_dict'])
        self.target_qf1.load_state_dict(ss['target_qf1_state_dict'])
        self.qf2.load_state_dict(ss['qf2_state_dict'])
        self.qf2_optimizer.load_state_dict(ss['qf2_optim_state_dict'])
        self.target_qf2.load_state_dict(ss['target_qf2_state_dict'])
        if self.use_automatic_entropy_tuning:
            self.log_alpha.copy_(torch.as_tensor(ss['log_alpha']).to(self.log_alpha.device).reshape(1))
            self.alpha_optimizer.load_state_dict(ss['alpha_optim_state_dict'])
        self.eval_statistics = ss['eval_statistics']
        self._n_train_steps_total = ss['_n_train_steps_total']
        self._engine.set_train_steps(self._n_train_steps_total)
        self._need_to_update_eval_statistics = ss['_need_to_update_eval_statistics']
