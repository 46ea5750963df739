"""Drop-in for the reference's ``optimistic_exploration.py`` (:7-196).

``get_optimistic_exploration_action(ob_np, policy=, qfs=, trainer=, hyper_params=, deterministic=)``
returns ``(np.float32[A], {})`` like the reference, computed by ONE fused kernel
(``oac_explore``): policy forward, critic forwards, closed-form gradient of
Q_UB = mu_Q + beta_UB * sigma_Q with respect to the pre-tanh mean, the KL-constrained
shift ``sqrt(2 delta) Sigma g / (sqrt(g^T Sigma g) + 1e-5)`` and the final TanhNormal sample.
The observation goes up and the action comes back through pinned host buffers.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import OacExploreArgs
from .networks import MakeDeterministic, _device

_STATE = {}


def _buffers(O, A, n):
    key = (O, A, n, torch.cuda.current_device())
    st = _STATE.get(key)
    if st is None:
        dev = _device()
        st = dict(ob_host=torch.zeros((n, O), dtype=torch.float32).pin_memory(),
                  ob_dev=torch.zeros((n, O), dtype=torch.float32, device=dev),
                  out_dev=torch.zeros((3, n, A), dtype=torch.float32, device=dev),
                  out_host=torch.zeros((3, n, A), dtype=torch.float32).pin_memory(),
                  eps_dev=torch.zeros((n, A), dtype=torch.float32, device=dev),
                  args={}, lib=_lib.lib(), byref=C.byref)
        st['ob_np'], st['out_np'] = st['ob_host'].numpy(), st['out_host'].numpy()
        _STATE[key] = st
    return st


_ZERO_COPY_MAX_OBS = 16      # up to this many observations travel through mapped pinned memory (no copy calls)


def _explore_args(policy, qfs, trainer, deterministic, n, st):
    """The OacExploreArgs of one (policy, critics, mode, n) configuration, built once: only the noise pointer, the RNG
    offset and beta / delta change between calls.  Re-built when a net is re-bound to a new arena (engine resize)."""
    key = (id(policy), policy._arena.data_ptr(), tuple((id(q), q._arena.data_ptr()) for q in qfs),
           type(trainer), getattr(trainer, 'delta_index', None), bool(deterministic))
    a = st['args'].get(key)
    if a is not None:
        return a
    a = OacExploreArgs()
    a.policy, a.policy_lay = policy._rel()
    nq = len(qfs)
    if nq > 16:
        raise NotImplementedError("more than 16 critics")
    for i, q in enumerate(qfs):
        a.q[i], lay = q._rel()
    a.q_lay, a.n_q = lay, nq
    # every critic applies its OWN ``positive`` flags (networks.py:69-75): bit (net * n_heads + head)
    n_heads = qfs[0].output_size
    a.exp_mask = 0
    for i, q in enumerate(qfs):
        if q.output_size != n_heads or q.hidden != qfs[0].hidden:
            raise NotImplementedError("critics of different shapes in one exploration call")
        a.exp_mask |= (q._exp_mask() << (i * n_heads)) & 0xffffffff
    if trainer is not None and hasattr(trainer, 'delta_index'):
        a.mode, a.quantile_index = _lib.EXPLORE_QUANTILE, trainer.delta_index      # ParticleTrainer.predict
    elif trainer is not None:
        if not _predict_is_twin(trainer):
            # e.g. GaussianTrainer.predict(obs, action, std=True) (trainer/gaussian_trainer.py:161): the reference's call
            # ``trainer.predict(..., upper_bound=True, beta_UB=...)`` (:38-39) raises TypeError there, and so do we
            raise TypeError("%s.predict() got an unexpected keyword argument 'upper_bound'" % type(trainer).__name__)
        a.mode = _lib.EXPLORE_TWIN          # SACTrainer.predict (trainer/trainer.py:105-123): trainer.qfs[0], [1]
    elif nq >= 2 and not deterministic:
        a.mode = _lib.EXPLORE_TWIN          # the try branch (:42-46): qfs[0], qfs[1] only
    else:
        a.mode = _lib.EXPLORE_ENSEMBLE      # the except branch (:47-58): mean + beta * unbiased std
    a.deterministic = int(bool(deterministic))
    a.n_obs = n
    zero_copy = n <= _ZERO_COPY_MAX_OBS
    # unified addressing: a pinned host allocation is device-accessible under the same pointer, so for a few
    # observations the kernel reads the observation from, and writes the action to, host memory directly
    a.obs = (st['ob_host'] if zero_copy else st['ob_dev']).data_ptr()
    out = st['out_host'] if zero_copy else st['out_dev']
    a.action, a.mu_E, a.grad = out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr()
    st['args'] = {key: a}                   # one live configuration per (O, A, n)
    return a


def _predict_is_twin(trainer):
    """True when ``trainer.predict`` is the twin upper bound of trainer/trainer.py:105-123."""
    import inspect
    try:
        return 'upper_bound' in inspect.signature(trainer.predict).parameters and len(trainer.qfs) >= 2
    except (TypeError, ValueError, AttributeError):
        return False


def _noise_key(policy, rng_seed):
    """(Philox key, call counter) of the final ``TanhNormal(mu_E, std).sample()`` draw (:92-94).  The reference draws from
    torch's global generator, so runs with different ``torch.manual_seed`` are independent and a resumed run does not
    replay its noise.  Here the key follows ``torch.initial_seed()`` (i.e. ``--seed``) unless given, and the call counter
    lives on the POLICY object: it is per policy, and it is pickled with the policy in the algorithm snapshot
    (rl_algorithm.py:220-239), so a resumed run continues the stream instead of restarting it."""
    if rng_seed is None:
        rng_seed = torch.initial_seed() & 0x7fffffffffffffff
    calls = getattr(policy, '_explore_calls', 0)
    policy._explore_calls = calls + 1
    return rng_seed, calls


def explore_batch(obs, policy, qfs, hyper_params, trainer=None, deterministic=False, eps=None, rng_seed=None):
    """Vectorised entry point: ``obs`` [n, O] numpy -> (actions [n, A], mu_E [n, A], grad [n, A])
    float32 numpy.  One launch for all observations."""
    if isinstance(policy, MakeDeterministic):
        policy = policy.stochastic_policy
    if trainer is not None:
        qfs = trainer.qfs                   # trainer.predict evaluates the trainer's own critics (:38-39)
    policy._ensure_bound()
    for q in qfs:
        q._ensure_bound()
    obs = np.asarray(obs)
    n, O = obs.shape
    A = policy.action_dim
    st = _buffers(O, A, n)
    a = _explore_args(policy, qfs, trainer, deterministic, n, st)
    a.beta_UB, a.delta = float(hyper_params['beta_UB']), float(hyper_params['delta'])
    zero_copy = n <= _ZERO_COPY_MAX_OBS
    st['ob_np'][...] = obs                                # f64 -> f32 (ptu.from_numpy, :22)
    if not zero_copy:
        st['ob_dev'].copy_(st['ob_host'], non_blocking=True)
    if eps is not None:
        st['eps_dev'].copy_(torch.as_tensor(np.asarray(eps, dtype=np.float32)).reshape(n, A))
        a.eps = st['eps_dev'].data_ptr()
    else:
        a.eps = None
    a.rng_seed, a.rng_offset = _noise_key(policy, rng_seed)
    stream = torch.cuda.current_stream()
    rc = st['lib'].oac_explore(st['byref'](a), C.c_void_p(stream.cuda_stream))
    if rc:
        _lib.check(rc, "oac_explore")
    if not zero_copy:
        st['out_host'].copy_(st['out_dev'], non_blocking=True)
    stream.synchronize()
    res = st['out_np']
    return res[0].copy(), res[1].copy(), res[2].copy()


def get_optimistic_exploration_action(ob_np, policy=None, qfs=None, trainer=None, hyper_params=None,
                                      deterministic=False, eps=None):
    """optimistic_exploration.py:7-11 (dispatch), :14-109 (stochastic), :111-196 (deterministic).  The sampling noise
    follows ``torch.manual_seed`` (see ``_noise_key``); ``eps`` injects the N(0,1) draw instead (parity tests)."""
    assert ob_np.ndim == 1
    for key in ('beta_UB', 'delta', 'share_layers'):
        hyper_params[key]                                   # same KeyError contract as :18-20
    ac, _, _ = explore_batch(ob_np[None], policy, qfs, hyper_params, trainer=trainer,
                             deterministic=deterministic, eps=None if eps is None else np.asarray(eps)[None])
    return ac[0], {}


def get_optimistic_exploration_action_stochastic(ob_np, policy=None, qfs=None, hyper_params=None, trainer=None):
    return get_optimistic_exploration_action(ob_np, policy, qfs, trainer, hyper_params, deterministic=False)


def get_optimistic_exploration_action_deterministic(ob_np, policy=None, qfs=None, hyper_params=None, trainer=None):
    return get_optimistic_exploration_action(ob_np, policy, qfs, trainer, hyper_params, deterministic=True)


class GroupExplorer(object):
    """Per-seed batched exploration for the rollout (``path_collector.py:214-232`` calls
    ``get_optimistic_exploration_action`` once per environment step and per seed process).

    ONE ``oac_explore`` launch serves the current observations of all seeds of a ``SACSeedGroup``: observation ``i``
    is evaluated with the policy and critics of seed slot ``slots[i]`` (``OacExploreArgs.obs_group``: a per-observation
    index into the ``[n_seeds, param_floats]`` parameter arena), so several environments per seed are possible too.
    Observations go up and actions come back through a ring of ``depth`` pinned host buffers with one CUDA event each:
    ``submit`` returns at once and ``collect`` waits for that submission only, so the caller can step one half of its
    environments on the CPU while the other half's actions are being computed (double buffering).
    """

    def __init__(self, group, hyper_params, max_obs=None, depth=2, rng_seed=None):
        self.group = group
        e = group.engine
        self.O, self.A = group.O, group.A
        self.max_obs = int(max_obs or group.n_seeds)
        self.beta_UB, self.delta = float(hyper_params['beta_UB']), float(hyper_params['delta'])
        hyper_params['share_layers']                      # same KeyError contract as optimistic_exploration.py:18-20
        self._lib = _lib.lib()
        dev = e.device
        n, O, A = self.max_obs, self.O, self.A
        self.depth = depth
        self._slots = []
        for _ in range(depth):
            d = dict(obs_host=torch.zeros((n, O), dtype=torch.float32).pin_memory(),
                     grp_host=torch.zeros((n,), dtype=torch.int32).pin_memory(),
                     out_host=torch.zeros((3, n, A), dtype=torch.float32).pin_memory(),
                     obs_dev=torch.zeros((n, O), dtype=torch.float32, device=dev),
                     grp_dev=torch.zeros((n,), dtype=torch.int32, device=dev),
                     eps_dev=torch.zeros((n, A), dtype=torch.float32, device=dev),
                     out_dev=torch.zeros((3, n, A), dtype=torch.float32, device=dev),
                     event=torch.cuda.Event(), busy=False, n=0)
            d['obs_np'], d['grp_np'], d['out_np'] = d['obs_host'].numpy(), d['grp_host'].numpy(), d['out_host'].numpy()
            self._slots.append(d)
        self._next = 0
        nets = group.nets[0]
        a = OacExploreArgs()
        a.policy, a.policy_lay = nets['policy']._rel()
        a.q[0], lay = nets['qf1']._rel()
        a.q[1], _ = nets['qf2']._rel()
        a.q_lay, a.n_q, a.mode = lay, 2, _lib.EXPLORE_TWIN          # SACTrainer critics: the twin branch (:42-46)
        a.exp_mask, a.deterministic = 0, 0
        a.group_stride = e.lay.param_floats
        self._args = a
        if rng_seed is None:
            rng_seed = (torch.initial_seed() ^ (0x9E3779B97F4A7C15 * (1 + group.seed_ids[0]))) & 0x7fffffffffffffff
        self._rng_seed, self._calls = rng_seed, 0

    def submit(self, obs, slots=None, eps=None):
        """obs [n, O] numpy; slots [n] seed slots (default: observation i belongs to seed slot i).  Returns a ticket."""
        obs = np.asarray(obs)
        n = obs.shape[0]
        if n > self.max_obs:
            raise ValueError("more observations than max_obs")
        t = self._next
        d = self._slots[t]
        if d['busy']:
            raise RuntimeError("GroupExplorer ring full: collect() the oldest ticket first")
        self._next = (t + 1) % self.depth
        d['obs_np'][:n] = obs                                    # f64 -> f32 (ptu.from_numpy, :22)
        d['grp_np'][:n] = np.arange(n) if slots is None else np.asarray(slots)
        stream = torch.cuda.current_stream()
        d['obs_dev'][:n].copy_(d['obs_host'][:n], non_blocking=True)
        d['grp_dev'][:n].copy_(d['grp_host'][:n], non_blocking=True)
        a = self._args
        a.n_obs, a.obs, a.obs_group = n, d['obs_dev'].data_ptr(), d['grp_dev'].data_ptr()
        a.beta_UB, a.delta = self.beta_UB, self.delta
        if eps is not None:
            d['eps_dev'][:n].copy_(torch.as_tensor(np.asarray(eps, dtype=np.float32)).reshape(n, self.A))
            a.eps = d['eps_dev'].data_ptr()
        else:
            a.eps = None
        a.rng_seed, a.rng_offset = self._rng_seed, self._calls
        self._calls += 1
        out = d['out_dev']
        a.action, a.mu_E, a.grad = out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr()
        rc = self._lib.oac_explore(C.byref(a), C.c_void_p(stream.cuda_stream))
        if rc:
            _lib.check(rc, "oac_explore")
        d['out_host'].copy_(out, non_blocking=True)
        d['event'].record(stream)
        d['busy'], d['n'] = True, n
        return t

    def collect(self, ticket):
        """(actions [n, A], mu_E [n, A]) float32 numpy of a submission."""
        d = self._slots[ticket]
        if not d['busy']:
            raise RuntimeError("ticket already collected")
        d['event'].synchronize()
        d['busy'] = False
        n = d['n']
        return d['out_np'][0, :n].copy(), d['out_np'][1, :n].copy()

    def actions(self, obs, slots=None, eps=None):
        return self.collect(self.submit(obs, slots, eps))[0]
