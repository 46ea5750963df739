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
                  calls=0, args={}, lib=_lib.lib(), byref=C.byref)
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
    a.exp_mask = qfs[0]._exp_mask()
    if trainer is not None and hasattr(trainer, 'delta_index'):
        a.mode, a.quantile_index = _lib.EXPLORE_QUANTILE, trainer.delta_index      # ParticleTrainer.predict
    elif trainer is not None or (nq >= 2 and not deterministic):
        a.mode = _lib.EXPLORE_TWIN          # SACTrainer.predict / the try branch (:42-46): qfs[0], qfs[1] only
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


def explore_batch(obs, policy, qfs, hyper_params, trainer=None, deterministic=False, eps=None, rng_seed=0):
    """Vectorised entry point: ``obs`` [n, O] numpy -> (actions [n, A], mu_E [n, A], grad [n, A])
    float32 numpy.  One launch for all observations."""
    if isinstance(policy, MakeDeterministic):
        policy = policy.stochastic_policy
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
    a.rng_seed, a.rng_offset = rng_seed, st['calls']
    st['calls'] += 1
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
    """optimistic_exploration.py:7-11 (dispatch), :14-109 (stochastic), :111-196 (deterministic)."""
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
