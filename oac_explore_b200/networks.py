"""Host-side mirrors of the reference's network objects (networks.py:14-79,154-161 and
trainer/policies.py:195-316,486-513).

They are thin handles: the weights live in a device arena (a trainer's, or a private one
for a free-standing net), initialised exactly like the reference (same torch RNG
consumption, so ``torch.manual_seed(s)`` gives the same initial weights as the
reference), and every forward pass runs the library's fused kernels
(``oac_q_forward`` / ``oac_policy_forward``).  There is no torch/CPU compute path.
"""
import ctypes as C
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from ._lib import OacNetLayout

LOG_SIG_MAX = 2
LOG_SIG_MIN = -20


def _pad4(n):
    return (n + 3) & ~3


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("oac_explore_b200 needs a CUDA device (B200, sm_100a); no CPU fallback exists")
    return torch.device("cuda:%d" % torch.cuda.current_device())


def from_numpy(x):
    """utils/pytorch_util.py:76-77."""
    return torch.from_numpy(np.ascontiguousarray(x)).float().to(_device())


class Mlp(object):
    """networks.py:16-79.  Two equal hidden layers (main.py's ``[M] * N`` with N = 2)."""
    KIND = _lib.NET_Q

    def __init__(self, hidden_sizes, output_size, input_size, init_w=3e-3, b_init_value=0.1, bias=None,
                 positive=False, train_bias=True, _extra_head=0):
        hidden_sizes = list(hidden_sizes)
        if len(hidden_sizes) != 2 or hidden_sizes[0] != hidden_sizes[1]:
            raise NotImplementedError("the fused kernels are built for two equal hidden layers "
                                      "(main.py: --num_layers 2); got %r" % (hidden_sizes,))
        self.input_size, self.output_size = input_size, output_size
        self.hidden = hidden_sizes[0]
        self.positive = positive
        self.train_bias = train_bias
        # --- initial values on the host, consuming torch's RNG like networks.py:42-60 ---
        cpu = OrderedDict()
        in_size = input_size
        for i, h in enumerate(hidden_sizes):
            fc = torch.nn.Linear(in_size, h)
            bound = 1. / np.sqrt(fc.weight.size(0))          # ptu.fanin_init uses size[0] (= out features)
            fc.weight.data.uniform_(-bound, bound)
            fc.bias.data.fill_(b_init_value)
            cpu['fc%d.weight' % i], cpu['fc%d.bias' % i] = fc.weight.data, fc.bias.data
            in_size = h
        last = torch.nn.Linear(in_size, output_size)
        last.weight.data.uniform_(-init_w, init_w)
        if bias is None:
            last.bias.data.uniform_(-init_w, init_w)
        elif isinstance(bias, np.ndarray):
            last.bias.data = torch.from_numpy(bias.astype(np.float32)).reshape(-1)
        else:
            last.bias.data.fill_(float(bias))
        cpu['last_fc.weight'], cpu['last_fc.bias'] = last.weight.data, last.bias.data
        if _extra_head:
            hd = torch.nn.Linear(in_size, _extra_head)
            hd.weight.data.uniform_(-init_w, init_w)
            hd.bias.data.uniform_(-init_w, init_w)
            cpu['last_fc_log_std.weight'], cpu['last_fc_log_std.bias'] = hd.weight.data, hd.bias.data
        self._init = cpu
        self._views = None          # state_dict-named device views once bound
        self._arena = None          # tensor whose data_ptr is the layout's base
        self._lay = None
        self._own = None

    # ---- binding ----------------------------------------------------------------
    def _n_out(self):
        return self.output_size

    def _bind(self, views, arena_row, lay, copy_from=None):
        src = copy_from if copy_from is not None else (self._views if self._views is not None else self._init)
        for k, v in views.items():
            v.copy_(src[k].to(v.device))
        self._views, self._arena, self._lay = views, arena_row, lay
        self._init = None

    def _ensure_bound(self):
        if self._views is not None:
            return
        H, K, NO = self.hidden, self.input_size, self._n_out()
        lay = OacNetLayout()
        lay.kind, lay.in_dim, lay.in_ld, lay.hidden, lay.n_out, lay.trainable = self.KIND, K, _pad4(K), H, NO, 0
        cur = 0
        for name, n in (("off_w0", H * lay.in_ld), ("off_b0", H), ("off_w1", H * H), ("off_b1", H),
                        ("off_w2", NO * H), ("off_b2", NO)):
            setattr(lay, name, cur)
            cur += _pad4(n)
        lay.size = cur
        self._own = torch.zeros(cur, dtype=torch.float32, device=_device())
        b = self._own
        views = OrderedDict()
        views['fc0.weight'] = b[lay.off_w0:lay.off_w0 + H * lay.in_ld].view(H, lay.in_ld)[:, :K]
        views['fc0.bias'] = b[lay.off_b0:lay.off_b0 + H]
        views['fc1.weight'] = b[lay.off_w1:lay.off_w1 + H * H].view(H, H)
        views['fc1.bias'] = b[lay.off_b1:lay.off_b1 + H]
        w2 = b[lay.off_w2:lay.off_w2 + NO * H].view(NO, H)
        b2 = b[lay.off_b2:lay.off_b2 + NO]
        if self.KIND == _lib.NET_POLICY:
            A = NO // 2
            views['last_fc.weight'], views['last_fc.bias'] = w2[:A], b2[:A]
            views['last_fc_log_std.weight'], views['last_fc_log_std.bias'] = w2[A:], b2[A:]
        else:
            views['last_fc.weight'], views['last_fc.bias'] = w2, b2
        self._bind(views, self._own, lay)

    def _rel(self):
        """(device pointer of this net's first float, layout with offsets relative to it): lets a
        kernel take several same-shaped nets as plain base pointers + ONE layout."""
        self._ensure_bound()
        l, r = self._lay, OacNetLayout()
        for f, _ in OacNetLayout._fields_:
            setattr(r, f, getattr(l, f))
        for f in ("off_w0", "off_b0", "off_w1", "off_b1", "off_w2", "off_b2"):
            setattr(r, f, getattr(l, f) - l.off_w0)
        return self._arena.data_ptr() + 4 * l.off_w0, r

    # ---- nn.Module-like surface (what main.py / rl_algorithm.py / snapshots touch) --------
    def state_dict(self):
        src = self._views if self._views is not None else self._init
        return OrderedDict((k, v.detach()) for k, v in src.items())

    def load_state_dict(self, sd, strict=True):
        self._ensure_bound()
        for k, v in self._views.items():
            v.copy_(torch.as_tensor(sd[k]).to(v.device, torch.float32).reshape(v.shape))

    def parameters(self):
        self._ensure_bound()
        return list(self._views.values())

    def to(self, *a, **k):
        return self

    def train(self, mode=True):
        return self

    def eval(self):
        return self

    def _exp_mask(self):
        if not self.positive:
            return 0
        if isinstance(self.positive, (list, tuple)):
            return sum(1 << i for i, v in enumerate(self.positive) if v)
        return (1 << self.output_size) - 1

    def forward(self, input, return_preactivations=False):
        if return_preactivations:
            raise NotImplementedError("return_preactivations is not on the hot path")
        self._ensure_bound()
        x = input
        if x.dim() == 1:
            x = x[None]
        x = x.to(_device(), torch.float32)
        if x.stride(-1) != 1:
            x = x.contiguous()
        n = x.shape[0]
        out = torch.empty((n, self.output_size), dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().oac_q_forward(_lib.ptr(self._arena), C.byref(self._lay), _lib.ptr(x), x.stride(0), n,
                                            self._exp_mask(), _lib.ptr(out), _lib.current_stream()), "oac_q_forward")
        return out

    def __call__(self, *a, **k):
        return self.forward(*a, **k)


class FlattenMlp(Mlp):
    """networks.py:154-161: concatenate the inputs along dim 1, then the MLP."""

    def forward(self, *inputs, **kwargs):
        return super().forward(torch.cat(inputs, dim=1), **kwargs)


class TanhGaussianPolicy(Mlp):
    """trainer/policies.py:195-316."""
    KIND = _lib.NET_POLICY

    def __init__(self, hidden_sizes, obs_dim, action_dim, std=None, init_w=1e-3, bias=None, **kwargs):
        if std is not None:
            raise NotImplementedError("fixed-std policies (ddpg_noisy) are outside the OAC hot path")
        if bias is not None:
            bias = np.arctanh(bias)
        super().__init__(hidden_sizes, input_size=obs_dim, output_size=action_dim, init_w=init_w, bias=bias,
                         _extra_head=action_dim, **kwargs)
        self.action_dim = action_dim
        self.std = None
        self.log_std = None
        self.policies_list = [self]

    def _n_out(self):
        return 2 * self.action_dim

    def get_action(self, obs_np, deterministic=False):
        actions = self.get_actions(obs_np[None], deterministic=deterministic)
        return actions[0, :], {}

    _ACT_ZERO_COPY_MAX = 16

    def get_actions(self, obs_np, deterministic=False):
        """numpy observations -> numpy actions (trainer/policies.py:253-258).  Up to a few rows (the per-environment-step
        case) travel through mapped pinned memory: the kernel reads the observation from, and writes the action to, host
        memory directly, so a call is one launch and one stream synchronisation."""
        obs_np = np.asarray(obs_np)
        n = obs_np.shape[0]
        if n > self._ACT_ZERO_COPY_MAX:
            out = self.forward(from_numpy(obs_np), deterministic=deterministic)[0]
            return out.to('cpu').numpy()
        self._ensure_bound()
        A, O = self.action_dim, self.input_size
        st = getattr(self, '_act_state', None)
        if st is None or st['n'] != n:
            st = dict(n=n, obs=torch.zeros((n, O), dtype=torch.float32).pin_memory(),
                      act=torch.zeros((n, A), dtype=torch.float32).pin_memory(), lib=_lib.lib())
            st['obs_np'], st['act_np'] = st['obs'].numpy(), st['act'].numpy()
            self._act_state = st
        st['obs_np'][...] = obs_np                              # f64 -> f32 (ptu.from_numpy)
        eps = None if deterministic else torch.randn((n, A), dtype=torch.float32, device=_device())
        stream = torch.cuda.current_stream()
        rc = st['lib'].oac_policy_forward(
            _lib.ptr(self._arena), C.byref(self._lay), _lib.ptr(st['obs']), O, n, _lib.ptr(eps),
            _lib.ptr(st['act']), None, None, None, None, None, C.c_void_p(stream.cuda_stream))
        if rc:
            _lib.check(rc, "oac_policy_forward")
        stream.synchronize()
        return st['act_np'].copy()

    def forward(self, obs, reparameterize=True, deterministic=False, return_log_prob=False):
        self._ensure_bound()
        single = obs.dim() == 1
        x = (obs[None] if single else obs).to(_device(), torch.float32)
        if x.stride(-1) != 1:
            x = x.contiguous()
        n, A = x.shape[0], self.action_dim
        dev = x.device
        mk = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        action, mean, log_std, std, pre = mk(n, A), mk(n, A), mk(n, A), mk(n, A), mk(n, A)
        lp = mk(n, 1)
        eps = None if deterministic else torch.randn((n, A), dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().oac_policy_forward(
            _lib.ptr(self._arena), C.byref(self._lay), _lib.ptr(x), x.stride(0), n, _lib.ptr(eps),
            _lib.ptr(action), _lib.ptr(mean), _lib.ptr(log_std), _lib.ptr(std), _lib.ptr(pre), _lib.ptr(lp),
            _lib.current_stream()), "oac_policy_forward")
        if deterministic or not return_log_prob:
            # log_prob None -> zeros_like(action), pre_tanh -> mean   (policies.py:309-311)
            lp, pre = torch.zeros_like(action), mean
        outs = (action, mean, log_std, lp, std, pre)
        if single:
            outs = tuple(o[0] for o in outs)
        return outs

    def reset(self):
        pass


class MakeDeterministic(object):
    """trainer/policies.py:486-513."""

    def __init__(self, stochastic_policy):
        self.stochastic_policy = stochastic_policy

    def get_action(self, observation, deterministic=True):
        return self.stochastic_policy.get_action(observation, deterministic=True)

    def get_actions(self, obs_np, deterministic=True):
        return self.stochastic_policy.get_actions(obs_np, deterministic=True)

    def reset(self):
        pass

    def forward(self, obs, reparameterize=True, deterministic=False, return_log_prob=False):
        return self.stochastic_policy.forward(obs, reparameterize=reparameterize, deterministic=True,
                                              return_log_prob=return_log_prob)

    __call__ = forward

    def load_state_dict(self, *a, **k):
        return self.stochastic_policy.load_state_dict(*a, **k)

    def state_dict(self, *a, **k):
        return self.stochastic_policy.state_dict(*a, **k)

    def parameters(self):
        return self.stochastic_policy.parameters()

    def to(self, *a, **k):
        return self


def get_policy_producer(obs_dim, action_dim, hidden_sizes, clip=True, std=None):
    """main.py:44-94 (the TanhGaussianPolicy branch)."""
    def policy_producer(deterministic=False, bias=None, **unused):
        policy = TanhGaussianPolicy(obs_dim=obs_dim, action_dim=action_dim, hidden_sizes=hidden_sizes,
                                    bias=bias, std=std)
        if deterministic:
            policy = MakeDeterministic(policy)
        return policy
    return policy_producer


def get_q_producer(obs_dim, action_dim, hidden_sizes, output_size=1):
    """main.py:97-106."""
    def q_producer(bias=None, positive=False, train_bias=True):
        return FlattenMlp(input_size=obs_dim + action_dim, output_size=output_size, hidden_sizes=hidden_sizes,
                          bias=bias, positive=positive, train_bias=train_bias)
    return q_producer
