"""CPU oracle: a from-scratch restatement of the reference's per-gradient-step hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module, and only as the checker or the CPU baseline.  Nothing under
``oac_explore_b200/`` imports it; the product path fails loudly without the
CUDA library.

Parity pinning: the reference (amarildolikmeta/oac-explore) ships NO tests and
NO golden vectors, so this restatement is pinned against
the reference code itself, imported unmodified in the build container
(``oracle/ref_import.py``): ``tests/golden/make_golden.py`` ran the real
``SACTrainer`` / ``ParticleTrainer`` / ``GaussianTrainer`` /
``get_optimistic_exploration_action`` / ``ReplayBuffer`` and committed their
inputs+outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks
this file against them on every run and ``tests/test_oracle_vs_reference.py``
re-runs the live comparison whenever ``/root/reference`` is present.

All arithmetic is torch CPU in the dtype of the tensors handed in (fp32 to
mirror the reference, fp64 to arbitrate).  Citations are relative to the
reference root.

Stale-graph semantics ("Mode A"): the reference pins
torch 1.4 where ``policy_loss.backward()`` (trainer/trainer.py:209) runs after
the Q optimizers stepped (:202,:206) and therefore multiplies by the POST-step
Q weights in the dX products while ReLU masks come from the PRE-step forward.
``sac_step(..., mode="A")`` states that explicitly; ``mode="B"`` uses pre-step
weights (what a "fixed" reference would do).
"""
from collections import OrderedDict
import math

import numpy as np
import torch

LOG_SIG_MAX = 2.0   # trainer/policies.py:10
LOG_SIG_MIN = -20.0  # trainer/policies.py:11
TANH_EPS = 1e-6     # trainer/policies.py:127


# --------------------------------------------------------------------------
# TF32 arithmetic model (for the tcgen05 kind::tf32 path only)
# --------------------------------------------------------------------------
# The reference computes every nn.Linear product in fp32.  The CUDA GEMM_TF32 path feeds the tensor
# cores with operands rounded to tf32 (10 mantissa bits; the TMA unit rounds to nearest on the way into
# shared memory) and accumulates in fp32.  ``tf32_mode`` restates exactly that arithmetic on the CPU --
# both operands of the forward product, of the input-gradient product (dY W) and of the weight-gradient
# product (dY^T X, and the bias gradient as the column sums of the ROUNDED dY, which is what the all-ones
# MMA computes) are rounded, products and sums stay fp32 -- so the CUDA path can be held to rel <= 1e-3 on
# gradients and weights too, not only on values: against THIS model ReLU masks agree, against the fp32
# oracle they flip for the few units whose pre-activation sits within the rounding error of zero.
# Which products are GEMM stages (rounded) and which are fused into the fp32 "glue" kernels (exact) depends on the regime
# the program builder picks (oac_explore_b200/csrc/program.cu); hidden layers are GEMM stages in every regime:
#   "trunk" : few seeds (< 2048 batch rows per launch).  Head layers (critic AND policy) are exact in the forward and in
#             their input gradient, and so is the action-column product dh1 W0[:, O:O+A] of the policy loss; a head's
#             WEIGHT gradient is a GEMM stage.
#   "many"  : many seeds (the batched-seed program).  The policy head, its backward and dQ/da are GEMM stages too; the
#             critic head's forward and its input gradients (dq W3: critic_head / rank1_mask kernels) stay exact fp32.
#   "chain" : "many" where the forward layers of a 128-row strip run as one strip-fused launch (gemm_chain.cuh, small
#             seed groups): the hidden activations are stored tf32-ROUNDED, so the critic head's fp32 dot product sees
#             rounded h2 rows (every other consumer rounds them itself or only tests their sign).
#   "all"   : every Linear rounded (the idealised model; no kernel regime is exactly this).
_TF32 = {"mode": None}


def round_tf32(x):
    """fp32 -> tf32, round to nearest (ties away from zero, cvt.rna.tf32.f32), kept in an fp32 container."""
    if x.dtype != torch.float32:
        return x
    i = x.detach().contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


class tf32_mode(object):
    """``with orc.tf32_mode("all"): orc.sac_step(...)`` (see the comment above)."""

    def __init__(self, mode="all"):
        assert mode in (None, "all", "trunk", "many", "chain")
        self.mode = mode

    def __enter__(self):
        self.prev = _TF32["mode"]
        _TF32["mode"] = self.mode
        return self

    def __exit__(self, *a):
        _TF32["mode"] = self.prev
        return False


class _LinearTF32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, r_fwd, r_dx, r_dw):
        ctx.save_for_backward(x, w)
        ctx.flags = (r_dx, r_dw)
        if r_fwd == "x":                        # the input arrives rounded (stored so), the weights are used exactly
            xr, wr = round_tf32(x), w
        else:
            xr, wr = (round_tf32(x), round_tf32(w)) if r_fwd else (x, w)
        y = xr @ wr.t()
        return y + b if b is not None else y

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        r_dx, r_dw = ctx.flags
        gr = round_tf32(gy)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = (gr @ round_tf32(w)) if r_dx else (gy @ w)
        if ctx.needs_input_grad[1]:
            gw = (gr.t() @ round_tf32(x)) if r_dw else (gy.t() @ x)
        if ctx.needs_input_grad[2]:
            gb = (gr if r_dw else gy).sum(dim=0)
        return gx, gw, gb, None, None, None


def linear(x, w, b, head=None):
    """y = x W^T + b (nn.Linear).  fp32 / fp64 exact unless a ``tf32_mode`` is active.  head: None (hidden layer),
    'q' (critic output layer) or 'pi' (policy output layers)."""
    mode = _TF32["mode"]
    if mode is None or x.dtype != torch.float32:
        return x @ w.t() + b
    exact = (head is not None and mode == "trunk") or (head == 'q' and mode in ("many", "chain"))
    r_fwd = "x" if (head == 'q' and mode == "chain") else (not exact)
    return _LinearTF32.apply(x, w, b, r_fwd, not exact, True)


def matmul_dx(g, w, exact_modes=()):
    """grad_in = grad_out @ W (Linear backward, written out by hand in ``_q_dx_action``); exact fp32 in the tf32 regimes
    listed in ``exact_modes`` (the product is fused into a glue kernel there)."""
    mode = _TF32["mode"]
    if mode is None or g.dtype != torch.float32 or mode in exact_modes or (mode == "chain" and "many" in exact_modes):
        return g @ w
    return round_tf32(g) @ round_tf32(w)


# --------------------------------------------------------------------------
# replay buffer  (replay_buffer.py:8-148, 151-203)
# --------------------------------------------------------------------------
class ReplayBuffer(object):
    """replay_buffer.py:8-148.  float64 ring arrays, uint8 terminals."""

    def __init__(self, max_replay_buffer_size, ob_dim, ac_dim):
        n = int(max_replay_buffer_size)
        self._max_replay_buffer_size = n
        self._observations = np.zeros((n, ob_dim))       # :32
        self._next_obs = np.zeros((n, ob_dim))           # :37
        self._actions = np.zeros((n, ac_dim))            # :38
        self._rewards = np.zeros((n, 1))                 # :42
        self._terminals = np.zeros((n, 1), dtype='uint8')  # :45
        self._top = 0
        self._size = 0

    def add_sample(self, observation, action, reward, next_observation, terminal, **kw):
        # replay_buffer.py:88-99
        t = self._top
        self._observations[t] = observation
        self._actions[t] = action
        self._rewards[t] = reward
        self._terminals[t] = terminal
        self._next_obs[t] = next_observation
        self._advance()

    def _advance(self):
        # replay_buffer.py:101-104
        self._top = (self._top + 1) % self._max_replay_buffer_size
        if self._size < self._max_replay_buffer_size:
            self._size += 1

    def draw_indices(self, batch_size):
        # replay_buffer.py:107 -- global numpy MT19937 stream, with replacement
        return np.random.randint(0, self._size, batch_size)

    def gather(self, indices):
        # replay_buffer.py:108-114
        return dict(
            observations=self._observations[indices],
            actions=self._actions[indices],
            rewards=self._rewards[indices],
            terminals=self._terminals[indices],
            next_observations=self._next_obs[indices],
        )

    def random_batch(self, batch_size):
        return self.gather(self.draw_indices(batch_size))

    def num_steps_can_sample(self):
        return self._size


class ReplayBufferCount(ReplayBuffer):
    """replay_buffer.py:151-203 (uniform sampling branch :186)."""

    def __init__(self, max_replay_buffer_size, ob_dim, ac_dim):
        super().__init__(max_replay_buffer_size, ob_dim, ac_dim)
        self._counts = np.zeros((int(max_replay_buffer_size), 1))  # :164

    def add_sample(self, observation, action, reward, next_observation, terminal, **kw):
        self._counts[self._top] = 0  # :177
        super().add_sample(observation, action, reward, next_observation, terminal)

    def gather(self, indices):
        batch = super().gather(indices)
        batch['counts'] = np.copy(self._counts[indices])  # :193
        self._counts[indices] += 1  # :195 -- fancy "+=": a duplicated index increments ONCE
        return batch


def np_to_torch_batch(np_batch, dtype=torch.float32):
    """utils/core.py:40-61: ``torch.from_numpy(x).float()`` per array (bool -> int first)."""
    out = {}
    for k, v in np_batch.items():
        if not isinstance(v, np.ndarray) or v.dtype == np.dtype('O'):
            continue
        if v.dtype == np.bool_:
            v = v.astype(int)
        out[k] = torch.from_numpy(v).to(dtype)
    return out


# --------------------------------------------------------------------------
# networks  (networks.py:17-79,154-161; trainer/policies.py:195-316)
# --------------------------------------------------------------------------
def init_mlp(input_size, hidden_sizes, output_size, init_w=3e-3, b_init_value=0.1,
             bias=None, extra_head=None):
    """networks.py:17-60.  Consumes torch's global RNG exactly like the reference:
    each ``nn.Linear`` constructor draws its default init, then hidden weights are
    re-drawn U(+-1/sqrt(size[0])) (utils/pytorch_util.py:17-26 -- ``size[0]`` is
    OUT features), hidden biases = 0.1, last layer U(+-init_w) (bias fixed when
    ``bias`` is given).  ``extra_head`` adds ``last_fc_log_std``
    (trainer/policies.py:236-241).  Returns an OrderedDict with state_dict keys."""
    p = OrderedDict()
    in_size = input_size
    for i, h in enumerate(hidden_sizes):
        fc = torch.nn.Linear(in_size, h)
        bound = 1.0 / np.sqrt(fc.weight.size(0))
        fc.weight.data.uniform_(-bound, bound)
        fc.bias.data.fill_(b_init_value)
        p['fc%d.weight' % i] = fc.weight.data.clone()
        p['fc%d.bias' % i] = fc.bias.data.clone()
        in_size = h
    last = torch.nn.Linear(in_size, output_size)
    last.weight.data.uniform_(-init_w, init_w)
    if bias is None:
        last.bias.data.uniform_(-init_w, init_w)
    elif isinstance(bias, np.ndarray):
        last.bias.data = torch.from_numpy(bias.astype(np.float32))
    else:
        last.bias.data.fill_(float(bias))
    p['last_fc.weight'] = last.weight.data.clone()
    p['last_fc.bias'] = last.bias.data.clone()
    if extra_head is not None:
        hd = torch.nn.Linear(in_size, extra_head)
        hd.weight.data.uniform_(-init_w, init_w)
        hd.bias.data.uniform_(-init_w, init_w)
        p['last_fc_log_std.weight'] = hd.weight.data.clone()
        p['last_fc_log_std.bias'] = hd.bias.data.clone()
    return p


def init_q(obs_dim, act_dim, hidden=(256, 256), output_size=1, bias=None):
    """main.py:97-106 get_q_producer -> FlattenMlp(init_w=3e-3)."""
    return init_mlp(obs_dim + act_dim, hidden, output_size, init_w=3e-3, bias=bias)


def init_policy(obs_dim, act_dim, hidden=(256, 256)):
    """main.py:44-94 -> TanhGaussianPolicy(init_w=1e-3) (trainer/policies.py:214-241)."""
    return init_mlp(obs_dim, hidden, act_dim, init_w=1e-3, extra_head=act_dim)


def _n_hidden(p):
    n = 0
    while 'fc%d.weight' % n in p:
        n += 1
    return n


def mlp_trunk(p, x):
    """networks.py:63-67: (Linear+ReLU) per hidden layer.  Returns list of activations."""
    hs = []
    h = x
    for i in range(_n_hidden(p)):
        h = torch.relu(linear(h, p['fc%d.weight' % i], p['fc%d.bias' % i]))
        hs.append(h)
    return hs


def q_forward(p, obs, act, positive=None, return_hidden=False):
    """FlattenMlp.forward (networks.py:154-161 cat; :62-79 MLP; :69-75 ``positive`` -> exp)."""
    x = torch.cat([obs, act], dim=1)
    hs = mlp_trunk(p, x)
    out = linear(hs[-1], p['last_fc.weight'], p['last_fc.bias'], head='q')
    if positive is not None and positive is not False:
        if isinstance(positive, (list, tuple)):
            cols = [torch.exp(out[:, i]) if v else out[:, i] for i, v in enumerate(positive)]
            out = torch.stack(cols, dim=1)
        else:
            out = torch.exp(out)
    if return_hidden:
        return out, hs
    return out


def policy_forward(p, obs, eps=None, deterministic=False):
    """TanhGaussianPolicy.forward (trainer/policies.py:260-316) with
    reparameterize=True, return_log_prob=True.  ``eps`` replaces the N(0,1) draw of
    TanhNormal.rsample (:179-187).  Returns the reference's 6-tuple."""
    hs = mlp_trunk(p, obs)
    h = hs[-1]
    mean = linear(h, p['last_fc.weight'], p['last_fc.bias'], head='pi')
    log_std = linear(h, p['last_fc_log_std.weight'], p['last_fc_log_std.bias'], head='pi')
    log_std = torch.clamp(log_std, LOG_SIG_MIN, LOG_SIG_MAX)
    std = torch.exp(log_std)
    if deterministic:
        action = torch.tanh(mean)              # :285-287
        log_prob = torch.zeros_like(action)    # :309-311
        pre_tanh = mean
    else:
        if eps is None:
            eps = torch.normal(torch.zeros_like(mean), torch.ones_like(std))
        z = mean + std * eps                   # :179-187
        action = torch.tanh(z)
        # Normal(mean,std).log_prob(z) (torch.distributions) - log(1 - a^2 + eps)  (:147-160)
        normal_lp = -((z - mean) ** 2) / (2 * std ** 2) - torch.log(std) - math.log(math.sqrt(2 * math.pi))
        log_prob = normal_lp - torch.log(1 - action * action + TANH_EPS)
        log_prob = log_prob.sum(dim=1, keepdim=True)  # :304
        pre_tanh = z
    return action, mean, log_std, log_prob, std, pre_tanh


# --------------------------------------------------------------------------
# Adam (torch 1.4 torch/optim/adam.py as used at trainer/trainer.py:75-91), Polyak
# --------------------------------------------------------------------------
class Adam(object):
    """lr given, betas (0.9,0.999), eps 1e-8, no weight decay, no amsgrad.
    ``m = b1*m + (1-b1)*g; v = b2*v + (1-b2)*g*g;
    p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)``.
    Parameters whose gradient is None are skipped (no state, no step count)."""

    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-8):
        self.params = params  # dict name -> tensor (updated in place)
        self.lr, self.b1, self.b2, self.eps = lr, betas[0], betas[1], eps
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}
        self.t = {k: 0 for k in params}

    def step(self, grads):
        for k, g in grads.items():
            if g is None:
                continue
            self.t[k] += 1
            t = self.t[k]
            self.m[k].mul_(self.b1).add_(g, alpha=1 - self.b1)
            self.v[k].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            bc1 = 1 - self.b1 ** t
            bc2 = 1 - self.b2 ** t
            denom = (self.v[k].sqrt() / math.sqrt(bc2)).add_(self.eps)
            self.params[k].addcdiv_(self.m[k], denom, value=-(self.lr / bc1))


def soft_update(source, target, tau):
    """utils/pytorch_util.py:5-9: target <- target*(1-tau) + source*tau (that order)."""
    for k in target:
        target[k].copy_(target[k] * (1.0 - tau) + source[k] * tau)


def _grads(loss, params, retain_graph=False):
    names = list(params.keys())
    gs = torch.autograd.grad(loss, [params[k] for k in names], allow_unused=True,
                             retain_graph=retain_graph)
    return dict(zip(names, gs))


def _leafs(params):
    return OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in params.items())


# --------------------------------------------------------------------------
# SAC / OAC update  (trainer/trainer.py:15-97 ctor, :126-224 step)
# --------------------------------------------------------------------------
class SACState(object):
    """Weights + optimizers of SACTrainer.  Construction order policy, qf1, qf2,
    target_qf1, target_qf2 (trainer/trainer.py:58-71): targets are INDEPENDENT
    random inits, not copies."""

    def __init__(self, obs_dim, act_dim, hidden=(256, 256), policy_lr=3e-4, qf_lr=3e-4,
                 discount=0.99, reward_scale=1.0, soft_target_tau=5e-3, target_update_period=1,
                 use_automatic_entropy_tuning=True, target_entropy=None, dtype=torch.float32):
        self.obs_dim, self.act_dim = obs_dim, act_dim
        self.policy = init_policy(obs_dim, act_dim, hidden)
        self.qf1 = init_q(obs_dim, act_dim, hidden)
        self.qf2 = init_q(obs_dim, act_dim, hidden)
        self.target_qf1 = init_q(obs_dim, act_dim, hidden)
        self.target_qf2 = init_q(obs_dim, act_dim, hidden)
        self.log_alpha = {'log_alpha': torch.zeros(1)}
        self.discount, self.reward_scale = discount, reward_scale
        self.tau, self.period = soft_target_tau, target_update_period
        self.auto_alpha = use_automatic_entropy_tuning
        self.target_entropy = target_entropy if target_entropy else -float(act_dim)  # :38-44
        self.policy_lr, self.qf_lr = policy_lr, qf_lr
        self.n_steps = 0
        self.to(dtype)

    def nets(self):
        return OrderedDict(policy=self.policy, qf1=self.qf1, qf2=self.qf2,
                           target_qf1=self.target_qf1, target_qf2=self.target_qf2)

    def to(self, dtype):
        for net in list(self.nets().values()) + [self.log_alpha]:
            for k in net:
                net[k] = net[k].to(dtype)
        self.opt_policy = Adam(self.policy, self.policy_lr)
        self.opt_qf1 = Adam(self.qf1, self.qf_lr)
        self.opt_qf2 = Adam(self.qf2, self.qf_lr)
        self.opt_alpha = Adam(self.log_alpha, self.policy_lr)
        return self

    def load(self, nets):
        """nets: dict name -> state_dict-like mapping (tensors or numpy)."""
        for name, sd in nets.items():
            dst = self.log_alpha if name == 'log_alpha' else getattr(self, name)
            if name == 'log_alpha':
                sd = {'log_alpha': sd}
            for k in dst:
                dst[k].copy_(torch.as_tensor(np.asarray(sd[k])).to(dst[k].dtype).reshape(dst[k].shape))


def _q_dx_action(p, hs, dq, obs_dim, mode_b=False):
    """dQ/d(action) chain with explicit weights ``p`` and saved activations ``hs``:
    Linear backward grad_in = grad_out @ W, ReLU backward grad * (out > 0).  (tf32 regimes: the head product is a GEMM
    stage of the many-seed program in mode A only -- in mode B the critic_head kernel forms it; the action columns of the
    fc0 product are formed inside the policy_grad kernel in the few-seed regime.)"""
    n = len(hs)
    g = matmul_dx(dq, p['last_fc.weight'], exact_modes=("trunk", "many") if mode_b else ("trunk",))
    for i in range(n - 1, -1, -1):
        g = g * (hs[i] > 0).to(g.dtype)
        g = matmul_dx(g, p['fc%d.weight' % i], exact_modes=("trunk",) if i == 0 else ())
    return g[:, obs_dim:]


def sac_step(st, batch, eps_pi, eps_next, mode="A", deterministic=False):
    """One SACTrainer.train_from_torch (trainer/trainer.py:126-224).  Mutates ``st``.
    Returns a dict of the intermediates the reference logs (:230-279) plus the
    gradients, for parity checks."""
    obs, actions = batch['observations'], batch['actions']
    next_obs, rewards, terminals = batch['next_observations'], batch['rewards'], batch['terminals']
    B = obs.shape[0]
    out = {}

    # --- policy forward on obs, alpha update (:136-149) ---
    pol = _leafs(st.policy)
    a_pi, mean, log_std, log_pi, std, _ = policy_forward(pol, obs, eps_pi, deterministic)
    if st.auto_alpha:
        la = st.log_alpha['log_alpha'].detach().clone().requires_grad_(True)
        alpha_loss = -(la * (log_pi + st.target_entropy).detach()).mean()
        g_la, = torch.autograd.grad(alpha_loss, la)
        st.opt_alpha.step({'log_alpha': g_la})
        alpha = st.log_alpha['log_alpha'].exp()   # POST-step alpha (:147)
    else:
        alpha_loss = torch.zeros(())
        alpha = torch.zeros(1, dtype=obs.dtype)   # the fork uses 0 (:148-149)

    # --- Q(s, a_pi) with PRE-step weights: values + saved activations (:151-154) ---
    with torch.no_grad():
        q1_new, hs1 = q_forward(st.qf1, obs, a_pi.detach(), return_hidden=True)
        q2_new, hs2 = q_forward(st.qf2, obs, a_pi.detach(), return_hidden=True)
    q_new = torch.min(q1_new, q2_new)
    policy_loss_value = (alpha * log_pi.detach() - q_new).mean()

    # --- Q losses (:168-195) ---
    qf1 = _leafs(st.qf1)
    qf2 = _leafs(st.qf2)
    q1_pred = q_forward(qf1, obs, actions)
    q2_pred = q_forward(qf2, obs, actions)
    with torch.no_grad():
        a_next, _, _, log_pi_next, _, _ = policy_forward(st.policy, next_obs, eps_next, deterministic)
        tq = torch.min(q_forward(st.target_qf1, next_obs, a_next),
                       q_forward(st.target_qf2, next_obs, a_next)) - alpha * log_pi_next
        q_target = st.reward_scale * rewards + (1. - terminals) * st.discount * tq
    qf1_loss = ((q1_pred - q_target) ** 2).mean()
    qf2_loss = ((q2_pred - q_target) ** 2).mean()
    g_q1 = _grads(qf1_loss, qf1)
    g_q2 = _grads(qf2_loss, qf2)
    st.opt_qf1.step(g_q1)   # :200-202
    st.opt_qf2.step(g_q2)   # :204-206

    # --- policy loss backward (:208-210) through the Q graph built at :151-154 ---
    # torch.min backward (1.4): grad goes to the first arg where q1 <= q2.
    sel1 = (q1_new <= q2_new).to(obs.dtype)
    dq = torch.full_like(q_new, -1.0 / B)   # d(policy_loss)/d(q_new)
    w1, w2 = (st.qf1, st.qf2) if mode == "A" else ({k: v.detach() for k, v in qf1.items()},
                                                    {k: v.detach() for k, v in qf2.items()})
    with torch.no_grad():
        g_a = _q_dx_action(w1, hs1, dq * sel1, st.obs_dim, mode == "B") + \
              _q_dx_action(w2, hs2, dq * (1 - sel1), st.obs_dim, mode == "B")
    surrogate = (alpha.detach() * log_pi).mean() + (a_pi * g_a).sum()
    g_pi = _grads(surrogate, pol)
    st.opt_policy.step(g_pi)

    # --- Polyak (:215-224) ---
    if st.n_steps % st.period == 0:
        soft_update(st.qf1, st.target_qf1, st.tau)
        soft_update(st.qf2, st.target_qf2, st.tau)
    st.n_steps += 1

    out.update(q1_pred=q1_pred.detach(), q2_pred=q2_pred.detach(), q_target=q_target,
               log_pi=log_pi.detach(), policy_mean=mean.detach(), policy_log_std=log_std.detach(),
               qf1_loss=qf1_loss.detach(), qf2_loss=qf2_loss.detach(),
               policy_loss=policy_loss_value.detach(), alpha=alpha.detach().clone(),
               alpha_loss=alpha_loss.detach(), a_pi=a_pi.detach(), q_new=q_new,
               grad_qf1=g_q1, grad_qf2=g_q2, grad_policy=g_pi, grad_action=g_a)
    return out


# --------------------------------------------------------------------------
# P-OAC update  (trainer/particle_trainer_oac.py:13-113 ctor, :169-324 step)
# --------------------------------------------------------------------------
class ParticleState(object):
    """ParticleTrainer(OAC) state.  ctor order: SACTrainer.__init__ first (policy, qf1, qf2,
    target_qf1, target_qf2, log_alpha -- all created, the four SAC Q nets then unused,
    :46-58), then per estimator a (qf, tf) pair with last-layer bias
    ``linspace(q_min,q_max,P)`` (:75,:100-113)."""

    def __init__(self, obs_dim, act_dim, n_estimators=10, share_layers=True, hidden=(256, 256),
                 policy_lr=3e-4, qf_lr=3e-4, discount=0.99, reward_scale=1.0, soft_target_tau=5e-3,
                 target_update_period=1, use_automatic_entropy_tuning=True, target_entropy=None,
                 delta=0.95, q_min=0.0, q_max=100.0, counts=False, deterministic=False,
                 dtype=torch.float32, std_soft_update=False, std_soft_update_prob=0.0):
        self.obs_dim, self.act_dim = obs_dim, act_dim
        self.std_soft_update, self.std_soft_update_prob = std_soft_update, std_soft_update_prob
        self.policy = init_policy(obs_dim, act_dim, hidden)
        q_out = n_estimators if share_layers else 1
        for _ in range(4):  # the SAC twin nets, created and dropped (trainer/trainer.py:68-71)
            init_q(obs_dim, act_dim, hidden, q_out)
        self.log_alpha = {'log_alpha': torch.zeros(1)}
        P = n_estimators
        quantiles = [i * 1. / (P - 1) for i in range(P)]
        self.delta_index = next(i for i, q in enumerate(quantiles) if q >= delta)  # :60-67
        init_vals = np.linspace(q_min, q_max, P)
        self.P, self.share_layers = P, share_layers
        self.qfs, self.tfs = [], []
        if share_layers:
            self.qfs.append(init_q(obs_dim, act_dim, hidden, P, bias=init_vals))
            self.tfs.append(init_q(obs_dim, act_dim, hidden, P, bias=init_vals))
        else:
            for i in range(P):
                self.qfs.append(init_q(obs_dim, act_dim, hidden, 1, bias=init_vals[i]))
                self.tfs.append(init_q(obs_dim, act_dim, hidden, 1, bias=init_vals[i]))
        self.discount, self.reward_scale = discount, reward_scale
        self.tau, self.period = soft_target_tau, target_update_period
        self.auto_alpha = use_automatic_entropy_tuning
        self.target_entropy = target_entropy if target_entropy else -float(act_dim)
        self.policy_lr, self.qf_lr = policy_lr, qf_lr
        self.counts, self.deterministic = counts, deterministic
        self.n_steps = 0
        self.to(dtype)

    def to(self, dtype):
        for net in [self.policy, self.log_alpha] + self.qfs + self.tfs:
            for k in net:
                net[k] = net[k].to(dtype)
        self.opt_policy = Adam(self.policy, self.policy_lr)
        self.opt_alpha = Adam(self.log_alpha, self.policy_lr)
        self.opt_qfs = [Adam(q, self.qf_lr) for q in self.qfs]
        return self


def _particles(nets, obs, act, share_layers):
    """:185-191 -- stack to [P,B,1] (shared trunk: one [B,P] net, permuted)."""
    outs = [q_forward(q, obs, act) for q in nets]
    qs = torch.stack(outs, dim=0)
    if share_layers:
        qs = qs.permute(2, 1, 0)
    return qs


def poac_step(st, batch, eps_next, eps_pi):
    """ParticleTrainer.train_from_torch (trainer/particle_trainer_oac.py:169-324).
    NB the noise order is next_obs FIRST (:193), then obs (:271)."""
    obs, actions = batch['observations'], batch['actions']
    next_obs, rewards, terminals = batch['next_observations'], batch['rewards'], batch['terminals']
    out = {}
    qfs = [_leafs(q) for q in st.qfs]
    qs = _particles(qfs, obs, actions, st.share_layers)           # [P,B,1]
    sorted_qs, _ = torch.sort(qs, dim=0)                            # :192
    with torch.no_grad():
        a_next, *_ = policy_forward(st.policy, next_obs, eps_next, st.deterministic)
        tqs = _particles(st.tfs, next_obs, a_next, st.share_layers)
        tq_sorted, _ = torch.sort(tqs, dim=0)                       # :202
        q_target = st.reward_scale * rewards + (1. - terminals) * st.discount * tq_sorted  # :207-208
        if st.std_soft_update:                                      # :210-219
            cur = sorted_qs.detach()
            q_target = st.std_soft_update_prob * q_target + \
                (1 - st.std_soft_update_prob) * (cur - cur.mean(dim=0) + q_target.mean(dim=0))
        if st.counts:                                               # :220-224
            factor = (batch['counts'] == 0).to(obs.dtype)
            sq = sorted_qs.detach()
            q_target = q_target * factor + (1 - factor) * (sq - sq.mean(dim=0) + q_target.mean(dim=0))
    losses = [((sorted_qs[i] - q_target[i]) ** 2).mean() for i in range(st.P)]   # :249,:258
    if st.share_layers:
        g = _grads(sum(losses), qfs[0])
        st.opt_qfs[0].step(g)                                       # :253-257
        out['grad_qf'] = [g]
    else:
        # each optimizer zeroes its grads right before "its" loss (:261-264): net i keeps
        # only d(loss_i)/d(net_i); later losses' contributions are wiped before use.
        gs = [_grads(losses[i], qfs[i], retain_graph=True) for i in range(st.P)]
        for i in range(st.P):
            st.opt_qfs[i].step(gs[i])
        out['grad_qf'] = gs

    # --- policy / alpha (:271-300): fresh forward through the UPDATED Q nets ---
    pol = _leafs(st.policy)
    a_pi, mean, log_std, log_pi, *_ = policy_forward(pol, obs, eps_pi, st.deterministic)
    if st.auto_alpha:
        la = st.log_alpha['log_alpha'].detach().clone().requires_grad_(True)
        alpha_loss = -(la * (log_pi + st.target_entropy).detach()).mean()
        g_la, = torch.autograd.grad(alpha_loss, la)
        st.opt_alpha.step({'log_alpha': g_la})
        alpha = st.log_alpha['log_alpha'].exp().detach()
    else:
        alpha = torch.zeros(1, dtype=obs.dtype)
    pi_qs = _particles(st.qfs, obs, a_pi, st.share_layers)
    q_new = torch.sort(pi_qs, dim=0)[0][0]                          # lowest particle (:294-295)
    policy_loss = (alpha * log_pi - q_new).mean()
    g_pi = _grads(policy_loss, pol)
    st.opt_policy.step(g_pi)
    if st.n_steps % st.period == 0:                                 # :320-324
        for q, t in zip(st.qfs, st.tfs):
            soft_update(q, t, st.tau)
    st.n_steps += 1
    out.update(sorted_qs=sorted_qs.detach(), q_target=q_target, qf_losses=torch.stack(losses).detach(),
               policy_loss=policy_loss.detach(), alpha=alpha.clone(), log_pi=log_pi.detach(),
               policy_mean=mean.detach(), policy_log_std=log_std.detach(), grad_policy=g_pi)
    return out


# --------------------------------------------------------------------------
# G-OAC update  (trainer/gaussian_trainer.py:14-160 ctor, :177-388 step)
# --------------------------------------------------------------------------
def norm_ppf(delta):
    """scipy.stats.norm.ppf (gaussian_trainer.py:68) via torch's erfinv."""
    return float(math.sqrt(2.0) * torch.erfinv(torch.tensor(2.0 * delta - 1.0, dtype=torch.float64)))


class GaussianState(object):
    """GaussianTrainer state, shared 2-head critic (share_layers=True, :92-99) or separate
    mean/std nets (:100-112).  ctor order: SACTrainer.__init__ (policy + 4 dropped Q nets +
    log_alpha), q, q_target, [std, std_target], target_policy (:143)."""

    def __init__(self, obs_dim, act_dim, share_layers=True, hidden=(256, 256), policy_lr=3e-4,
                 qf_lr=3e-4, std_lr=3e-5, discount=0.99, reward_scale=1.0, soft_target_tau=5e-3,
                 target_update_period=1, delta=0.95, q_min=0.0, q_max=100.0, counts=False,
                 dtype=torch.float32):
        self.obs_dim, self.act_dim = obs_dim, act_dim
        self.policy = init_policy(obs_dim, act_dim, hidden)
        q_out = 2 if share_layers else 1
        for _ in range(4):
            init_q(obs_dim, act_dim, hidden, q_out)
        self.log_alpha = {'log_alpha': torch.zeros(1)}   # created, never stepped
        self.standard_bound = norm_ppf(delta)
        mean = (q_max + q_min) / 2
        std = (q_max - q_min) / np.sqrt(12)
        self.std_init = float(std)
        self.share_layers = share_layers
        if share_layers:
            b = np.array([mean, np.log(std)])
            self.q = init_q(obs_dim, act_dim, hidden, 2, bias=b)
            self.q_target = init_q(obs_dim, act_dim, hidden, 2, bias=b)
            self.std = self.std_target = None
        else:
            self.q = init_q(obs_dim, act_dim, hidden, 1, bias=mean)
            self.q_target = init_q(obs_dim, act_dim, hidden, 1, bias=mean)
            self.std = init_q(obs_dim, act_dim, hidden, 1, bias=np.log(std))
            self.std_target = init_q(obs_dim, act_dim, hidden, 1, bias=np.log(std))
        self.target_policy = init_policy(obs_dim, act_dim, hidden)
        self.discount, self.reward_scale = discount, reward_scale
        self.tau, self.period = soft_target_tau, target_update_period
        self.policy_lr, self.qf_lr, self.std_lr = policy_lr, qf_lr, std_lr
        self.counts = counts
        self.n_steps = 0
        self.to(dtype)

    def all_nets(self):
        nets = [self.policy, self.q, self.q_target, self.target_policy, self.log_alpha]
        if not self.share_layers:
            nets += [self.std, self.std_target]
        return nets

    def to(self, dtype):
        for net in self.all_nets():
            for k in net:
                net[k] = net[k].to(dtype)
        self.opt_policy = Adam(self.policy, self.policy_lr)
        self.opt_q = Adam(self.q, self.qf_lr)
        self.opt_target_policy = Adam(self.target_policy, self.policy_lr)
        if not self.share_layers:
            self.opt_std = Adam(self.std, self.std_lr)
        return self


def goac_step(st, batch):
    """GaussianTrainer.train_from_torch with the defaults main.py uses for g-oac:
    deterministic=True (no sampling, no entropy term), mean_update=False,
    use_target_policy=False, ensemble=False (gaussian_trainer.py:177-388)."""
    obs, actions = batch['observations'], batch['actions']
    next_obs, rewards, terminals = batch['next_observations'], batch['rewards'], batch['terminals']
    out = {}
    pos = [False, True]
    with torch.no_grad():
        a_next, *_ = policy_forward(st.policy, next_obs, None, True)          # :203-205
    q = _leafs(st.q)
    if st.share_layers:
        pred = q_forward(q, obs, actions, positive=pos)                        # :189
        std_preds, q_preds = pred[:, 1:2], pred[:, 0:1]                         # :212-213
        with torch.no_grad():
            tq = q_forward(st.q_target, next_obs, a_next, positive=pos)         # :207
            std_target = (1. - terminals) * st.discount * tq[:, 1:2]            # :217
            if st.counts:                                                       # :224-228
                factor = (batch['counts'] == 0).to(obs.dtype)
                std_target = std_target * factor + (1 - factor) * std_preds.detach()
            q_target = st.reward_scale * rewards + (1. - terminals) * st.discount * tq[:, 0:1]
            std_target = torch.clamp(std_target, 0, st.std_init)                # :233
        q_loss = ((q_preds - q_target) ** 2).mean()
        std_loss = ((std_preds - std_target) ** 2).mean()
        g = _grads(q_loss + std_loss, q)
        st.opt_q.step(g)                                                        # :239-241
        out['grad_q'] = g
    else:
        q_preds = q_forward(q, obs, actions)
        with torch.no_grad():
            tq = q_forward(st.q_target, next_obs, a_next)
            q_target = st.reward_scale * rewards + (1. - terminals) * st.discount * tq
        q_loss = ((q_preds - q_target) ** 2).mean()
        g = _grads(q_loss, q)
        st.opt_q.step(g)                                                        # :246-248
        sn = _leafs(st.std)
        std_preds = q_forward(sn, obs, actions, positive=True)                  # :251
        with torch.no_grad():
            std_target = (1. - terminals) * st.discount * \
                q_forward(st.std_target, next_obs, a_next, positive=True)
            if st.counts:
                factor = (batch['counts'] == 0).to(obs.dtype)
                std_target = std_target * factor + (1 - factor) * std_preds.detach()
            std_target = torch.clamp(std_target, 0, st.std_init)                # :267
        std_loss = ((std_preds - std_target) ** 2).mean()
        gs = _grads(std_loss, sn)
        st.opt_std.step(gs)
        out['grad_q'], out['grad_std'] = g, gs

    def ub_of(a):
        if st.share_layers:
            qq = q_forward(st.q, obs, a, positive=pos)
            return qq[:, 0:1], qq[:, 1:2]
        return q_forward(st.q, obs, a), q_forward(st.std, obs, a, positive=True)

    # --- policy on the upper bound (:339-356), through the UPDATED critic ---
    pol = _leafs(st.policy)
    a_pi, mean, log_std, *_ = policy_forward(pol, obs, None, True)
    qv, sv = ub_of(a_pi)
    upper_bound = qv + st.standard_bound * sv
    policy_loss = (-upper_bound).mean()
    g_pi = _grads(policy_loss, pol)
    st.opt_policy.step(g_pi)
    # --- target policy on the mean (:361-373) ---
    tp = _leafs(st.target_policy)
    a_tp, *_ = policy_forward(tp, obs, None, True)
    qv_tp, _ = ub_of(a_tp)
    tp_loss = (-qv_tp).mean()
    g_tp = _grads(tp_loss, tp)
    st.opt_target_policy.step(g_tp)
    if st.n_steps % st.period == 0:                                             # :377-388
        soft_update(st.q, st.q_target, st.tau)
        if not st.share_layers:
            soft_update(st.std, st.std_target, st.tau)
    st.n_steps += 1
    out.update(q_preds=q_preds.detach(), std_preds=std_preds.detach(), q_target=q_target,
               std_target=std_target, q_loss=q_loss.detach(), std_loss=std_loss.detach(),
               policy_loss=(-policy_loss).detach(), target_policy_loss=tp_loss.detach(),
               policy_mean=mean.detach(), policy_log_std=log_std.detach(),
               grad_policy=g_pi, grad_target_policy=g_tp)
    return out


# --------------------------------------------------------------------------
# optimistic exploration  (optimistic_exploration.py:14-196)
# --------------------------------------------------------------------------
def explore(ob, policy, qfs, beta_UB, delta, share_layers=False, eps_sample=None,
            deterministic=False, positive=None, trainer=None, positives=None):
    """get_optimistic_exploration_action (stochastic :14-109 / deterministic :111-196).
    ``ob`` is an unbatched [O] tensor.  The policy's own (discarded) rsample draw at :27 is not
    modelled -- it only advances the RNG.
    ``trainer=None``: dispatch quirk (:41-58): with >=2 nets only qfs[0], qfs[1] are used (twin
    formula); a single multi-head net takes the mean / unbiased-std branch.  The deterministic
    variant always takes the ensemble branch and returns un-squashed mu_E (:181).
    ``trainer=dict(kind=...)`` restates ``Q_UB = trainer.predict(ob[None], a[None], upper_bound=True,
    beta_UB=beta_UB)`` (:38-39 / :135-136):
      kind='sac'      SACTrainer.predict (trainer/trainer.py:105-123): twin formula on trainer.qfs[0:2]
      kind='particle' ParticleTrainer.predict (trainer/particle_trainer_oac.py:147-167): the
                      ``delta_index``-th smallest particle (sort over the particle axis); beta_UB unused
      kind='gaussian' GaussianTrainer.predict(obs, action, std=True) has no ``upper_bound`` keyword
                      (trainer/gaussian_trainer.py:161): the reference raises TypeError, so do we.
    ``positives``: one ``positive`` spec per critic (networks.py:69-75), e.g. [False, True] for G-OAC's
    separate mean / std nets; ``positive`` applies one spec to every critic.
    Returns (action_or_muE, mu_E, grad)."""
    _, mu_T, _, _, std, _ = policy_forward(policy, ob[None], None, True)
    mu_T = mu_T[0].detach().clone().requires_grad_(True)
    std = std[0].detach()
    a = torch.tanh(mu_T)
    pos = positives if positives is not None else [positive] * len(qfs)
    qf = lambda i: q_forward(qfs[i], ob[None], a[None], positive=pos[i])
    Q_UB = None
    if trainer is not None:
        kind = trainer['kind']
        if kind == 'sac':
            Q1, Q2 = qf(0), qf(1)
            mu_Q = (Q1 + Q2) / 2.0
            sigma_Q = torch.abs(Q1 - Q2) / 2.0
        elif kind == 'particle':
            qs = torch.stack([qf(i) for i in range(len(qfs))], dim=0)
            if trainer.get('share_layers', share_layers):
                qs = qs.permute(2, 1, 0)
            Q_UB = torch.sort(qs, dim=0)[0][trainer['delta_index']]
        else:
            raise TypeError("predict() got an unexpected keyword argument 'upper_bound'")
    elif (not deterministic) and len(qfs) >= 2:
        Q1, Q2 = qf(0), qf(1)
        mu_Q = (Q1 + Q2) / 2.0
        sigma_Q = torch.abs(Q1 - Q2) / 2.0
    else:
        qs = torch.stack([qf(i) for i in range(len(qfs))], dim=0)
        if share_layers:
            qs = qs.permute(2, 1, 0)
        mu_Q = torch.mean(qs, dim=0)
        sigma_Q = torch.std(qs, dim=0)
    if Q_UB is None:
        Q_UB = mu_Q + beta_UB * sigma_Q
    grad, = torch.autograd.grad(Q_UB.sum(), mu_T)
    if deterministic:
        denom = torch.sqrt(torch.sum(grad ** 2)) + 10e-6          # :160-164
        mu_E = mu_T.detach() + math.sqrt(2.0 * delta) * grad / denom
        return mu_E, mu_E, grad
    Sigma = std ** 2
    denom = torch.sqrt(torch.sum(grad ** 2 * Sigma)) + 10e-6      # :76-80
    mu_E = mu_T.detach() + math.sqrt(2.0 * delta) * Sigma * grad / denom   # :83-87
    if eps_sample is None:
        eps_sample = torch.normal(torch.zeros_like(mu_E), torch.ones_like(std))
    ac = torch.tanh(mu_E + std * eps_sample)                       # TanhNormal.sample :162-173
    return ac, mu_E, grad
