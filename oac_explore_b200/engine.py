"""Device arenas + handle for one (or a group of) trainer(s).

PyTorch owns every byte (SURVEY.md section 8b "Ownership"): the parameter arena, the two
Adam moment arenas, the activation workspace, the IO slice (batch rows, noise, per-sample
outputs) and the int32 step counters are plain CUDA tensors; the library borrows their
pointers.  ``net_views`` exposes the arena as ``state_dict``-shaped tensor views
(``fc0.weight`` ... ``last_fc_log_std.bias``) so snapshots and parity checkers see the
reference's layout.
"""
import ctypes as C
from collections import OrderedDict

import torch

from . import _lib
from ._lib import OacConfig, OacLayout, OacBuffers


def make_config(algo, obs_dim, act_dim, hidden, batch, n_seeds=1, n_particles=0, share_layers=False,
                deterministic=False, auto_alpha=True, counts=False, train_bias=True, stale_graph_mode="A",
                target_update_period=1, gemm_path=_lib.GEMM_FP32, discount=0.99, reward_scale=1.0,
                soft_target_tau=1e-2, policy_lr=1e-3, qf_lr=1e-3, std_lr=3e-5, target_entropy=None,
                standard_bound=0.0, std_init=0.0, betas=(0.9, 0.999), adam_eps=1e-8, rng_seed=0,
                std_soft_update=False, std_soft_update_prob=0.0):
    c = OacConfig()
    c.algo, c.obs_dim, c.act_dim, c.hidden, c.batch, c.n_seeds = algo, obs_dim, act_dim, hidden, batch, n_seeds
    c.n_particles, c.share_layers, c.deterministic = n_particles, int(share_layers), int(deterministic)
    c.auto_alpha, c.counts, c.train_bias = int(auto_alpha), int(counts), int(train_bias)
    c.stale_graph_mode = 0 if stale_graph_mode in ("A", 0) else 1
    c.target_update_period, c.gemm_path = target_update_period, gemm_path
    c.discount, c.reward_scale, c.soft_target_tau = discount, reward_scale, soft_target_tau
    c.policy_lr, c.qf_lr, c.std_lr = policy_lr, qf_lr, std_lr
    c.target_entropy = float(target_entropy) if target_entropy is not None else -float(act_dim)
    c.standard_bound, c.std_init = standard_bound, std_init
    c.adam_beta1, c.adam_beta2, c.adam_eps = betas[0], betas[1], adam_eps
    c.rng_seed = rng_seed
    c.std_soft_update, c.std_soft_update_prob = int(bool(std_soft_update)), float(std_soft_update_prob)
    return c


class Engine(object):
    def __init__(self, cfg, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("oac_explore_b200 needs a CUDA device (B200, sm_100a); no CPU fallback exists")
        self.lib = _lib.lib()
        self.cfg = cfg
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.lay = OacLayout()
        _lib.check(self.lib.oac_trainer_layout(C.byref(cfg), C.byref(self.lay)), "oac_trainer_layout")
        L, S = self.lay, cfg.n_seeds
        z = lambda n: torch.zeros((S, max(int(n), 4)), dtype=torch.float32, device=self.device)
        self.params, self.adam_m, self.adam_v = z(L.param_floats), z(L.adam_floats), z(L.adam_floats)
        self.work, self.io = z(L.work_floats), z(L.io_floats)
        self.counters = torch.zeros((S, L.n_counters), dtype=torch.int32, device=self.device)
        # the step also stores its scalars (alpha, alpha loss, mean log pi, its own 1-based step number) into this mapped
        # pinned host tensor, slot = parity of the steps taken before it: valid once the step has completed (an event or a
        # stream synchronisation), and still intact while the NEXT step is in flight (``step_scalars``)
        self.host_scalars = torch.zeros((2, S, 16), dtype=torch.float32).pin_memory()
        self._host_scalars_np = self.host_scalars.numpy()
        self.steps = 0               # host mirror of the device's train-step counter (CNT_TRAIN_STEPS, equal for all seeds)
        buf = OacBuffers(_lib.ptr(self.params), _lib.ptr(self.adam_m), _lib.ptr(self.adam_v),
                         _lib.ptr(self.work), _lib.ptr(self.io), _lib.ptr(self.counters), _lib.ptr(self.host_scalars))
        h = C.c_void_p()
        _lib.check(self.lib.oac_trainer_create(C.byref(cfg), C.byref(buf), C.byref(h)), "oac_trainer_create")
        self.handle = h
        self.B, self.O, self.A, self.H = cfg.batch, cfg.obs_dim, cfg.act_dim, cfg.hidden
        self.launches_per_step = self.lib.oac_trainer_launches_per_step(self.handle)
        self.ws_stages = self.lib.oac_trainer_ws_stages(self.handle)
        self.n_stats = self.lib.oac_trainer_stats_count(self.handle)
        self._host_stats = None

    def __del__(self):
        h = getattr(self, "handle", None)
        if h is not None and h.value:
            self.lib.oac_trainer_destroy(h)
            self.handle = None

    # ---- views -----------------------------------------------------------------
    def net_layout(self, i):
        return self.lay.nets[i]

    def net_views(self, i, seed=0, arena=None):
        """OrderedDict of state_dict-named views of net ``i`` (arena: params | adam_m | adam_v)."""
        n = self.lay.nets[i]
        base = (self.params if arena is None else arena)[seed]
        H = n.hidden
        out = OrderedDict()
        if n.kind == _lib.NET_SCALAR:
            out['log_alpha'] = base[n.off_w0:n.off_w0 + 1]
            return out
        out['fc0.weight'] = base[n.off_w0:n.off_w0 + H * n.in_ld].view(H, n.in_ld)[:, :n.in_dim]
        out['fc0.bias'] = base[n.off_b0:n.off_b0 + H]
        out['fc1.weight'] = base[n.off_w1:n.off_w1 + H * H].view(H, H)
        out['fc1.bias'] = base[n.off_b1:n.off_b1 + H]
        w2 = base[n.off_w2:n.off_w2 + n.n_out * H].view(n.n_out, H)
        b2 = base[n.off_b2:n.off_b2 + n.n_out]
        if n.kind == _lib.NET_POLICY:
            A = n.n_out // 2
            out['last_fc.weight'], out['last_fc.bias'] = w2[:A], b2[:A]
            out['last_fc_log_std.weight'], out['last_fc_log_std.bias'] = w2[A:], b2[A:]
        else:
            out['last_fc.weight'], out['last_fc.bias'] = w2, b2
        return out

    def net_base_ptr(self, seed=0):
        return self.params[seed].data_ptr()

    def io_view(self, off, shape, seed=0):
        n = 1
        for s in shape:
            n *= s
        return self.io[seed, off:off + n].view(*shape)

    def x_block(self, block, seed=0):
        L = self.lay
        x = self.io_view(L.off_x, (L.x_rows, L.x_ld), seed)
        return x[block * self.B:(block + 1) * self.B]

    def scalars(self, seed=0):
        return self.io_view(self.lay.off_scalars, (16,), seed)

    # ---- batch upload (train_from_torch path) ---------------------------------------
    def load_batch(self, obs, actions, rewards, terminals, next_obs, counts=None, seed=0, with_tp=False):
        """Device-to-device copy of five fp32 tensors into the X blocks / IO slots."""
        O, A = self.O, self.A
        self.x_block(1, seed)[:, :O].copy_(obs)
        xb2 = self.x_block(2, seed)
        xb2[:, :O].copy_(obs)
        xb2[:, O:O + A].copy_(actions)
        self.x_block(3, seed)[:, :O].copy_(next_obs)
        if with_tp:
            self.x_block(0, seed)[:, :O].copy_(obs)
        L = self.lay
        self.io_view(L.off_rewards, (self.B,), seed).copy_(rewards.reshape(-1))
        self.io_view(L.off_terminals, (self.B,), seed).copy_(terminals.reshape(-1))
        if counts is not None:
            self.io_view(L.off_counts, (self.B,), seed).copy_(counts.reshape(-1))

    def set_eps(self, eps_obs, eps_next, seed=0):
        e = self.io_view(self.lay.off_eps, (2, self.B, self.A), seed)
        if eps_obs is not None:
            e[0].copy_(eps_obs)
        if eps_next is not None:
            e[1].copy_(eps_next)

    def step(self, external_eps=False):
        rc = self.lib.oac_trainer_step(self.handle, 1 if external_eps else 0, _lib.current_stream())
        if rc:
            _lib.check(rc, "oac_trainer_step")
        self.steps += 1

    def set_train_steps(self, n):
        """Set the device's train-step counter of every seed (snapshot restore) together with its host mirror."""
        self.counters[:, _lib.CNT_TRAIN_STEPS] = int(n)
        self.steps = int(n)

    def step_scalars(self, step=None):
        """[n_seeds, 16] numpy view of the scalars written by the 1-based train step ``step`` (default: the last one
        enqueued): 0 alpha, 1 alpha loss, 2 mean log pi, 3 the step number.  The caller waits for that step first (event
        or stream synchronisation); one later step may be in flight meanwhile (two slots)."""
        step = self.steps if step is None else step
        return self._host_scalars_np[(step - 1) & 1]

    def stats_device(self, out=None):
        """[n_seeds, n_stats] fp32 DEVICE tensor: the ``eval_statistics`` vector of every seed for the last step, reduced
        by one kernel (``oac_trainer_stats``; key order in include/oac_b200.h).  The all-gather payload."""
        if out is None:
            out = torch.empty((self.cfg.n_seeds, self.n_stats), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.oac_trainer_stats(self.handle, _lib.ptr(out), out.stride(0), _lib.current_stream()),
                   "oac_trainer_stats")
        return out

    def stats_host(self):
        """The same vector as a [n_seeds, n_stats] numpy array: the kernel stores straight into mapped pinned host memory,
        so this is one launch + one stream synchronisation (no device-to-host copy calls)."""
        if self._host_stats is None:
            self._host_stats = torch.zeros((self.cfg.n_seeds, self.n_stats), dtype=torch.float32).pin_memory()
        self.stats_device(self._host_stats)
        torch.cuda.current_stream().synchronize()
        return self._host_stats.numpy().copy()

    def profile(self, iters=20):
        """Per-stage mean milliseconds (stage-by-stage launches, CUDA events): [(name, ms, is_gemm, flops)]."""
        n_max = 64
        ms = (C.c_float * n_max)()
        isg = (C.c_int32 * n_max)()
        fl = (C.c_double * n_max)()
        names = (C.c_char_p * n_max)()
        n = C.c_int32(0)
        _lib.check(self.lib.oac_trainer_profile(self.handle, iters, n_max, ms, isg, fl, names, C.byref(n),
                                                _lib.current_stream()), "oac_trainer_profile")
        self.steps = int(self.counters[0, _lib.CNT_TRAIN_STEPS].item())     # the profile loop ran whole steps
        return [(names[i].decode(), float(ms[i]), bool(isg[i]), float(fl[i])) for i in range(n.value)]
