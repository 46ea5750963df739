"""GPU-resident replay buffer: drop-in for the reference's ``replay_buffer.py``.

Same constructor and methods (``add_sample``, ``add_path(s)``, ``random_batch``,
``num_steps_can_sample``, ``get_dataset``, ``get_diagnostics``, ``end_epoch``,
``get_snapshot``, ``restore_from_snapshot``; replay_buffer.py:8-148, 151-203).  The store is
struct-of-arrays fp32 in HBM (the f64->f32 cast of utils/core.py:45 happens once, at insert);
``random_batch`` draws its indices from the SAME global numpy stream as the reference
(``np.random.randint(0, size, B)``, replay_buffer.py:107) and gathers them with one coalesced
kernel (``oac_replay_gather``), straight into the attached trainer's batch rows when there is
one.  Without an attached trainer it returns the reference's numpy dict (f64 / uint8), so the
reference's own trainers keep working.
"""
import ctypes as C
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from ._lib import OacReplayStore, OacBatchDst
from .networks import _device


def get_dim(space):
    """utils/env_utils.py:8-18 (Box / Discrete / Tuple / flat_dim)."""
    if hasattr(space, 'low') and hasattr(space.low, 'size'):
        return space.low.size
    if hasattr(space, 'n'):
        return space.n
    if hasattr(space, 'spaces'):
        return sum(get_dim(s) for s in space.spaces)
    if hasattr(space, 'flat_dim'):
        return space.flat_dim
    raise TypeError("Unknown space: {}".format(space))


class ReplayBuffer(object):
    STAGE_ROWS = 4096

    def __init__(self, max_replay_buffer_size, ob_space, action_space):
        self._ob_space, self._action_space = ob_space, action_space
        O, A = get_dim(ob_space), get_dim(action_space)
        N = int(max_replay_buffer_size)
        self._ob_dim, self._ac_dim = O, A
        self._max_replay_buffer_size = N
        self._lib = _lib.lib()
        dev = _device()
        self._device = dev
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
        self._observations, self._next_obs, self._actions = z(N, O), z(N, O), z(N, A)
        self._rewards, self._terminals = z(N, 1), z(N, 1)
        self._counts_dev = None
        self._top = 0
        self._size = 0
        # host staging for add_sample: rows = obs | action | reward | terminal | next_obs
        self._W = 2 * O + A + 2
        # two pinned staging buffers (+ their device twins): while the copy / scatter of one is in flight the host
        # fills the other; a buffer is only rewritten after the event recorded behind its scatter has completed
        self._stages = [torch.zeros((self.STAGE_ROWS, self._W), dtype=torch.float32).pin_memory() for _ in range(2)]
        self._stages_dev = [torch.zeros((self.STAGE_ROWS, self._W), dtype=torch.float32, device=dev) for _ in range(2)]
        self._stage_events = [None, None]
        self._stage_cur = 0
        self._stage = self._stages[0]
        self._stage_np = self._stage.numpy()
        self._n_staged = 0
        self._stage_top = 0
        # ``random_batch`` with an attached trainer returns VIEWS of the trainer's batch rows (the next call overwrites
        # them in place; the reference returns fresh arrays).  Set True when the caller keeps batches, e.g.
        # rl_algorithm.py:163-165 ``save_sampled_data``: every returned tensor is then a private clone.
        self.return_copies = False
        self._idx_host = None
        self._idx_dev = None
        self._trainer = None
        self._desc_cache = {}
        self._view_cache = None

    # ---- insert path (replay_buffer.py:50-104) -------------------------------------
    def add_path(self, path):
        """replay_buffer.py:57-82, without the per-sample Python loop: the whole path is packed into the pinned staging
        rows with five vectorised numpy stores ([T, 2O+A+2] fp32: obs | action | reward | terminal | next_obs) and goes to
        the device as one copy + one scatter kernel (ring wrap and count zeroing on the device)."""
        if hasattr(self._action_space, 'n') and not hasattr(self._action_space, 'low'):
            raise AssertionError("discrete action spaces are not supported (replay_buffer.py:91)")
        O, A = self._ob_dim, self._ac_dim
        T = len(path["observations"])
        obs = np.asarray(path["observations"], dtype=np.float64).reshape(T, O)
        act = np.asarray(path["actions"], dtype=np.float64).reshape(T, A)
        rew = np.asarray(path["rewards"], dtype=np.float64).reshape(T)
        nob = np.asarray(path["next_observations"], dtype=np.float64).reshape(T, O)
        term = np.asarray(path["terminals"]).reshape(T).astype(np.uint8)        # uint8 store (replay_buffer.py:45)
        t = 0
        while t < T:
            if self._n_staged == 0:
                self._stage_top = self._top
            n = min(T - t, self.STAGE_ROWS - self._n_staged)
            rows = self._stage_np[self._n_staged:self._n_staged + n]
            rows[:, :O] = obs[t:t + n]
            rows[:, O:O + A] = act[t:t + n]
            rows[:, O + A] = rew[t:t + n]
            rows[:, O + A + 1] = term[t:t + n]
            rows[:, O + A + 2:] = nob[t:t + n]
            self._n_staged += n
            self._top = (self._top + n) % self._max_replay_buffer_size
            self._size = min(self._size + n, self._max_replay_buffer_size)
            t += n
            if self._n_staged == self.STAGE_ROWS:
                self._flush()

    def add_paths(self, paths):
        for path in paths:
            self.add_path(path)

    def add_sample(self, observation, action, reward, next_observation, terminal, env_info=None, **kwargs):
        if hasattr(self._action_space, 'n') and not hasattr(self._action_space, 'low'):
            raise AssertionError("discrete action spaces are not supported (replay_buffer.py:91)")
        if self._n_staged == 0:
            self._stage_top = self._top
        O, A = self._ob_dim, self._ac_dim
        row = self._stage_np[self._n_staged]
        row[:O] = np.asarray(observation, dtype=np.float64).reshape(-1)
        row[O:O + A] = np.asarray(action, dtype=np.float64).reshape(-1)
        row[O + A] = np.asarray(reward, dtype=np.float64).reshape(-1)[0]
        row[O + A + 1] = np.uint8(np.asarray(terminal).reshape(-1)[0])   # uint8 store (replay_buffer.py:45)
        row[O + A + 2:] = np.asarray(next_observation, dtype=np.float64).reshape(-1)
        self._n_staged += 1
        self._advance()
        if self._n_staged == self.STAGE_ROWS:
            self._flush()

    def _advance(self):
        self._top = (self._top + 1) % self._max_replay_buffer_size
        if self._size < self._max_replay_buffer_size:
            self._size += 1

    def _flush(self):
        n = self._n_staged
        if n == 0:
            return
        stream = torch.cuda.current_stream()
        k = self._stage_cur
        dev = self._stages_dev[k]
        dev[:n].copy_(self._stage[:n], non_blocking=True)
        _lib.check(self._lib.oac_replay_add(
            _lib.ptr(self._observations), _lib.ptr(self._next_obs), _lib.ptr(self._actions),
            _lib.ptr(self._rewards), _lib.ptr(self._terminals), _lib.ptr(self._counts_dev),
            self._max_replay_buffer_size, self._ob_dim, self._ac_dim, _lib.ptr(dev), n,
            self._stage_top, C.c_void_p(stream.cuda_stream)), "oac_replay_add")
        ev = self._stage_events[k] or torch.cuda.Event()
        ev.record(stream)
        self._stage_events[k] = ev
        # switch to the other staging buffer; wait only if ITS previous flush is still in flight (no host stall otherwise)
        k ^= 1
        if self._stage_events[k] is not None:
            self._stage_events[k].synchronize()
        self._stage_cur = k
        self._stage = self._stages[k]
        self._stage_np = self._stage.numpy()
        self._n_staged = 0

    # ---- sample path (replay_buffer.py:106-115) ----------------------------------------
    def attach(self, trainer):
        """Later ``random_batch`` calls gather straight into ``trainer``'s batch rows."""
        self._trainer = trainer

    def _store(self):
        return OacReplayStore(_lib.ptr(self._observations), _lib.ptr(self._next_obs), _lib.ptr(self._actions),
                              _lib.ptr(self._rewards), _lib.ptr(self._terminals), _lib.ptr(self._counts_dev),
                              self._max_replay_buffer_size, self._ob_dim, self._ac_dim)

    def _draw_indices(self, batch_size):
        return np.random.randint(0, self._size, batch_size)

    # The host runs ahead of the device (nothing on this path synchronises), so the indices go through a ring of pinned
    # slots; a slot group is only reused after the gathers that read it have completed (one event per group).  Every slot
    # has a device twin, filled by an asynchronous copy on a COPY STREAM of its own while the previous update still runs:
    # the gather then reads its indices from HBM (reading them in place from the pinned slot -- unified addressing -- puts a
    # PCIe round trip at the head of every step's dependency chain; that form is still used when the stream is idle, i.e.
    # when there is nothing to overlap with, and always with ``zero_copy_indices = True``).
    _RING_GROUPS, _RING_GROUP_SLOTS = 4, 256
    zero_copy_indices = False

    def _upload_indices(self, indices):
        n = len(indices)
        if self._idx_host is None or self._idx_host.shape[1] < n:
            if self._idx_host is not None:
                # queued gather kernels may still read the old pinned ring by pointer (torch's host allocator does not
                # know about that use): drain the device before the ring is dropped
                torch.cuda.synchronize()
            slots = self._RING_GROUPS * self._RING_GROUP_SLOTS
            self._idx_host = torch.zeros((slots, n), dtype=torch.int64).pin_memory()
            self._idx_np = self._idx_host.numpy()
            self._idx_dev = torch.zeros((slots, n), dtype=torch.int64, device=self._device)
            self._idx_copy_stream = torch.cuda.Stream(device=self._device)
            self._idx_copied = [torch.cuda.Event() for _ in range(8)]
            self._idx_slot = 0
            self._idx_events = [None] * self._RING_GROUPS
        slot = self._idx_slot
        g, in_g = divmod(slot, self._RING_GROUP_SLOTS)
        if in_g == 0 and self._idx_events[g] is not None:
            self._idx_events[g].synchronize()               # the device is done with this group's previous round
        self._idx_np[slot, :n] = indices
        self._idx_last_slot = slot
        self._idx_slot = (slot + 1) % (self._RING_GROUPS * self._RING_GROUP_SLOTS)
        main = torch.cuda.current_stream()
        if self.zero_copy_indices or main.query():
            # nothing is running that the copy could overlap (a caller that synchronises after every update): the gather
            # reads the pinned slot in place, which costs one PCIe round trip but no copy launch + cross-stream wait
            return self._idx_host[slot]
        ev = self._idx_copied[slot & 7]
        with torch.cuda.stream(self._idx_copy_stream):
            self._idx_dev[slot].copy_(self._idx_host[slot], non_blocking=True)
            ev.record(self._idx_copy_stream)
        main.wait_event(ev)
        return self._idx_dev[slot]

    def _indices_consumed(self):
        """Call after the kernel reading the last uploaded slot has been launched."""
        g, in_g = divmod(self._idx_last_slot, self._RING_GROUP_SLOTS)
        if in_g == self._RING_GROUP_SLOTS - 1:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self._idx_events[g] = ev

    def _gather_desc(self, engine, seed, n_seeds):
        key = (id(engine), seed, n_seeds)
        d = self._desc_cache.get(key)
        if d is None:
            L = engine.lay
            dst = OacBatchDst()
            dst.x = engine.io[seed, L.off_x:].data_ptr()
            dst.x_ld = L.x_ld
            goac = engine.cfg.algo == _lib.ALGO_GOAC
            dst.obs_blocks[0], dst.obs_blocks[1], dst.obs_blocks[2] = 1, 2, (0 if goac else -1)
            dst.act_block, dst.next_block = 2, 3
            dst.rewards = engine.io[seed, L.off_rewards:].data_ptr()
            dst.terminals = engine.io[seed, L.off_terminals:].data_ptr()
            dst.counts = engine.io[seed, L.off_counts:].data_ptr() if self._counts_dev is not None else None
            dst.n_seeds, dst.seed_stride = n_seeds, L.io_floats
            d = (self._store(), dst, engine)          # keeps the engine alive while cached
            self._desc_cache = {key: d}
        return d

    def gather_into(self, engine, indices_dev, batch_size, seed=0, n_seeds=1):
        st, dst, _ = self._gather_desc(engine, seed, n_seeds)
        rc = self._lib.oac_replay_gather(C.byref(st), C.c_void_p(indices_dev.data_ptr()), batch_size, C.byref(dst),
                                         _lib.current_stream())
        if rc:
            _lib.check(rc, "oac_replay_gather")

    def _resident_views(self, e, batch_size):
        key = (id(e), batch_size)
        if self._view_cache is None or self._view_cache[0] != key:
            O, A = self._ob_dim, self._ac_dim
            xb2, L = e.x_block(2), e.lay
            batch = dict(observations=xb2[:, :O], actions=xb2[:, O:O + A],
                         rewards=e.io_view(L.off_rewards, (batch_size, 1)),
                         terminals=e.io_view(L.off_terminals, (batch_size, 1)),
                         next_observations=e.x_block(3)[:, :O], _oac_resident=e)
            if self._counts_dev is not None:
                batch['counts'] = e.io_view(L.off_counts, (batch_size, 1))
            self._view_cache = (key, batch)
        return dict(self._view_cache[1])

    def random_batch(self, batch_size):
        if self._n_staged:
            self._flush()
        indices = self._draw_indices(batch_size)
        idx_dev = self._upload_indices(indices)
        tr = self._trainer
        if tr is not None:
            tr._ensure_engine(batch_size)
            e = tr._engine
            self.gather_into(e, idx_dev, batch_size)
            self._indices_consumed()
            batch = self._resident_views(e, batch_size)
            if self.return_copies:
                batch = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
                batch.pop('_oac_resident')             # clones are not the trainer's rows any more: normal upload path
            return batch
        out = self.gather_dense(idx_dev, batch_size)
        self._indices_consumed()
        return self._numpy_batch(out)

    def gather_dense(self, idx_dev, batch_size):
        """Five dense fp32 device tensors (the "fast" return type of SURVEY.md section 8b)."""
        O, A, dev = self._ob_dim, self._ac_dim, self._device
        mk = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        out = dict(observations=mk(batch_size, O), actions=mk(batch_size, A), rewards=mk(batch_size, 1),
                   terminals=mk(batch_size, 1), next_observations=mk(batch_size, O))
        if self._counts_dev is not None:
            out['counts'] = mk(batch_size, 1)
        st = self._store()
        _lib.check(self._lib.oac_replay_gather_dense(
            C.byref(st), _lib.ptr(idx_dev), batch_size, _lib.ptr(out['observations']), _lib.ptr(out['actions']),
            _lib.ptr(out['rewards']), _lib.ptr(out['terminals']), _lib.ptr(out['next_observations']),
            _lib.ptr(out.get('counts')), _lib.current_stream()), "oac_replay_gather_dense")
        return out

    @staticmethod
    def _numpy_batch(dev_batch):
        """The reference's return type: float64 arrays, uint8 terminals (replay_buffer.py:32-45)."""
        out = {}
        for k, v in dev_batch.items():
            a = v.to('cpu').numpy()
            out[k] = a.astype(np.uint8) if k == 'terminals' else a.astype(np.float64)
        return out

    # ---- misc (replay_buffer.py:117-148) -----------------------------------------------
    def get_dataset(self):
        self._flush()
        return self._observations[:self._size].to('cpu').numpy().astype(np.float64)

    def num_steps_can_sample(self):
        return self._size

    def get_diagnostics(self):
        return OrderedDict([('size', self._size)])

    def end_epoch(self, epoch):
        return

    _SNAP = ('_observations', '_next_obs', '_actions', '_rewards', '_terminals')

    def get_snapshot(self):
        self._flush()
        ss = {k: getattr(self, k).to('cpu').numpy() for k in self._SNAP}
        ss['_terminals'] = ss['_terminals'].astype(np.uint8)
        ss['_top'], ss['_size'] = self._top, self._size
        return ss

    def restore_from_snapshot(self, ss):
        self._n_staged = 0
        for key in ss.keys():
            assert hasattr(self, key) or key == '_counts'
            if key in ('_top', '_size'):
                setattr(self, key, int(ss[key]))
            elif key == '_counts':
                self._counts_dev.copy_(torch.as_tensor(np.asarray(ss[key], dtype=np.float32)).to(self._device))
            else:
                dst = getattr(self, key)
                dst.copy_(torch.as_tensor(np.asarray(ss[key], dtype=np.float32)).to(self._device).reshape(dst.shape))


class ReplayBufferCount(ReplayBuffer):
    """replay_buffer.py:151-203: per-slot sample counts, returned with the batch and bumped
    once per distinct sampled slot; ``add_sample`` zeroes the slot."""

    def __init__(self, max_replay_buffer_size, ob_space, action_space, priority_sample=False):
        super().__init__(max_replay_buffer_size, ob_space, action_space)
        self._counts_dev = torch.zeros((int(max_replay_buffer_size), 1), dtype=torch.float32, device=self._device)
        self.priority_sample = priority_sample

    @property
    def _counts(self):
        self._flush()
        return self._counts_dev.to('cpu').numpy().astype(np.float64)

    def _draw_indices(self, batch_size):
        if self.priority_sample:
            # replay_buffer.py:181-185 (host-side: needs the whole count vector)
            probs = 1 / (self._counts[:self._size] + 1)
            probs /= probs.sum()
            return np.random.choice(np.arange(self._size), size=batch_size, p=probs[:, 0])
        return np.random.randint(0, self._size, batch_size)

    def get_snapshot(self):
        ss = super().get_snapshot()
        ss['_counts'] = self._counts
        return ss
