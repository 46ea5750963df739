"""Import the UNMODIFIED reference (amarildolikmeta/oac-explore) in-process.

TEST INFRASTRUCTURE ONLY.  Usable only where ``/root/reference`` exists (the
build container); it does not travel to the GPU box.  It is used by
``tests/golden/make_golden.py`` to generate the committed golden vectors and by
``tests/test_oracle_vs_reference.py`` (skipped when the reference is absent) to
pin ``oracle/oac_oracle.py`` against the real code.

Nothing under ``oac_explore_b200/`` imports this file.

The reference needs three third-party modules that are not installed here
(SURVEY.md section 8c): ``gym`` (replay_buffer.py:3, utils/env_utils.py:3-4,
envs/__init__.py:1), ``matplotlib.pyplot`` (utils/core.py:2) and ``gtimer``
(rl_algorithm.py:8).  Tiny stand-ins are injected into ``sys.modules`` before
the import; no reference file is edited or copied.

Semantics patch ("Mode A", torch-1.4-literal): the reference pins torch==1.4.0
(requirements.txt:4) where ``Adam.step`` wrote ``p.data`` without bumping the
autograd version counter, so ``policy_loss.backward()`` at
trainer/trainer.py:209 silently used the *post-step* Q weights for the dX
products.  On torch>=1.5 the same line raises.  ``mode_a(trainer)`` reproduces
the 1.4 behaviour by wrapping each optimizer's ``step`` in
``torch.autograd._unsafe_preserve_version_counter``.
"""
import contextlib
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("OAC_REFERENCE_ROOT", "/root/reference")
# the byte-compiled copy of the same modules (oracle/build_ref.py): what travels to the GPU box
COMPILED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "trainer", "trainer.py"))


def compiled_available():
    from oracle import build_ref
    return build_ref.available()


def import_root():
    """Where the reference is imported from: its sources when present, else oracle/_ref (compiled from them)."""
    if reference_available():
        return REFERENCE_ROOT
    if compiled_available():
        return COMPILED_ROOT
    return None


def _install_shims():
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")

        class Env(object):
            metadata = {}

            def seed(self, seed=None):
                return [seed]

        class Space(object):
            pass

        class Box(Space):
            def __init__(self, low, high, shape=None, dtype=np.float32):
                if shape is None:
                    low = np.asarray(low, dtype=dtype)
                    high = np.asarray(high, dtype=dtype)
                    shape = low.shape
                else:
                    low = np.full(shape, low, dtype=dtype)
                    high = np.full(shape, high, dtype=dtype)
                self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

            def sample(self):
                return np.random.uniform(self.low, self.high).astype(self.dtype)

        class Discrete(Space):
            def __init__(self, n):
                self.n = n
                self.shape = ()

        class Tuple(Space):
            def __init__(self, spaces):
                self.spaces = spaces

        spaces = types.ModuleType("gym.spaces")
        spaces.Box, spaces.Discrete, spaces.Tuple, spaces.Space = Box, Discrete, Tuple, Space
        utils = types.ModuleType("gym.utils")
        seeding = types.ModuleType("gym.utils.seeding")

        def np_random(seed=None):
            return np.random.RandomState(seed), seed

        seeding.np_random = np_random

        class EzPickle(object):
            def __init__(self, *a, **k):
                pass

        utils.seeding, utils.EzPickle = seeding, EzPickle
        envs = types.ModuleType("gym.envs")
        registration = types.ModuleType("gym.envs.registration")
        registration.register = lambda *a, **k: None
        envs.registration = registration
        gym.Env, gym.spaces, gym.utils, gym.envs = Env, spaces, utils, envs
        gym.make = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("gym shim: no gym.make"))
        gym.Wrapper = Env
        for name, mod in (("gym", gym), ("gym.spaces", spaces), ("gym.utils", utils),
                          ("gym.utils.seeding", seeding), ("gym.envs", envs),
                          ("gym.envs.registration", registration)):
            sys.modules[name] = mod
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if "gtimer" not in sys.modules:
        gt = types.ModuleType("gtimer")
        gt.timed_for = lambda it, **k: it
        gt.stamp = lambda *a, **k: None

        class _Times(object):
            class stamps(object):
                itrs = {}
            total = 0.0
        gt.get_times = lambda: _Times
        sys.modules["gtimer"] = gt


class _CompiledFinder(object):
    """sys.meta_path finder for the byte-compiled reference modules under oracle/_ref (``<path>.bin`` = .pyc bytes)."""

    def __init__(self, root):
        self.root = root

    def find_spec(self, fullname, path=None, target=None):
        import importlib.machinery
        import importlib.util
        rel = fullname.replace(".", os.sep)
        pkg = os.path.join(self.root, rel, "__init__.bin")
        mod = os.path.join(self.root, rel + ".bin")
        if os.path.isfile(pkg):
            loader = importlib.machinery.SourcelessFileLoader(fullname, pkg)
            return importlib.util.spec_from_file_location(fullname, pkg, loader=loader,
                                                          submodule_search_locations=[os.path.dirname(pkg)])
        if os.path.isfile(mod):
            loader = importlib.machinery.SourcelessFileLoader(fullname, mod)
            return importlib.util.spec_from_file_location(fullname, mod, loader=loader)
        return None


_REF = None


def load_reference():
    """Returns a namespace holding the reference's hot-path modules."""
    global _REF
    if _REF is not None:
        return _REF
    root = import_root()
    if root is None:
        raise RuntimeError("reference not present at %s and oracle/_ref not built" % REFERENCE_ROOT)
    _install_shims()
    # The reference uses top-level module names (``trainer``, ``utils``, ...),
    # so its root has to lead sys.path while it is imported.
    compiled = root == COMPILED_ROOT
    if compiled:
        finder = _CompiledFinder(root)
        sys.meta_path.insert(0, finder)      # stays installed: the reference also imports lazily inside functions
    else:
        sys.path.insert(0, root)
    try:
        import importlib
        ns = types.SimpleNamespace()
        ns.ptu = importlib.import_module("utils.pytorch_util")
        ns.core = importlib.import_module("utils.core")
        ns.networks = importlib.import_module("networks")
        ns.policies = importlib.import_module("trainer.policies")
        ns.trainer = importlib.import_module("trainer.trainer")
        ns.particle_trainer_oac = importlib.import_module("trainer.particle_trainer_oac")
        ns.gaussian_trainer = importlib.import_module("trainer.gaussian_trainer")
        ns.replay_buffer = importlib.import_module("replay_buffer")
        ns.optimistic_exploration = importlib.import_module("optimistic_exploration")
        ns.gym = sys.modules["gym"]
    finally:
        if not compiled:
            sys.path.remove(root)
    ns.root = root
    ns.ptu.set_gpu_mode(False)
    _REF = ns
    return ns


def make_spaces(obs_dim, act_dim):
    ref = load_reference()
    Box = ref.gym.spaces.Box
    return Box(-np.inf, np.inf, shape=(obs_dim,)), Box(-1.0, 1.0, shape=(act_dim,))


def make_producers(obs_dim, act_dim, hidden=(256, 256), q_out=1):
    """Same objects main.py:44-106 builds (get_policy_producer / get_q_producer)."""
    ref = load_reference()

    def policy_producer(deterministic=False, bias=None, **_):
        p = ref.policies.TanhGaussianPolicy(obs_dim=obs_dim, action_dim=act_dim,
                                            hidden_sizes=list(hidden), bias=bias, std=None)
        if deterministic:
            p = ref.policies.MakeDeterministic(p)
        return p

    def q_producer(bias=None, positive=False, train_bias=True):
        return ref.networks.FlattenMlp(input_size=obs_dim + act_dim, output_size=q_out,
                                       hidden_sizes=list(hidden), bias=bias, positive=positive,
                                       train_bias=train_bias)

    return policy_producer, q_producer


def mode_a(trainer):
    """Make every optimizer of ``trainer`` step like torch 1.4 (no version bump)."""
    import torch
    seen = set()
    for name, opt in list(vars(trainer).items()):
        opts = opt if isinstance(opt, (list, tuple)) else [opt]
        for o in opts:
            if not isinstance(o, torch.optim.Optimizer) or id(o) in seen:
                continue
            seen.add(id(o))
            params = [p for g in o.param_groups for p in g["params"]]
            orig = o.step

            def step(*a, _orig=orig, _params=params, **k):
                with torch.autograd._unsafe_preserve_version_counter(tuple(_params)):
                    return _orig(*a, **k)

            o.step = step
    return trainer


@contextlib.contextmanager
def injected_noise(eps_list):
    """Replace the N(0,1) draws of ``TanhNormal.rsample`` (trainer/policies.py:179-187)
    with the given tensors, in call order, so fp32/fp64/our kernels see one noise."""
    ref = load_reference()
    import torch
    queue = list(eps_list)
    cls = ref.policies.TanhNormal
    orig = cls.rsample

    def rsample(self, return_pretanh_value=False):
        eps = queue.pop(0).to(self.normal_mean.dtype)
        z = self.normal_mean + self.normal_std * eps
        z.requires_grad_()
        if return_pretanh_value:
            return torch.tanh(z), z
        return torch.tanh(z)

    cls.rsample = rsample
    try:
        yield
    finally:
        cls.rsample = orig
