"""Recipe for ``oracle/_ref``: the reference's OWN hot-path modules, byte-compiled from the sources where they lie.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see the header of oracle/oac_oracle.py for who may use ``oracle/``).

The reference is pure Python, so "building" it is ``py_compile``: every module on the path named by
BASELINE.json's north_star (and the modules those import) is compiled from ``/root/reference/<path>.py`` to
``oracle/_ref/<path>.bin`` (the bytes of a CPython .pyc; the extension is .bin because the GPU-box snapshot drops
``*.pyc`` files -- oracle/ref_import.py loads them through importlib's SourcelessFileLoader).  No reference SOURCE is
copied into the repository:
``oracle/_ref/`` holds compiled artefacts only, is listed in .gitignore (never in history) and not in .gpurunignore,
so it travels to the GPU box like our own built ``.so``.  There ``bench.py --impl reference`` and the ``cpu_baseline``
leg import it (through oracle/ref_import.py, with the gym / matplotlib / gtimer stand-ins and the torch-1.4 "mode A"
optimizer patch) and time the UNMODIFIED reference code: ReplayBuffer.random_batch -> np_to_pytorch_batch ->
SACTrainer.train_from_torch.

    python oracle/build_ref.py            # needs /root/reference; run by __graft_entry__.build() when it is present
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SRC = os.environ.get("OAC_REFERENCE_ROOT", "/root/reference")

# the hot path and its import closure
MODULES = [
    "replay_buffer.py", "networks.py", "optimistic_exploration.py",
    "trainer/__init__.py", "trainer/trainer.py", "trainer/policies.py", "trainer/mixture_same_family.py",
    "trainer/particle_trainer_oac.py", "trainer/gaussian_trainer.py",
    "utils/__init__.py", "utils/core.py", "utils/pytorch_util.py", "utils/eval_util.py", "utils/pythonplusplus.py",
    "utils/misc.py", "utils/env_utils.py",
]


def build(force=False):
    if not os.path.isfile(os.path.join(SRC, "trainer", "trainer.py")):
        return False
    stamp = os.path.join(OUT, "PYTHON_TAG")
    tag = sys.implementation.cache_tag
    if not force and os.path.isfile(stamp) and open(stamp).read().strip() == tag and \
            all(os.path.isfile(os.path.join(OUT, m[:-3] + ".bin")) for m in MODULES):
        return True
    for m in MODULES:
        dst = os.path.join(OUT, m[:-3] + ".bin")
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(os.path.join(SRC, m), cfile=dst, dfile=m, doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(stamp, "w") as f:
        f.write(tag + "\n")
    return True


def available():
    """True when the compiled reference is usable by THIS interpreter."""
    stamp = os.path.join(OUT, "PYTHON_TAG")
    return os.path.isfile(stamp) and open(stamp).read().strip() == sys.implementation.cache_tag and \
        os.path.isfile(os.path.join(OUT, "trainer", "trainer.bin"))


if __name__ == "__main__":
    ok = build(force=True)
    print("oracle/_ref: %s" % ("built %d modules" % len(MODULES) if ok else "reference sources not found at " + SRC))
