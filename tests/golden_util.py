"""Loading helpers for tests/golden/*.npz (written by tests/golden/make_golden.py)."""
import os
from collections import OrderedDict

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name)))


def digest(x):
    x = np.asarray(x.detach().cpu() if isinstance(x, torch.Tensor) else x, dtype=np.float64).ravel()
    return np.array([x.sum(), (x * x).sum(), x[0], x[x.size // 2], x[-1]])


def net_from(g, prefix):
    """state_dict-like OrderedDict of the arrays stored under ``prefix/``."""
    out = OrderedDict()
    for k in g:
        if k.startswith(prefix + '/'):
            out[k[len(prefix) + 1:]] = torch.from_numpy(np.array(g[k]))
    return out


def set_net(dst, src):
    for k in dst:
        dst[k].copy_(src[k].to(dst[k].dtype).reshape(dst[k].shape))


def assert_digest_close(x, d, rtol, what):
    got = digest(x)
    # sum can cancel: scale its tolerance by the L2 mass
    scale = np.sqrt(max(d[1], 1e-30) * np.asarray(x.detach().cpu() if isinstance(x, torch.Tensor) else x).size)
    assert abs(got[0] - d[0]) <= rtol * scale + 1e-12, "%s sum %r vs %r" % (what, got[0], d[0])
    assert abs(got[1] - d[1]) <= 2 * rtol * abs(d[1]) + 1e-20, "%s sumsq %r vs %r" % (what, got[1], d[1])
    for i in (2, 3, 4):
        assert abs(got[i] - d[i]) <= rtol * max(abs(d[i]), np.sqrt(d[1] / max(1, np.asarray(got).size))) + 1e-6, \
            "%s elem %r vs %r" % (what, got[i], d[i])
