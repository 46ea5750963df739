"""CPU checks of the C-ABI library: it loads, exports every symbol include/oac_b200.h declares, and its
memory plan matches the reference's parameter counts.  No compute call (no GPU here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "oac_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(oac_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from oac_explore_b200 import _lib
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), n
    assert set(_lib.EXPORTS) == set(names)
    assert L.oac_abi_version() == 4


def test_ctypes_structs_match_header_sizes():
    from oac_explore_b200 import _lib
    # field-by-field mirrors of the C structs (natural alignment on both sides)
    assert C.sizeof(_lib.OacNetLayout) == 6 * 4 + 7 * 8
    assert C.sizeof(_lib.OacConfig) == 16 * 4 + 14 * 4 + 8
    assert C.sizeof(_lib.OacBuffers) == 7 * 8
    # OacExploreArgs: policy*, layout, q[48], layout, 8 x 32-bit (n_q .. n_obs), obs*, eps*, 2 u64, 3 out*, obs_group*, group_stride
    lay = C.sizeof(_lib.OacNetLayout)
    assert C.sizeof(_lib.OacExploreArgs) == 8 + lay + 48 * 8 + lay + 8 * 4 + 2 * 8 + 2 * 8 + 3 * 8 + 2 * 8


@pytest.mark.parametrize("algo,kw,n_nets,n_train", [
    (0, {}, 6, 4), (1, dict(n_particles=10, share_layers=True), 4, 3),
    (1, dict(n_particles=10, share_layers=False), 22, 12), (2, dict(share_layers=True), 5, 4),
    (2, dict(share_layers=False), 7, 5)])
def test_layout_matches_reference_parameter_counts(algo, kw, n_nets, n_train):
    from oac_explore_b200 import _lib
    from oac_explore_b200.engine import make_config
    O, A, H, B = 376, 17, 256, 256
    cfg = make_config(algo, O, A, H, B, **kw)
    lay = _lib.OacLayout()
    _lib.check(_lib.lib().oac_trainer_layout(C.byref(cfg), C.byref(lay)))
    assert lay.n_nets == n_nets and lay.n_trainable == n_train
    live = 0
    for i in range(lay.n_nets):
        n = lay.nets[i]
        if n.kind == _lib.NET_SCALAR:
            continue
        live_i = H * n.in_dim + H + H * H + H + n.n_out * H + n.n_out
        live += live_i
        assert n.in_ld % 4 == 0 and n.in_ld >= n.in_dim and n.size >= live_i
        for off in (n.off_w0, n.off_b0, n.off_w1, n.off_b1, n.off_w2, n.off_b2):
            assert off % 4 == 0                      # 16-byte aligned blocks (cp.async / float4)
    if algo == 0:
        # Q net 166 913 parameters, policy 171 042
        assert live == 171042 + 4 * 166913
    assert lay.x_rows == 4 * B and lay.x_ld % 4 == 0 and lay.x_ld >= O + A
    assert lay.adam_floats < lay.param_floats


def test_invalid_config_is_reported_not_crashed():
    from oac_explore_b200 import _lib
    from oac_explore_b200.engine import make_config
    cfg = make_config(1, 376, 17, 256, 256, n_particles=40, share_layers=True)
    lay = _lib.OacLayout()
    rc = _lib.lib().oac_trainer_layout(C.byref(cfg), C.byref(lay))
    assert rc != 0 and b"n_particles" in _lib.lib().oac_last_error_string()
    with pytest.raises(RuntimeError):
        _lib.check(rc, "oac_trainer_layout")


def test_product_path_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from oac_explore_b200.engine import Engine, make_config
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(make_config(0, 11, 3, 32, 32))
    from oac_explore_b200.networks import get_q_producer
    q = get_q_producer(11, 3, [32, 32])()
    with pytest.raises(RuntimeError):
        q(torch.zeros(2, 11), torch.zeros(2, 3))


def test_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "oac_explore_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
            assert "oracle" not in src, f
