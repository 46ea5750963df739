"""ctypes binding of liboac_b200.so (include/oac_b200.h).

The CUDA library is the product: there is no CPU fallback.  Importing this module
raises if the shared object is missing, and every call raises ``RuntimeError`` with
``oac_last_error_string()`` on a non-zero return code.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboac_b200.so")

OAC_MAX_NETS = 48
ALGO_SAC, ALGO_POAC, ALGO_GOAC = 0, 1, 2
GEMM_FP32, GEMM_TF32, GEMM_TF32X3 = 0, 1, 2
NET_POLICY, NET_Q, NET_SCALAR = 0, 1, 2
EXPLORE_TWIN, EXPLORE_ENSEMBLE, EXPLORE_QUANTILE = 0, 1, 2
# counters (oac_internal.h)
CNT_TRAIN_STEPS, CNT_OPT0 = 0, 1
CNT_RNG_LO, CNT_RNG_HI = 22, 23      # per-seed 64-bit key mixed into the Philox stream of the step's rsample noise
SC_ALPHA, SC_ALPHA_LOSS, SC_MEAN_LOGPI = 0, 1, 2


class OacConfig(C.Structure):
    _fields_ = [
        ("algo", C.c_int32), ("obs_dim", C.c_int32), ("act_dim", C.c_int32), ("hidden", C.c_int32),
        ("batch", C.c_int32), ("n_seeds", C.c_int32), ("n_particles", C.c_int32),
        ("share_layers", C.c_int32), ("deterministic", C.c_int32), ("auto_alpha", C.c_int32),
        ("counts", C.c_int32), ("train_bias", C.c_int32), ("stale_graph_mode", C.c_int32),
        ("target_update_period", C.c_int32), ("gemm_path", C.c_int32), ("std_soft_update", C.c_int32),
        ("discount", C.c_float), ("reward_scale", C.c_float), ("soft_target_tau", C.c_float),
        ("policy_lr", C.c_float), ("qf_lr", C.c_float), ("std_lr", C.c_float),
        ("target_entropy", C.c_float), ("standard_bound", C.c_float), ("std_init", C.c_float),
        ("adam_beta1", C.c_float), ("adam_beta2", C.c_float), ("adam_eps", C.c_float),
        ("std_soft_update_prob", C.c_float), ("reserved2", C.c_float),
        ("rng_seed", C.c_uint64),
    ]


class OacNetLayout(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("in_dim", C.c_int32), ("in_ld", C.c_int32), ("hidden", C.c_int32),
        ("n_out", C.c_int32), ("trainable", C.c_int32),
        ("off_w0", C.c_int64), ("off_b0", C.c_int64), ("off_w1", C.c_int64), ("off_b1", C.c_int64),
        ("off_w2", C.c_int64), ("off_b2", C.c_int64), ("size", C.c_int64),
    ]


class OacLayout(C.Structure):
    _fields_ = [
        ("n_nets", C.c_int32), ("n_trainable", C.c_int32),
        ("nets", OacNetLayout * OAC_MAX_NETS),
        ("param_floats", C.c_int64), ("adam_floats", C.c_int64), ("work_floats", C.c_int64),
        ("io_floats", C.c_int64),
        ("n_counters", C.c_int32), ("x_rows", C.c_int32), ("x_ld", C.c_int32),
        ("off_x", C.c_int64), ("off_rewards", C.c_int64), ("off_terminals", C.c_int64),
        ("off_counts", C.c_int64), ("off_eps", C.c_int64), ("off_log_pi", C.c_int64),
        ("off_mean", C.c_int64), ("off_log_std", C.c_int64), ("off_q_pred", C.c_int64),
        ("off_q_target", C.c_int64), ("off_q_new", C.c_int64), ("off_scalars", C.c_int64),
        ("nq", C.c_int32), ("reserved1", C.c_int32),
    ]


class OacBuffers(C.Structure):
    _fields_ = [("params", C.c_void_p), ("adam_m", C.c_void_p), ("adam_v", C.c_void_p),
                ("work", C.c_void_p), ("io", C.c_void_p), ("counters", C.c_void_p), ("host_scalars", C.c_void_p)]


class OacReplayStore(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("next_obs", C.c_void_p), ("actions", C.c_void_p),
                ("rewards", C.c_void_p), ("terminals", C.c_void_p), ("counts", C.c_void_p),
                ("capacity", C.c_int64), ("obs_dim", C.c_int32), ("act_dim", C.c_int32)]


class OacBatchDst(C.Structure):
    _fields_ = [("x", C.c_void_p), ("x_ld", C.c_int32), ("obs_blocks", C.c_int32 * 3),
                ("act_block", C.c_int32), ("next_block", C.c_int32),
                ("rewards", C.c_void_p), ("terminals", C.c_void_p), ("counts", C.c_void_p),
                ("n_seeds", C.c_int32), ("reserved", C.c_int32), ("seed_stride", C.c_int64)]


class OacExploreArgs(C.Structure):
    _fields_ = [
        ("policy", C.c_void_p), ("policy_lay", OacNetLayout),
        ("q", C.c_void_p * OAC_MAX_NETS), ("q_lay", OacNetLayout), ("n_q", C.c_int32),
        ("mode", C.c_int32), ("deterministic", C.c_int32), ("quantile_index", C.c_int32),
        ("exp_mask", C.c_uint32), ("beta_UB", C.c_float), ("delta", C.c_float),
        ("n_obs", C.c_int32), ("obs", C.c_void_p), ("eps", C.c_void_p),
        ("rng_seed", C.c_uint64), ("rng_offset", C.c_uint64),
        ("action", C.c_void_p), ("mu_E", C.c_void_p), ("grad", C.c_void_p),
        ("obs_group", C.c_void_p), ("group_stride", C.c_int64),
    ]


# every symbol include/oac_b200.h declares
EXPORTS = [
    "oac_last_error_string", "oac_abi_version",
    "oac_replay_gather", "oac_replay_gather_dense", "oac_replay_add",
    "oac_trainer_layout", "oac_trainer_create", "oac_trainer_destroy", "oac_trainer_step",
    "oac_trainer_launches_per_step", "oac_trainer_ws_stages", "oac_trainer_stats", "oac_trainer_stats_count", "oac_trainer_profile", "oac_gemm_debug", "oac_gemm_debug_kernel",
    "oac_policy_forward", "oac_q_forward", "oac_explore",
]

_lib = None


def lib():
    """Loads the shared library once; raises loudly when it is missing or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            "liboac_b200.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    for name in EXPORTS:
        if not hasattr(L, name):
            raise RuntimeError("liboac_b200.so lacks symbol %s (stale build?)" % name)
    L.oac_last_error_string.restype = C.c_char_p
    L.oac_abi_version.restype = C.c_int
    vp, i32, i64, u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32
    L.oac_replay_gather.argtypes = [C.POINTER(OacReplayStore), vp, i32, C.POINTER(OacBatchDst), vp]
    L.oac_replay_gather_dense.argtypes = [C.POINTER(OacReplayStore), vp, i32, vp, vp, vp, vp, vp, vp, vp]
    L.oac_replay_add.argtypes = [vp, vp, vp, vp, vp, vp, i64, i32, i32, vp, i32, i64, vp]
    L.oac_trainer_layout.argtypes = [C.POINTER(OacConfig), C.POINTER(OacLayout)]
    L.oac_trainer_create.argtypes = [C.POINTER(OacConfig), C.POINTER(OacBuffers), C.POINTER(vp)]
    L.oac_trainer_destroy.argtypes = [vp]
    L.oac_trainer_step.argtypes = [vp, i32, vp]
    L.oac_trainer_launches_per_step.argtypes = [vp]
    L.oac_trainer_ws_stages.argtypes = [vp]
    L.oac_trainer_stats_count.argtypes = [vp]
    L.oac_trainer_stats.argtypes = [vp, vp, i32, vp]
    L.oac_trainer_profile.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp]
    L.oac_gemm_debug.argtypes = [i32, i32, i32, i32, i32, i32, vp, i32, vp, i32, vp, i32, vp, i32, vp]
    L.oac_policy_forward.argtypes = [vp, C.POINTER(OacNetLayout), vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    L.oac_q_forward.argtypes = [vp, C.POINTER(OacNetLayout), vp, i32, i32, u32, vp, vp]
    L.oac_explore.argtypes = [C.POINTER(OacExploreArgs), vp]
    if L.oac_abi_version() != 4:
        raise RuntimeError("liboac_b200.so ABI version mismatch")
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        msg = lib().oac_last_error_string().decode("utf-8", "replace")
        raise RuntimeError("%s failed (code %d): %s" % (what or "liboac_b200 call", rc, msg))


def ptr(t):
    """Raw device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


_RAW_STREAM = None


def current_stream():
    """The current CUDA stream as a cudaStream_t.  torch.cuda.current_stream() builds a Stream object through several
    layers of Python (~5 us, twice per update on the hot path); the raw getter behind it costs a fraction of that."""
    global _RAW_STREAM
    import torch
    if _RAW_STREAM is None:
        raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)
        dev = getattr(torch._C, "_cuda_getDevice", None)
        _RAW_STREAM = (raw, dev) if (raw is not None and dev is not None) else False
    if _RAW_STREAM:
        return C.c_void_p(_RAW_STREAM[0](_RAW_STREAM[1]()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
