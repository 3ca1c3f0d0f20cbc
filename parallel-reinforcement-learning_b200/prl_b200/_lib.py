"""ctypes binding of libprl_b200.so (the C ABI declared in include/prl_b200.h).

There is no CPU fallback: if the shared library is missing, or a call returns a non-zero status, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PRL_B200_LIB") or os.path.join(_HERE, "libprl_b200.so")   # (PRL_B200_LIB: A/B experiments with another build)
TEST_LIB_PATH = os.path.join(_HERE, "libprl_b200_test.so")

ENV_IDS = {"CartPole-v1": 0, "Pendulum-v1": 1, "Acrobot-v1": 2, "MountainCar-v0": 3, "MountainCarContinuous-v0": 4}
ACT_I32, ACT_I64, ACT_F32 = 0, 1, 2

_vp, _i32, _i64, _u64, _u32, _f32, _f64, _sz = (C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_uint32, C.c_float,
                                                C.c_double, C.c_size_t)
_ip = C.POINTER(C.c_int)

# name -> (restype, argtypes); every symbol include/prl_b200.h declares
PROTOTYPES = {
    "prl_last_error": (C.c_char_p, []),
    "prl_version": (_i32, []),
    "prl_env_info": (_i32, [_i32, _ip, _ip, _ip, _ip, _ip]),
    "prl_policy_param_count": (_i64, [_i32, _i32, _i32]),
    "prl_rnd_param_count": (_i64, [_i32, _i32]),
    "prl_scan_ws_bytes": (_sz, [_i64]),
    "prl_test_sincos": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "prl_test_pow2": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "prl_test_philox": (_i32, [_u64, _u32, _u32, _u32, _u32, _vp, _vp]),
    "prl_test_umma": (_i32, [_i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "prl_env_reset": (_i32, [_i32, _i32, _u64, _u64, _vp, _vp, _vp, _vp, _vp]),
    "prl_env_reset_numpy": (_i32, [_i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "prl_test_pcg64": (_i32, [_vp, _i32, _i32, _vp, _vp]),
    "prl_env_set_state": (_i32, [_i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "prl_env_get_state": (_i32, [_i32, _i32, _vp, _vp, _vp]),
    "prl_env_step": (_i32, [_i32, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "prl_compact_indices": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _sz, _vp]),
    "prl_gather_rows": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "prl_mask_update": (_i32, [_vp, _vp, _vp, _i32, _vp]),
    "prl_buffer_append": (_i32, [_i32, _i32, _i32, _vp, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "prl_buffer_transfer": (_i32, [_i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp,
                                   _vp, _sz, _vp]),
    "prl_buffer_transfer_ex": (_i32, [_i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp,
                                      _vp, _sz, _vp]),
    "prl_policy_act": (_i32, [_vp, _i32, _i32, _i32, _f32, _vp, _vp, _i64, _u64, _u64, _vp, _vp, _vp]),
    "prl_policy_evaluate": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "prl_rollout": (_i32, [_i32, _i32, _i32, _vp, _f32, _u64, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "prl_rollout_eval": (_i32, [_i32, _i32, _i32, _vp, _f32, _u64, _u64] + [_vp] * 12 + [_i32, _vp, _vp]),
    "prl_rollout_score_ws_doubles": (_sz, [_i32]),
    "prl_gae": (_i32, [_vp, _vp, _vp, _vp, _f64, _f64, _i64, _vp, _vp, _sz, _vp]),
    "prl_gae_ws_bytes": (_sz, [_i64]),
    "prl_gae_columns": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _f64, _f64, _vp, _vp]),
    "prl_adv_normalize": (_i32, [_vp, _vp, _i64, _vp, _vp, _i32, _vp]),
    "prl_update_ws_floats": (_sz, [_i32, _i32, _i32, _i64]),
    "prl_ppo_grad": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i64, _f32, _f32, _vp, _vp, _vp, _sz, _vp]),
    "prl_ppo_grad_tc_supported": (_i32, [_i32, _i32, _i32]),
    "prl_update_tc_ws_floats": (_sz, [_i32, _i32, _i32, _i64]),
    "prl_ppo_grad_tc": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i64, _f32, _f32, _vp, _vp, _vp, _sz, _vp]),
    "prl_ppo_grad_tc_status": (_i32, [_vp, _vp, _vp]),
    "prl_ppo_step_tc": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i64, _f32, _f32, _vp, _vp, _vp, _vp, _vp, _f32, _f32, _f32,
                               _vp, _vp, _sz, _vp]),
    "prl_ppo_step_tc_p2p": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i64, _f32, _f32, _vp, _vp, _vp, _vp, _vp, _f32, _f32,
                                   _f32, _vp, _vp, _i32, _i32, _vp, _sz, _vp]),
    "prl_p2p_exchange_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "prl_p2p_alloc": (_i32, [_sz, C.POINTER(_vp)]),
    "prl_p2p_free": (_i32, [_vp]),
    "prl_p2p_get_handle": (_i32, [_vp, C.c_char_p]),
    "prl_p2p_open_handle": (_i32, [C.c_char_p, C.POINTER(_vp)]),
    "prl_p2p_close_handle": (_i32, [_vp]),
    "prl_adamw_step_dev": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _f32, _f32, _f32, _vp, _vp]),
    "prl_adamw_step": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _f32, _f32, _f32, _vp, _vp]),
    "prl_rnd_intrinsic": (_i32, [_vp, _vp, _i32, _i32, _vp, _i64, _f32, _vp, _vp, _vp]),
    "prl_rnd_grad": (_i32, [_vp, _vp, _i32, _i32, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
}

TEST_FUNCTIONS = ("prl_test_sincos", "prl_test_pow2", "prl_test_philox", "prl_test_umma", "prl_test_pcg64")   # libprl_b200_test.so

_lib = None
_test_lib = None


class PrlError(RuntimeError):
    pass


def load_test_library(path: str | None = None):
    """The parity-test hooks (include/prl_b200_test.h): a separate library, never loaded by the product path."""
    global _test_lib
    if _test_lib is None:
        path = path or TEST_LIB_PATH
        if not os.path.exists(path):
            raise PrlError(f"{path} is missing - run `python __graft_entry__.py` (build()) first")
        lib = C.CDLL(path)
        for name in TEST_FUNCTIONS:
            fn_ = getattr(lib, name)
            fn_.restype, fn_.argtypes = PROTOTYPES[name]
        lib.prl_test_last_error.restype = C.c_char_p
        _test_lib = lib
    return _test_lib


def load_library(path: str | None = None):
    """dlopen the in-tree shared library and attach prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise PrlError(f"{path} is missing - run `python __graft_entry__.py` (build()) first; there is no CPU fallback")
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        if name in TEST_FUNCTIONS:
            continue
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def require_cuda():
    """The torch device the kernels run on.  Raises (no CPU fallback) when the library or a CUDA device is missing."""
    import torch

    load_library()
    if not torch.cuda.is_available():
        raise PrlError("prl_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


# kernels launched by one call of each entry point (csrc/*.cu) - used to count launches, e.g. by bench.py
KERNELS_PER_CALL = {
    "prl_test_sincos": 1, "prl_test_pow2": 1, "prl_test_philox": 1, "prl_test_umma": 1, "prl_env_reset": 1, "prl_env_reset_numpy": 1, "prl_test_pcg64": 1, "prl_env_set_state": 1,
    "prl_env_get_state": 1, "prl_env_step": 1, "prl_compact_indices": 2, "prl_gather_rows": 1, "prl_mask_update": 1,
    "prl_buffer_append": 1, "prl_buffer_transfer": 4, "prl_buffer_transfer_ex": 4, "prl_policy_act": 1, "prl_policy_evaluate": 1, "prl_rollout": 1, "prl_rollout_eval": 1,
    "prl_gae": 1, "prl_gae_columns": 1, "prl_adv_normalize": 1, "prl_ppo_grad": 2, "prl_ppo_grad_tc": 2, "prl_ppo_step_tc": 1, "prl_ppo_step_tc_p2p": 1, "prl_adamw_step": 1, "prl_adamw_step_dev": 1,
    "prl_rnd_intrinsic": 1, "prl_rnd_grad": 2,
}
CALL_COUNTS: dict[str, int] = {}
_profile = None  # name -> [(start event, end event)] while profile_calls() is active


def profile_calls(enable: bool):
    """Time every entry-point call with CUDA events on the current (= launching) stream.  Returns the previous
    record: {name: [(torch.cuda.Event, torch.cuda.Event), ...]}."""
    global _profile
    prev = _profile
    _profile = {} if enable else None
    return prev


def call(name: str, *args):
    """Invoke a status-returning entry point; raise PrlError with prl_last_error() on failure."""
    is_test = name in TEST_FUNCTIONS
    lib = load_test_library() if is_test else load_library()
    CALL_COUNTS[name] = CALL_COUNTS.get(name, 0) + 1
    import torch

    if _profile is not None and not torch.cuda.is_current_stream_capturing():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        _profile.setdefault(name, []).append((e0, e1))
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        raise PrlError(f"{name} failed ({rc}): {(lib.prl_test_last_error() if is_test else lib.prl_last_error()).decode()}")


def launches(counts: dict[str, int] | None = None) -> int:
    counts = CALL_COUNTS if counts is None else counts
    return sum(KERNELS_PER_CALL.get(k, 1) * v for k, v in counts.items())


def fn(name: str):
    return getattr(load_library(), name)


def env_info(env_id: str):
    v = [C.c_int() for _ in range(5)]
    call("prl_env_info", ENV_IDS[env_id], *[C.byref(x) for x in v])
    S, O, A, cont, ms = [x.value for x in v]
    return dict(S=S, O=O, A=A, continuous=bool(cont), max_steps=ms, code=ENV_IDS[env_id])
