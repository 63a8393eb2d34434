"""ctypes binding of libdmp.so (include/dmp.h).  There is NO fallback: if the CUDA library is
missing or does not export the expected ABI, importing this module raises."""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libdmp.so")

ABI_VERSION = 4
OK, EINVAL, ECUDA = 0, 1, 2
OBS_F32, OBS_F64, OBS_I16, OBS_REC, OBS_BITS = 0, 1, 2, 3, 4
BITS_REWARDS = (0.0, 1.0, 5.0, 10.0, -1.0, -100.0, 0.0, 0.0)     # reward code of a DMP_OBS_BITS trailer -> value
REC_DONE, REC_SATURATED = 1, 2
F_AUTORESET, F_NORMALISE, F_TSLOT1, F_NO_L2_HINT = 1, 2, 4, 8
F_NO_PDL, F_TILE_LDST, F_GENERIC, F_ROLLOUT_K1 = 16, 32, 64, 128
F_RESET_OBS = 256
# tuning switches by name (BatchedDMPEnv(tuning=...)); decided once on the host, passed as DmpIO.flags bits
TUNING_FLAGS = {"no_l2_hint": F_NO_L2_HINT, "no_pdl": F_NO_PDL, "tile_ldst": F_TILE_LDST, "generic": F_GENERIC,
                "rollout_k1": F_ROLLOUT_K1}
PLAN_PHILOX, PLAN_SEQUENTIAL, PLAN_KEEP = 0, 1, 2
ACT_UNIFORM, ACT_REF3D = 0, 1
ERR_ACTION, ERR_STEPSIZE, ERR_PLANIDX, ERR_OVERFLOW = 1, 2, 4, 8
T_INIT = 0xFFFFFFFFFFFFFFFF
SEED_DEFAULT = 0x534E4143


class DmpState(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("dynamic", C.c_int32), ("n_plans", C.c_int32), ("total_step", C.c_int32),
        ("plan_mode", C.c_int32), ("action_dist", C.c_int32),
        ("n_envs", C.c_int64), ("env_base", C.c_int64), ("seed", C.c_uint64), ("t", C.c_uint64), ("t_dev", C.c_void_p),
        ("cells", C.c_void_p), ("aux", C.c_void_p), ("plans", C.c_void_p), ("plan_total", C.c_void_p),
        ("ep_cnt", C.c_void_p), ("ep_len", C.c_void_p), ("ep_ret", C.c_void_p), ("ep_iou", C.c_void_p),
        ("err", C.c_void_p),
    ]


class DmpIO(C.Structure):
    _fields_ = [
        ("actions", C.c_void_p), ("step_sizes", C.c_void_p), ("next_plan", C.c_void_p),
        ("obs", C.c_void_p), ("reward", C.c_void_p), ("done", C.c_void_p),
        ("obs_kind", C.c_int32), ("flags", C.c_int32),
    ]


class DmpLayout(C.Structure):
    _fields_ = [
        ("cells_bytes", C.c_int64), ("aux_bytes", C.c_int64), ("plan_row_bytes", C.c_int64),
        ("obs_dim", C.c_int32), ("n_actions", C.c_int32), ("grid_rows", C.c_int32), ("grid_cols", C.c_int32),
        ("total_step_static", C.c_int32), ("total_step_dynamic", C.c_int32),
        ("rec_bytes", C.c_int32), ("bits_bytes", C.c_int32),
    ]


_P = C.c_void_p
_SIGNATURES = {
    "dmp_abi_version": (C.c_int, []),
    "dmp_last_error": (C.c_int, []),
    "dmp_layout": (C.c_int, [C.c_int, C.c_int64, C.POINTER(DmpLayout)]),
    "dmp_plan_static": (C.c_int, [C.c_int, C.c_int, _P, _P, _P]),
    "dmp_plans_pack": (C.c_int, [C.c_int, _P, C.c_int, _P, _P, _P]),
    "dmp_plans_generate": (C.c_int, [C.c_int, C.c_int, C.c_uint64, C.c_int64, C.c_int, _P, C.c_int, _P, _P, _P, _P, _P]),
    "dmp_plans_from_state": (C.c_int, [C.POINTER(DmpState), _P, _P, _P]),
    "dmp_reset": (C.c_int, [C.POINTER(DmpState), _P, _P, C.c_uint64, _P, C.c_int, _P]),
    "dmp_step": (C.c_int, [C.POINTER(DmpState), C.POINTER(DmpIO), _P]),
    "dmp_rollout": (C.c_int, [C.POINTER(DmpState), C.POINTER(DmpIO), C.c_int, _P]),
    "dmp_records_unpack": (C.c_int, [C.c_int, C.c_int, _P, C.c_int64, _P, C.c_int, _P, _P, _P, _P]),
    "dmp_iou": (C.c_int, [C.POINTER(DmpState), _P, _P]),
    "dmp_stats_scratch_bytes": (C.c_int64, [C.c_int64]),
    "dmp_stats_reduce": (C.c_int, [C.POINTER(DmpState), _P, _P, _P]),
    "dmp_stats_clear": (C.c_int, [C.POINTER(DmpState), _P]),
    "dmp_export_state": (C.c_int, [C.POINTER(DmpState), _P, _P, _P, _P]),
    "dmp_import_state": (C.c_int, [C.POINTER(DmpState), _P, _P, _P, _P]),
    "dmp_stage_move": (C.c_int, [C.POINTER(DmpState), C.POINTER(DmpIO), _P, _P]),
    "dmp_stage_deposit": (C.c_int, [C.POINTER(DmpState), C.POINTER(DmpIO), _P, _P]),
    "dmp_stage_observe": (C.c_int, [C.POINTER(DmpState), C.POINTER(DmpIO), _P, _P]),
    "dmp_stage_reward": (C.c_int, [C.POINTER(DmpState), C.POINTER(DmpIO), _P, _P]),
    "dmp_stage_done_reset": (C.c_int, [C.POINTER(DmpState), C.POINTER(DmpIO), _P, _P]),
}
EXPORTS = tuple(_SIGNATURES)


class DmpError(RuntimeError):
    pass


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "snac_b200: %s is missing -- build it with `python snac_b200/build.py` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise ImportError("snac_b200: libdmp.so does not export %s" % name) from e
        fn.restype, fn.argtypes = res, args
    v = lib.dmp_abi_version()
    if v != ABI_VERSION:
        raise ImportError("snac_b200: libdmp.so ABI %d, expected %d" % (v, ABI_VERSION))
    return lib


lib = _load()


def check(rc: int, what: str = "libdmp") -> None:
    if rc == OK:
        return
    if rc == EINVAL:
        raise DmpError("%s: invalid argument" % what)
    raise DmpError("%s: CUDA error %d" % (what, lib.dmp_last_error()))
