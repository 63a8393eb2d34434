"""ctypes wrapper of oracle/dmp_oracle.c (TEST INFRASTRUCTURE ONLY): N oracle envs with the vector
env's auto-reset/statistics semantics, fast enough for 10^5-env parity cases and for a multi-threaded
CPU baseline."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import dmp_oracle as O
from .build import build_oracle


class OrcCfg(C.Structure):
    _fields_ = [("dim", C.c_int32), ("dynamic", C.c_int32), ("n_plans", C.c_int32), ("total_step", C.c_int32),
                ("plans", C.c_void_p), ("total_brick", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_oracle())
        _lib.orc_sizeof_env.restype = C.c_int
        _lib.orc_rollout.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class COracleBatch:
    def __init__(self, dim, dynamic, n, plan_choose=0, plans=None, sequential=False):
        self.dim, self.dynamic, self.n, self.sequential = dim, bool(dynamic), int(n), bool(sequential)
        if dynamic:
            self.plans = np.ascontiguousarray(plans, dtype=np.float64)
        else:
            self.plans = np.ascontiguousarray(O.static_plan(dim, plan_choose)[None], dtype=np.float64)
        self.n_plans = len(self.plans)
        self.total_brick = np.array([O.total_brick_of(dim, p, dynamic) for p in self.plans], dtype=np.float64)
        self.D = O.SPEC[dim]["obs_dim"]
        self.cfg = OrcCfg(dim, int(dynamic), self.n_plans, O.SPEC[dim]["total_step"][1 if dynamic else 0],
                          self.plans.ctypes.data, self.total_brick.ctypes.data)
        self.envs = np.zeros(self.n * lib().orc_sizeof_env(), dtype=np.uint8)
        self.ep_cnt = np.zeros(n, np.int64)
        self.ep_len = np.zeros(n, np.int64)
        self.ep_ret = np.zeros(n, np.float64)
        self.ep_iou = np.zeros(n, np.float64)

    def reset(self, plan_idx=None):
        obs = np.zeros((self.n, self.D))
        p = None if plan_idx is None else np.ascontiguousarray(plan_idx, dtype=np.int32)
        lib().orc_reset_all(C.byref(self.cfg), _p(self.envs), C.c_int64(self.n), _p(p), _p(obs))
        return obs

    def rollout(self, actions, sizes, next_plan=None, auto_reset=True, normalise=False, want_obs=True):
        K = actions.shape[0]
        a = np.ascontiguousarray(actions, dtype=np.uint8)
        s = np.ascontiguousarray(sizes, dtype=np.uint8)
        p = None if next_plan is None else np.ascontiguousarray(next_plan, dtype=np.int32)
        obs = np.zeros((K, self.n, self.D)) if want_obs else None
        rew = np.zeros((K, self.n), np.float32)
        done = np.zeros((K, self.n), np.uint8)
        err = lib().orc_rollout(C.byref(self.cfg), _p(self.envs), C.c_int64(self.n), C.c_int(K), _p(a), _p(s), _p(p),
                                C.c_int(int(self.sequential)), C.c_int(int(auto_reset)), C.c_int(int(normalise)),
                                _p(obs), _p(rew), _p(done), _p(self.ep_cnt), _p(self.ep_len), _p(self.ep_ret), _p(self.ep_iou))
        return obs, rew, done.astype(bool), err

    def export(self):
        G = 34 if self.dim == 1 else 676
        grid = np.zeros((self.n, G), np.int32)
        sc = np.zeros((self.n, 8), np.int32)
        lib().orc_export(C.byref(self.cfg), _p(self.envs), C.c_int64(self.n), _p(grid), _p(sc))
        return grid.reshape((self.n, 1, 34) if self.dim == 1 else (self.n, 26, 26)), sc

    def iou(self):
        out = np.zeros(self.n)
        lib().orc_iou_all(C.byref(self.cfg), _p(self.envs), C.c_int64(self.n), _p(out))
        return out
