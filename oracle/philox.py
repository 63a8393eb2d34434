"""Philox4x32-10 counter-based RNG, numpy restatement (TEST INFRASTRUCTURE ONLY).

The CUDA kernels (snac_b200/csrc/dmp_rng.cuh) generate the reference's two stochastic
draws -- ``np.random.randint(1, 4)`` per step (e.g. Env/2D/DMP_Env_2D_static.py:97) and
``np.random.randint(0, len(dataset))`` per reset (e.g. Env/2D/DMP_Env_2D_dynamic_usedata_plan.py:36)
-- plus the synthetic random actions of throughput runs from a counter-based stream so that
results do not depend on how envs are sharded over GPUs.  This file states the same stream
on the CPU so parity tests can replay it through the oracle.

Stream definition (the single source of truth for both sides):
    key     = (seed & 0xffffffff, seed >> 32)
    In-step draws -- one Philox block serves FOUR consecutive steps of an env:
      counter = (env_id & 0xffffffff, env_id >> 32, (t >> 2) & 0xffffffff, (t >> 2) >> 32)
                env_id = GLOBAL env index, t = global step index of the vector env
      x0..x3  = philox4x32_10(counter, key);   w = x[t & 3]
      step_size = 1 + (((w & 0xffff) * 3) >> 16)             in {1,2,3}
      action    = ((w >> 16) * A) >> 16                      uniform over A actions (16-bit resolution)
                | ref3d: v = ((w >> 16) * 20) >> 16; v < 16 ? v >> 2 : v - 12
                  (p = [.2,.2,.2,.2,.05,.05,.05,.05], Env/3D/DMP_simulator_3d_static_circle.py:361-362)
    Plan of an env that auto-resets in step t (consumed only then) -- a block of its own, keyed apart:
      plan_idx  = mulhi32(philox4x32_10((env_id lo, env_id hi, t lo, t hi), key ^ PLAN_KEY).x2, n_plans)
Explicit resets (dmp_reset with plan_idx == NULL, plan_mode Philox) draw
    plan_idx  = mulhi32(x3, n_plans)   with counter t = 0xffffffffffffffff - (global step index at reset time)
under the plain key, far away from the in-step block counters t >> 2.
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
SEED_DEFAULT = 0x534E4143          # "SNAC"
T_INIT = 0xFFFFFFFFFFFFFFFF


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """All arguments uint32 arrays (broadcastable) or ints; returns four uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) & MASK for v in (c0, c1, c2, c3))
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(v.astype(np.uint32) for v in (c0, c1, c2, c3))


def mulhi32(x, n: int):
    return ((np.asarray(x, dtype=np.uint64) * np.uint64(n)) >> np.uint64(32)).astype(np.int64)


PLAN_KEY = 0x504C414E5F4B4559      # "PLAN_KEY": xor-ed into the seed for the auto-reset plan draw


def draws(seed: int, env_ids, t: int, n_actions: int, n_plans: int = 1, ref3d: bool = False, with_plan: bool = True):
    """(step_size, action, plan_idx) int64 arrays for global env ids at global step t
    (with_plan=False skips the plan block and returns plan_idx = None: static envs never consume it)."""
    env_ids = np.asarray(env_ids, dtype=np.uint64)
    t = int(t) & 0xFFFFFFFFFFFFFFFF
    tb = t >> 2
    x = philox4x32_10(env_ids & MASK, env_ids >> np.uint64(32), tb & 0xFFFFFFFF, tb >> 32,
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    w = x[t & 3].astype(np.int64)
    lo, hi = w & 0xFFFF, w >> 16
    step_size = 1 + ((lo * 3) >> 16)
    if ref3d:
        v = (hi * 20) >> 16
        action = np.where(v < 16, v >> 2, v - 12)
    else:
        action = (hi * n_actions) >> 16
    if not with_plan:
        return step_size, action, None
    pk = (seed ^ PLAN_KEY) & 0xFFFFFFFFFFFFFFFF
    _, _, x2, _ = philox4x32_10(env_ids & MASK, env_ids >> np.uint64(32), t & 0xFFFFFFFF, t >> 32,
                                pk & 0xFFFFFFFF, (pk >> 32) & 0xFFFFFFFF)
    plan_idx = mulhi32(x2, max(int(n_plans), 1))
    return step_size, action, plan_idx


def reset_draw(seed: int, env_ids, t_now: int, n_plans: int):
    """Plan index drawn by an explicit reset issued when the vector env's step index is t_now."""
    env_ids = np.asarray(env_ids, dtype=np.uint64)
    t = (T_INIT - int(t_now)) & 0xFFFFFFFFFFFFFFFF
    _, _, _, x3 = philox4x32_10(env_ids & MASK, env_ids >> np.uint64(32), t & 0xFFFFFFFF, t >> 32,
                                seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return mulhi32(x3, max(int(n_plans), 1))
