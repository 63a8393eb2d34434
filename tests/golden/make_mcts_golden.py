#!/usr/bin/env python
"""Golden traces of the reference's tree-search env variants (Env/*/…MCTS*.py): ``reset() -> (state, obs)``,
``step(a) -> (state, obs, reward, done)`` and the functional ``transition(state, a) -> (state', obs, reward, done)``
that script/MCTS/utils/uct.py expands tree nodes with.  ``state = (position, environment_memory, count_brick,
count_step)``.  Runs the UNMODIFIED reference classes (build container only) and writes tests/golden/mcts_golden.npz.

    python tests/golden/make_mcts_golden.py
Protocol per case (what the replaying tests repeat call for call): ``np.random.seed(seed)``; construct; ``reset()``;
actions come from ``RandomState(seed + 1000).choice(A, p)``; every step is ``env.step(a)``; after every EXPAND-th step the
current ``env.state`` is expanded with ``transition(copy of state, b)`` for every action b (one global-RNG step-size draw
each, in that order); an episode that ends is followed by ``reset()``.  Step sizes are recorded by peeking the global
stream (state saved / restored around the peek) so that the oracle can replay without numpy's generator."""
from __future__ import annotations

import copy
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refload  # noqa: E402

CASES = [
    # name, dim, kind, ctor kwargs, seed, total steps, expand every, action probabilities
    ("1d_static_p0", "1D", "mcts_static", dict(plan_choose=0), 101, 900, 9, [.2, .3, .5]),
    ("1d_static_p2", "1D", "mcts_static", dict(plan_choose=2), 102, 500, 7, [.3, .3, .4]),
    ("1d_dynamic", "1D", "mcts_dynamic", dict(split="test"), 103, 900, 9, [.2, .3, .5]),
    ("2d_static_dense", "2D", "mcts_static", dict(plan_choose=0), 201, 700, 7, [.15, .2, .2, .1, .35]),
    ("2d_static_sparse", "2D", "mcts_static", dict(plan_choose=1), 202, 400, 7, [.1, .25, .25, .1, .3]),
    ("2d_dynamic_dense", "2D", "mcts_dynamic", dict(density="dense", split="val"), 203, 700, 7, [.15, .2, .2, .1, .35]),
    ("3d_static_dense", "3D", "mcts_static", dict(plan_choose=0), 301, 600, 5, [.1, .15, .15, .1, .1, .15, .15, .1]),
    ("3d_static_sparse", "3D", "mcts_static", dict(plan_choose=1), 302, 400, 5, [.2, .2, .2, .2, .05, .05, .05, .05]),
    ("3d_dynamic_dense", "3D", "mcts_dynamic", dict(density="dense", split="test"), 303, 600, 5, [.1, .15, .15, .1, .1, .15, .15, .1]),
    ("3d_dynamic_sparse", "3D", "mcts_dynamic", dict(density="sparse", split="val"), 304, 400, 5, [.1, .1, .1, .1, .15, .15, .15, .15]),
]


def make(dim, kind, kw):
    cls = refload.load_class(dim, kind)
    if kind == "mcts_static":
        return cls(plan_choose=kw["plan_choose"])
    return cls(data_path=refload.dataset_path(dim, kw.get("density", "dense"), kw["split"]))


def peek_step_size():
    st = np.random.get_state()
    s = int(np.random.randint(1, 4))
    np.random.set_state(st)
    return s


def pos2(p):
    return [int(p), 0] if np.isscalar(p) or isinstance(p, (int, np.integer)) else [int(p[0]), int(p[1])]


def record(name, dim, kind, kw, seed, T, expand, p):
    np.random.seed(seed)
    env = make(dim, kind, kw)
    arng = np.random.RandomState(seed + 1000)
    A = len(p)
    state, o = env.reset()
    R = dict(reset_at=[0], reset_idx=[-1 if getattr(env, "index_random", None) is None else int(env.index_random)],
             reset_obs=[np.asarray(o, np.float64).reshape(-1)], reset_tb=[float(env.total_brick)],
             act=[], size=[], obs=[], rew=[], rint=[], done=[], pos=[], cb=[], cs=[], grid=[],
             x_at=[], x_act=[], x_size=[], x_obs=[], x_rew=[], x_rint=[], x_done=[], x_pos=[], x_cb=[], x_cs=[], x_grid=[],
             x_inplace=[])
    for t in range(T):
        a = int(arng.choice(A, p=p))
        s = peek_step_size()
        state, o, r, d = env.step(a)
        R["act"].append(a); R["size"].append(s); R["obs"].append(np.asarray(o, np.float64).reshape(-1))
        R["rew"].append(float(r)); R["rint"].append(isinstance(r, int)); R["done"].append(bool(d))
        R["pos"].append(pos2(state[0])); R["cb"].append(int(state[2])); R["cs"].append(int(state[3]))
        R["grid"].append(np.asarray(state[1], np.float64).copy())
        assert state is env.state and not np.shares_memory(state[1], env.environment_memory)
        if (t + 1) % expand == 0:
            for b in range(A):
                sin = copy.deepcopy(env.state)
                s = peek_step_size()
                sout, o2, r2, d2 = env.transition(sin, b)
                R["x_at"].append(t); R["x_act"].append(b); R["x_size"].append(s)
                R["x_obs"].append(np.asarray(o2, np.float64).reshape(-1)); R["x_rew"].append(float(r2))
                R["x_rint"].append(isinstance(r2, int)); R["x_done"].append(bool(d2))
                R["x_pos"].append(pos2(sout[0])); R["x_cb"].append(int(sout[2])); R["x_cs"].append(int(sout[3]))
                R["x_grid"].append(np.asarray(sout[1], np.float64).copy())
                R["x_inplace"].append(sout[1] is sin[1])             # the reference mutates the caller's array
        if d:
            state, o = env.reset()
            R["reset_at"].append(t + 1)
            R["reset_idx"].append(-1 if getattr(env, "index_random", None) is None else int(env.index_random))
            R["reset_obs"].append(np.asarray(o, np.float64).reshape(-1)); R["reset_tb"].append(float(env.total_brick))
    out = {}
    for k, v in R.items():
        a = np.asarray(v)
        if k.endswith("grid"):
            assert np.array_equal(a, a.astype(np.int16))
            a = a.astype(np.int16)
        elif k.endswith("obs"):
            assert np.array_equal(a, a.astype(np.int16))
            a = a.astype(np.int16)
        out["%s/%s" % (name, k)] = a
    out["%s/meta" % name] = np.asarray([seed, T, expand, A], dtype=np.int64)
    out["%s/p" % name] = np.asarray(p, dtype=np.float64)
    print("%-20s steps %4d  episodes %3d  expansions %4d  in-place %s" % (
        name, T, len(R["reset_at"]), len(R["x_at"]), all(R["x_inplace"])))
    return out


def main():
    out = {}
    for c in CASES:
        out.update(record(*c))
    path = os.path.join(HERE, "mcts_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
