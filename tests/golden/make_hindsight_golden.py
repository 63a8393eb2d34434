#!/usr/bin/env python
"""Golden traces of the reference's *_hindsight_replay classes and of the hindsight relabelling its learners do
(script/DRQN_hindsight/*/DRQN_hindsight_*.py): run an episode, reset a second env, overwrite its ``plan`` with what the
first one built (by assignment in 1D, in place in 2D/3D), replay the same actions / step sizes, keep the rewards.
Runs the UNMODIFIED reference classes (build container only) and writes tests/golden/hindsight_golden.npz.

    python tests/golden/make_hindsight_golden.py
Protocol per case: np.random.seed(seed) once; everything the classes draw comes from that global stream, in order."""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refload  # noqa: E402

REF3D_P = [.2, .2, .2, .2, .05, .05, .05, .05]

CASES = [
    # name, dim, kind, ctor kwargs, seed, episodes, max steps, action probabilities
    ("1d_generator", "1D", "hindsight_dynamic", {}, 11, 4, 160, [.25, .35, .4]),
    ("2d_dataset_dense", "2D", "hindsight_dynamic", dict(density="dense", split="train"), 21, 4, 160, [.15, .2, .2, .1, .35]),
    ("2d_dataset_sparse", "2D", "hindsight_dynamic", dict(density="sparse", split="val"), 22, 3, 120, [.15, .2, .2, .1, .35]),
    ("3d_dataset_dense", "3D", "hindsight_dynamic", dict(density="dense", split="train"), 31, 4, 200, [.1, .15, .15, .1, .1, .15, .15, .1]),
    ("1d_static", "1D", "hindsight_static", dict(plan_choose=1), 41, 3, 160, [.25, .35, .4]),
    ("2d_static", "2D", "hindsight_static", dict(plan_choose=1), 42, 3, 160, [.15, .2, .2, .1, .35]),
    ("3d_static", "3D", "hindsight_static", dict(plan_choose=0), 43, 3, 200, [.1, .15, .15, .1, .1, .15, .15, .1]),
]


def make(dim, kind, kw):
    cls = refload.load_class(dim, kind)
    if kind == "hindsight_static":
        return cls(plan_choose=kw["plan_choose"])
    if dim == "1D":
        return cls()
    return cls(data_path=refload.dataset_path(dim, kw["density"], kw["split"]))


def raw(o):
    return np.asarray(o[0] if isinstance(o, list) else o, dtype=np.float64).reshape(-1)


def relabel(env_h, env, dim):
    """What the learners do between env_hindsight.reset() and the replay."""
    h = env.HALF_WINDOW_SIZE
    if dim == "1D":                                          # script/DRQN_hindsight/1d/DRQN_hindsight_1D_dynamic.py:255
        env_h.plan = env.environment_memory[0, h:h + env.plan_width]
    else:                                                    # script/DRQN_hindsight/2d/DRQN_hindsight_2D_dynamic.py:274-278
        env_h.plan[h:h + env.plan_height, h:h + env.plan_width] = env.environment_memory[h:h + env.plan_height, h:h + env.plan_width]
        env_h.input_plan = env_h.plan[h:h + env.plan_height, h:h + env.plan_width]


def record(name, dim, kind, kw, seed, episodes, T, p):
    np.random.seed(seed)
    env, env_h = make(dim, kind, kw), make(dim, kind, kw)
    arng = np.random.RandomState(seed + 1000)
    out = dict(seed=seed, n_episodes=episodes)
    for ep in range(episodes):
        o = env.reset()
        acts, sizes, obs, rew, rint, done = [], [], [raw(o)], [], [], []
        plan0 = np.asarray(env.plan, dtype=np.float64).copy()
        tb0 = float(env.total_brick)
        idx0 = -1 if getattr(env, "index_random", None) is None else int(env.index_random)
        one_hot = np.asarray(env.one_hot, dtype=np.float64) if getattr(env, "one_hot", None) is not None else np.zeros(3)
        for t in range(T):
            a, s = int(arng.choice(len(p), p=p)), int(arng.randint(1, 4))
            o, r, d = env.step(a, s)
            acts.append(a); sizes.append(s); obs.append(raw(o)); rew.append(float(r)); rint.append(isinstance(r, int)); done.append(bool(d))
            if d:
                break
        grid = np.asarray(env.environment_memory, dtype=np.float64).copy()
        # hindsight replay against the achieved structure
        oh = env_h.reset()
        idx_h = -1 if getattr(env_h, "index_random", None) is None else int(env_h.index_random)
        tb_h = float(env_h.total_brick)
        relabel(env_h, env, dim)
        h_obs, h_rew, h_rint, h_done = [raw(oh)], [], [], []
        for a, s in zip(acts, sizes):
            o, r, d = env_h.step(a, s)
            h_obs.append(raw(o)); h_rew.append(float(r)); h_rint.append(isinstance(r, int)); h_done.append(bool(d))
        k = "ep%d_" % ep
        out.update({k + "plan": plan0, k + "total_brick": tb0, k + "plan_idx": idx0, k + "one_hot": one_hot,
                    k + "actions": np.asarray(acts, np.uint8), k + "sizes": np.asarray(sizes, np.uint8),
                    k + "obs": np.stack(obs), k + "reward": np.asarray(rew), k + "reward_is_int": np.asarray(rint),
                    k + "done": np.asarray(done), k + "final_grid": grid,
                    k + "h_plan_idx": idx_h, k + "h_total_brick": tb_h, k + "h_obs": np.stack(h_obs),
                    k + "h_reward": np.asarray(h_rew), k + "h_reward_is_int": np.asarray(h_rint), k + "h_done": np.asarray(h_done),
                    k + "h_final_grid": np.asarray(env_h.environment_memory, dtype=np.float64).copy(),
                    k + "h_iou": float(env_h.iou()) if hasattr(env_h, "iou") and dim != "2D" else float("nan")})
        print("%s ep %d: %d steps, return %.0f, hindsight return %.0f, bricks %d" % (
            name, ep, len(acts), sum(rew), sum(h_rew), int(getattr(env, "count_brick", None) or getattr(env, "conut_brick", 0))))
    return out


def main():
    allout = {}
    for name, dim, kind, kw, seed, episodes, T, p in CASES:
        for k, v in record(name, dim, kind, kw, seed, episodes, T, p).items():
            allout["%s/%s" % (name, k)] = v
    path = os.path.join(HERE, "hindsight_golden.npz")
    np.savez_compressed(path, **allout)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
