"""Generate the committed golden fixtures by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Outputs (all under tests/golden/):
    plans_packed.npz      the reference's .pkl plan datasets re-encoded compactly
                          (1D: uint8 heights [n,30]; 2D/3D: np.packbits of the 20x20 interior mask)
    trace_<case>.npz      step-for-step traces of the reference env classes
    kat.json              one-episode known answers (SURVEY.md App. B protocol)

Trace protocol: np.random.seed(seed); env = RefClass(...); obs = env.reset(); then T steps with
actions from np.random.RandomState(seed+1); after every step the reference's own draw
``env.step_size`` (and ``env.index_random`` after every reset) is recorded, so the trace can be
replayed deterministically through ``step(action, step_size)`` / ``reset(plan_idx)`` entry points.
The env is reset after ``done`` (the reference leaves that to the caller).

Static 2D/3D cases run the reference with matplotlib's CirclePolygon replaced by the oracle's
restated 20-gon (matplotlib is not installed) -- their *plan* is therefore not pinned by these
traces, their step logic is.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refload  # noqa: E402

REF3D_P = [.2, .2, .2, .2, .05, .05, .05, .05]

CASES = [
    # name, dim, kind, ctor kwargs, seed, T, action mode
    ("1d_static_p0", "1D", "static", dict(plan_choose=0), 101, 2400, "uniform"),
    ("1d_static_p1", "1D", "static", dict(plan_choose=1), 102, 2400, "uniform"),
    ("1d_static_p2", "1D", "static", dict(plan_choose=2), 103, 2400, "uniform"),
    ("1d_static_p2_dropheavy", "1D", "static", dict(plan_choose=2), 104, 2400, "dropheavy"),
    ("1d_dynamic", "1D", "dynamic", dict(density="dense", split="train"), 105, 2400, "uniform"),
    ("1d_dynamic_dropheavy", "1D", "dynamic", dict(density="dense", split="test"), 106, 2400, "dropheavy"),
    ("2d_static_dense", "2D", "static", dict(plan_choose=0), 201, 2400, "uniform"),
    ("2d_static_sparse", "2D", "static", dict(plan_choose=1), 202, 2400, "uniform"),
    ("2d_dynamic_dense", "2D", "dynamic", dict(density="dense", split="train"), 203, 2400, "uniform"),
    ("2d_dynamic_sparse", "2D", "dynamic", dict(density="sparse", split="train"), 204, 2400, "uniform"),
    ("3d_static_dense", "3D", "static", dict(plan_choose=0), 301, 2400, "uniform"),
    ("3d_static_sparse_refp", "3D", "static", dict(plan_choose=1), 302, 3000, "ref3d"),
    ("3d_static_dense_refp", "3D", "static", dict(plan_choose=0), 303, 3000, "ref3d"),
    ("3d_dynamic_dense", "3D", "dynamic", dict(density="dense", split="train"), 304, 2400, "uniform"),
    ("3d_dynamic_sparse_refp", "3D", "dynamic", dict(density="sparse", split="train"), 305, 3000, "ref3d"),
    ("3d_dynamic_dense_refp", "3D", "dynamic", dict(density="dense", split="val"), 306, 3000, "ref3d"),
    # directed, closed-loop scripted policies: they read the reference env's own state
    ("1d_static_p0_greedy", "1D", "static", dict(plan_choose=0), 401, 1800, "greedy"),
    ("1d_dynamic_greedy", "1D", "dynamic", dict(density="dense", split="val"), 402, 1800, "greedy"),
    ("2d_static_dense_greedy", "2D", "static", dict(plan_choose=0), 403, 1500, "greedy"),
    ("2d_dynamic_sparse_greedy", "2D", "dynamic", dict(density="sparse", split="test"), 404, 1500, "greedy"),
    ("3d_static_dense_builder", "3D", "static", dict(plan_choose=0), 405, 4000, "builder"),
    ("3d_static_sparse_builder", "3D", "static", dict(plan_choose=1), 406, 4000, "builder"),
    ("3d_dynamic_dense_builder", "3D", "dynamic", dict(density="dense", split="test"), 407, 4000, "builder"),
    ("3d_dynamic_sparse_builder", "3D", "dynamic", dict(density="sparse", split="val"), 408, 4000, "builder"),
]


def iou_2d(env):
    """The reference's 2D IoU lives in render() (Env/2D/DMP_Env_2D_static.py:169-175) and in each
    learner script (script/DQN/2d/DQN_2d_static.py:62-70); same formula, evaluated on the
    reference env's own arrays."""
    p = env.plan[3:23, 3:23].astype(bool)
    g = env.environment_memory[3:23, 3:23].astype(bool)
    union = (p + g).sum()
    return float((p * g).sum() / float(union)) if union else float("nan")


def make_env(dim, kind, kw):
    cls = refload.load_class(dim, kind)
    if kind == "static":
        return cls(plan_choose=kw["plan_choose"])
    return cls(data_path=refload.dataset_path(dim, kw["density"], kw["split"]))


def scripted_action(env, dim, rng, mode):
    """Closed-loop policies for the directed traces.  They only READ the reference env's public state
    (position_memory, environment_memory, plan) -- the env itself is untouched."""
    if dim == "1D":
        # fill each column to its plan height (rewards 1 ... 1, 10), sometimes overbuild (-1), sweep right/left
        pos = env.position_memory[-1]
        h, p = env.environment_memory[0, pos], env.plan[pos - 2]
        if h < p or rng.rand() < 0.02:
            return 2
        return int(rng.choice([0, 1], p=[0.3, 0.7])) if pos < 31 else 0
    if dim == "2D":
        r, c = env.position_memory[-1]
        on_plan = env.plan[r, c] > 0
        if (on_plan and env.environment_memory[r, c] == 0 and rng.rand() < 0.9) or rng.rand() < 0.03:
            return 4                                   # first brick (5.0), double drops and off-plan drops (0)
        return int(rng.randint(4))
    # 3D builder: build on adjacent plan cells below target, occasionally overbuild / build at random,
    # otherwise move in a random direction (legal or not: collisions and walls get exercised)
    r, c = env.position_memory[-1]
    g, plan = env.environment_memory, env.plan
    nbr = [(r, c - 1), (r, c + 1), (r + 1, c), (r - 1, c)]
    cand = [i for i, (a, b) in enumerate(nbr) if g[a, b] != -1 and g[a, b] < plan[a, b]]
    free = [i for i, (a, b) in enumerate(nbr) if g[a, b] == 0]
    u = rng.rand()
    if cand and len(free) > 1 and u < 0.55:
        return 4 + int(rng.choice(cand))
    if u > 0.985:
        return 4 + int(rng.randint(4))                 # stray build: overbuild (-1), off-plan, against the wall
    if free and u < 0.97:
        return int(rng.choice(free))
    return int(rng.randint(4))


def draw_action(rng, mode, n_actions):
    if mode == "uniform":
        return int(rng.randint(n_actions))
    if mode == "ref3d":
        return int(rng.choice(8, p=REF3D_P))
    if mode == "dropheavy":            # mostly drops: exercises the brick-budget termination in 1D
        return int(rng.choice(3, p=[.05, .05, .9]))
    raise ValueError(mode)


def record(name, dim, kind, kw, seed, T, mode):
    env = make_env(dim, kind, kw)
    dyn = kind == "dynamic"
    np.random.seed(seed)
    rng = np.random.RandomState(seed + 1)
    A = env.action_dim
    D = env.state_dim
    out = dict(actions=np.zeros(T, np.uint8), step_sizes=np.zeros(T, np.uint8),
               obs=np.zeros((T, D), np.int16), obs_norm=np.zeros((T, 2), np.float64),
               reward=np.zeros(T, np.float64), reward_is_int=np.zeros(T, bool),
               done=np.zeros(T, bool), iou=np.zeros(T, np.float64),
               pos=np.zeros((T, 2), np.int16), count_brick=np.zeros(T, np.int32),
               count_step=np.zeros(T, np.int32), total_brick=np.zeros(T, np.float64))
    reset_plan_idx, final_grids, reset_obs = [], [], []

    def raw_obs(o):
        """Integer-valued raw observation row (window + raw counters) as int16."""
        if dyn and dim == "1D":
            arr = o[0]
        elif dyn:
            arr = np.array(o[0], dtype=np.float64)
            arr[0, -2], arr[0, -1] = env.count_brick, env.count_step
        else:
            arr = o
        a16 = arr.astype(np.int16)
        assert np.array_equal(a16.astype(np.float64), arr)
        return a16[0]

    def do_reset():
        o = env.reset()
        reset_plan_idx.append(int(env.index_random) if dyn else 0)
        reset_obs.append(raw_obs(o))
        return o

    def brick_count():
        return env.conut_brick if dim == "1D" else env.count_brick

    do_reset()
    for t in range(T):
        a = scripted_action(env, dim, rng, mode) if mode in ("greedy", "builder") else draw_action(rng, mode, A)
        o, r, d = env.step(a)
        out["actions"][t], out["step_sizes"][t] = a, env.step_size
        out["obs"][t] = raw_obs(o)
        if dyn:
            nrm = o[1] if dim == "1D" else o[0]
            out["obs_norm"][t] = nrm[0, -2:]
        out["reward"][t], out["reward_is_int"][t], out["done"][t] = r, isinstance(r, int), d
        out["iou"][t] = env.iou() if dim != "2D" else iou_2d(env)
        p = env.position_memory[-1]
        out["pos"][t] = [p, 0] if dim == "1D" else p
        out["count_brick"][t], out["count_step"][t] = brick_count(), env.count_step
        out["total_brick"][t] = env.total_brick
        if d:
            final_grids.append(env.environment_memory.astype(np.int16).copy())
            do_reset()
    final_grids.append(env.environment_memory.astype(np.int16).copy())
    out["reset_plan_idx"] = np.asarray(reset_plan_idx, np.int32)
    out["reset_obs"] = np.asarray(reset_obs, np.int16)
    out["final_grids"] = np.asarray(final_grids, np.int16)
    out["meta"] = np.asarray(json.dumps(dict(name=name, dim=dim, kind=kind, kw=kw, seed=seed, T=T, mode=mode)))
    np.savez_compressed(os.path.join(HERE, "trace_%s.npz" % name), **out)
    rv, rc = np.unique(out["reward"], return_counts=True)
    print("%-28s T=%d episodes=%d return=%.1f rewards=%s maxh=%d" % (
        name, T, int(out["done"].sum()), out["reward"].sum(), dict(zip(rv.tolist(), rc.tolist())),
        int(out["final_grids"].max())))


def record_lnet(name, dim, plan_choose, seed, T, mode):
    """Traces of the *_Lnet observation-format variants: 1D appends the position,
    2D uses 2 for the frame and returns [normalised obs, position], 3D adopts the dynamic termination rules."""
    env = refload.load_class(dim, "lnet")(plan_choose=plan_choose)
    np.random.seed(seed)
    rng = np.random.RandomState(seed + 1)
    D = 8 if dim == "1D" else 51
    out = dict(actions=np.zeros(T, np.uint8), step_sizes=np.zeros(T, np.uint8), obs=np.zeros((T, D), np.float64),
               reward=np.zeros(T, np.float64), done=np.zeros(T, bool), pos=np.zeros((T, 2), np.int16))
    reset_obs = []

    def flat(o):
        return (o if dim == "1D" else o[0])[0].astype(np.float64)

    reset_obs.append(flat(env.reset()))
    for t in range(T):
        a = scripted_action(env, dim, rng, mode) if mode in ("greedy", "builder") else draw_action(rng, mode, env.action_dim)
        o, r, d = env.step(a)
        out["actions"][t], out["step_sizes"][t] = a, env.step_size
        out["obs"][t], out["reward"][t], out["done"][t] = flat(o), r, d
        p = env.position_memory[-1]
        out["pos"][t] = [p, 0] if dim == "1D" else p
        if dim != "1D":
            assert list(o[1]) == list(p)
        if d:
            reset_obs.append(flat(env.reset()))
    out["reset_obs"] = np.asarray(reset_obs)
    out["final_grid"] = env.environment_memory.astype(np.int16)
    out["meta"] = np.asarray(json.dumps(dict(name=name, dim=dim, kind="lnet", kw=dict(plan_choose=plan_choose), seed=seed, T=T, mode=mode)))
    np.savez_compressed(os.path.join(HERE, "lnet_%s.npz" % name), **out)
    print("%-28s T=%d episodes=%d return=%.1f" % (name, T, int(out["done"].sum()), out["reward"].sum()))


LNET_CASES = [("1d_p1", "1D", 1, 501, 1600, "greedy"), ("2d_dense", "2D", 0, 502, 1300, "greedy"),
              ("3d_dense", "3D", 0, 503, 3000, "builder"), ("3d_sparse_uniform", "3D", 1, 504, 1200, "uniform")]


def pack_plans():
    out = {}
    for split in ("train", "val", "test"):
        p = np.asarray(refload.load_dataset("1D", split=split))
        assert np.array_equal(p, p.astype(np.uint8))
        out["1d_%s" % split] = p.astype(np.uint8)
        for dim in ("2D", "3D"):
            for dens in ("dense", "sparse"):
                p = np.asarray(refload.load_dataset(dim, dens, split))
                z = 1.0 if dim == "2D" else 6.0
                assert set(np.unique(p)) <= {0.0, z}
                inner = p[:, 3:23, 3:23]
                assert inner.sum() == p.sum()          # borders are all zero
                out["%s_%s_%s" % (dim.lower(), dens, split)] = np.packbits((inner > 0).reshape(len(p), 400), axis=1)
    np.savez_compressed(os.path.join(HERE, "plans_packed.npz"), **out)
    print("plans_packed.npz:", {k: v.shape for k, v in out.items()})


def kat():
    """SURVEY.md App. B protocol: seed(7), actions from RandomState(1234), one episode."""
    rows = []
    specs = [("1D", "static", dict(plan_choose=0)), ("1D", "static", dict(plan_choose=1)),
             ("1D", "static", dict(plan_choose=2)), ("1D", "dynamic", dict(density="dense", split="test")),
             ("2D", "static", dict(plan_choose=0)), ("2D", "static", dict(plan_choose=1)),
             ("2D", "dynamic", dict(density="dense", split="test")),
             ("3D", "static", dict(plan_choose=0)), ("3D", "static", dict(plan_choose=1)),
             ("3D", "dynamic", dict(density="dense", split="test"))]
    for dim, kind, kw in specs:
        cls = refload.load_class(dim, kind)
        if kind == "static":
            env = cls(plan_choose=kw["plan_choose"])
        else:
            env = cls(data_path=refload.dataset_path(dim, kw["density"], kw["split"]), random_choose_paln=False)
        np.random.seed(7)
        rng = np.random.RandomState(1234)
        env.reset()
        ret, steps, acts, sizes = 0.0, 0, [], []
        while True:
            a = int(rng.choice(8, p=REF3D_P)) if dim == "3D" else int(rng.randint(env.action_dim))
            _, r, d = env.step(a)
            acts.append(a)
            sizes.append(int(env.step_size))
            ret += r
            steps += 1
            if d:
                break
        p = env.position_memory[-1]
        rows.append(dict(dim=dim, kind=kind, kw=kw, steps=steps, ret=ret,
                         bricks=int(env.conut_brick if dim == "1D" else env.count_brick),
                         total_brick=float(env.total_brick),
                         pos=[int(p)] if dim == "1D" else [int(p[0]), int(p[1])],
                         iou=float(env.iou() if dim != "2D" else iou_2d(env)),
                         actions=acts, step_sizes=sizes))
        print(dim, kind, kw, steps, ret, rows[-1]["bricks"], rows[-1]["total_brick"], rows[-1]["pos"], rows[-1]["iou"])
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(rows, f)


if __name__ == "__main__":
    assert refload.available(), "needs the reference tree"
    pack_plans()
    for c in CASES:
        record(*c)
    for c in LNET_CASES:
        record_lnet(*c)
    kat()
