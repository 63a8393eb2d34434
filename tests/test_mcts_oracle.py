"""The oracle stepped from ARBITRARY initial states against the reference's tree-search variants
(Env/*/...MCTS*.py: ``step`` and the functional ``transition(state, action)``), replaying
tests/golden/mcts_golden.npz.  CPU only."""
import numpy as np
import pytest

from mcts_cases import MCTS_CASES, load_mcts_case
from oracle import dmp_oracle as O


def set_state(env, dim, pos, grid, cb, cs):
    env.grid = np.asarray(grid, dtype=np.float64).copy()
    env.pos = int(pos[0]) if dim == 1 else [int(pos[0]), int(pos[1])]
    env.count_brick, env.count_step = int(cb), int(cs)


def check(env, dim, o, r, d, g, pre, i):
    assert np.array_equal(o[0], g[pre + "obs"][i].astype(np.float64)), (pre, i)
    assert r == g[pre + "rew"][i] and d == bool(g[pre + "done"][i]), (pre, i, r, d)
    pos = [env.pos, 0] if dim == 1 else env.pos
    assert list(pos) == list(g[pre + "pos"][i])
    assert env.count_brick == g[pre + "cb"][i] and env.count_step == g[pre + "cs"][i]
    assert np.array_equal(env.grid, g[pre + "grid"][i].astype(np.float64))


@pytest.mark.parametrize("name", sorted(MCTS_CASES))
def test_oracle_matches_reference_mcts_variants(name):
    g, dim, dynamic, plan_choose, plans = load_mcts_case(name)
    env = O.make_env(dim, dynamic, plan_choose=plan_choose, plans=plans)
    twin = O.make_env(dim, dynamic, plan_choose=plan_choose, plans=plans)       # expands tree nodes
    resets = {int(t): k for k, t in enumerate(g["reset_at"])}
    xs_by_t = {}
    for j, t in enumerate(g["x_at"]):
        xs_by_t.setdefault(int(t), []).append(j)

    def do_reset(k):
        idx = int(g["reset_idx"][k])
        o = env.reset(max(idx, 0))
        twin.reset(max(idx, 0))
        assert np.array_equal(o[0], g["reset_obs"][k].astype(np.float64))
        assert env.total_brick == g["reset_tb"][k]

    do_reset(0)
    for t in range(len(g["act"])):
        o, r, d = env.step(int(g["act"][t]), int(g["size"][t]))
        check(env, dim, o, r, d, g, "", t)
        # the MCTS variants' move reward: python int 0 from step() in every class; transition() returns
        # the float 0.0 in 1D / 3D and the int 0 in 2D (Env/1D/DMP_Env_1D_static_MCTS.py:141, Env/2D/DMP_ENV_2D_static_MCTS.py:166)
        for j in xs_by_t.get(t, []):
            set_state(twin, dim, g["pos"][t], g["grid"][t], g["cb"][t], g["cs"][t])
            o2, r2, d2 = twin.step(int(g["x_act"][j]), int(g["x_size"][j]))
            check(twin, dim, o2, r2, d2, g, "x_", j)
        if d:
            do_reset(resets[t + 1])
