"""Tree-search drop-in classes (snac_b200.compat.*_MCTS*) against traces of the reference's Env/*/...MCTS*.py classes
(tests/golden/mcts_golden.npz): same seed, same sequence of reset / step / transition calls, same global-RNG draws --
states, observations, rewards (values AND python types) and done flags must be identical.  Also the batched expansion
(transition_batch / BatchedDMPEnv.transition_dense: many tree nodes per launch) against the oracle."""
import copy

import numpy as np
import pytest
import torch

from mcts_cases import MCTS_CASES, load_mcts_case
from oracle import dmp_oracle as O

pytestmark = pytest.mark.gpu


def make_compat(name):
    import snac_b200 as S
    g, dim, dynamic, plan_choose, plans = load_mcts_case(name)
    cls = {(1, False): S.deep_mobile_printing_1d1r_MCTS, (1, True): S.deep_mobile_printing_1d1r_MCTS_obs,
           (2, False): S.deep_mobile_printing_2d1r_MCTS, (2, True): S.deep_mobile_printing_2d1r_MCTS_dynamic,
           (3, False): S.deep_mobile_printing_3d1r_MCTS, (3, True): S.deep_mobile_printing_3d1r_MCTS_dynamic}[(dim, dynamic)]
    env = cls(plans=plans) if dynamic else cls(plan_choose=plan_choose)
    return g, dim, dynamic, env


def state_matches(state, dim, g, pre, i):
    pos, grid, cb, cs = state
    assert ([pos, 0] if dim == 1 else list(pos)) == list(g[pre + "pos"][i]), (pre, i)
    assert cb == g[pre + "cb"][i] and cs == g[pre + "cs"][i], (pre, i)
    assert grid.dtype == np.float64 and np.array_equal(grid, g[pre + "grid"][i].astype(np.float64)), (pre, i)


@pytest.mark.parametrize("name", sorted(MCTS_CASES))
def test_mcts_classes_replay_reference(name):
    g, dim, dynamic, env = make_compat(name)
    seed, T, expand, A = [int(v) for v in g["meta"]]
    p = g["p"]
    np.random.seed(seed)
    arng = np.random.RandomState(seed + 1000)
    k = 0
    state, o = env.reset()
    assert np.array_equal(np.asarray(o).reshape(-1), g["reset_obs"][0].astype(np.float64))
    assert float(env.total_brick) == g["reset_tb"][0]
    j = 0
    for t in range(T):
        a = int(arng.choice(A, p=p))
        assert a == g["act"][t]
        state, o, r, d = env.step(a)
        assert env.step_size == g["size"][t]
        assert o.shape == (1, env.state_dim) and np.array_equal(o[0], g["obs"][t].astype(np.float64)), t
        assert r == g["rew"][t] and isinstance(r, int) == bool(g["rint"][t]), (t, r)
        assert d == bool(g["done"][t]) and isinstance(d, bool)
        state_matches(state, dim, g, "", t)
        assert state is env.state and not np.shares_memory(state[1], env.environment_memory)
        if (t + 1) % expand == 0:
            for b in range(A):
                sin = copy.deepcopy(env.state)
                sout, o2, r2, d2 = env.transition(sin, b)
                assert g["x_at"][j] == t and g["x_act"][j] == b
                assert np.array_equal(o2[0], g["x_obs"][j].astype(np.float64)), (t, b)
                assert r2 == g["x_rew"][j] and isinstance(r2, int) == bool(g["x_rint"][j]), (t, b, r2)
                assert d2 == bool(g["x_done"][j])
                state_matches(sout, dim, g, "x_", j)
                assert (sout[1] is sin[1]) == bool(g["x_inplace"][j])
                j += 1
            # expanding must not disturb the episode
            state_matches(env.state, dim, g, "", t)
        if d:
            k += 1
            state, o = env.reset()
            assert g["reset_at"][k] == t + 1
            if g["reset_idx"][k] >= 0:
                assert env.index_random == g["reset_idx"][k]
            assert np.array_equal(np.asarray(o).reshape(-1), g["reset_obs"][k].astype(np.float64))
            assert float(env.total_brick) == g["reset_tb"][k]
    assert j == len(g["x_at"])


@pytest.mark.parametrize("name", ["1d_static_p0", "2d_dynamic_dense", "3d_static_dense", "3d_dynamic_sparse"])
def test_transition_batch_one_launch(name):
    """All recorded expansions of a case as ONE transition_batch call (one kernel launch) == the reference, one by one."""
    g, dim, dynamic, env = make_compat(name)
    np.random.seed(int(g["meta"][0]))
    env.reset()
    # nodes of the first episode only: later episodes of the dataset classes run against other plans
    first_end = int(g["reset_at"][1]) if len(g["reset_at"]) > 1 else len(g["act"])
    js = [j for j in range(len(g["x_at"])) if g["x_at"][j] < first_end]
    assert len(js) >= 8
    states = []
    for j in js:
        t = int(g["x_at"][j])
        pos = int(g["pos"][t][0]) if dim == 1 else [int(g["pos"][t][0]), int(g["pos"][t][1])]
        states.append((pos, g["grid"][t].astype(np.float64), int(g["cb"][t]), int(g["cs"][t])))
    out = env.transition_batch(states, [int(g["x_act"][j]) for j in js], [int(g["x_size"][j]) for j in js])
    for (sout, o, r, d), j in zip(out, js):
        state_matches(sout, dim, g, "x_", j)
        assert np.array_equal(o[0], g["x_obs"][j].astype(np.float64))
        assert r == g["x_rew"][j] and isinstance(r, int) == bool(g["x_rint"][j]) and d == bool(g["x_done"][j])


@pytest.mark.parametrize("dim,dynamic", [(1, False), (2, False), (2, True), (3, False), (3, True)])
def test_transition_dense_random_nodes_vs_oracle(dim, dynamic):
    """BatchedDMPEnv.transition_dense on 3 000 random (reachable-looking) tree nodes against the oracle stepped from the
    same states: heights / occupancy, positions and counters drawn at random, every action, every step size."""
    from conftest import load_plans
    from snac_b200 import BatchedDMPEnv
    n = 3000
    rng = np.random.RandomState(77 + dim)
    plans = load_plans(dim, "dense", "train")[:40] if dynamic else None
    env = BatchedDMPEnv(dim, dynamic=dynamic, plan_choose=0, plans=plans, num_envs=n, device="cuda",
                        obs_dtype=torch.float64, auto_reset=False)
    spec = O.SPEC[dim]
    W, HW, A = spec["width"], spec["hw"], spec["actions"]
    T = spec["total_step"][1 if dynamic else 0]
    if dim == 1:
        grid = np.full((n, 1, W + 2 * HW), -1.0)
        grid[:, :, HW:HW + W] = rng.randint(0, 32, size=(n, 1, W)) * (rng.rand(n, 1, W) < 0.6)
        pos = rng.randint(HW, HW + W, size=n)
    else:
        grid = np.full((n, W + 2 * HW, W + 2 * HW), -1.0)
        hi = 2 if dim == 2 else 9
        grid[:, HW:HW + W, HW:HW + W] = rng.randint(0, hi, size=(n, W, W)) * (rng.rand(n, W, W) < (0.5 if dim == 2 else 0.35))
        pos = rng.randint(HW, HW + W, size=(n, 2))
    cb = rng.randint(0, 200, size=n)
    cs = np.where(rng.rand(n) < 0.1, T - 1, rng.randint(0, T - 1, size=n))
    cb = np.where(rng.rand(n) < 0.05, 2000, cb)                      # some nodes are one brick from the budget
    acts = rng.randint(0, A, size=n)
    sizes = rng.randint(1, 4, size=n)
    pidx = rng.randint(0, len(plans), size=n) if dynamic else np.zeros(n, dtype=np.int64)
    npos, ngrid, ncb, ncs, obs, rew, done = env.transition_dense(pos, grid, cb, cs, acts, sizes, plan_idx=pidx)
    npos, ngrid, ncb, ncs = npos.cpu().numpy(), ngrid.cpu().numpy(), ncb.cpu().numpy(), ncs.cpu().numpy()
    obs, rew, done = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()
    iou = env.iou().cpu().numpy()
    ref = O.make_env(dim, dynamic, plan_choose=0, plans=plans)
    for i in range(n):
        ref.reset(int(pidx[i]))
        ref.grid = grid[i].copy()
        ref.pos = int(pos[i]) if dim == 1 else [int(pos[i, 0]), int(pos[i, 1])]
        ref.count_brick, ref.count_step = int(cb[i]), int(cs[i])
        o, r, d = ref.step(int(acts[i]), int(sizes[i]))
        assert np.array_equal(o[0], obs[i]), i
        assert r == rew[i] and d == bool(done[i]), (i, r, rew[i], d, done[i])
        assert np.array_equal(ref.grid.reshape(ngrid[i].shape), ngrid[i]), i
        assert (ref.pos == npos[i]) if dim == 1 else (list(ref.pos) == list(npos[i])), i
        assert ref.count_brick == ncb[i] and ref.count_step == ncs[i]
        ri = ref.iou()
        assert ri == iou[i] or (np.isnan(ri) and np.isnan(iou[i])), (i, ri, iou[i])
