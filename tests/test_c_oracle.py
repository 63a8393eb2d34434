"""The C restatement (oracle/dmp_oracle.c) against the golden traces of the unmodified reference and
against the python oracle on long Philox-driven vector runs.  CPU only."""
import numpy as np
import pytest

from conftest import TRACE_NAMES, load_plans, load_trace, trace_env_spec
from oracle import dmp_oracle as O
from oracle import philox
from oracle.c_oracle import COracleBatch
from oracle_batch import OracleBatch, philox_rollout


@pytest.mark.parametrize("name", TRACE_NAMES)
def test_c_oracle_replays_reference_trace(name):
    tr = load_trace(name)
    dim, dynamic, plan_choose, plans = trace_env_spec(tr["meta"])
    cb = COracleBatch(dim, dynamic, 1, plan_choose, plans)
    T = len(tr["actions"])
    resets = tr["reset_plan_idx"]
    nxt = np.zeros((T, 1), np.int32)
    nd = np.cumsum(tr["done"])
    for t in range(T):
        if tr["done"][t]:
            nxt[t, 0] = resets[nd[t]]
    o = cb.reset([resets[0]])
    assert np.array_equal(o[0], tr["reset_obs"][0].astype(np.float64))
    obs, rew, done, err = cb.rollout(tr["actions"][:, None], tr["step_sizes"][:, None], nxt)
    assert err == 0
    assert np.array_equal(obs[:, 0], tr["obs"].astype(np.float64))
    assert np.array_equal(rew[:, 0].astype(np.float64), tr["reward"])
    assert np.array_equal(done[:, 0], tr["done"])
    ends = np.nonzero(tr["done"])[0]
    assert cb.ep_cnt[0] == len(ends) and cb.ep_len[0] == tr["count_step"][ends].sum()
    want_iou = 0.0
    for e in ends:
        want_iou += tr["iou"][e]
    assert cb.ep_iou[0] == want_iou or np.isnan(want_iou)
    g, sc = cb.export()
    assert np.array_equal(g[0].reshape(tr["final_grids"][-1].shape), tr["final_grids"][-1])
    if dynamic:   # normalised counters
        cb2 = COracleBatch(dim, dynamic, 1, plan_choose, plans)
        cb2.reset([resets[0]])
        obs2, _, _, _ = cb2.rollout(tr["actions"][:, None], tr["step_sizes"][:, None], nxt, normalise=True)
        assert np.array_equal(obs2[:, 0, -2:], tr["obs_norm"])


@pytest.mark.parametrize("dim,dynamic,density,ref3d", [(1, False, None, False), (1, True, "dense", False),
                                                        (2, False, None, False), (2, True, "sparse", False),
                                                        (3, False, None, True), (3, True, "dense", False)])
def test_c_oracle_equals_python_oracle_on_philox_streams(dim, dynamic, density, ref3d):
    n, K, seed, base = 24, 900, 77, 1000
    plans = load_plans(dim, density, "val") if dynamic else None
    ob = OracleBatch(dim, dynamic, n, 0, plans)
    cb = COracleBatch(dim, dynamic, n, 0, plans)
    p0 = philox.reset_draw(seed, np.arange(base, base + n), 0, ob.n_plans) if dynamic else None
    assert np.array_equal(ob.reset(p0), cb.reset(p0))
    A = O.SPEC[dim]["actions"]
    r_obs, r_rew, r_done, acts, sizes = philox_rollout(ob, K, seed, base, 0, A, ref3d, normalise=dynamic)
    nxt = None
    if dynamic:
        nxt = np.stack([philox.draws(seed, np.arange(base, base + n), k, A, ob.n_plans, ref3d)[2] for k in range(K)]).astype(np.int32)
    obs, rew, done, err = cb.rollout(acts, sizes, nxt, normalise=dynamic)
    assert err == 0
    assert np.array_equal(obs, r_obs) and np.array_equal(rew, r_rew) and np.array_equal(done, r_done)
    assert np.array_equal(cb.ep_cnt, ob.ep_cnt) and np.array_equal(cb.ep_len, ob.ep_len)
    assert np.array_equal(cb.ep_ret, ob.ep_ret) and np.array_equal(cb.ep_iou, ob.ep_iou, equal_nan=True)
    g1, s1 = ob.export()
    g2, s2 = cb.export()
    assert np.array_equal(g1.reshape(g2.shape), g2) and np.array_equal(s1[:, :6], s2[:, :6])
    assert np.array_equal(ob.iou(), cb.iou(), equal_nan=True)
