"""The *_hindsight_replay classes of snac_b200.compat and the learners' hindsight relabelling (overwrite ``env.plan``
between reset() and the replay, by assignment or in place) against traces of the unmodified reference
(tests/golden/make_hindsight_golden.py) -- same numpy seed, same global-RNG consumption, same rewards."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_plans

pytestmark = pytest.mark.gpu

CASES = {
    # name: (dim, ctor, probabilities of the recorded action stream, max steps)
    "1d_generator": ("1D", lambda S: S.deep_mobile_printing_1d1r_hindsight(), [.25, .35, .4], 160),
    "2d_dataset_dense": ("2D", lambda S: S.deep_mobile_printing_2d1r_hindsight(plans=load_plans(2, "dense", "train"), plan_choose=0),
                         [.15, .2, .2, .1, .35], 160),
    "2d_dataset_sparse": ("2D", lambda S: S.deep_mobile_printing_2d1r_hindsight(plans=load_plans(2, "sparse", "val"), plan_choose=1),
                          [.15, .2, .2, .1, .35], 120),
    "3d_dataset_dense": ("3D", lambda S: S.deep_mobile_printing_3d1r_hindsight(plans=load_plans(3, "dense", "train")),
                         [.1, .15, .15, .1, .1, .15, .15, .1], 200),
    "1d_static": ("1D", lambda S: S.deep_mobile_printing_1d1r_hindsight_static(plan_choose=1), [.25, .35, .4], 160),
    "2d_static": ("2D", lambda S: S.deep_mobile_printing_2d1r_hindsight_static(plan_choose=1), [.15, .2, .2, .1, .35], 160),
    "3d_static": ("3D", lambda S: S.deep_mobile_printing_3d1r_hindsight_static(plan_choose=0),
                  [.1, .15, .15, .1, .1, .15, .15, .1], 200),
}


def raw(o):
    return np.asarray(o[0] if isinstance(o, list) else o, dtype=np.float64).reshape(-1)


def relabel(env_h, env, dim):
    h = env.HALF_WINDOW_SIZE
    if dim == "1D":
        env_h.plan = env.environment_memory[0, h:h + env.plan_width]
    else:
        env_h.plan[h:h + env.plan_height, h:h + env.plan_width] = env.environment_memory[h:h + env.plan_height, h:h + env.plan_width]
        env_h.input_plan = env_h.plan[h:h + env.plan_height, h:h + env.plan_width]


@pytest.mark.parametrize("name", sorted(CASES))
def test_hindsight_classes_and_relabelling_follow_the_reference(name):
    import snac_b200 as S
    z = np.load(os.path.join(GOLDEN, "hindsight_golden.npz"))
    g = lambda k: z["%s/%s" % (name, k)]
    dim, ctor, p, T = CASES[name]
    seed = int(g("seed"))
    np.random.seed(seed)
    env, env_h = ctor(S), ctor(S)
    arng = np.random.RandomState(seed + 1000)
    for ep in range(int(g("n_episodes"))):
        k = "ep%d_" % ep
        o = env.reset()
        assert np.array_equal(np.asarray(env.plan, dtype=np.float64), g(k + "plan")), (name, ep)
        assert float(env.total_brick) == float(g(k + "total_brick"))
        if int(g(k + "plan_idx")) >= 0:
            assert env.index_random == int(g(k + "plan_idx"))
        if name == "1d_generator":
            assert np.array_equal(np.asarray(env.one_hot, dtype=np.float64), g(k + "one_hot"))
            assert isinstance(o, list) and len(o) == 2 and np.array_equal(o[1], env.plan)
        assert np.array_equal(raw(o), g(k + "obs")[0])
        acts, sizes = g(k + "actions"), g(k + "sizes")
        for t in range(len(acts)):
            a, s = int(arng.choice(len(p), p=p)), int(arng.randint(1, 4))
            assert (a, s) == (int(acts[t]), int(sizes[t]))
            o, r, d = env.step(a, s)
            assert np.array_equal(raw(o), g(k + "obs")[t + 1]), (name, ep, t)
            assert r == g(k + "reward")[t] and d == bool(g(k + "done")[t]), (name, ep, t, r, d)
            if dim != "3D":
                assert isinstance(r, int) == bool(g(k + "reward_is_int")[t]), (name, ep, t, r)
        assert d or len(acts) == T
        assert np.array_equal(env.environment_memory, g(k + "final_grid"))
        # hindsight replay against what was built
        oh = env_h.reset()
        if int(g(k + "h_plan_idx")) >= 0:
            assert env_h.index_random == int(g(k + "h_plan_idx"))
        assert float(env_h.total_brick) == float(g(k + "h_total_brick"))
        assert np.array_equal(raw(oh), g(k + "h_obs")[0])
        relabel(env_h, env, dim)
        for t in range(len(acts)):
            o, r, d = env_h.step(int(acts[t]), int(sizes[t]))
            assert np.array_equal(raw(o), g(k + "h_obs")[t + 1]), (name, ep, t)
            assert r == g(k + "h_reward")[t] and d == bool(g(k + "h_done")[t]), (name, ep, t, r, d)
        assert np.array_equal(env_h.environment_memory, g(k + "h_final_grid"))
        if dim != "2D":
            want = float(g(k + "h_iou"))
            assert env_h.iou() == want or (np.isnan(want) and np.isnan(env_h.iou()))
