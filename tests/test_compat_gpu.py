"""The reference-facing classes (snac_b200.compat) replay the reference's golden traces when driven
exactly like the reference: np.random.seed(s); env.reset(); env.step(a) -- same global-RNG draws."""
import numpy as np
import pytest
import torch

from conftest import TRACE_NAMES, load_plans, load_trace

pytestmark = pytest.mark.gpu


def make_compat(meta):
    import snac_b200 as S
    dim, kind, kw = meta["dim"], meta["kind"], meta["kw"]
    if kind == "static":
        cls = {"1D": S.deep_mobile_printing_1d1r, "2D": S.deep_mobile_printing_2d1r, "3D": S.deep_mobile_printing_3d1r}[dim]
        return cls(plan_choose=kw["plan_choose"])
    cls = {"1D": S.deep_mobile_printing_1d1r_dynamic, "2D": S.deep_mobile_printing_2d1r_dynamic,
           "3D": S.deep_mobile_printing_3d1r_dynamic}[dim]
    return cls(plans=load_plans(int(dim[0]), kw["density"], kw["split"]))


UNIFORM_TRACES = [n for n in TRACE_NAMES if not any(k in n for k in ("dropheavy", "refp", "greedy", "builder"))]


@pytest.mark.parametrize("name", UNIFORM_TRACES)
def test_scalar_classes_follow_global_numpy_rng(name):
    tr = load_trace(name)
    meta = tr["meta"]
    dyn = meta["kind"] == "dynamic"
    dim = meta["dim"]
    env = make_compat(meta)
    np.random.seed(meta["seed"])
    rng = np.random.RandomState(meta["seed"] + 1)
    ri = 0
    o = env.reset()
    if dyn:
        assert env.index_random == tr["reset_plan_idx"][ri]
    T = 700
    for t in range(T):
        a = int(rng.randint(env.action_dim))
        assert a == tr["actions"][t]
        o, r, d = env.step(a)
        assert env.step_size == tr["step_sizes"][t]
        raw = o if not dyn else (o[0] if dim == "1D" else None)
        if raw is not None:
            assert raw.dtype == np.float64 and raw.shape == (1, env.state_dim)
            assert np.array_equal(raw[0], tr["obs"][t].astype(np.float64)), t
        if dyn:
            nrm = o[1] if dim == "1D" else o[0]
            assert np.array_equal(nrm[0, :-2], tr["obs"][t, :-2].astype(np.float64))
            assert np.array_equal(nrm[0, -2:], tr["obs_norm"][t]), (t, nrm[0, -2:], tr["obs_norm"][t])
            if dim != "1D":
                assert list(o[2]) == list(tr["pos"][t]) and o[1].shape == (20, 20)
        assert r == tr["reward"][t] and isinstance(r, int) == bool(tr["reward_is_int"][t]), (t, r)
        assert d == bool(tr["done"][t]) and isinstance(d, bool)
        assert env.count_step == tr["count_step"][t] and env.count_brick == tr["count_brick"][t]
        assert env.total_brick == tr["total_brick"][t]
        if dim == "1D":
            assert env.conut_brick == env.count_brick and env.position_memory[-1] == tr["pos"][t][0]
        else:
            assert list(env.position_memory[-1]) == list(tr["pos"][t])
        if t % 50 == 0 and dim != "2D":
            assert env.iou() == tr["iou"][t] or np.isnan(tr["iou"][t])
        if d:
            ri += 1
            env.reset()
            if dyn:
                assert env.index_random == tr["reset_plan_idx"][ri]


def test_vectorized_wrapper_shapes_like_multiprocess_py():
    """multiprocess.py:78-87 prints (N,1,D), (N,), (N,) -- SURVEY.md 3.1."""
    import snac_b200 as S
    for cls, D in ((S.deep_mobile_printing_1d1r, 7), (S.deep_mobile_printing_2d1r, 51), (S.deep_mobile_printing_3d1r, 51)):
        env = S.VectorizedEnvWrapper(cls(plan_choose=0), num_envs=5)
        o = env.reset()
        assert o.shape == (5, 1, D) and o.dtype == np.float64
        for _ in range(20):
            o, r, d = env.step(np.random.randint(3, size=5))
        assert o.shape == (5, 1, D) and r.shape == (5,) and d.shape == (5,) and d.dtype == bool
        one = env.reset_at(2)
        assert one.shape == (1, D)


def test_vectorized_wrapper_equals_independent_reference_style_envs():
    """N wrapper envs == N independent scalar envs fed the same per-env step sizes (App. C.6)."""
    import snac_b200 as S
    from oracle import dmp_oracle as O
    n = 7
    env = S.VectorizedEnvWrapper(S.deep_mobile_printing_2d1r(plan_choose=1), num_envs=n)
    orcs = [O.make_env(2, False, plan_choose=1) for _ in range(n)]
    np.random.seed(11)
    o = env.reset()
    for i, e in enumerate(orcs):
        assert np.array_equal(o[i], e.reset())
    rng = np.random.RandomState(5)
    for t in range(300):
        acts = rng.randint(5, size=n)
        state = np.random.get_state()
        sizes = np.random.randint(1, 4, size=n)
        np.random.set_state(state)
        o, r, d = env.step(acts)
        for i, e in enumerate(orcs):
            oo, rr, dd = e.step(int(acts[i]), int(sizes[i]))
            assert np.array_equal(o[i], oo) and r[i] == rr and d[i] == dd
            if dd:
                e.reset()
                env.reset_at(i)


@pytest.mark.parametrize("dim,dynamic", [(2, False), (3, True), (1, True)])
@pytest.mark.parametrize("mapped", [False, True], ids=["staged", "mapped"])
def test_vectorized_wrapper_shard_pipeline_is_invisible(dim, dynamic, mapped):
    """The staged wrapper runs a step as a pipeline over device shards (numpy step sizes drawn shard by shard while the
    previous shard's results are copied out): same arrays, same consumption of the global numpy stream as one shard -- and
    as the mapped small-batch path."""
    import snac_b200 as S
    n, T = 1003, 40
    def proto():
        if not dynamic:
            return {1: S.deep_mobile_printing_1d1r, 2: S.deep_mobile_printing_2d1r, 3: S.deep_mobile_printing_3d1r}[dim](plan_choose=0)
        cls = {1: S.deep_mobile_printing_1d1r_dynamic, 2: S.deep_mobile_printing_2d1r_dynamic, 3: S.deep_mobile_printing_3d1r_dynamic}[dim]
        return cls(plans=load_plans(dim, "dense", "val"))
    a = S.VectorizedEnvWrapper(proto(), num_envs=n, auto_reset=True, mapped=False, shards=4)
    b = S.VectorizedEnvWrapper(proto(), num_envs=n, auto_reset=True, mapped=mapped, shards=1)
    assert a.n_shards == 4 and b.n_shards == 1 and [c for _, c in a._bounds] == [251, 251, 251, 250]
    A = a.vec.action_dim
    rng = np.random.RandomState(3)
    np.random.seed(21)
    oa = a.reset()
    sa = np.random.get_state()
    np.random.seed(21)
    ob_ = b.reset()
    assert np.array_equal(oa, ob_) and np.array_equal(sa[1], np.random.get_state()[1]) and sa[2] == np.random.get_state()[2]
    for t in range(T):
        acts = rng.randint(A, size=n)
        state = np.random.get_state()
        oa, ra, da = a.step(acts)
        after = np.random.get_state()
        np.random.set_state(state)
        ob_, rb, db = b.step(acts)
        assert np.array_equal(oa, ob_) and np.array_equal(ra, rb) and np.array_equal(da, db), t
        assert np.array_equal(after[1], np.random.get_state()[1]) and after[2] == np.random.get_state()[2]
    assert da.any() or dim != 3                               # 3D episodes are short: auto-resets happened
    one = a.reset_at(700)
    assert one.shape == (1, a._D)


def test_multiprocess_cli_config1():
    """BASELINE config 1: --env 1DStatic --plan_type 2 --num_envs 5."""
    from snac_b200.compat import main
    o, r, d = main(["--env", "1DStatic", "--plan_type", "2", "--num_envs", "5"])
    assert o.shape == (5, 1, 7) and r.shape == (5,) and d.shape == (5,)
    assert main([]) is None and main(["--env", "2DStatic"]) is None


def test_invalid_action_raises_unboundlocalerror():
    import snac_b200 as S
    env = S.deep_mobile_printing_2d1r(plan_choose=0)
    env.reset()
    with pytest.raises(UnboundLocalError):
        env.step(5)
    assert env.count_step == 1          # the reference increments before it fails
    o, r, d = env.step(4)
    assert r == 5.0 or r == 0


@pytest.mark.parametrize("name", ["1d_p1", "2d_dense", "3d_dense", "3d_sparse_uniform"])
def test_lnet_observation_variants_replay_reference(name):
    """*_Lnet classes vs traces of the unmodified reference Lnet classes."""
    import json
    import os
    import snac_b200 as S
    from conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, "lnet_%s.npz" % name))
    meta = json.loads(str(z["meta"]))
    cls = {"1D": S.deep_mobile_printing_1d1r_Lnet, "2D": S.deep_mobile_printing_2d1r_Lnet,
           "3D": S.deep_mobile_printing_3d1r_Lnet}[meta["dim"]]
    env = cls(plan_choose=meta["kw"]["plan_choose"])
    flat = lambda o: (o if meta["dim"] == "1D" else o[0])[0]
    ri = 0
    assert np.array_equal(flat(env.reset()), z["reset_obs"][ri])
    for t in range(len(z["actions"])):
        o, r, d = env.step(int(z["actions"][t]), int(z["step_sizes"][t]))
        assert np.array_equal(flat(o), z["obs"][t]), (t, flat(o), z["obs"][t])
        assert r == z["reward"][t] and d == bool(z["done"][t]), t
        if meta["dim"] != "1D":
            assert list(o[1]) == list(z["pos"][t])
        if d:
            ri += 1
            assert np.array_equal(flat(env.reset()), z["reset_obs"][ri])
    assert np.array_equal(env.environment_memory.astype(np.int16).reshape(z["final_grid"].shape), z["final_grid"])


def test_flat_observation_wrappers_like_the_sac_and_ppo_env_copies():
    """script/SAC/environments/*.py return (D,) observations; script/PPO/*/DMP_*.py add gym spaces and a 4-tuple."""
    import snac_b200 as S
    for cls, D, A in ((S.deep_mobile_printing_1d1r, 7, 3), (S.deep_mobile_printing_2d1r, 51, 5), (S.deep_mobile_printing_3d1r, 51, 8)):
        np.random.seed(5)
        base = cls(plan_choose=0)
        want = [base.reset()] + [base.step(a % A)[0] for a in range(12)]
        np.random.seed(5)
        sac, ppo = S.FlatObsEnv(cls(plan_choose=0)), S.FlatObsEnv(cls(plan_choose=0), gym_api=True)
        o = sac.reset()
        assert o.shape == (D,) and np.array_equal(o, want[0][0])
        for a in range(12):
            o, r, d = sac.step(a % A)
            assert o.shape == (D,) and np.array_equal(o, want[a + 1][0])
        assert ppo.reset().shape == (D,)
        o, r, d, info = ppo.step(0)
        assert o.shape == (D,) and info == {} and isinstance(d, bool)
        assert ppo.action_space.n == A and ppo.observation_space.shape == (D,) and ppo.observation_space.contains(o)
        assert ppo.total_step == base.total_step and ppo.HALF_WINDOW_SIZE == base.HALF_WINDOW_SIZE
