"""``random_choose_paln=False``: the dataset classes walk their plan list in order and wrap around
(Env/2D/DMP_Env_2D_dynamic_usedata_plan.py:39-44; every evaluation script uses it with the ``test`` pkl, e.g.
script/DRQN/2d/test_DRQN_2d_dynamic.py:38).  DMP_PLAN_SEQUENTIAL does that inside the kernels; SURVEY.md App. C.5 asks
for the wrap-around after 50 resets."""
import numpy as np
import pytest
import torch

from conftest import load_plans
from oracle import dmp_oracle as O
from oracle_batch import OracleBatch, philox_rollout

pytestmark = pytest.mark.gpu
SEED = 0x534E4143


def make_gpu(dim, n, plans, **kw):
    from snac_b200.vecenv import BatchedDMPEnv
    return BatchedDMPEnv(dim, dynamic=True, plans=plans, num_envs=n, random_choose_paln=False, **kw)


@pytest.mark.parametrize("dim,density", [(1, "dense"), (2, "dense"), (2, "sparse"), (3, "dense"), (3, "sparse")])
@pytest.mark.parametrize("mode", ["rollout", "step"])
def test_in_kernel_auto_reset_walks_the_test_set_in_order_and_wraps(dim, density, mode):
    plans = load_plans(dim, density, "test")
    assert len(plans) == 50
    n, T = 41, 9                                              # short episodes: 50 resets come round within ~460 steps
    K = T * 54
    env = make_gpu(dim, n, plans, auto_reset=True, seed=SEED, env_base=3, total_step=T, normalise=True,
                   obs_dtype=torch.float64)
    ob = OracleBatch(dim, True, n, 0, plans, sequential=True)
    for e in ob.envs:
        e.total_step = T
    o = env.reset()                                           # first reset of every env: plan 0 (index_for_non_random = 0)
    assert np.array_equal(o.cpu().numpy(), ob.reset(np.zeros(n, np.int32)))
    assert (env.export_state()["scalars"][:, 4] == 0).all()
    r_obs, r_rew, r_done, _, _ = philox_rollout(ob, K, SEED, 3, 0, O.SPEC[dim]["actions"], normalise=True)
    if mode == "rollout":
        obs, rew, done = env.rollout(K)
    else:
        outs = [[x.clone() for x in env.step(None)] for _ in range(K)]
        obs, rew, done = [torch.stack([o[i] for o in outs]) for i in range(3)]
    torch.cuda.synchronize()
    assert np.array_equal(obs.cpu().numpy(), r_obs)
    assert np.array_equal(rew.cpu().numpy(), r_rew) and np.array_equal(done.cpu().numpy(), r_done)
    g_ref, sc_ref = ob.export()
    st = env.export_state()
    assert np.array_equal(st["grid"].cpu().numpy().reshape(g_ref.shape), g_ref)
    assert np.array_equal(st["scalars"].cpu().numpy()[:, :6], sc_ref[:, :6])          # incl. plan index and budget
    cnt, ln, ret, iou = [x.cpu().numpy() for x in env.episode_stats()]
    assert np.array_equal(cnt, ob.ep_cnt) and np.array_equal(ln, ob.ep_len) and np.array_equal(ret, ob.ep_ret)
    assert np.array_equal(iou, ob.ep_iou, equal_nan=True)
    assert ob.ep_cnt.min() >= 51                              # every env has wrapped past plan 49 back to plan 0
    assert np.array_equal(sc_ref[:, 4], ob.ep_cnt % 50)
    env.check_errors()


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_caller_side_resets_follow_the_same_order(dim):
    """reset() without plan_idx: plan 0 first, then +1 per reset, wrapping -- also for a masked first reset
    (reset_at before any full reset) and through the raw C ABI on zero-initialised state."""
    plans = load_plans(dim, "dense", "test")
    n = 6
    env = make_gpu(dim, n, plans, seed=SEED)
    m = np.zeros(n, np.uint8)
    m[[1, 4]] = 1
    env.reset(mask=m)                                         # envs 1 and 4 start at plan 0 ...
    env.reset()                                               # ... and move on to plan 1 while the others start at 0
    want = np.where(m, 1, 0)
    assert np.array_equal(env.export_state()["scalars"][:, 4].cpu().numpy(), want)
    for i in range(1, 120):
        env.reset()
        assert np.array_equal(env.export_state()["scalars"][:, 4].cpu().numpy(), (want + i) % 50)
        if i == 60:
            assert np.array_equal(env.plan_totals().cpu().numpy()[(want + i) % 50],
                                  env.export_state()["scalars"][:, 5].cpu().numpy())


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_scalar_classes_sequential_mode_like_the_reference(dim):
    """The drop-in classes with random_choose_paln=False: 52 episodes over the 50-plan test set against the oracle, the
    plan attribute following the dataset in order and index_for_non_random wrapping (…usedata_plan.py:39-44)."""
    import snac_b200 as S
    plans = load_plans(dim, "dense", "test")
    cls = {1: S.deep_mobile_printing_1d1r_dynamic, 2: S.deep_mobile_printing_2d1r_dynamic,
           3: S.deep_mobile_printing_3d1r_dynamic}[dim]
    env = cls(plans=plans, random_choose_paln=False)
    orc = O.make_env(dim, True, plans=plans)
    rng = np.random.RandomState(4)
    A = O.SPEC[dim]["actions"]
    np.random.seed(8)
    for ep in range(52):
        o = env.reset()
        oo = orc.reset(ep % 50)
        assert env.index_for_non_random == (ep + 1) % 50
        assert np.array_equal(env.plan, plans[ep % 50])
        first = o[0] if dim == 1 else o[0]
        assert np.array_equal(np.asarray(first), oo if dim == 1 else orc.obs_normalised())
        for t in range(12):
            a = int(rng.randint(A))
            state = np.random.get_state()
            s = int(np.random.randint(1, 4))
            np.random.set_state(state)
            o, r, d = env.step(a)
            o2, r2, d2 = orc.step(a, s)
            got = np.asarray(o[0])
            assert np.array_equal(got, o2 if dim == 1 else orc.obs_normalised()) and r == r2 and d == d2, (ep, t)
            if d:
                break
