"""BASELINE.json's configurations at their FULL sizes (65 536 1D / 1 048 576 2D / 262 144 3D envs): the CUDA rollout
against the compiled oracle (oracle/dmp_oracle.c, itself pinned to the reference's golden traces) on the same Philox
draws -- rewards and done flags of every step, final grids / positions / counters, per-env episode statistics, IoU,
and the observations of the last step, all with exact equality."""
import numpy as np
import pytest
import torch

from conftest import load_plans
from oracle import dmp_oracle as O
from oracle import philox
from oracle.c_oracle import COracleBatch

pytestmark = pytest.mark.gpu
SEED = 0x534E4143

CASES = [
    # name, dim, dynamic, density, n, K, total_step override (None = the reference's), ref3d actions
    ("cfg2_1d_dynamic", 1, True, "dense", 65536, 48, None, False),
    ("cfg2_1d_dynamic_short_episodes", 1, True, "dense", 65536, 96, 40, False),
    ("cfg3_2d_static_dense", 2, False, None, 1048576, 24, None, False),
    ("cfg3_2d_static_dense_short_episodes", 2, False, None, 1048576, 40, 16, False),
    ("cfg4_2d_dynamic_dense_short_episodes", 2, True, "dense", 1048576, 40, 16, False),
    ("cfg5_3d_static_dense", 3, False, None, 262144, 64, None, False),
    ("cfg5_3d_dynamic_dense_ref_actions", 3, True, "dense", 262144, 64, None, True),
]


@pytest.mark.parametrize("name,dim,dynamic,density,n,K,total_step,ref3d", CASES, ids=[c[0] for c in CASES])
def test_full_size_rollout_equals_compiled_oracle(name, dim, dynamic, density, n, K, total_step, ref3d):
    import psutil
    if psutil.virtual_memory().available < 24 * 2**30 and n > 300000:
        pytest.skip("needs ~15 GB of host memory for the oracle's 10^6 float64 grids")
    from snac_b200.vecenv import BatchedDMPEnv
    plans = load_plans(dim, density, "train") if dynamic else None
    A = O.SPEC[dim]["actions"]
    env = BatchedDMPEnv(dim, dynamic=dynamic, plan_choose=0, plans=plans, num_envs=n, auto_reset=True, seed=SEED,
                        action_dist="ref3d" if ref3d else "uniform", total_step=total_step)
    cb = COracleBatch(dim, dynamic, n, 0, plans)
    if total_step is not None:
        cb.cfg.total_step = total_step
    ids = np.arange(n)
    p0 = philox.reset_draw(SEED, ids, 0, cb.n_plans).astype(np.int32) if dynamic else None
    o0 = env.reset().cpu().numpy().astype(np.float64)
    assert np.array_equal(o0, cb.reset(p0))
    acts = np.empty((K, n), np.uint8)
    sizes = np.empty((K, n), np.uint8)
    nxt = np.empty((K, n), np.int32) if dynamic else None
    for k in range(K):
        s, a, p = philox.draws(SEED, ids, k, A, cb.n_plans, ref3d, with_plan=dynamic)
        acts[k], sizes[k] = a, s
        if dynamic:
            nxt[k] = p
    # K - 1 steps in one launch without observations, the last step with them
    _, rew, done = env.rollout(K - 1, materialise_obs=False)
    obs_l, rew_l, done_l = env.rollout(1)
    torch.cuda.synchronize()
    _, r_rew, r_done, err = cb.rollout(acts[:K - 1], sizes[:K - 1], None if nxt is None else nxt[:K - 1], want_obs=False)
    assert err == 0
    r_obs_l, r_rew_l, r_done_l, err = cb.rollout(acts[K - 1:], sizes[K - 1:], None if nxt is None else nxt[K - 1:])
    assert err == 0
    assert np.array_equal(rew.cpu().numpy(), r_rew)
    assert np.array_equal(done.cpu().numpy(), r_done)
    assert np.array_equal(obs_l.cpu().numpy().astype(np.float64), r_obs_l)
    assert np.array_equal(rew_l.cpu().numpy(), r_rew_l) and np.array_equal(done_l.cpu().numpy(), r_done_l)
    st = env.export_state()
    g_ref, sc_ref = cb.export()
    assert np.array_equal(st["grid"].cpu().numpy().reshape(g_ref.shape), g_ref)
    assert np.array_equal(st["scalars"].cpu().numpy()[:, :6], sc_ref[:, :6])
    cnt, ln, ret, iou = [x.cpu().numpy() for x in env.episode_stats()]
    assert np.array_equal(cnt, cb.ep_cnt) and np.array_equal(ln, cb.ep_len)
    assert np.array_equal(ret, cb.ep_ret)
    assert np.array_equal(iou, cb.ep_iou, equal_nan=True)
    assert np.array_equal(env.iou().cpu().numpy(), cb.iou(), equal_nan=True)
    if total_step is not None or dim == 3:
        assert cb.ep_cnt.min() >= 1 if dim != 3 else cb.ep_cnt.sum() > n // 4     # episodes finished and were reset
    stats = env.stats().cpu().numpy()
    assert stats[2] == cb.ep_cnt.sum() and stats[3] == cb.ep_len.sum() and stats[0] == cb.ep_ret.sum()
    env.check_errors()
