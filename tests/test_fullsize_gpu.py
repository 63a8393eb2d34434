"""BASELINE.json's configurations at their FULL sizes (65 536 1D / 1 048 576 2D / 262 144 3D envs): the CUDA rollout
against the compiled oracle (oracle/dmp_oracle.c, itself pinned to the reference's golden traces) on the same Philox
draws -- observations, rewards and done flags of EVERY step (one K-step launch with everything materialised: exactly the
launch bench.py times, bulk copy-out of 8 192 blocks x K observation tiles included), final grids / positions /
counters, per-env episode statistics, IoU, all with exact equality."""
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
    ("cfg3_2d_static_dense", 2, False, None, 1048576, 16, None, False),
    ("cfg3_2d_static_dense_short_episodes", 2, False, None, 1048576, 24, 10, False),
    ("cfg4_2d_dynamic_dense_short_episodes", 2, True, "dense", 1048576, 24, 10, False),
    # the 8-GPU shard of cfg 3 / cfg 4 runs in 224-thread blocks (one wave instead of 1.15): full and ragged
    ("cfg3_2d_static_dense_shard_of_8", 2, False, None, 131072, 24, 10, False),
    ("cfg4_2d_dynamic_dense_ragged_wide_blocks", 2, True, "dense", 120001, 24, 10, False),
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
    # ONE launch of K steps with observations, rewards and done flags materialised; the oracle is stepped one vector step
    # at a time so that only one step's float64 observations (428 MB at 1 048 576 envs) are alive on the host
    obs, rew, done = env.rollout(K)
    torch.cuda.synchronize()
    for k in range(K):
        r_obs, r_rew, r_done, err = cb.rollout(acts[k:k + 1], sizes[k:k + 1], None if nxt is None else nxt[k:k + 1])
        assert err == 0
        assert np.array_equal(obs[k].cpu().numpy(), r_obs[0].astype(np.float32)), (name, k)
        assert np.array_equal(rew[k].cpu().numpy(), r_rew[0]) and np.array_equal(done[k].cpu().numpy(), r_done[0]), (name, k)
    del obs
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


@pytest.mark.parametrize("dim,n,total_step", [(2, 1048576, 10), (3, 262144, None)], ids=["cfg3_2d_bits_host", "cfg5_3d_bits_host"])
def test_full_size_host_steps_with_bit_records_equal_compiled_oracle(dim, n, total_step):
    """The call bench.py's `e2e` times, at BASELINE's size: HostStepper.step(numpy actions) -> bit records in pinned host
    memory (2D: written by the kernel itself through mapped memory), every step against the compiled oracle."""
    import psutil
    if psutil.virtual_memory().available < 24 * 2**30 and n > 300000:
        pytest.skip("needs ~15 GB of host memory for the oracle's 10^6 float64 grids")
    from snac_b200.compat import HostStepper
    from snac_b200.vecenv import BatchedDMPEnv, unpack_bits
    A, T = O.SPEC[dim]["actions"], 14
    env = BatchedDMPEnv(dim, plan_choose=0, num_envs=n, auto_reset=True, seed=SEED, obs_dtype="bits", total_step=total_step)
    cb = COracleBatch(dim, False, n, 0, None)
    if total_step is not None:
        cb.cfg.total_step = total_step
    o0, _, _, _ = unpack_bits(env.reset(), dim)
    assert np.array_equal(o0, cb.reset(None))
    hs = HostStepper(env)
    assert hs.mapped == (dim == 2) and hs.d2h_bytes == n * (16 if dim == 2 else 32)
    rng = np.random.RandomState(17)
    for t in range(T):
        acts = rng.randint(0, A, size=n).astype(np.uint8)
        sizes = rng.randint(1, 4, size=n).astype(np.uint8)
        rec = hs.step(acts, sizes)
        r_obs, r_rew, r_done, err = cb.rollout(acts[None], sizes[None], None)
        o, r, d, sat = unpack_bits(rec, dim, np.float32)
        assert err == 0 and not sat.any()
        assert np.array_equal(o, r_obs[0].astype(np.float32)), t
        assert np.array_equal(r, r_rew[0]) and np.array_equal(d, r_done[0]), t
    assert cb.ep_cnt.sum() > 0
    env.check_errors()
