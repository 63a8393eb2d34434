"""Packed step records (DMP_OBS_REC and the bit-packed DMP_OBS_BITS, include/dmp.h) and the host-buffer callers built on them: the same steps as the
numeric observation kinds, bit for bit, against the oracle; HostStepper / VectorizedEnvWrapper return what a direct
device-side step returns."""
import numpy as np
import pytest
import torch

from conftest import load_plans
from oracle import dmp_oracle as O
from oracle import philox
from oracle_batch import OracleBatch, philox_rollout

pytestmark = pytest.mark.gpu
SEED = 0x534E4143


def make_gpu(dim, dynamic, n, plan_choose=0, plans=None, **kw):
    from snac_b200.vecenv import BatchedDMPEnv
    return BatchedDMPEnv(dim, dynamic=dynamic, plan_choose=plan_choose, plans=plans, num_envs=n, **kw)


CASES = [
    # dim, dynamic, density, ref3d, n (even / odd: an odd n leaves the record rows of odd steps 8 B aligned only), K
    (1, False, None, False, 200, 800), (1, True, "dense", False, 131, 900),
    (2, False, None, False, 256, 650), (2, True, "sparse", False, 97, 400),
    (3, False, None, True, 160, 500), (3, True, "dense", False, 75, 300),
]


@pytest.mark.parametrize("dim,dynamic,density,ref3d,n,K", CASES)
def test_record_rollout_matches_oracle(dim, dynamic, density, ref3d, n, K):
    from snac_b200.vecenv import record_dtype, unpack_records
    plans = load_plans(dim, density, "train") if dynamic else None
    kw = dict(auto_reset=True, env_base=77, seed=SEED, action_dist="ref3d" if ref3d else "uniform")
    env = make_gpu(dim, dynamic, n, 0, plans, obs_dtype="record", **kw)
    ob = OracleBatch(dim, dynamic, n, 0, plans)
    p0 = philox.reset_draw(SEED, np.arange(77, 77 + n), 0, ob.n_plans) if dynamic else None
    o0, r0, d0, s0 = unpack_records(env.reset(), dim)
    assert np.array_equal(o0, ob.reset(p0)) and not r0.any() and not d0.any() and not s0.any()
    assert env.reset().shape == (n, record_dtype(dim).itemsize)
    r_obs, r_rew, r_done, _, _ = philox_rollout(ob, K, SEED, 77, 0, O.SPEC[dim]["actions"], ref3d)
    K1 = K - 40
    rec, rew, done = env.rollout(K1)                          # one launch ...
    assert rec.dtype == torch.uint8 and rec.shape == (K1, n, record_dtype(dim).itemsize)
    o, r, d, sat = unpack_records(rec, dim)
    assert np.array_equal(o, r_obs[:K1]) and np.array_equal(r, r_rew[:K1]) and np.array_equal(d, r_done[:K1]) and not sat.any()
    assert np.array_equal(rew.cpu().numpy(), r_rew[:K1]) and np.array_equal(done.cpu().numpy(), r_done[:K1])
    for k in range(K1, K):                                    # ... then single steps (3D: the other kernel)
        rec, rew, done = env.step(None)
        o, r, d, _ = unpack_records(rec, dim)
        assert np.array_equal(o, r_obs[k]) and np.array_equal(r, r_rew[k]) and np.array_equal(d, r_done[k]), k
    g_ref, sc_ref = ob.export()
    st = env.export_state()
    assert np.array_equal(st["grid"].cpu().numpy().reshape(g_ref.shape), g_ref)
    assert np.array_equal(st["scalars"].cpu().numpy()[:, :6], sc_ref[:, :6])
    assert np.array_equal(env.episode_stats()[0].cpu().numpy(), ob.ep_cnt) and ob.ep_cnt.sum() > 0
    env.check_errors()


BITS_CASES = [
    # dim, dynamic, density, ref3d, n (ragged last warp / block), K
    (2, False, None, False, 256, 650), (2, True, "sparse", False, 97, 400), (2, True, "dense", False, 333, 500),
    (3, False, None, True, 160, 500), (3, True, "dense", False, 75, 300), (3, True, "sparse", True, 201, 400),
]


@pytest.mark.parametrize("dim,dynamic,density,ref3d,n,K", BITS_CASES)
def test_bit_record_rollout_matches_oracle(dim, dynamic, density, ref3d, n, K):
    """DMP_OBS_BITS: 49 two-bit (2D) / four-bit (3D) window codes + 12-bit counters + reward code + done in 16 / 32 bytes per
    env-step, against the oracle: one rollout launch, then single steps (3D: the other kernel); the device-side expansion
    (dmp_records_unpack) must give the same rows as the numpy one."""
    from snac_b200.vecenv import bits_bytes, unpack_bits, unpack_records_device
    plans = load_plans(dim, density, "train") if dynamic else None
    kw = dict(auto_reset=True, env_base=77, seed=SEED, action_dist="ref3d" if ref3d else "uniform")
    env = make_gpu(dim, dynamic, n, 0, plans, obs_dtype="bits", **kw)
    ob = OracleBatch(dim, dynamic, n, 0, plans)
    p0 = philox.reset_draw(SEED, np.arange(77, 77 + n), 0, ob.n_plans) if dynamic else None
    o0, r0, d0, s0 = unpack_bits(env.reset(), dim)
    assert env.obs_row == bits_bytes(dim) and env.reset().shape == (n, bits_bytes(dim))
    assert np.array_equal(o0, ob.reset(p0)) and not r0.any() and not d0.any() and not s0.any()
    r_obs, r_rew, r_done, _, _ = philox_rollout(ob, K, SEED, 77, 0, O.SPEC[dim]["actions"], ref3d)
    assert r_obs[..., :49].max() < 14                         # random policies never stack that high: every record is exact
    K1 = K - 40
    rec, rew, done = env.rollout(K1)
    assert rec.dtype == torch.uint8 and rec.shape == (K1, n, bits_bytes(dim))
    o, r, d, sat = unpack_bits(rec, dim)
    assert np.array_equal(o, r_obs[:K1]) and np.array_equal(r, r_rew[:K1]) and np.array_equal(d, r_done[:K1]) and not sat.any()
    assert np.array_equal(rew.cpu().numpy(), r_rew[:K1]) and np.array_equal(done.cpu().numpy(), r_done[:K1])
    for dt in (torch.float32, torch.float64, torch.int16):
        od, rd, dd, sd = unpack_records_device(rec, dim, "bits", dt)
        assert od.dtype == dt and np.array_equal(od.cpu().numpy().astype(np.float64), o)
        assert np.array_equal(rd.cpu().numpy(), r) and np.array_equal(dd.cpu().numpy(), d) and not sd.any()
    for k in range(K1, K):
        rec, rew, done = env.step(None)
        o, r, d, _ = unpack_bits(rec, dim)
        assert np.array_equal(o, r_obs[k]) and np.array_equal(r, r_rew[k]) and np.array_equal(d, r_done[k]), k
    g_ref, sc_ref = ob.export()
    st = env.export_state()
    assert np.array_equal(st["grid"].cpu().numpy().reshape(g_ref.shape), g_ref)
    assert np.array_equal(st["scalars"].cpu().numpy()[:, :6], sc_ref[:, :6])
    assert np.array_equal(env.episode_stats()[0].cpu().numpy(), ob.ep_cnt) and ob.ep_cnt.sum() > 0
    env.check_errors()


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_device_unpack_of_byte_records(dim):
    """dmp_records_unpack on DMP_OBS_REC records == unpack_records on the host."""
    from snac_b200.vecenv import unpack_records, unpack_records_device
    env = make_gpu(dim, False, 1000, 0, auto_reset=True, seed=SEED, obs_dtype="record", total_step=30)
    env.reset()
    rec, _, _ = env.rollout(64)
    o, r, d, sat = unpack_records(rec, dim, np.float32)
    od, rd, dd, sd = unpack_records_device(rec, dim, "record", torch.float32)
    assert np.array_equal(od.cpu().numpy(), o) and np.array_equal(rd.cpu().numpy(), r)
    assert np.array_equal(dd.cpu().numpy(), d) and np.array_equal(sd.cpu().numpy(), sat) and d.any()


def test_bit_record_counters_saturate_with_flag():
    """12-bit counters: an env stepped on after done (no auto-reset) passes 4 095 steps; the record saturates and says so."""
    from snac_b200.vecenv import unpack_bits
    env = make_gpu(2, False, 40, 0, auto_reset=False, seed=SEED, obs_dtype="bits")
    env.reset()
    env.rollout(4090, materialise_obs=False)
    rec, _, _ = env.rollout(10)
    o, r, d, sat = unpack_bits(rec, 2)
    steps = np.arange(4091, 4101)[:, None]
    assert np.array_equal(o[..., 50], np.minimum(steps, 4095) * np.ones((1, 40))) and np.array_equal(sat, (steps > 4095) * np.ones((1, 40), bool))


def test_bits_have_no_1d_form():
    from snac_b200 import _lib as L
    with pytest.raises(ValueError):
        make_gpu(1, False, 4, obs_dtype="bits")
    with pytest.raises(ValueError):
        make_gpu(2, True, 4, plans=load_plans(2, "dense", "val"), obs_dtype="bits", normalise=True)
    lay = L.DmpLayout()
    for dim, b in ((1, 0), (2, 16), (3, 32)):
        L.check(L.lib.dmp_layout(dim, 8, lay))
        assert lay.bits_bytes == b


@pytest.mark.parametrize("dim,n", [(1, 128 * 148 * 4), (2, 128 * 148 * 4), (3, 128 * 148 * 4), (2, 118400)])
def test_records_equal_float_observations_at_scale(dim, n):
    """Same seed, same launch shape as the throughput runs (full blocks, 1D: the specialised instantiation; 2D at 118 400
    envs: the 224-thread blocks of the 8-GPU shard, last block ragged)."""
    from snac_b200.vecenv import unpack_records, unpack_records_device
    K = 40
    a = make_gpu(dim, False, n, 0, auto_reset=True, seed=SEED, obs_dtype="record", total_step=25)
    b = make_gpu(dim, False, n, 0, auto_reset=True, seed=SEED, obs_dtype=torch.float32, total_step=25)
    a.reset(), b.reset()
    rec, ra, da = a.rollout(K)
    ob_, rb, db = b.rollout(K)
    torch.cuda.synchronize()
    o, r, d, sat = unpack_records(rec, dim, np.float32)
    assert np.array_equal(o, ob_.cpu().numpy()) and np.array_equal(r, rb.cpu().numpy()) and np.array_equal(d, db.cpu().numpy())
    assert torch.equal(ra, rb) and torch.equal(da, db) and not sat.any() and d.any()
    assert torch.equal(a.get_state()["cells"], b.get_state()["cells"])
    if dim != 1:                                              # the bit records of the same steps
        c = make_gpu(dim, False, n, 0, auto_reset=True, seed=SEED, obs_dtype="bits", total_step=25)
        c.reset()
        bits, rc, dc = c.rollout(K)
        ou, ru, du, su = unpack_records_device(bits, dim, "bits", torch.float32)
        assert torch.equal(ou, ob_) and torch.equal(ru, rb) and torch.equal(du, db) and not su.any()
        assert torch.equal(rc, rb) and torch.equal(dc, db) and torch.equal(c.get_state()["cells"], b.get_state()["cells"])


def test_records_reject_normalised_counters_and_stage_kernels():
    from snac_b200 import _lib as L
    with pytest.raises(ValueError):
        make_gpu(2, True, 4, plans=load_plans(2, "dense", "val"), obs_dtype="record", normalise=True)
    env = make_gpu(2, False, 4, obs_dtype="record")
    env.reset()
    with pytest.raises(L.DmpError):
        env.step_staged(None)


# ------------------------------------------------------------------------------------------------
# host-buffer callers
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,kind", [(1, torch.float32), (2, torch.float32), (2, torch.int16), (3, torch.float64),
                                      (1, "record"), (2, "record"), (3, "record"), (2, "bits"), (3, "bits")])
@pytest.mark.parametrize("mapped", [False, True, "out"], ids=["staged", "mapped", "mapped_out"])
def test_host_stepper_returns_what_a_device_step_returns(dim, kind, mapped):
    from snac_b200.compat import HostStepper
    from snac_b200.vecenv import unpack_records
    n, T = 300, 60
    A = O.SPEC[dim]["actions"]
    env = make_gpu(dim, False, n, 0, auto_reset=True, seed=SEED, obs_dtype=kind)
    ref = make_gpu(dim, False, n, 0, auto_reset=True, seed=SEED, obs_dtype=torch.float64)
    env.reset(), ref.reset()
    assert HostStepper(env).mapped == (HostStepper(env)._total <= HostStepper.MAPPED_MAX_BYTES or (kind == "bits" and dim == 2) or
                                      (kind == "record" and dim == 1))                      # default: small results, 16 B records
    hs = HostStepper(env, mapped=mapped)                     # mapped: the kernel reads / writes pinned host memory itself
    assert hs.d2h_bytes == (n * env.obs_row if kind in ("record", "bits") else
                            hs._off_done + n) and hs.h2d_bytes == n
    rng = np.random.RandomState(1)
    prev = None
    for t in range(T):
        acts = rng.randint(A, size=n).astype(np.uint8)
        sizes = rng.randint(1, 4, size=n).astype(np.uint8) if t % 2 else None      # injected and in-kernel step sizes
        out = hs.step(acts, sizes)
        o2, r2, d2 = ref.step(acts, sizes)
        if kind == "bits":
            o, r, d, _ = unpack_records(out, dim)
            assert out.dtype == np.uint8 and out.shape == (n, env.obs_row)
        elif kind == "record":
            o, r, d, _ = unpack_records(out, dim)
            assert out.dtype.names and out.shape == (n,)
        else:
            o, r, d = out
            assert o.shape == (n, env.obs_dim) and r.dtype == np.float32 and d.dtype == np.bool_
        assert np.array_equal(np.asarray(o, np.float64), o2.cpu().numpy()) and np.array_equal(r, r2.cpu().numpy())
        assert np.array_equal(d, d2.cpu().numpy())
        if prev is not None and kind not in ("record", "bits"):     # the previous step's arrays are still intact (two buffers)
            assert np.array_equal(prev[0], prev[1])
        if kind not in ("record", "bits"):
            prev = (o, o.copy())
    env.check_errors()


@pytest.mark.parametrize("mapped", [False, True], ids=["staged", "mapped"])
def test_vectorized_wrapper_philox_mode_and_views(mapped):
    """step_size_rng="philox": no host draws; results equal a BatchedDMPEnv stepped with the same actions; the returned
    arrays are float64 / float64 / bool views that survive one more step.  Staged copies and mapped host buffers."""
    import snac_b200 as S
    n = 64
    w = S.VectorizedEnvWrapper(S.deep_mobile_printing_2d1r(plan_choose=0), num_envs=n, step_size_rng="philox", mapped=mapped)
    assert w.mapped == mapped
    ref = make_gpu(2, False, n, 0, obs_dtype=torch.float64)
    o = w.reset()
    assert np.array_equal(o[:, 0, :], ref.reset().cpu().numpy())
    rng = np.random.RandomState(2)
    state = np.random.get_state()[1].copy()
    keep = None
    for t in range(50):
        acts = rng.randint(5, size=n)
        o, r, d = w.step(acts)
        o2, r2, d2 = ref.step(acts.astype(np.uint8))
        assert o.shape == (n, 1, 51) and o.dtype == np.float64 and r.dtype == np.float64 and d.dtype == np.bool_
        assert np.array_equal(o[:, 0, :], o2.cpu().numpy()) and np.array_equal(r, r2.cpu().numpy().astype(np.float64))
        assert np.array_equal(d, d2.cpu().numpy())
        if keep is not None:
            assert np.array_equal(keep[0], keep[1])
        keep = (o, o.copy())
    assert np.array_equal(np.random.get_state()[1], state)    # the global numpy stream was not touched


# ------------------------------------------------------------------------------------------------
# counters are 16 bits in the packed state: stepping on after done without a reset latches DMP_ERR_OVERFLOW
# ------------------------------------------------------------------------------------------------
def test_counter_overflow_is_latched_not_silent():
    env = make_gpu(1, False, 33, 2, auto_reset=False, seed=SEED)
    env.reset()
    env.rollout(65000, materialise_obs=False)
    env.check_errors()                                       # 65 000 steps: still representable
    assert int(env.export_state()["scalars"][:, 3].min()) == 65000
    env.rollout(600, materialise_obs=False)
    with pytest.raises(OverflowError):
        env.check_errors()
    assert int(env.export_state()["scalars"][:, 3].max()) == 65535      # saturated, not wrapped
