"""Parity of the CUDA path (through the C ABI) with the oracle and the reference's golden traces.
Bit-exact: grids, positions, counters, observations, done flags, rewards (small integers in f32).
IoU / normalised counters: IEEE fp64 division on both sides -> compared exactly, with a 1e-6
relative fallback stated in BASELINE.json's north_star."""
import numpy as np
import pytest
import torch

from conftest import TRACE_NAMES, load_plans, load_trace, trace_env_spec
from oracle import dmp_oracle as O
from oracle import philox
from oracle_batch import OracleBatch, philox_rollout

pytestmark = pytest.mark.gpu

SEED = 0x534E4143


def make_gpu(dim, dynamic, n, plan_choose=0, plans=None, **kw):
    from snac_b200.vecenv import BatchedDMPEnv
    return BatchedDMPEnv(dim, dynamic=dynamic, plan_choose=plan_choose, plans=plans, num_envs=n, **kw)


# ------------------------------------------------------------------------------------------------
# 1. golden traces of the unmodified reference, replayed step by step (caller-side reset, N = 1)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", TRACE_NAMES)
def test_reference_trace_step_mode(name):
    tr = load_trace(name)
    dim, dynamic, plan_choose, plans = trace_env_spec(tr["meta"])
    env = make_gpu(dim, dynamic, 1, plan_choose, plans, obs_dtype=torch.float64)
    resets = tr["reset_plan_idx"]
    ri = 0
    o = env.reset(plan_idx=[resets[ri]])
    assert np.array_equal(o.cpu().numpy()[0], tr["reset_obs"][ri].astype(np.float64))
    T = min(len(tr["actions"]), 1200)
    ep = 0
    for t in range(T):
        o, r, d = env.step([tr["actions"][t]], [tr["step_sizes"][t]])
        o, r, d = o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy()
        assert o.dtype == np.float64
        assert np.array_equal(o[0], tr["obs"][t].astype(np.float64)), (t, o[0], tr["obs"][t])
        assert r[0] == tr["reward"][t], (t, r, tr["reward"][t])
        assert bool(d[0]) == bool(tr["done"][t]), t
        if t % 97 == 0 or d[0]:
            iou = env.iou().cpu().numpy()[0]
            assert iou == tr["iou"][t] or (np.isnan(iou) and np.isnan(tr["iou"][t])), (t, iou, tr["iou"][t])
            st = env.export_state()
            sc = st["scalars"].cpu().numpy()[0]
            assert list(sc[:2]) == list(tr["pos"][t]) and sc[2] == tr["count_brick"][t] and sc[3] == tr["count_step"][t]
            assert sc[5] == tr["total_brick"][t]
        if d[0]:
            g = env.export_state()["grid"].cpu().numpy()[0]
            assert np.array_equal(g.reshape(tr["final_grids"][ep].shape), tr["final_grids"][ep])
            ep += 1
            ri += 1
            o = env.reset(plan_idx=[resets[ri]])
            assert np.array_equal(o.cpu().numpy()[0], tr["reset_obs"][ri].astype(np.float64))
    env.check_errors()


# ------------------------------------------------------------------------------------------------
# 2. the same traces in ONE rollout launch with in-kernel auto-reset (K = T), all obs dtypes
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", TRACE_NAMES)
@pytest.mark.parametrize("obs_dtype", [torch.float32, torch.int16, torch.float64])
def test_reference_trace_rollout_mode(name, obs_dtype):
    tr = load_trace(name)
    dim, dynamic, plan_choose, plans = trace_env_spec(tr["meta"])
    env = make_gpu(dim, dynamic, 1, plan_choose, plans, obs_dtype=obs_dtype, auto_reset=True)
    T = len(tr["actions"])
    resets = tr["reset_plan_idx"]
    nxt = np.zeros((T, 1), np.int32)
    nd = np.cumsum(tr["done"])                      # number of dones up to and including t
    for t in range(T):
        if tr["done"][t]:
            nxt[t, 0] = resets[nd[t]]
    env.reset(plan_idx=[resets[0]])
    obs, rew, done = env.rollout(T, actions=tr["actions"][:, None], step_sizes=tr["step_sizes"][:, None], next_plan=nxt)
    obs, rew, done = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()
    assert np.array_equal(obs[:, 0, :].astype(np.float64), tr["obs"].astype(np.float64))
    assert np.array_equal(rew[:, 0].astype(np.float64), tr["reward"])
    assert np.array_equal(done[:, 0], tr["done"])
    # per-env episode statistics == what the reference trace implies
    cnt, ln, ret, iou = [x.cpu().numpy() for x in env.episode_stats()]
    n_ep = int(tr["done"].sum())
    assert cnt[0] == n_ep
    ends = np.nonzero(tr["done"])[0]
    assert ln[0] == int(tr["count_step"][ends].sum())
    starts = np.concatenate([[0], ends[:-1] + 1])
    want_ret = sum(tr["reward"][s:e + 1].sum() for s, e in zip(starts, ends))
    assert ret[0] == want_ret
    want_iou = 0.0
    for e in ends:
        want_iou += tr["iou"][e]
    if np.isnan(want_iou):
        assert np.isnan(iou[0])
    else:
        assert iou[0] == want_iou
    g = env.export_state()["grid"].cpu().numpy()[0]
    assert np.array_equal(g.reshape(tr["final_grids"][-1].shape), tr["final_grids"][-1])
    stats = env.stats().cpu().numpy()
    assert stats[2] == n_ep and stats[3] == ln[0] and stats[0] == want_ret
    env.check_errors()


# ------------------------------------------------------------------------------------------------
# 3. vector runs with in-kernel Philox draws vs N independent oracle envs
# ------------------------------------------------------------------------------------------------
CASES = [
    # dim, dynamic, plan_choose, density, ref3d, n, K
    (1, False, 2, None, False, 200, 800),
    (1, True, 0, "dense", False, 200, 800),
    (2, False, 0, None, False, 200, 700),
    (2, False, 1, None, False, 131, 400),
    (2, True, 0, "dense", False, 200, 700),
    (2, True, 0, "sparse", False, 97, 400),
    (3, False, 0, None, False, 150, 300),
    (3, False, 1, None, True, 100, 1400),
    (3, True, 0, "dense", False, 150, 300),
    (3, True, 0, "sparse", True, 100, 1100),
]


@pytest.mark.parametrize("dim,dynamic,plan_choose,density,ref3d,n,K", CASES)
def test_philox_vector_rollout_matches_oracle(dim, dynamic, plan_choose, density, ref3d, n, K):
    plans = load_plans(dim, density, "train") if dynamic else None
    env_base = 12345
    env = make_gpu(dim, dynamic, n, plan_choose, plans, auto_reset=True, env_base=env_base, seed=SEED,
                   action_dist="ref3d" if ref3d else "uniform", normalise=dynamic, obs_dtype=torch.float64)
    ob = OracleBatch(dim, dynamic, n, plan_choose, plans)
    # initial reset: Philox plan draw on the device, replayed on the host
    p0 = philox.reset_draw(SEED, np.arange(env_base, env_base + n), 0, ob.n_plans) if dynamic else None
    o_gpu = env.reset().cpu().numpy()
    o_ref = ob.reset(p0)
    assert np.array_equal(o_gpu, o_ref)
    # K steps: half in one rollout launch, half step by step (exercises the host-side t counter)
    K1 = K // 2
    A = O.SPEC[dim]["actions"]
    r_obs, r_rew, r_done, _, _ = philox_rollout(ob, K, SEED, env_base, 0, A, ref3d, normalise=dynamic)
    obs, rew, done = env.rollout(K1)
    assert np.array_equal(obs.cpu().numpy(), r_obs[:K1])
    assert np.array_equal(rew.cpu().numpy(), r_rew[:K1])
    assert np.array_equal(done.cpu().numpy(), r_done[:K1])
    for k in range(K1, K):
        o, r, d = env.step(None)
        if k % 50 == 0 or k == K - 1:
            assert np.array_equal(o.cpu().numpy(), r_obs[k]), k
            assert np.array_equal(r.cpu().numpy(), r_rew[k]), k
            assert np.array_equal(d.cpu().numpy(), r_done[k]), k
    st = env.export_state()
    g_ref, sc_ref = ob.export()
    assert np.array_equal(st["grid"].cpu().numpy().reshape(g_ref.shape), g_ref)
    assert np.array_equal(st["scalars"].cpu().numpy()[:, :6], sc_ref[:, :6])
    assert np.array_equal(st["ret"].cpu().numpy().astype(np.float64), ob.ret)
    cnt, ln, ret, iou = [x.cpu().numpy() for x in env.episode_stats()]
    assert np.array_equal(cnt, ob.ep_cnt) and np.array_equal(ln, ob.ep_len)
    assert np.array_equal(ret, ob.ep_ret)
    assert np.array_equal(iou, ob.ep_iou, equal_nan=True)
    assert ob.ep_cnt.sum() > 0
    cur = env.iou().cpu().numpy()
    assert np.array_equal(cur, ob.iou(), equal_nan=True)
    stats = env.stats().cpu().numpy()
    assert stats[2] == ob.ep_cnt.sum() and stats[3] == ob.ep_len.sum() and stats[0] == ob.ep_ret.sum()
    if not np.isnan(ob.ep_iou.sum()):
        assert abs(stats[1] - ob.ep_iou.sum()) <= 1e-9 * max(1.0, abs(ob.ep_iou.sum()))
    env.check_errors()


def test_shard_independence_2d():
    """Philox streams are keyed by the GLOBAL env id: two shards == one big vector env (8(e))."""
    n, K = 256, 300
    full = make_gpu(2, False, n, 0, auto_reset=True, env_base=0)
    a = make_gpu(2, False, n // 2, 0, auto_reset=True, env_base=0)
    b = make_gpu(2, False, n // 2, 0, auto_reset=True, env_base=n // 2)
    for e in (full, a, b):
        e.reset()
    of, rf, df = full.rollout(K)
    oa, ra, da = a.rollout(K)
    ob_, rb, db = b.rollout(K)
    assert torch.equal(of, torch.cat([oa, ob_], dim=1))
    assert torch.equal(rf, torch.cat([ra, rb], dim=1)) and torch.equal(df, torch.cat([da, db], dim=1))
    sf = full.stats().clone()
    assert torch.equal(sf[[0, 2, 3]], (a.stats() + b.stats())[[0, 2, 3]])


def test_static_plans_generated_on_device_match_oracle():
    for dim, choices in ((1, (0, 1, 2)), (2, (0, 1)), (3, (0, 1))):
        for pc in choices:
            env = make_gpu(dim, False, 1, pc)
            want = O.static_plan(dim, pc)
            got = env.plans_dense()[0]
            assert np.array_equal(got, want), (dim, pc)
            assert env.plan_totals().cpu().numpy()[0] == O.total_brick_of(dim, want, False)
    with pytest.raises(ValueError):
        make_gpu(2, False, 1, 2)
    with pytest.raises(ValueError):
        make_gpu(1, False, 1, 3)


def test_dataset_packing_matches_reference_budgets():
    for dim, dens in ((1, "dense"), (2, "dense"), (2, "sparse"), (3, "dense"), (3, "sparse")):
        plans = load_plans(dim, dens, "train")
        env = make_gpu(dim, True, 1, plans=plans)
        assert np.array_equal(env.plans_dense(), plans)
        want = [O.total_brick_of(dim, p, True) for p in plans]
        assert np.array_equal(env.plan_totals().cpu().numpy(), np.asarray(want))


def test_out_of_range_action_latches_error():
    env = make_gpu(2, False, 4, 0)
    env.reset()
    env.step([0, 1, 5, 4])
    with pytest.raises(UnboundLocalError):
        env.check_errors()


def test_step_after_done_keeps_mutating_like_reference():
    """No auto-reset: stepping past done is legal (multiprocess.py never resets, quirk Q3)."""
    env = make_gpu(1, False, 1, 2, obs_dtype=torch.float64)
    orc = O.make_env(1, False, plan_choose=2)
    env.reset()
    orc.reset()
    rng = np.random.RandomState(3)
    for t in range(900):
        a, s = int(rng.randint(3)), int(rng.randint(1, 4))
        o, r, d = env.step([a], [s])
        oo, orr, od = orc.step(a, s)
        if t % 25 == 0 or t > 740:
            assert np.array_equal(o.cpu().numpy(), oo) and r.item() == orr and bool(d.item()) == od, t


# ------------------------------------------------------------------------------------------------
# 4. the five standalone stage kernels (a)-(e) == the fused step, and == the oracle
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,dynamic,density,ref3d,n,K", [
    (1, False, None, False, 70, 800), (1, True, "dense", False, 70, 800),
    (2, False, None, False, 70, 700), (2, True, "sparse", False, 70, 400),
    (3, False, None, True, 70, 400), (3, True, "dense", False, 70, 200), (3, True, "sparse", True, 40, 1100)])
def test_stage_kernels_equal_fused_step_and_oracle(dim, dynamic, density, ref3d, n, K):
    plans = load_plans(dim, density, "val") if dynamic else None
    kw = dict(auto_reset=True, env_base=99, seed=SEED, action_dist="ref3d" if ref3d else "uniform",
              normalise=dynamic, obs_dtype=torch.float64)
    staged = make_gpu(dim, dynamic, n, 0, plans, **kw)
    fused = make_gpu(dim, dynamic, n, 0, plans, **kw)
    ob = OracleBatch(dim, dynamic, n, 0, plans)
    p0 = philox.reset_draw(SEED, np.arange(99, 99 + n), 0, ob.n_plans) if dynamic else None
    assert torch.equal(staged.reset(), fused.reset())
    ob.reset(p0)
    r_obs, r_rew, r_done, _, _ = philox_rollout(ob, K, SEED, 99, 0, O.SPEC[dim]["actions"], ref3d, normalise=dynamic)
    for k in range(K):
        o1, r1, d1 = staged.step_staged(None)
        o2, r2, d2 = fused.step(None)
        assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2), k
        if k % 40 == 0 or k == K - 1:
            assert np.array_equal(o1.cpu().numpy(), r_obs[k]) and np.array_equal(d1.cpu().numpy(), r_done[k]), k
    a, b = staged.export_state(), fused.export_state()
    assert torch.equal(a["grid"], b["grid"]) and torch.equal(a["scalars"], b["scalars"]) and torch.equal(a["ret"], b["ret"])
    for x, y in zip(staged.episode_stats(), fused.episode_stats()):
        assert torch.equal(x, y)
    assert np.array_equal(staged.episode_stats()[0].cpu().numpy(), ob.ep_cnt)
    assert np.array_equal(staged.episode_stats()[3].cpu().numpy(), ob.ep_iou, equal_nan=True)


def test_get_set_state_and_functional_transition():
    """state snapshots + transition(state, action) (the MCTS variants' surface)."""
    n = 64
    env = make_gpu(3, False, n, 0, auto_reset=True, seed=SEED)
    env.reset()
    env.rollout(15)
    s0 = env.get_state()
    o1, r1, d1 = [x.clone() for x in env.rollout(10)]
    s1 = env.get_state()
    env.rollout(7)                                   # wander off, then come back
    env.set_state(s0)
    o2, r2, d2 = env.rollout(10)
    assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2)
    assert torch.equal(env.get_state()["cells"], s1["cells"]) and torch.equal(env.get_state()["aux"], s1["aux"])
    # functional transition from s0 with explicit actions == stepping an env that is in s0
    acts = torch.randint(0, 8, (n,), dtype=torch.uint8, device=env.device)
    sizes = torch.randint(1, 4, (n,), dtype=torch.uint8, device=env.device)
    env.set_state(s1)                                # unrelated current state
    ns, o, r, d = env.transition(s0, acts, sizes)
    env.set_state(s0)
    o_, r_, d_ = env.step(acts, sizes)
    assert torch.equal(o, o_) and torch.equal(r, r_) and torch.equal(d, d_)
    assert torch.equal(ns["cells"], env.get_state()["cells"])


# ------------------------------------------------------------------------------------------------
# 7. the two 3D kernels (rollout: whole nibble maps cached in shared memory; single step: only the rows a step can look
#    at) and their tuning variants are interchangeable and equal the oracle; the "tall env" path (a height >= 15
#    somewhere: the env runs from its wide map in HBM) is exact
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dynamic", [False, True])
@pytest.mark.parametrize("K", [1, 37])
def test_3d_kernels_agree(dynamic, K):
    plans = load_plans(3, "dense", "train") if dynamic else None
    n = 333                                                   # ragged last warp
    outs = {}
    # "d": the default dispatch (rollout kernel for K > 1, single-step kernel for K = 1); "c": the rollout kernel for
    # every K; "l": load/store copy-out of the observation tile instead of the bulk async copy; "p": no programmatic
    # dependent launch
    variants = {"d": (), "c": ("rollout_k1",), "l": ("tile_ldst",), "p": ("no_pdl",), "cl": ("rollout_k1", "tile_ldst")}
    reps = 120 // K + 1
    for kind, tuning in variants.items():
        env = make_gpu(3, dynamic, n, 0, plans, auto_reset=True, env_base=99, seed=SEED, normalise=dynamic,
                       obs_dtype=torch.float64 if dynamic else torch.float32, tuning=tuning)
        env.reset()
        res = [[x.clone() for x in env.rollout(K)] for _ in range(reps)]
        torch.cuda.synchronize()
        st = env.get_state()
        # heights as exported, the nibble maps behind the wide maps (the wide map of an env that is not tall is scratch),
        # scalar state incl. the tall flags
        outs[kind] = (res, torch.cat([env.export_state()["grid"].reshape(-1), st["cells"][n * 800:].to(torch.int32)]),
                      st["aux"].clone(), [x.clone() for x in env.episode_stats()])
        env.check_errors()
    for kind in variants:
        if kind == "d":
            continue
        for a, b in zip(outs["d"][0], outs[kind][0]):
            for x, y in zip(a, b):
                assert torch.equal(x, y), kind
        assert torch.equal(outs["d"][1], outs[kind][1]) and torch.equal(outs["d"][2], outs[kind][2]), kind
        for x, y in zip(outs["d"][3], outs[kind][3]):
            assert torch.equal(x, y), kind
    ob = OracleBatch(3, dynamic, n, 0, plans)
    ob.reset(philox.reset_draw(SEED, np.arange(99, 99 + n), 0, ob.n_plans) if dynamic else None)
    r_obs, r_rew, r_done, _, _ = philox_rollout(ob, reps * K, SEED, 99, 0, 8, normalise=dynamic)
    obs = torch.cat([r[0] for r in outs["d"][0]]).cpu().numpy().astype(np.float64)
    assert np.array_equal(obs, r_obs)
    assert np.array_equal(torch.cat([r[1] for r in outs["d"][0]]).cpu().numpy(), r_rew)
    assert np.array_equal(torch.cat([r[2] for r in outs["d"][0]]).cpu().numpy(), r_done)


@pytest.mark.parametrize("kind", ["c", "d1", "c1", "cr", "d1r", "cb", "d1b"])
@pytest.mark.parametrize("dynamic", [False, True])
def test_3d_tall_columns_match_oracle(kind, dynamic):
    """Heights around the nibble paths' threshold (13..17: an env turns "tall" at 15), around the record bytes' (252..256)
    and far beyond (300, 40000) next to the agent: builds on top of them -- including the brick that makes an env tall
    inside a launch --, walks blocked by them and windows over them must equal the oracle's."""
    plans = load_plans(3, "dense", "train") if dynamic else None
    n, K = 70, 48
    rng = np.random.RandomState(11)
    # "c": one rollout launch; "d1": step by step through the single-step kernel; "c1": step by step through the rollout
    # kernel; a trailing "r": packed step records (window bytes saturate at 255 and the record says so)
    # a trailing "b": bit records (window codes saturate at 15 = "height 14 or more" and the record says so)
    records, bits = kind.endswith("r"), kind.endswith("b")
    env = make_gpu(3, dynamic, n, 0, plans, auto_reset=False, seed=SEED,
                   obs_dtype="record" if records else ("bits" if bits else torch.float32),
                   tuning=("rollout_k1",) if kind.startswith("c1") else ())
    ob = OracleBatch(3, dynamic, n, 0, plans)
    p0 = rng.randint(ob.n_plans, size=n).astype(np.int32) if dynamic else None
    env.reset(plan_idx=p0)
    ob.reset(p0)
    g, sc = ob.export()
    tall_values = [12, 13, 13, 14, 14, 14, 15, 16, 17, 127, 252, 253, 254, 255, 256, 300, 40000]
    for i, e in enumerate(ob.envs):
        r, c = int(rng.randint(5, 21)), int(rng.randint(5, 21))
        e.pos = [r, c]
        if hasattr(e, "position_memory"):
            e.position_memory = [[r, c]]
        for _ in range(rng.randint(0, 4)):                    # some envs stay short
            dr, dc = [(0, 1), (0, -1), (1, 0), (-1, 0), (1, 1), (2, 0)][rng.randint(6)]
            e.grid[r + dr, c + dc] = tall_values[rng.randint(len(tall_values))]
    g, sc = ob.export()
    env.import_state(grid=g, scalars=sc)
    acts = rng.choice(8, size=(K, n), p=[.08, .08, .08, .08, .17, .17, .17, .17]).astype(np.uint8)
    sizes = rng.randint(1, 4, size=(K, n)).astype(np.uint8)
    r_obs, r_rew, r_done, _, _ = philox_rollout(ob, K, SEED, 0, 0, 8, auto_reset=False, actions=acts, step_sizes=sizes)

    ta, ts = torch.as_tensor(acts, device=env.device), torch.as_tensor(sizes, device=env.device)
    if kind[:2] in ("d1", "c1"):                              # step by step
        outs = [[x.clone() for x in env.step(ta[k], ts[k])] for k in range(K)]
        obs, rew, done = [torch.stack([o[i] for o in outs]) for i in range(3)]
    else:
        obs, rew, done = env.rollout(K, actions=ta, step_sizes=ts)
    torch.cuda.synchronize()
    if records:
        from snac_b200.vecenv import unpack_records
        o, r, d, sat = unpack_records(obs, 3)
        # exact wherever the window holds nothing above 253; saturated bytes read 254 and the record is flagged
        assert np.array_equal(o[..., :49], np.minimum(r_obs[..., :49], 254)) and np.array_equal(o[..., 49:], r_obs[..., 49:])
        assert (sat | (r_obs[..., :49].max(axis=-1) < 254)).all() and sat.any()
        assert np.array_equal(r, r_rew) and np.array_equal(d, r_done)
    elif bits:
        from snac_b200.vecenv import unpack_bits, unpack_records_device
        o, r, d, sat = unpack_bits(obs, 3)
        assert np.array_equal(o[..., :49], np.minimum(r_obs[..., :49], 14)) and np.array_equal(o[..., 49:], r_obs[..., 49:])
        assert np.array_equal(sat, r_obs[..., :49].max(axis=-1) >= 14) and sat.any() and not sat.all()
        assert np.array_equal(r, r_rew) and np.array_equal(d, r_done)
        od, rd, dd, sd = unpack_records_device(obs, 3, "bits", torch.float64)      # the device-side expansion agrees
        assert np.array_equal(od.cpu().numpy(), o) and np.array_equal(rd.cpu().numpy(), r)
        assert np.array_equal(dd.cpu().numpy(), d) and np.array_equal(sd.cpu().numpy(), sat)
    else:
        assert np.array_equal(obs.cpu().numpy().astype(np.float64), r_obs)
    assert np.array_equal(rew.cpu().numpy(), r_rew)
    assert np.array_equal(done.cpu().numpy(), r_done)
    g_ref, sc_ref = ob.export()
    st = env.export_state()
    assert np.array_equal(st["grid"].cpu().numpy().reshape(g_ref.shape), g_ref)
    assert np.array_equal(st["scalars"].cpu().numpy()[:, :4], sc_ref[:, :4])
    # the nibble maps behind the wide maps (208 B per env, low nibble = even cell) are min(height, 15) whichever kernel
    # wrote them, and exactly the envs holding a height >= 15 carry the tall flag (bit 7 of aux.x)
    raw = env.get_state()["cells"].cpu().numpy()
    packed = raw[n * 800:].reshape(n, 208)
    assert not packed[:, 200:].any()
    shadow = np.stack([packed[:, :200] & 15, packed[:, :200] >> 4], axis=-1).reshape(n, 20, 20)
    interior = g_ref.reshape(n, 26, 26)[:, 3:23, 3:23]
    assert np.array_equal(shadow, np.minimum(interior, 15).astype(np.uint8))
    flags = (env.get_state()["aux"].cpu().numpy().view(np.uint32).reshape(n, 4)[:, 0] >> 7) & 1
    assert np.array_equal(flags.astype(bool), interior.reshape(n, -1).max(axis=1) >= 15)
    assert g_ref.max() > 40000 or g_ref.max() >= 254
    # some env crossed the threshold INSIDE the run (started at <= 14 everywhere, ended tall)
    assert ((g.reshape(n, -1).max(axis=1) <= 14) & flags.astype(bool)).any()
    assert np.array_equal(env.iou().cpu().numpy(), ob.iou(), equal_nan=True)


# ------------------------------------------------------------------------------------------------
# 8. 2D single steps leave through a bulk (TMA) copy of the warp tile: equal to the oracle, to the load/store copy-out
#    (tuning switch "tile_ldst") and independent of the alignment of the caller's observation buffer
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dynamic", [False, True])
def test_2d_bulk_copy_out_step_mode(dynamic):
    plans = load_plans(2, "sparse", "train") if dynamic else None
    n, K = 1000 + 13, 90                                      # ragged last warp
    ob = OracleBatch(2, dynamic, n, 0, plans)
    envs = {}
    for kind in ("float", "unaligned", "tma"):
        envs[kind] = make_gpu(2, dynamic, n, 0, plans, auto_reset=True, env_base=3, seed=SEED, obs_dtype=torch.float32,
                              total_step=35, tuning=("tile_ldst",) if kind == "float" else ())
    for e in ob.envs:
        e.total_step = 35
    p0 = philox.reset_draw(SEED, np.arange(3, 3 + n), 0, ob.n_plans) if dynamic else None
    ob.reset(p0)
    for env in envs.values():
        env.reset()
    r_obs, r_rew, r_done, _, _ = philox_rollout(ob, K, SEED, 3, 0, 5)
    raw = torch.zeros(n * 51 + 1, dtype=torch.float32, device="cuda")
    odd = raw[1:].view(1, n, 51)                              # 4 B aligned only
    for k in range(K):
        o1, r1, d1 = envs["tma"].step(None)
        o2, r2, d2 = envs["float"].step(None)
        e3 = envs["unaligned"]
        o3, r3, d3 = e3.rollout(1, out=(odd, e3._reward[None], e3._done[None]))           # default mode, 4 B aligned buffer
        assert np.array_equal(o1.cpu().numpy().astype(np.float64), r_obs[k]), k
        assert torch.equal(o1, o2) and torch.equal(o1, o3[0]), k
        assert np.array_equal(r1.cpu().numpy(), r_rew[k]) and np.array_equal(d1.cpu().numpy(), r_done[k])
        assert torch.equal(r1, r2) and torch.equal(d1, d2) and torch.equal(r1, r3[0]) and torch.equal(d1, d3[0])
    assert ob.ep_cnt.sum() > 0
    for env in envs.values():
        env.check_errors()


# ------------------------------------------------------------------------------------------------
# 9. 1D rollouts: ragged last warp, both block shapes (32 / 128 envs per block), every observation dtype -- equal to the
#    oracle and independent of how a rollout is cut into launches
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dynamic,n", [(False, 32 * 37 + 20), (True, 128 * 148 * 4 + 64)])
def test_1d_rollout_shapes_and_cuts(dynamic, n):
    plans = load_plans(1, "dense", "train") if dynamic else None
    K = 23 if n > 10000 else 70
    ref = None
    if n < 10000:
        ob = OracleBatch(1, dynamic, n, 2, plans)
        for e in ob.envs:
            e.total_step = 30
        ob.reset(philox.reset_draw(SEED, np.arange(5, 5 + n), 0, ob.n_plans) if dynamic else None)
        ref = philox_rollout(ob, K, SEED, 5, 0, 3)
    outs = {}
    for dt in (torch.float32, torch.int16, torch.float64):
        # "b": one launch; "l": the same steps cut into launches of 1, 3, 7, ...
        for mode in ("b", "l"):
            env = make_gpu(1, dynamic, n, 2, plans, auto_reset=True, env_base=5, seed=SEED, obs_dtype=dt, total_step=30)
            env.reset()
            if mode == "b":
                o, r, d = env.rollout(K)
            else:
                parts, left, c = [], K, 1
                while left:
                    kk = min(left, c)
                    parts.append([x.clone() for x in env.rollout(kk)])
                    left -= kk
                    c = c * 2 + (c & 1)
                o, r, d = (torch.cat([p[i] for p in parts]) for i in range(3))
            o2, r2, d2 = env.rollout(3)                                   # K below the block length
            torch.cuda.synchronize()
            env.check_errors()
            outs[(dt, mode)] = (o, r, d, o2, r2, d2, env.export_state())
        a = outs[(dt, "b")]
        for other in ("l",):
            b = outs[(dt, other)]
            for x, y in zip(a[:6], b[:6]):
                assert torch.equal(x, y), (dt, other)
            assert torch.equal(a[6]["grid"], b[6]["grid"]) and torch.equal(a[6]["scalars"], b[6]["scalars"])
        if ref is not None:
            assert np.array_equal(a[0].cpu().numpy().astype(np.float64), ref[0]), dt
            assert np.array_equal(a[1].cpu().numpy(), ref[1]) and np.array_equal(a[2].cpu().numpy(), ref[2])


# ------------------------------------------------------------------------------------------------
# 10. the 1D kernel's FAST instantiation (launch-uniform branches folded: Philox draws, everything materialised,
#     auto-reset, raw counters, full blocks) against the generic one (tuning switch "generic") and the oracle
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dynamic,n", [(False, 32 * 64), (True, 32 * 33), (False, 128 * 600), (True, 128 * 592)])
def test_1d_fast_path_equals_generic(dynamic, n):
    plans = load_plans(1, "dense", "train") if dynamic else None
    K = 75
    outs = {}
    for dt in (torch.float32, torch.int16, torch.float64, "record"):
        for fast in ("1", "0"):
            env = make_gpu(1, dynamic, n, 1, plans, auto_reset=True, env_base=9, seed=SEED, obs_dtype=dt, total_step=40,
                           tuning=() if fast == "1" else ("generic",))
            env.reset()
            o, r, d = env.rollout(K)
            o2, r2, d2 = env.rollout(2)
            torch.cuda.synchronize()
            env.check_errors()
            outs[(dt, fast)] = (o, r, d, o2, r2, d2, env.export_state(), [x.clone() for x in env.episode_stats()])
        a, b = outs[(dt, "1")], outs[(dt, "0")]
        for x, y in zip(a[:6], b[:6]):
            assert torch.equal(x, y), dt
        assert torch.equal(a[6]["grid"], b[6]["grid"]) and torch.equal(a[6]["scalars"], b[6]["scalars"])
        for x, y in zip(a[7], b[7]):
            assert torch.equal(x, y)
        assert int(a[7][0].sum()) > 0                         # episodes did finish (auto-reset path exercised)
    if n < 4000:
        ob = OracleBatch(1, dynamic, n, 1, plans)
        for e in ob.envs:
            e.total_step = 40
        ob.reset(philox.reset_draw(SEED, np.arange(9, 9 + n), 0, ob.n_plans) if dynamic else None)
        ref = philox_rollout(ob, K, SEED, 9, 0, 3)
        a = outs[(torch.float32, "1")]
        assert np.array_equal(a[0].cpu().numpy().astype(np.float64), ref[0])
        assert np.array_equal(a[1].cpu().numpy(), ref[1]) and np.array_equal(a[2].cpu().numpy(), ref[2])


# ------------------------------------------------------------------------------------------------
# 9. DMP_F_RESET_OBS (BatchedDMPEnv(reset_obs=True)): the row written for an env that finishes in a step is the observation
#    its reset returns -- every kernel, rollout and single steps, observations and packed records
# ------------------------------------------------------------------------------------------------
RESET_OBS_CASES = [
    # dim, dynamic, density, ref3d, obs kind, normalise
    (1, False, None, False, torch.float32, False), (1, True, "dense", False, "record", False),
    (2, False, None, False, torch.float32, False), (2, True, "sparse", False, torch.float64, True),
    (2, False, None, False, "record", False),
    (3, False, None, False, torch.float32, False), (3, True, "dense", True, torch.float64, True),
    (3, True, "dense", False, "record", False),
    (2, True, "dense", False, "bits", False), (3, False, None, True, "bits", False),
]


@pytest.mark.parametrize("dim,dynamic,density,ref3d,kind,normalise", RESET_OBS_CASES)
def test_reset_observation_mode_matches_oracle(dim, dynamic, density, ref3d, kind, normalise):
    from snac_b200.vecenv import unpack_records
    plans = load_plans(dim, density, "train") if dynamic else None
    n, K, T = 131, 120, (None if dim == 3 else 17)            # ragged last warp; short 1D / 2D episodes so that resets happen
    env = make_gpu(dim, dynamic, n, 0, plans, auto_reset=True, reset_obs=True, env_base=31, seed=SEED, obs_dtype=kind,
                   normalise=normalise, total_step=T, action_dist="ref3d" if ref3d else "uniform")
    ob = OracleBatch(dim, dynamic, n, 0, plans)
    if T is not None:
        for e in ob.envs:
            e.total_step = T
    p0 = philox.reset_draw(SEED, np.arange(31, 31 + n), 0, ob.n_plans) if dynamic else None
    env.reset()
    ob.reset(p0)
    r_obs, r_rew, r_done, _, _ = philox_rollout(ob, K, SEED, 31, 0, O.SPEC[dim]["actions"], ref3d, normalise=normalise,
                                                reset_obs=True)
    assert r_done.sum() > n // 2                              # resets happened all over the batch
    K1 = K - 30
    obs, rew, done = env.rollout(K1)                          # one launch, then single steps (3D: the other kernel)
    outs = [(obs, rew, done)] + [tuple(x.clone()[None] for x in env.step(None)) for _ in range(K1, K)]
    obs, rew, done = [torch.cat([o[i] for o in outs]) for i in range(3)]
    if kind in ("record", "bits"):
        o, r, d, _ = unpack_records(obs, dim)
        assert np.array_equal(o, r_obs) and np.array_equal(r, r_rew) and np.array_equal(d, r_done)
    else:
        assert np.array_equal(obs.cpu().numpy().astype(np.float64), r_obs.astype(np.float32 if kind == torch.float32 else np.float64))
    assert np.array_equal(rew.cpu().numpy(), r_rew) and np.array_equal(done.cpu().numpy(), r_done)
    g_ref, sc_ref = ob.export()
    st = env.export_state()
    assert np.array_equal(st["grid"].cpu().numpy().reshape(g_ref.shape), g_ref)
    assert np.array_equal(st["scalars"].cpu().numpy()[:, :6], sc_ref[:, :6])
    cnt, ln, ret, iou = [x.cpu().numpy() for x in env.episode_stats()]
    assert np.array_equal(cnt, ob.ep_cnt) and np.array_equal(ln, ob.ep_len) and np.array_equal(ret, ob.ep_ret)
    assert np.array_equal(iou, ob.ep_iou, equal_nan=True)
    env.check_errors()


def test_reset_obs_needs_auto_reset():
    with pytest.raises(ValueError):
        make_gpu(2, False, 4, reset_obs=True)
