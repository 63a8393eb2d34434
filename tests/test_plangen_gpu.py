"""On-device plan generators (dmp_plans_generate / dmp_plans_from_state, through the C ABI) against the oracle
(oracle/plangen.py, pinned to cv2 and to the reference generators) and the committed golden vectors."""
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import plangen as G
from oracle_batch import OracleBatch, philox_rollout
from test_plangen_oracle import enumerate_triples, pack13

pytestmark = pytest.mark.gpu
SEED = 0x534E4143


@pytest.fixture(scope="module")
def gold():
    z = np.load(os.path.join(GOLDEN, "plangen_golden.npz"))
    return {k: z[k] for k in z.files}


def gen(dim, n, pc=0, **kw):
    from snac_b200.vecenv import generate_plans
    return generate_plans(dim, n, pc, **kw)


def words13(table):
    return table.cpu().numpy().view(np.uint32).reshape(len(table), 16)[:, :13]


def raw_generate(dim, pc, draws, max_attempts=1):
    """dmp_plans_generate without the Python wrapper's error check (rejected single attempts are expected here)."""
    import ctypes as C
    from snac_b200 import _lib as L
    n = len(draws)
    d = torch.as_tensor(np.ascontiguousarray(draws, np.int32), device="cuda")
    row = 64 if dim == 2 else 400
    table = torch.zeros((n, row), dtype=torch.uint8, device="cuda")
    totals = torch.zeros(n, dtype=torch.int32, device="cuda")
    att = torch.zeros(n, dtype=torch.int32, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    rc = L.lib.dmp_plans_generate(dim, pc, C.c_uint64(0), 0, n, d.data_ptr(), max_attempts, table.data_ptr(),
                                  totals.data_ptr(), att.data_ptr(), err.data_ptr(),
                                  torch.cuda.current_stream().cuda_stream)
    assert rc == L.OK
    torch.cuda.synchronize()
    return table, totals, att, int(err.item())


def test_sampled_triples_match_cv2(gold):
    sx, sy = gold["sample_x"].astype(np.int32), gold["sample_y"].astype(np.int32)
    draws = np.concatenate([sx, sy], 1)
    for pc, key in ((1, "sample_sparse"), (0, "sample_dense")):
        table, totals, att, _ = raw_generate(2, pc, draws)
        assert np.array_equal(words13(table), gold[key])
        bits = np.unpackbits(gold[key].view(np.uint8), axis=1, bitorder="little").sum(1)
        assert np.array_equal(totals.cpu().numpy(), np.maximum(bits, 30))
        t3, tot3, _, _ = raw_generate(3, pc, draws)
        m = np.unpackbits(gold[key].view(np.uint8), axis=1, bitorder="little")[:, :400]
        assert np.array_equal(t3.cpu().numpy(), m * 6)
        assert np.array_equal(tot3.cpu().numpy(), bits * 6)


def test_all_vertex_triples_digest(gold):
    """Every one of the 10 746 800 vertex triples of the 20x20 grid: SHA-256 of the CUDA generator's mask table equals
    the digest of the table that was compared with cv2.polylines / cv2.fillPoly pixel by pixel."""
    xs, ys = enumerate_triples()
    draws = np.concatenate([xs, ys], 1).astype(np.int32)
    for pc, key in ((1, "sha256_sparse"), (0, "sha256_dense")):
        h = hashlib.sha256()
        step = 1 << 21
        for lo in range(0, len(draws), step):
            table, _, _, _ = raw_generate(2, pc, draws[lo:lo + step])
            h.update(np.ascontiguousarray(words13(table)).tobytes())
        assert h.hexdigest() == str(gold[key])


@pytest.mark.parametrize("dens,pc", [("dense", 0), ("sparse", 1)])
def test_reference_create_plan_streams(gold, dens, pc):
    """The unmodified reference's create_plan() for 300 numpy seeds: same retry loop, same plan, same area."""
    plans = np.unpackbits(gold["ref2d_%s_plans" % dens])[:300 * 400].reshape(300, 20, 20)
    verts, natt, areas = gold["ref2d_%s_verts" % dens], gold["ref2d_%s_attempts" % dens], gold["ref2d_%s_area" % dens]
    table, totals, att = gen(2, 300, pc, draws=verts)
    assert np.array_equal(words13(table), pack13(plans))
    assert np.array_equal(att.cpu().numpy(), natt)
    assert np.array_equal(totals.cpu().numpy(), np.maximum(areas, 30).astype(np.int32))


def test_reference_sinusoids(gold):
    table, totals, aux = gen(1, len(gold["sin_params"]), draws=gold["sin_params"])
    t = table.cpu().numpy()
    assert np.array_equal(t[:, :30], gold["sin_plans"])
    assert not t[:, 30:].any()
    assert np.array_equal(totals.cpu().numpy(), gold["sin_plans"].astype(np.int64).sum(1))
    assert np.array_equal(aux.cpu().numpy(), gold["sin_params"])


@pytest.mark.parametrize("pc", [0, 1])
def test_philox_triangles_match_oracle(pc):
    ids = np.arange(5000, 5000 + 1500)
    masks, areas, att = G.generate_2d(SEED, ids, pc, max_attempts=256)
    table, totals, natt = gen(2, len(ids), pc, seed=SEED, first_id=5000)
    assert np.array_equal(words13(table), pack13(masks))
    assert np.array_equal(natt.cpu().numpy(), att)
    assert np.array_equal(totals.cpu().numpy(), np.maximum(areas, 30))
    # plan ids, not table positions, key the stream: a shifted window reproduces the overlap
    t2, _, _ = gen(2, 100, pc, seed=SEED, first_id=5700)
    assert torch.equal(t2, table[700:800])


def test_philox_sinusoids_match_oracle():
    ids = np.arange(123, 123 + 4096)
    k1, k2, ph = G.philox_sin_params(SEED, ids)
    table, totals, aux = gen(1, len(ids), seed=SEED, first_id=123)
    a = aux.cpu().numpy()
    assert np.array_equal(a[:, 0], k1) and np.array_equal(a[:, 1], k2.astype(np.float64)) and np.array_equal(a[:, 2], ph)
    ref = np.stack([G.plan_1d_sin(k1[i], int(k2[i]), ph[i]) for i in range(len(ids))])
    assert np.array_equal(table.cpu().numpy()[:, :30].astype(np.float64), ref)
    assert np.array_equal(totals.cpu().numpy(), ref.sum(1).astype(np.int64))


def test_errors():
    with pytest.raises(ValueError):
        gen(2, 4, 2)
    bad = np.zeros((4, 1, 6), np.int32)
    bad[2, 0, 1] = 20
    with pytest.raises(ValueError):
        gen(2, 4, 0, draws=bad)
    with pytest.raises(ValueError):                      # a degenerate triangle is rejected and there is no second draw
        gen(2, 4, 0, draws=np.zeros((4, 1, 6), np.int32))


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_generated_plan_env_matches_oracle(dim):
    """A dynamic env whose plan table comes from the on-device generator steps exactly like oracle envs built on
    the same plans in the reference's array format."""
    from snac_b200.vecenv import BatchedDMPEnv
    n, K = 96, 400 if dim != 3 else 150
    env = BatchedDMPEnv(dim, dynamic=True, plans="generate", n_plans=37, plan_choose=0, num_envs=n, auto_reset=True,
                        obs_dtype=torch.float64, seed=SEED, plan_id_base=11)
    plans = env.plans_dense()
    ob = OracleBatch(dim, True, n, 0, plans)
    from oracle import philox
    p0 = philox.reset_draw(SEED, np.arange(n), 0, 37)
    o = env.reset()
    assert np.array_equal(o.cpu().numpy(), ob.reset(p0))
    from oracle import dmp_oracle as O
    r_obs, r_rew, r_done, _, _ = philox_rollout(ob, K, SEED, 0, 0, O.SPEC[dim]["actions"])
    obs, rew, done = env.rollout(K)
    assert np.array_equal(obs.cpu().numpy(), r_obs)
    assert np.array_equal(rew.cpu().numpy(), r_rew)
    assert np.array_equal(done.cpu().numpy(), r_done)
    assert np.array_equal(env.iou().cpu().numpy(), ob.iou(), equal_nan=True)
    env.check_errors()


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_hindsight_relabelling(dim):
    """env_hindsight.plan = what the agent built (script/DRQN_hindsight/1d/DRQN_hindsight_1D_static.py:242-245):
    replaying the same actions against the relabelled plan gives the oracle's rewards for that plan."""
    from snac_b200.vecenv import BatchedDMPEnv
    from oracle import dmp_oracle as O
    n, K = 64, 60
    A = O.SPEC[dim]["actions"]
    rng = np.random.RandomState(3)
    acts = rng.randint(A, size=(K, n)).astype(np.uint8)
    if dim != 3:
        acts[rng.rand(K, n) < 0.4] = A - 1               # plenty of drops
    sizes = rng.randint(1, 4, size=(K, n)).astype(np.uint8)
    env = BatchedDMPEnv(dim, plan_choose=0, num_envs=n, obs_dtype=torch.float64)
    env.reset()
    env.rollout(K, actions=acts, step_sizes=sizes)
    table, totals = env.hindsight_plans()
    st = env.export_state()
    grid = st["grid"].cpu().numpy()
    interior = grid[:, 0, 2:32] if dim == 1 else grid[:, 3:23, 3:23]
    t = table.cpu().numpy()
    if dim == 1:
        assert np.array_equal(t[:, :30], interior)
    elif dim == 2:
        assert np.array_equal(t.view(np.uint32).reshape(n, 16)[:, :13], pack13(interior))
    else:
        assert np.array_equal(t.reshape(n, 20, 20), interior)
    assert np.array_equal(totals.cpu().numpy(), interior.reshape(n, -1).sum(1))
    # replay against the achieved structure (budget = the plan's own total, as reset() would compute it)
    if dim == 2:
        totals = torch.clamp(totals, min=30)
    keep = totals.cpu().numpy() > 0
    hs = BatchedDMPEnv(dim, dynamic=True, plans=(table, totals), num_envs=n, obs_dtype=torch.float64,
                       total_step=env.total_step, dynamic_rules=False)
    hs.reset(plan_idx=np.arange(n, dtype=np.int32))
    _, rew, done = hs.rollout(K, actions=acts, step_sizes=sizes)
    dense = hs.plans_dense()
    rew, done = rew.cpu().numpy(), done.cpu().numpy()
    for i in np.nonzero(keep)[0][:24]:
        e = O.make_env(dim, False, plan_choose=0)
        e.reset(0)
        e.plan = dense[i].copy() if dim != 1 else dense[i].copy()
        e.plans = [e.plan]
        e.total_brick = float(totals[i].item())
        for k in range(K):
            _, r, d = e.step(int(acts[k, i]), int(sizes[k, i]))
            assert r == rew[k, i] and d == done[k, i], (dim, i, k, r, rew[k, i], d, done[k, i])
