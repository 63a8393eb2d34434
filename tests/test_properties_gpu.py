"""Property tests: invariants of the CUDA path on random configurations -- sizes, seeds, action
mixes, injected or in-kernel draws -- that hold whatever the trajectory: positions stay inside the plan area, counters count,
2D cells are 0/1, IoU lies in [0, 1], rewards come from the classes' reward sets, and the observation a step returns IS the
window of the state it leaves behind (cut here, on the host, from the exported grid at the exported position)."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

pytestmark = pytest.mark.gpu

HW = {1: 2, 2: 3, 3: 3}
REWARDS = {1: {0.0, -1.0, 1.0, 10.0}, 2: {0.0, 5.0}, 3: {0.0, -1.0, 1.0, 10.0, -100.0}}


def window_of(dim, grid, pos):
    """Observation window cut from exported padded grids [n, ...] at positions [n, 2] (numpy restatement of observation_)."""
    n = len(grid)
    if dim == 1:
        g = grid.reshape(n, 34)
        return np.stack([g[i, pos[i, 0] - 2:pos[i, 0] + 3] for i in range(n)]).astype(np.float64)
    g = grid.reshape(n, 26, 26)
    return np.stack([g[i, pos[i, 0] - 3:pos[i, 0] + 4, pos[i, 1] - 3:pos[i, 1] + 4].reshape(-1) for i in range(n)]).astype(np.float64)


@settings(max_examples=int(__import__("os").environ.get("SNAC_PROPERTY_EXAMPLES", "40")), deadline=None, derandomize=__import__("os").environ.get("SNAC_PROPERTY_RANDOM", "0") != "1", suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(dim=st.sampled_from([1, 2, 3]), n=st.integers(1, 700), K=st.integers(1, 90), seed=st.integers(0, 2**31 - 1),
       inject=st.booleans(), drop_bias=st.floats(0.0, 0.9), plan_choose=st.integers(0, 1), step_mode=st.booleans())
def test_invariants_hold_on_random_configurations(dim, n, K, seed, inject, drop_bias, plan_choose, step_mode):
    from snac_b200.vecenv import BatchedDMPEnv
    A = {1: 3, 2: 5, 3: 8}[dim]
    env = BatchedDMPEnv(dim, plan_choose=plan_choose, num_envs=n, auto_reset=False, seed=seed, obs_dtype=torch.float64)
    o0 = env.reset()
    rng = np.random.RandomState(seed % 9973)
    acts = sizes = None
    if inject:                                               # a mix that favours drops / builds by `drop_bias`
        n_move = {1: 2, 2: 4, 3: 4}[dim]
        p = np.r_[np.full(n_move, (1 - drop_bias) / n_move), np.full(A - n_move, drop_bias / (A - n_move))]
        acts = torch.as_tensor(rng.choice(A, size=(K, n), p=p).astype(np.uint8), device=env.device)
        sizes = torch.as_tensor(rng.randint(1, 4, size=(K, n)).astype(np.uint8), device=env.device)
    if step_mode:
        outs = [[x.clone() for x in env.step(None if acts is None else acts[k], None if sizes is None else sizes[k])] for k in range(K)]
        obs, rew, done = [torch.stack([o[i] for o in outs]) for i in range(3)]
    else:
        obs, rew, done = env.rollout(K, actions=acts, step_sizes=sizes)
    torch.cuda.synchronize()
    obs, rew, done = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()
    ex = env.export_state()
    grid, sc = ex["grid"].cpu().numpy(), ex["scalars"].cpu().numpy()
    lo, hi = HW[dim], HW[dim] + (30 if dim == 1 else 20) - 1
    # positions inside the plan area; counters: every step counted, bricks never more than steps
    assert (sc[:, 0] >= lo).all() and (sc[:, 0] <= hi).all()
    if dim > 1:
        assert (sc[:, 1] >= lo).all() and (sc[:, 1] <= hi).all()
    assert (sc[:, 3] == K).all() and (sc[:, 2] >= 0).all() and (sc[:, 2] <= K).all()
    assert np.array_equal(obs[:, :, -1], np.broadcast_to(np.arange(1, K + 1)[:, None], (K, n)))      # count_step column
    assert (np.diff(obs[:, :, -2], axis=0) >= 0).all() and np.array_equal(obs[-1, :, -2], sc[:, 2])   # count_brick column
    # grids: the frame is -1 and never changes, interior cells are counts (2D: 0 / 1), their sum is what was laid
    g = grid.reshape(n, -1) if dim == 1 else grid.reshape(n, 26, 26)
    interior = g[:, 2:32] if dim == 1 else g[:, 3:23, 3:23]
    assert (interior >= 0).all() and (np.sort(np.unique(g[g < 0])) == [-1]).all()
    laid = interior.reshape(n, -1).sum(1)
    if dim == 2:                                             # a second drop on a cell is counted but leaves the cell at 1
        assert interior.max() <= 1 and (laid <= sc[:, 2]).all()
    else:                                                    # 1D / 3D: every counted brick raises one cell by one
        assert np.array_equal(laid, sc[:, 2])
    # the last observation is the window of the final state
    assert np.array_equal(obs[-1, :, :-2], window_of(dim, grid, sc[:, :2]))
    # rewards from the class's reward set; done is sticky only through the step limit / budget (never un-done by a move)
    assert set(np.unique(rew)) <= REWARDS[dim]
    iou = env.iou().cpu().numpy()
    assert ((iou >= 0) & (iou <= 1)).all()
    assert o0.shape == (n, obs.shape[-1])
    env.check_errors()


@settings(max_examples=int(__import__("os").environ.get("SNAC_PROPERTY_EXAMPLES", "30")), deadline=None, derandomize=__import__("os").environ.get("SNAC_PROPERTY_RANDOM", "0") != "1", suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(dim=st.sampled_from([2, 3]), n=st.integers(1, 600), K=st.integers(1, 120), seed=st.integers(0, 2**31 - 1),
       drop_bias=st.floats(0.0, 0.95), step_mode=st.booleans(), auto_reset=st.booleans(), reset_obs=st.booleans())
def test_bit_records_carry_what_the_float_rows_carry(dim, n, K, seed, drop_bias, step_mode, auto_reset, reset_obs):
    """DMP_OBS_BITS on random configurations: the expansion of the bit records equals the float64 rows of a twin env fed the
    same draws -- window cells clipped at 14 in 3D, with the saturation flag exactly on the records that were clipped -- on
    the host and on the device; rewards, done flags and the final state agree."""
    from snac_b200.vecenv import BatchedDMPEnv, unpack_bits, unpack_records_device
    A = {2: 5, 3: 8}[dim]
    kw = dict(plan_choose=seed & 1, num_envs=n, auto_reset=auto_reset, reset_obs=auto_reset and reset_obs, seed=seed)
    a = BatchedDMPEnv(dim, obs_dtype="bits", **kw)
    b = BatchedDMPEnv(dim, obs_dtype=torch.float64, **kw)
    a.reset(), b.reset()
    rng = np.random.RandomState(seed % 7919)
    p = np.r_[np.full(4, (1 - drop_bias) / 4), np.full(A - 4, drop_bias / (A - 4))]
    acts = torch.as_tensor(rng.choice(A, size=(K, n), p=p).astype(np.uint8), device=a.device)
    sizes = torch.as_tensor(rng.randint(1, 4, size=(K, n)).astype(np.uint8), device=a.device)

    def run(env):
        if step_mode:
            outs = [[x.clone() for x in env.step(acts[k], sizes[k])] for k in range(K)]
            return [torch.stack([o[i] for o in outs]) for i in range(3)]
        return env.rollout(K, actions=acts, step_sizes=sizes)
    rec, ra, da = run(a)
    obs, rb, db = run(b)
    torch.cuda.synchronize()
    o, r, d, sat = unpack_bits(rec, dim)
    ref = obs.cpu().numpy()
    clipped = np.concatenate([np.minimum(ref[..., :49], 14), ref[..., 49:]], axis=-1)
    assert np.array_equal(o, clipped) and np.array_equal(sat, ref[..., :49].max(axis=-1) >= 14)
    assert np.array_equal(r, rb.cpu().numpy()) and np.array_equal(d, db.cpu().numpy())
    assert torch.equal(ra, rb) and torch.equal(da, db)
    od, rd, dd, sd = unpack_records_device(rec, dim, "bits", torch.float64)
    assert np.array_equal(od.cpu().numpy(), o) and np.array_equal(rd.cpu().numpy(), r)
    assert np.array_equal(dd.cpu().numpy(), d) and np.array_equal(sd.cpu().numpy(), sat)
    ea, eb = a.export_state(), b.export_state()
    assert torch.equal(ea["grid"], eb["grid"]) and torch.equal(ea["scalars"], eb["scalars"])
    a.check_errors()
