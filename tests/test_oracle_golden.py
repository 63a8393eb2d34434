"""The oracle (oracle/dmp_oracle.py) against the golden traces recorded from the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, TRACE_NAMES, load_plans, load_trace, trace_env_spec
from oracle import dmp_oracle as O


def replay(tr):
    dim, dynamic, plan_choose, plans = trace_env_spec(tr["meta"])
    env = O.make_env(dim, dynamic, plan_choose=plan_choose, plans=plans)
    resets = list(tr["reset_plan_idx"])
    ri = 0
    o = env.reset(resets[ri])
    assert np.array_equal(o[0], tr["reset_obs"][ri].astype(np.float64))
    ep = 0
    T = len(tr["actions"])
    for t in range(T):
        o, r, d = env.step(int(tr["actions"][t]), int(tr["step_sizes"][t]))
        assert o.dtype == np.float64 and o.shape == (1, env.D)
        assert np.array_equal(o[0], tr["obs"][t].astype(np.float64)), (t, o[0], tr["obs"][t])
        assert r == tr["reward"][t] and isinstance(r, int) == bool(tr["reward_is_int"][t]), (t, r)
        assert d == bool(tr["done"][t]) and isinstance(d, bool)
        if dynamic:
            assert np.array_equal(env.obs_normalised()[0, -2:], tr["obs_norm"][t])
        assert env.count_brick == tr["count_brick"][t] and env.count_step == tr["count_step"][t]
        assert env.total_brick == tr["total_brick"][t]
        pos = [env.pos, 0] if dim == 1 else env.pos
        assert list(pos) == list(tr["pos"][t])
        iou = env.iou()
        assert iou == tr["iou"][t] or (np.isnan(iou) and np.isnan(tr["iou"][t]))
        if d:
            assert np.array_equal(env.grid.astype(np.int16).reshape(tr["final_grids"][ep].shape), tr["final_grids"][ep])
            ep += 1
            ri += 1
            o = env.reset(resets[ri])
            assert np.array_equal(o[0], tr["reset_obs"][ri].astype(np.float64))
    assert np.array_equal(env.grid.astype(np.int16).reshape(tr["final_grids"][ep].shape), tr["final_grids"][ep])


@pytest.mark.parametrize("name", TRACE_NAMES)
def test_oracle_replays_reference_trace(name):
    replay(load_trace(name))


def test_oracle_known_answers():
    """one seeded episode per env family, produced by the real reference."""
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        rows = json.load(f)
    assert len(rows) == 10
    for row in rows:
        dim = int(row["dim"][0])
        dynamic = row["kind"] == "dynamic"
        if dynamic:
            env = O.make_env(dim, True, plans=load_plans(dim, row["kw"]["density"], row["kw"]["split"]), sequential=True)
        else:
            env = O.make_env(dim, False, plan_choose=row["kw"]["plan_choose"])
        env.reset()
        ret, done = 0.0, False
        for a, s in zip(row["actions"], row["step_sizes"]):
            assert not done
            _, r, done = env.step(a, s)
            ret += r
        assert done
        assert env.count_step == row["steps"] and ret == row["ret"]
        assert env.count_brick == row["bricks"] and env.total_brick == row["total_brick"]
        assert ([env.pos] if dim == 1 else list(env.pos)) == row["pos"]
        assert env.iou() == row["iou"]


def test_static_polygon_matches_survey_bitmaps():
    """SURVEY.md App. A.4 row spans of the restated matplotlib CirclePolygon (parity unpinned)."""
    dense = {6: (11, 14), 7: (9, 16), 8: (8, 17), 9: (7, 18), 10: (7, 18), 11: (6, 19), 12: (6, 19), 13: (6, 19),
             14: (6, 19), 15: (7, 18), 16: (7, 18), 17: (8, 17), 18: (9, 16), 19: (11, 14)}
    m = O.circle_polygon_mask(0)
    assert m.sum() == 148
    for r in range(26):
        cols = np.nonzero(m[r])[0]
        if r in dense:
            assert (cols.min(), cols.max()) == dense[r] and len(cols) == dense[r][1] - dense[r][0] + 1
        else:
            assert len(cols) == 0
    sparse = {5: [10, 11, 12, 13, 14, 15], 6: [8, 9, 10, 15, 16, 17], 7: [7, 8, 17, 18], 8: [6, 7, 18, 19], 9: [6, 19],
              10: [5, 6, 19, 20], 11: [5, 20], 12: [5, 20], 13: [5, 20], 14: [5, 20], 15: [5, 6, 19, 20],
              16: [6, 19], 17: [6, 7, 18, 19], 18: [7, 8, 17, 18], 19: [8, 9, 10, 15, 16, 17], 20: [10, 11, 12, 13, 14, 15]}
    m = O.circle_polygon_mask(1)
    assert m.sum() == 60
    for r in range(26):
        assert list(np.nonzero(m[r])[0]) == sparse.get(r, [])
    assert O.total_brick_of(2, O.static_plan(2, 0), False) == 148 and O.total_brick_of(2, O.static_plan(2, 1), False) == 60
    assert O.total_brick_of(3, O.static_plan(3, 0), False) == 888 and O.total_brick_of(3, O.static_plan(3, 1), False) == 360
    assert [O.plan_1d_static(k).sum() for k in range(3)] == [600, 590, 600]
    with pytest.raises(ValueError):
        O.static_plan(1, 3)
    with pytest.raises(ValueError):
        O.static_plan(2, 2)


def test_static_polygon_matches_the_reference_figure():
    """External pin of the restated CirclePolygon: docs/2d_example_crop.png and the right panel of docs/3d_example_crop.png
    of the reference are matplotlib renderings of the sparse design on its 20 x 20 grid; tests/golden/make_polygon_pin.py
    decoded both cell by cell (they agree on the 315 cells both show; together 389 of 400 cells are not covered by labels /
    the robot, and all 60 ring cells are visible).  The oracle's sparse mask must equal the figures wherever they show the
    cell, and the dense mask (radius 7) must be exactly what the ring encloses."""
    z = np.load(os.path.join(GOLDEN, "sparse_ring_from_docs.npz"))
    cell, known = z["cell"].astype(bool), z["known"].astype(bool)
    assert known.sum() >= 389 and cell.sum() == 60
    ring = O.static_plan(2, 1)[3:23, 3:23] > 0
    dense = O.static_plan(2, 0)[3:23, 3:23] > 0
    assert np.array_equal(ring[known], cell[known])
    assert np.array_equal(ring[::-1][known], cell[known])            # the figure's y axis points up: same mask either way
    assert not (ring & ~known).any()                                   # every ring cell of the oracle is one the figures show
    # dense = inside the radius-7 polygon = the cells the ring (inside 8, not inside 7) encloses: flood fill from the centre
    fill, stack = np.zeros((20, 20), bool), [(10, 10)]
    while stack:
        i, j = stack.pop()
        if 0 <= i < 20 and 0 <= j < 20 and not fill[i, j] and not ring[i, j]:
            fill[i, j] = True
            stack += [(i + 1, j), (i - 1, j), (i, j + 1), (i, j - 1)]
    assert np.array_equal(fill, dense) and dense.sum() == 148 and ring.sum() == 60


def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10."""
    from oracle.philox import philox4x32_10
    h = lambda t: [int(v) for v in t]
    assert h(philox4x32_10(0, 0, 0, 0, 0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert h(philox4x32_10(*[0xffffffff] * 6)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert h(philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_out_of_range_action_raises_like_reference():
    for dim in (1, 2):
        env = O.make_env(dim, False, plan_choose=0)
        env.reset()
        with pytest.raises(UnboundLocalError):
            env.step(O.SPEC[dim]["actions"], 1)
