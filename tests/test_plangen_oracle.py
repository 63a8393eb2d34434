"""oracle/plangen.py (+ its C twin) against the committed golden vectors of tests/golden/make_plangen_golden.py:
cv2 masks of 4 096 sampled vertex triples, SHA-256 of the exhaustive mask tables, and the outputs of the unmodified
reference generators (create_plan of the 1D / 2D hindsight classes) replayed from their recorded numpy draws."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import plangen as G


@pytest.fixture(scope="module")
def gold():
    z = np.load(os.path.join(GOLDEN, "plangen_golden.npz"))
    return {k: z[k] for k in z.files}


def pack13(masks):
    flat = np.zeros((len(masks), 416), np.uint8)
    flat[:, :400] = np.asarray(masks).reshape(len(masks), 400)
    return np.packbits(flat, axis=1, bitorder="little").view(np.uint32)


def c_masks(xs, ys, dense):
    from oracle.build import build_oracle
    lib = C.CDLL(build_oracle())
    xs, ys = np.ascontiguousarray(xs, np.int32), np.ascontiguousarray(ys, np.int32)
    out = np.zeros((len(xs), 13), np.uint32)
    area = np.zeros(len(xs), np.int32)
    lib.orc_triangle_masks(C.c_int64(len(xs)), xs.ctypes.data_as(C.c_void_p), ys.ctypes.data_as(C.c_void_p), int(dense),
                           out.ctypes.data_as(C.c_void_p), area.ctypes.data_as(C.c_void_p))
    return out, area


def enumerate_triples():
    """Same enumeration as make_plangen_golden.enumerate_triples: p = y*20+x, p0 <= p1 <= p2."""
    a, b, c = np.meshgrid(np.arange(400, dtype=np.int16), np.arange(400, dtype=np.int16), np.arange(400, dtype=np.int16),
                          indexing="ij", sparse=True)
    keep = np.nonzero((a <= b) & (b <= c))
    P = np.stack(keep, 1).astype(np.int32)
    return P % 20, P // 20


def test_sampled_masks_python_and_c(gold):
    sx, sy = gold["sample_x"].astype(np.int64), gold["sample_y"].astype(np.int64)
    for dense, key in ((0, "sample_sparse"), (1, "sample_dense")):
        mine_c, area = c_masks(sx, sy, dense)
        assert np.array_equal(mine_c, gold[key])
        py = pack13(np.stack([G.triangle_mask(sx[i], sy[i], bool(dense)) for i in range(512)]))
        assert np.array_equal(py, gold[key][:512])
        bits = np.unpackbits(gold[key].view(np.uint8), axis=1, bitorder="little")
        assert np.array_equal(bits.sum(1), area)


def test_exhaustive_digest_c_oracle(gold):
    xs, ys = enumerate_triples()
    assert len(xs) == int(gold["n_triples"]) == 10746800
    for dense, key in ((0, "sha256_sparse"), (1, "sha256_dense")):
        m, _ = c_masks(xs, ys, dense)
        assert hashlib.sha256(m.tobytes()).hexdigest() == str(gold[key])


def test_reference_sinusoids(gold):
    for params, y in zip(gold["sin_params"][:500], gold["sin_plans"][:500]):
        assert np.array_equal(G.plan_1d_sin(params[0], int(params[1]), params[2]), y.astype(np.float64))


@pytest.mark.parametrize("dens,pc", [("dense", 0), ("sparse", 1)])
def test_reference_create_plan_replay(gold, dens, pc):
    plans = np.unpackbits(gold["ref2d_%s_plans" % dens])[:300 * 400].reshape(300, 20, 20)
    verts, natt, areas = gold["ref2d_%s_verts" % dens], gold["ref2d_%s_attempts" % dens], gold["ref2d_%s_area" % dens]
    for i in range(300):
        it = iter(verts[i])

        def draw():
            v = next(it)
            return v[:3], v[3:]
        plan, area, att = G.create_plan_2d(draw, pc)
        assert att == natt[i] and area == areas[i]
        assert np.array_equal(plan[3:23, 3:23], plans[i])
        assert plan.sum() == area


def test_philox_generator_is_self_consistent():
    masks, areas, att = G.generate_2d(0x534E4143, np.arange(100, 164), 0)
    assert (areas > 50).all() and (att >= 1).all()
    xs, ys = G.philox_vertices(0x534E4143, np.arange(100, 164), 0)
    assert xs.min() >= 0 and xs.max() <= 19 and ys.min() >= 0 and ys.max() <= 19
    first = att == 1
    for i in np.nonzero(first)[0][:8]:
        assert np.array_equal(masks[i], G.triangle_mask(xs[i], ys[i], True))
    k1, k2, ph = G.philox_sin_params(7, np.arange(50))
    assert ((k1 >= 3) & (k1 < 12)).all() and set(np.unique(k2)) <= {1, 2, 3} and (np.abs(ph) <= np.pi).all()
