#!/usr/bin/env python
"""Pins oracle/plangen.py (and its C twin) against the real thing and writes tests/golden/plangen_golden.npz.

Runs only in the build container (needs cv2 and /root/reference):
  1. cv2.line vs plangen.line_pixels for ALL 160 000 integer segments of the 20x20 grid;
  2. cv2.polylines (+ cv2.fillPoly) vs the C oracle for ALL 10 746 800 unordered vertex triples, sparse and
     dense, plus 200 000 random ORDERED triples (vertex order does not matter to cv2 either);
  3. the unmodified reference generators (create_plan of the 1D / 2D hindsight classes) for seeded numpy
     streams vs plangen.plan_1d_sin / plangen.create_plan_2d replaying the same streams.
Committed: SHA-256 of the two exhaustive mask tables (enumeration order: point index p = y*20+x,
p0 <= p1 <= p2 lexicographic), 4 096 explicit triples with their masks, and the reference generator outputs.

    python tests/golden/make_plangen_golden.py
"""
from __future__ import annotations

import ctypes as C
import hashlib
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "plangen_golden.npz")


def cv2_mask(xs, ys, dense):
    """The reference's own statements (Env/2D/DMP_Env_2D_dynamic_hindsight_replay_usedata.py:46-55)."""
    import cv2
    img_rgb = np.ones((20, 20, 3), np.uint8) * 255
    vertices = np.array([[xs[0], ys[0]], [xs[1], ys[1]], [xs[2], ys[2]]], np.int32)
    pts = vertices.reshape((-1, 1, 2))
    cv2.polylines(img_rgb, [pts], isClosed=True, color=(0, 0, 0))
    if dense:
        cv2.fillPoly(img_rgb, [pts], color=(0, 0, 0))
    return (img_rgb[:, :, 0] == 0)


def pack13(masks):
    """(n,20,20) bool -> (n,13) uint32, bit r*20+c."""
    flat = np.zeros((len(masks), 416), np.uint8)
    flat[:, :400] = masks.reshape(len(masks), 400)
    return np.packbits(flat, axis=1, bitorder="little").view(np.uint32)


def c_oracle_masks(xs, ys, dense):
    from oracle.build import build_oracle
    lib = C.CDLL(build_oracle())
    n = len(xs)
    out = np.zeros((n, 13), np.uint32)
    xs = np.ascontiguousarray(xs, np.int32)
    ys = np.ascontiguousarray(ys, np.int32)
    lib.orc_triangle_masks(C.c_int64(n), xs.ctypes.data_as(C.c_void_p), ys.ctypes.data_as(C.c_void_p), int(dense),
                           out.ctypes.data_as(C.c_void_p), None)
    return out


def enumerate_triples():
    """All multisets {p0 <= p1 <= p2} of grid points, p = y*20+x: arrays xs[n,3], ys[n,3] (n = 10 746 800)."""
    p0, p1, p2 = [], [], []
    for a in range(400):
        b = np.arange(a, 400)
        cnt = 400 - b                                   # number of c >= b
        bb = np.repeat(b, cnt)
        cc = np.concatenate([np.arange(x, 400) for x in b])
        p0.append(np.full(len(bb), a, np.int16)); p1.append(bb.astype(np.int16)); p2.append(cc.astype(np.int16))
    P = np.stack([np.concatenate(p0), np.concatenate(p1), np.concatenate(p2)], 1).astype(np.int32)
    return P % 20, P // 20


def _check_chunk(args):
    lo, hi, dense = args
    xs, ys = _G["xs"][lo:hi], _G["ys"][lo:hi]
    ref = np.zeros((hi - lo, 20, 20), bool)
    for i in range(hi - lo):
        ref[i] = cv2_mask(xs[i], ys[i], dense)
    return lo, pack13(ref)


_G = {}
MAX_ATT = 96                     # attempts kept per reference create_plan call (P(area > 50) is about 0.2 per draw)


def exhaustive(dense, procs):
    xs, ys = _G["xs"], _G["ys"]
    n = len(xs)
    mine = c_oracle_masks(xs, ys, dense)
    step = 20000
    jobs = [(lo, min(lo + step, n), dense) for lo in range(0, n, step)]
    bad = 0
    with mp.get_context("fork").Pool(procs) as pool:
        for lo, ref in pool.imap_unordered(_check_chunk, jobs, chunksize=4):
            d = (ref != mine[lo:lo + len(ref)]).any(1)
            if d.any():
                bad += int(d.sum())
                i = lo + int(np.argmax(d))
                print("MISMATCH dense=%d triple x=%s y=%s" % (dense, xs[i], ys[i]))
    return mine, bad


def main():
    import cv2
    from oracle import plangen as G
    from oracle import refload
    t0 = time.time()
    # 1. lines
    bad = 0
    for x1 in range(20):
        for y1 in range(20):
            for x2 in range(20):
                for y2 in range(20):
                    img = np.zeros((20, 20), np.uint8)
                    cv2.line(img, (x1, y1), (x2, y2), 1)
                    mine = np.zeros((20, 20), np.uint8)
                    for px, py in G.line_pixels(x1, y1, x2, y2):
                        mine[py, px] = 1
                    bad += not np.array_equal(img, mine)
    print("lines: 160000 segments, %d mismatches (%.0f s)" % (bad, time.time() - t0))
    assert bad == 0
    # 2. triangles, exhaustive
    _G["xs"], _G["ys"] = enumerate_triples()
    n = len(_G["xs"])
    assert n == 10746800
    procs = len(os.sched_getaffinity(0))
    digests = {}
    for dense in (0, 1):
        mine, bad = exhaustive(dense, procs)
        print("triangles dense=%d: %d triples, %d mismatches (%.0f s)" % (dense, n, bad, time.time() - t0))
        assert bad == 0
        digests[dense] = hashlib.sha256(mine.tobytes()).hexdigest()
    rng = np.random.RandomState(2024)
    oxs, oys = rng.randint(0, 20, size=(200000, 3)), rng.randint(0, 20, size=(200000, 3))
    for dense in (0, 1):
        mine = c_oracle_masks(oxs, oys, dense)
        ref = pack13(np.stack([cv2_mask(oxs[i], oys[i], dense) for i in range(len(oxs))]))
        assert np.array_equal(mine, ref), "ordered triples differ (dense=%d)" % dense
    print("ordered random triples OK (%.0f s)" % (time.time() - t0))
    sample = rng.choice(n, size=4096, replace=False)
    sx, sy = _G["xs"][sample], _G["ys"][sample]
    # 3. the reference generators themselves
    ref1d = refload.load_class("1D", "hindsight_dynamic")()
    p1, y1 = [], []
    np.random.seed(7)
    for _ in range(2000):
        y, area = ref1d.create_plan()
        assert area == float(np.sum(y))
        p1.append(ref1d.one_hot)
        y1.append(y)
        assert np.array_equal(G.plan_1d_sin(*ref1d.one_hot), y)
    out2 = {}
    for pc, dens in ((0, "dense"), (1, "sparse")):
        env = refload.load_class("2D", "hindsight_dynamic")(refload.dataset_path("2D", dens, "train"))
        assert env.plan_choose == pc
        plans, areas, verts, nverts = [], [], [], []
        for seed in range(300):
            np.random.seed(1000 + seed)
            plan, area = env.create_plan()
            np.random.seed(1000 + seed)                  # replay the same stream through the restated loop
            log = []

            def draw():
                x = np.random.randint(0, 20, size=3)
                y = np.random.randint(0, 20, size=3)
                log.append(np.concatenate([x, y]))
                return x, y
            plan2, area2, att = G.create_plan_2d(draw, pc)
            assert np.array_equal(plan, plan2) and float(area) == area2, (pc, seed)
            plans.append(plan[3:23, 3:23].astype(np.uint8)); areas.append(area)
            v = np.zeros((MAX_ATT, 6), np.int32)
            assert att <= MAX_ATT
            v[:att] = np.stack(log)
            verts.append(v); nverts.append(att)
        out2[dens] = (np.stack(plans), np.asarray(areas), np.stack(verts), np.asarray(nverts))
        print("reference create_plan 2D %s: 300 seeds OK, attempts max %d" % (dens, max(nverts)))
    np.savez_compressed(
        OUT,
        sha256_sparse=digests[0], sha256_dense=digests[1], n_triples=n,
        sample_x=sx.astype(np.int8), sample_y=sy.astype(np.int8),
        sample_sparse=c_oracle_masks(sx, sy, 0), sample_dense=c_oracle_masks(sx, sy, 1),
        sin_params=np.asarray(p1, np.float64), sin_plans=np.asarray(y1, np.uint8),
        ref2d_dense_plans=np.packbits(out2["dense"][0], axis=None), ref2d_dense_area=out2["dense"][1],
        ref2d_dense_verts=out2["dense"][2], ref2d_dense_attempts=out2["dense"][3],
        ref2d_sparse_plans=np.packbits(out2["sparse"][0], axis=None), ref2d_sparse_area=out2["sparse"][1],
        ref2d_sparse_verts=out2["sparse"][2], ref2d_sparse_attempts=out2["sparse"][3],
        cv2_version=cv2.__version__)
    print("wrote", OUT, "(%.0f s)" % (time.time() - t0))


if __name__ == "__main__":
    main()
