"""Random plan generators of the reference, restated on the CPU (TEST INFRASTRUCTURE ONLY).

Reference code followed (the stochastic draws are ARGUMENTS here, like everywhere in the oracle):

* 1D random sinusoid  -- ``deep_mobile_printing_1d1r_hindsight.create_plan``,
  Env/1D/DMP_Env_1D_dynamic_hindsight_replay.py:29-42:
      k_1 = uniform(3, 12); k_2 = randint(1, 4); phase = uniform(-1, 1) * pi
      y = round(k_1 * sin(2*pi/30 * (k_2 * x + phase)) + 20),  x = 0..29;   area = sum(y)
* 2D random triangle  -- ``create_plan`` of Env/2D/DMP_Env_2D_dynamic_hindsight_replay_usedata.py:37-59 and
  Env/2D/DMP_ENV_2D_dynamic_MCTS.py:40-62 (identical):
      repeat: x = randint(0, 20, size=3); y = randint(0, 20, size=3)
              cv2.polylines(white 20x20 image, triangle, closed)            (1-pixel, 8-connected outline)
              dense (plan_choose 0): cv2.fillPoly(same triangle)
              plan interior = 1 where the image is black;  total_area = number of ones
      until total_area > [50, 20][plan_choose]
  The 3D dataset classes have no generator of their own; their plans are these masks times z = 6
  (Env/3D/DMP_simulator_3d_dynamic_triangle_usedata.py:47-49 divides by z again for the budget).

cv2 (OpenCV, ``opencv-python`` in requirements.txt, unpinned; 4.13.0 in the build container) is a third-party
dependency, so its arithmetic is restated here:
  - ``line_pixels``: OpenCV's 8-connected LineIterator for integer end points (endpoints swapped so that x
    increases, then Bresenham with error term ``dx - 2*dy`` and the "err < 0" step rule);
  - dense triangles: every pixel fillPoly adds lies between the left-most and right-most OUTLINE pixel of its
    row, so polylines + fillPoly == per-row span fill of the outline.
PINNED: tests/golden/make_plangen_golden.py compares ``line_pixels`` with cv2.line for all 160 000 segments
of the 20x20 grid and ``triangle_mask`` with cv2.polylines(+fillPoly) for ALL 10 746 800 unordered vertex
triples (dense and sparse), and commits SHA-256 digests of both exhaustive mask tables, which the GPU tests
re-derive with the CUDA generator.
"""
from __future__ import annotations

import numpy as np

AREA_MIN = (50, 20)                 # create_plan: `while total_area <= area[plan_choose]`
PLAN_TAG = 0x504C414E               # "PLAN": fourth counter word of the plan-generation Philox stream


def plan_1d_sin(k_1: float, k_2: int, phase: float) -> np.ndarray:
    """Env/1D/DMP_Env_1D_dynamic_hindsight_replay.py:38-40 (same numpy expression)."""
    x = np.arange(30)
    return np.round((k_1 * np.sin(2 * np.pi / 30 * (k_2 * x + phase)) + 20))


def line_pixels(x1: int, y1: int, x2: int, y2: int):
    """Pixels of cv2.line(img, (x1,y1), (x2,y2), thickness=1, lineType=LINE_8), integer end points."""
    dx, dy = x2 - x1, y2 - y1
    if dx < 0:
        x1, y1, x2, y2, dx, dy = x2, y2, x1, y1, -dx, -dy
    ystep = 1
    if dy < 0:
        dy, ystep = -dy, -1
    out = []
    x, y = x1, y1
    if dy > dx:                                   # steep: one pixel per row
        err = dy - 2 * dx
        for _ in range(dy + 1):
            out.append((x, y))
            neg = err < 0
            err += -2 * dx + (2 * dy if neg else 0)
            y += ystep
            x += 1 if neg else 0
    else:                                         # shallow: one pixel per column
        err = dx - 2 * dy
        for _ in range(dx + 1):
            out.append((x, y))
            neg = err < 0
            err += -2 * dy + (2 * dx if neg else 0)
            x += 1
            y += ystep if neg else 0
    return out


def triangle_mask(xs, ys, dense: bool) -> np.ndarray:
    """(20,20) uint8 mask [row=y][col=x] of create_plan's image for one vertex triple."""
    img = np.zeros((20, 20), np.uint8)
    lo = np.full(20, 99)
    hi = np.full(20, -1)
    for i in range(3):
        j = (i + 1) % 3
        for px, py in line_pixels(int(xs[i]), int(ys[i]), int(xs[j]), int(ys[j])):
            img[py, px] = 1
            lo[py] = min(lo[py], px)
            hi[py] = max(hi[py], px)
    if dense:
        for r in range(20):
            if hi[r] >= 0:
                img[r, lo[r]:hi[r] + 1] = 1
    return img


def create_plan_2d(draw_vertices, plan_choose: int):
    """The reference's retry loop.  draw_vertices() -> (x[3], y[3]) is called once per attempt.
    Returns (26x26 float64 plan, total_area, attempts)."""
    attempts = 0
    while True:
        xs, ys = draw_vertices()
        attempts += 1
        m = triangle_mask(xs, ys, plan_choose == 0)
        area = int(m.sum())
        if area > AREA_MIN[plan_choose]:
            plan = np.zeros((26, 26))
            plan[3:23, 3:23] = m
            return plan, float(area), attempts


# ---- Philox stream of the on-device generators (single source of truth for kernel and tests) -------------------
def philox_vertices(seed: int, plan_ids, attempt: int):
    """counter = (plan_id lo, plan_id hi, attempt, "PLAN");  vertex i: x = ((x_i & 0xffff) * 20) >> 16,
    y = ((x_i >> 16) * 20) >> 16, i = 0..2."""
    from . import philox as P
    ids = np.asarray(plan_ids, dtype=np.uint64)
    w = P.philox4x32_10(ids & P.MASK, ids >> np.uint64(32), int(attempt) & 0xFFFFFFFF, PLAN_TAG,
                        seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    xs = np.stack([((w[i].astype(np.uint64) & np.uint64(0xFFFF)) * np.uint64(20)) >> np.uint64(16) for i in range(3)], -1)
    ys = np.stack([((w[i].astype(np.uint64) >> np.uint64(16)) * np.uint64(20)) >> np.uint64(16) for i in range(3)], -1)
    return xs.astype(np.int64), ys.astype(np.int64)


def philox_sin_params(seed: int, plan_ids):
    """k_1 = 3 + 9 * (x0 / 2^32);  k_2 = 1 + mulhi32(x1, 3);  phase = (x2 / 2^31 - 1) * pi   (fp64, IEEE ops)."""
    from . import philox as P
    ids = np.asarray(plan_ids, dtype=np.uint64)
    w = P.philox4x32_10(ids & P.MASK, ids >> np.uint64(32), 0, PLAN_TAG, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    k1 = 3.0 + 9.0 * (w[0].astype(np.float64) * 2.0 ** -32)
    k2 = 1 + P.mulhi32(w[1], 3)
    phase = (w[2].astype(np.float64) * 2.0 ** -31 - 1.0) * np.pi
    return k1, k2, phase


def generate_2d(seed: int, plan_ids, plan_choose: int, max_attempts: int = 64):
    """Masks (n,20,20), areas and attempt counts of the on-device Philox generator."""
    plan_ids = np.asarray(plan_ids, dtype=np.uint64)
    n = len(plan_ids)
    masks = np.zeros((n, 20, 20), np.uint8)
    areas = np.zeros(n, np.int64)
    att = np.zeros(n, np.int64)
    todo = np.arange(n)
    for a in range(max_attempts):
        if len(todo) == 0:
            break
        xs, ys = philox_vertices(seed, plan_ids[todo], a)
        keep = []
        for k, i in enumerate(todo):
            m = triangle_mask(xs[k], ys[k], plan_choose == 0)
            masks[i], areas[i], att[i] = m, int(m.sum()), a + 1
            if areas[i] <= AREA_MIN[plan_choose]:
                keep.append(i)
        todo = np.asarray(keep, dtype=np.int64)
    return masks, areas, att
