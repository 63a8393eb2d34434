"""CPU oracle for the SNAC mobile-construction envs (TEST INFRASTRUCTURE ONLY).

This module is a plain-Python/numpy restatement of the reference algorithm for the
hot path named in SURVEY.md section 8: the six base simulators under
``/root/reference/Env/{1D,2D,3D}``.  It is the checker for the CUDA path; it is
never imported by the product package ``snac_b200`` (only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it).

Deterministic entry points: the reference draws ``np.random.randint(1, 4)`` inside
``step`` and ``np.random.randint(0, len)`` inside ``reset``; here both draws are
arguments (``step(action, step_size)``, ``reset(plan_idx)``), exactly like the
reference's own ``*_hindsight_replay`` variants
(Env/1D/DMP_Env_1D_static_hindsight_replay.py:85-87,
Env/2D/DMP_Env_2D_static_hindsight_replay.py:95-97,
Env/3D/DMP_simulator_3d_static_circle_hindsight_replay.py:152-154).

Pinning (see tests/golden/make_golden.py, tests/test_oracle_golden.py):
  * 1D static/dynamic, 2D dynamic, 3D dynamic: pinned step-for-step against traces
    produced by the UNMODIFIED reference classes run in the build container.
  * 2D static / 3D static: the step logic is pinned the same way, but the static
    *plan* comes from ``matplotlib.patches.CirclePolygon`` (requirements.txt:3,
    unpinned; call sites Env/2D/DMP_Env_2D_static.py:43-48,
    Env/3D/DMP_simulator_3d_static_circle.py:54-59), a third-party dependency
    that is neither vendored in the reference nor installed here.  ``circle_polygon_mask``
    restates its published algorithm (regular 20-gon + crossing-number test).
    For that one function: PARITY UNPINNED (checked only against the row spans
    recorded in SURVEY.md App. A.4 and the areas 148/60).
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np

# --------------------------------------------------------------------------------------
# geometry constants (reference __init__ blocks)
#   1D: Env/1D/DMP_Env_1D_static.py:9-29            W=30, HW=2, T=750, A=3, D=7
#   2D: Env/2D/DMP_Env_2D_static.py:10-29           20x20, HW=3, T=600, A=5, D=51
#   3D: Env/3D/DMP_simulator_3d_static_circle.py:10-40   20x20, HW=3, z=6, T=1300 (static)
#       Env/3D/DMP_simulator_3d_dynamic_triangle_usedata.py:27   T=1000 (dynamic)
# --------------------------------------------------------------------------------------
SPEC = {
    1: dict(width=30, hw=2, actions=3, obs_dim=7, total_step=(750, 750)),
    2: dict(width=20, hw=3, actions=5, obs_dim=51, total_step=(600, 600)),
    3: dict(width=20, hw=3, actions=8, obs_dim=51, total_step=(1300, 1000)),
}
Z_HEIGHT = 6


# --------------------------------------------------------------------------------------
# plan generators
# --------------------------------------------------------------------------------------
def plan_1d_static(plan_choose: int) -> np.ndarray:
    """Env/1D/DMP_Env_1D_static.py:34-55 -- 0 sine, 1 Gaussian bump, 2 step curve."""
    if plan_choose == 0:
        x = np.arange(30)
        y = np.round(10 * np.sin(2 * np.pi / 30 * x) + 20)
    elif plan_choose == 1:
        x = np.linspace(-18.0, 18.0, 30)
        pdf = np.exp(-1 * (x ** 2) / (2 * 9)) / (math.sqrt(2 * np.pi) * 3)
        y = np.round(pdf * 100 + 17)
    elif plan_choose == 2:
        y = np.full(30, 15.0)
        for lo in (0, 10, 20):
            y[lo:lo + 5] = 25
    else:
        raise ValueError('0: Sin, 1: Gaussian, 2: Step')
    return y.astype(np.float64)


def _polygon_vertices(radius: float, resolution: int = 20):
    """matplotlib ``Path.unit_regular_polygon(20)`` scaled by radius about (12.5, 12.5):
    vertices at angle pi/2 + 2*pi*k/20, k = 0..20 (closed).  SURVEY.md App. A.4."""
    k = np.arange(resolution + 1)
    theta = 2 * np.pi / resolution * k + np.pi / 2
    return 12.5 + radius * np.cos(theta), 12.5 + radius * np.sin(theta)


def _point_in_polygon(px: float, py: float, vx, vy) -> bool:
    """Crossing-number test as used by matplotlib's ``point_in_path`` (pick radius 0)."""
    inside = False
    n = len(vx) - 1
    x0, y0 = vx[n - 1], vy[n - 1]
    f0 = y0 >= py
    for i in range(n):
        x1, y1 = vx[i], vy[i]
        f1 = y1 >= py
        if f0 != f1:
            if ((y1 - py) * (x0 - x1) >= (x1 - px) * (y0 - y1)) == f1:
                inside = not inside
        f0, x0, y0 = f1, x1, y1
    return inside


def circle_polygon_mask(plan_choose: int) -> np.ndarray:
    """26x26 {0,1} mask of Env/2D/DMP_Env_2D_static.py:31-52 (dense: inside R=7;
    sparse: inside R=8 and not inside R=7).  PARITY UNPINNED (matplotlib absent)."""
    if plan_choose == 0:
        r_out, r_in = 7, 0
    elif plan_choose == 1:
        r_out, r_in = 8, 7
    else:
        raise ValueError('0: Dense circle, 1: Sparse circle')
    ox, oy = _polygon_vertices(r_out)
    ix, iy = _polygon_vertices(r_in)
    m = np.zeros((26, 26), dtype=np.float64)
    for i in range(26):
        for j in range(26):
            if _point_in_polygon(i, j, ox, oy) and not (r_in > 0 and _point_in_polygon(i, j, ix, iy)):
                m[i, j] = 1.0
    return m


_MASK_CACHE: dict = {}


def static_plan(dim: int, plan_choose: int) -> np.ndarray:
    """Static plan in the reference's own array shape: 1D (30,), 2D/3D (26,26)."""
    key = (dim, plan_choose)
    if key not in _MASK_CACHE:
        if dim == 1:
            _MASK_CACHE[key] = plan_1d_static(plan_choose)
        else:
            m = circle_polygon_mask(plan_choose)
            _MASK_CACHE[key] = m * Z_HEIGHT if dim == 3 else m
    return _MASK_CACHE[key].copy()


def total_brick_of(dim: int, plan: np.ndarray, dynamic: bool) -> float:
    """Brick budget.  1D: sum(plan) (static :53,:68 / dynamic :44).  2D: max(area, 30)
    (static :55-57 / dynamic :38,45-46).  3D static: area*z (:62-64) with NO floor;
    3D dynamic: sum(plan/z)*z (Env/3D/...usedata.py:49)."""
    if dim == 1:
        return float(plan.sum())
    if dim == 2:
        return float(max(plan.sum(), 30.0))
    return float((plan / Z_HEIGHT).sum() * Z_HEIGHT)


# --------------------------------------------------------------------------------------
# the env
# --------------------------------------------------------------------------------------
class OracleEnv:
    """One scalar environment.  ``dim`` in {1,2,3}; ``dynamic`` picks the dataset-plan
    rules; ``plans`` is a sequence of reference-shaped plan arrays (static: length 1)."""

    def __init__(self, dim: int, dynamic: bool, plans: Sequence[np.ndarray],
                 sequential: bool = False):
        s = SPEC[dim]
        self.dim, self.dynamic = dim, bool(dynamic)
        self.W, self.HW, self.A, self.D = s['width'], s['hw'], s['actions'], s['obs_dim']
        self.total_step = s['total_step'][1 if dynamic else 0]
        self.plans = [np.asarray(p, dtype=np.float64) for p in plans]
        self.sequential = sequential
        self._seq = 0
        self.lo, self.hi = self.HW, self.W + self.HW - 1      # clamp bounds
        self.plan = None
        self.plan_idx = -1
        self.grid = None
        self.pos = None
        self.count_brick = 0
        self.count_step = 0
        self.total_brick = 0.0

    # -- reset --------------------------------------------------------------------
    def reset(self, plan_idx: Optional[int] = None) -> np.ndarray:
        """1D static :66-83 / dynamic :40-70; 2D static :54-76 / dynamic :34-66;
        3D static :67-86 / dynamic :45-75."""
        if plan_idx is None:
            if self.sequential:          # random_choose_paln=False branch, e.g. 2D dynamic :39-44
                plan_idx = self._seq
                self._seq = (self._seq + 1) % len(self.plans)
            else:
                plan_idx = 0
        self.plan_idx = int(plan_idx)
        self.plan = self.plans[self.plan_idx]
        self.total_brick = total_brick_of(self.dim, self.plan, self.dynamic)
        n = self.W + 2 * self.HW
        if self.dim == 1:
            g = np.zeros((1, n))
            g[:, :self.HW] = -1
            g[:, -self.HW:] = -1
            self.pos = self.HW
        else:
            g = np.zeros((n, n))
            g[:, :self.HW] = -1
            g[:, -self.HW:] = -1
            g[:self.HW, :] = -1
            g[-self.HW:, :] = -1
            self.pos = [self.HW, self.HW]
        self.grid = g
        self.count_brick = 0
        self.count_step = 0
        return self.obs()

    # -- observation --------------------------------------------------------------
    def window(self) -> np.ndarray:
        h = self.HW
        if self.dim == 1:
            return self.grid[:, self.pos - h:self.pos + h + 1]
        r, c = self.pos
        return self.grid[r - h:r + h + 1, c - h:c + h + 1].flatten().reshape(1, -1)

    def obs(self) -> np.ndarray:
        """(1, D) float64 = window, count_brick, count_step (raw counters)."""
        return np.hstack((self.window(), np.array([[self.count_brick]]), np.array([[self.count_step]])))

    def obs_normalised(self) -> np.ndarray:
        """Dynamic envs' normalised counters (1D dynamic :68-70, 2D dynamic :64-65)."""
        return np.hstack((self.window(), np.array([[self.count_brick / self.total_brick]]),
                          np.array([[self.count_step / self.total_step]])))

    def _clamp(self, v: int) -> int:
        """clip_position: 1D :57-64, 2D :84-93 (both axes use plan_width)."""
        if v <= self.lo:
            return self.lo
        if v >= self.hi:
            return self.hi
        return v

    # -- step ---------------------------------------------------------------------
    def step(self, action: int, step_size: int):
        self.count_step += 1
        if self.dim == 1:
            return self._step_1d(int(action), int(step_size))
        if self.dim == 2:
            return self._step_2d(int(action), int(step_size))
        return self._step_3d(int(action), int(step_size))

    def _step_1d(self, a, s):
        """Env/1D/DMP_Env_1D_static.py:85-136 (dynamic: ...usedata_plan.py:71-120)."""
        if a == 0 or a == 1:
            self.pos = self._clamp(self.pos - s if a == 0 else self.pos + s)
            return self.obs(), 0, bool(self.count_step >= self.total_step)
        if a != 2:
            raise UnboundLocalError("action out of range (reference leaves 'position' unbound)")
        self.count_brick += 1
        self.grid[0, self.pos] += 1
        if self.count_brick >= self.total_brick:
            return self.obs(), 0.0, True
        h, p = self.grid[0, self.pos], self.plan[self.pos - self.HW]
        reward = -1.0 if h > p else (10.0 if h == p else 1.0)
        return self.obs(), reward, bool(self.count_step >= self.total_step)

    def _step_2d(self, a, s):
        """Env/2D/DMP_Env_2D_static.py:95-154 (dynamic: ...usedata_plan.py:85-147)."""
        r, c = self.pos
        if a in (0, 1, 2, 3):
            if a == 0:
                c -= s
            elif a == 1:
                c += s
            elif a == 2:
                r += s
            else:
                r -= s
            self.pos = [self._clamp(r), self._clamp(c)]
            return self.obs(), 0, bool(self.count_step >= self.total_step)
        if a != 4:
            raise UnboundLocalError("action out of range (reference leaves 'position' unbound)")
        self.count_brick += 1
        self.grid[r, c] += 1.0
        if self.count_brick >= self.total_brick:
            if self.grid[r, c] > 1:
                self.grid[r, c] = 1.0
            return self.obs(), 0.0, True
        done = bool(self.count_step >= self.total_step)
        v, p = self.grid[r, c], self.plan[r, c]
        if v > p:
            reward = 0
        elif v == p:
            reward = 5.0
        else:                                  # unreachable with {0,1} plans
            raise UnboundLocalError('reward')
        if v > 1.0:
            self.grid[r, c] = 1.0
        return self.obs(), reward, done

    # 3D helpers -------------------------------------------------------------------
    _NBR = ((0, -1), (0, 1), (1, 0), (-1, 0))          # L, R, U(row+1), D(row-1)

    def _check_sur(self):
        """Env/3D/DMP_simulator_3d_static_circle.py:88-102."""
        r, c = self.pos
        chk = [0] * 8
        for i, (dr, dc) in enumerate(self._NBR):
            v = self.grid[r + dr, c + dc]
            if v == -1:
                chk[i] = 1
                chk[i + 4] = 1
            elif v > 0:
                chk[i] = 1
        return chk

    def _walk(self, a, s):
        """move_step :104-134 -- consecutive empty (==0) cells in direction a, at most s."""
        r, c = self.pos
        dr, dc = self._NBR[a]
        n = 0
        for i in range(1, s + 1):
            if self.grid[r + dr * i, c + dc * i] == 0:
                n += 1
            else:
                break
        return n

    def _reward_3d(self, tr, tc):
        """reward_check :232-239."""
        v, p = self.grid[tr, tc], self.plan[tr, tc]
        return -1.0 if v > p else (10.0 if v == p else 1.0)

    def _step_3d(self, a, s):
        """static: Env/3D/DMP_simulator_3d_static_circle.py:153-230;
        dynamic: Env/3D/DMP_simulator_3d_dynamic_triangle_usedata.py:142-231."""
        if a < 0 or a > 7:
            # reference: 'elif action > 3' swallows a>7 as an unbuilt brick; a<0 falls to the
            # illegal-move branch.  Keep both behaviours.
            pass
        chk = self._check_sur()
        boxed = chk[0:4] == [1, 1, 1, 1]
        r, c = self.pos
        if 0 <= a <= 3 and chk[a] == 0:
            n = self._walk(a, s)
            dr, dc = self._NBR[a]
            self.pos = [self._clamp(r + dr * n), self._clamp(c + dc * n)]
        elif a > 3:
            built = False
            tr = tc = None
            if a <= 7 and chk[a] == 0:
                dr, dc = self._NBR[a - 4]
                tr, tc = r + dr, c + dc
                self.count_brick += 1
                self.grid[tr, tc] += 1.0
                built = True
            if self.dynamic:
                if self._check_sur()[0:4] == [1, 1, 1, 1]:        # re-check AFTER placement (:199-206)
                    return self.obs(), -100.0, True
                if self.count_brick >= self.total_brick:           # :207-213
                    return self.obs(), 0.0, True
                if built:                                          # :214-221
                    return self.obs(), self._reward_3d(tr, tc), False
            else:
                if bool(self.count_brick >= self.total_brick) or boxed:   # :210-215 (pre-placement check)
                    return self.obs(), 0.0, True
                if built:                                          # :217-221 (step limit NOT tested)
                    return self.obs(), self._reward_3d(tr, tc), False
        # blocked move / unbuilt brick / completed move: common tail
        if self.dynamic:
            done = bool(self.count_step >= self.total_step)                     # :226
        else:
            done = bool(self.count_step >= self.total_step) or boxed            # :226 (static)
        return self.obs(), 0.0, done

    # -- IoU ----------------------------------------------------------------------
    def iou(self) -> float:
        """1D: Env/1D/DMP_Env_1D_static.py:138-151.  2D: render :169-175 (bool AND / OR).
        3D: Env/3D/DMP_simulator_3d_static_circle.py:257-276."""
        h = self.HW
        if self.dim == 1:
            g = self.grid[0][h:h + self.W]
            a1, a2 = sum(self.plan), sum(g)
            over = 0
            for i in range(self.W):
                if g[i] > self.plan[i]:
                    over += g[i] - self.plan[i]
            cross = a2 - over
            return cross / (a1 + a2 - cross)
        p = self.plan[h:h + self.W, h:h + self.W]
        g = self.grid[h:h + self.W, h:h + self.W]
        if self.dim == 2:
            pb, gb = p.astype(bool), g.astype(bool)
            return (pb * gb).sum() / float((pb + gb).sum())
        cross = float(np.minimum(g, p).sum())
        return cross / (self.total_brick + self.count_brick - cross)


def make_env(dim: int, dynamic: bool, plan_choose: int = 0, plans=None, sequential=False) -> OracleEnv:
    if dynamic:
        assert plans is not None, "dynamic envs need a plan dataset"
        return OracleEnv(dim, True, plans, sequential=sequential)
    return OracleEnv(dim, False, [static_plan(dim, plan_choose)])
