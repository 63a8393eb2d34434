"""Shared description of tests/golden/mcts_golden.npz (written by tests/golden/make_mcts_golden.py)."""
import os

import numpy as np

from conftest import GOLDEN, load_plans

# name: (dim, dynamic, plan_choose, (density, split))
MCTS_CASES = {
    "1d_static_p0": (1, False, 0, None),
    "1d_static_p2": (1, False, 2, None),
    "1d_dynamic": (1, True, 0, ("dense", "test")),
    "2d_static_dense": (2, False, 0, None),
    "2d_static_sparse": (2, False, 1, None),
    "2d_dynamic_dense": (2, True, 0, ("dense", "val")),
    "3d_static_dense": (3, False, 0, None),
    "3d_static_sparse": (3, False, 1, None),
    "3d_dynamic_dense": (3, True, 0, ("dense", "test")),
    "3d_dynamic_sparse": (3, True, 0, ("sparse", "val")),
}


def load_mcts_case(name):
    z = np.load(os.path.join(GOLDEN, "mcts_golden.npz"))
    pre = name + "/"
    g = {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}
    dim, dynamic, plan_choose, ds = MCTS_CASES[name]
    plans = load_plans(dim, ds[0], ds[1]) if dynamic else None
    return g, dim, dynamic, plan_choose, plans
