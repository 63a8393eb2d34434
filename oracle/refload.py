"""Import the UNMODIFIED reference env classes (TEST INFRASTRUCTURE ONLY).

Default root: ``/root/reference`` (the build container).  The GPU box has no reference tree, so nothing on the
``-m gpu`` / smoke path may call this.  Used by ``tests/golden/make_golden.py`` to produce the committed golden
traces, by the container-only live cross-check in ``tests/test_oracle_vs_reference.py`` and -- after ``use_root()`` has
pointed it at the files unpacked from ``oracle/_ref`` (oracle/stage_ref.py) -- by the CPU arm of ``bench.py``.

Two third-party imports of the reference are not installed here and are stubbed in
``sys.modules`` before the import: ``gym`` (only ``gym.Env`` / ``gym.Wrapper`` base
classes are used) and ``matplotlib`` (``pyplot`` for render(); ``patches.CirclePolygon``
for the static 2D/3D plan -- replaced by the oracle's restated 20-gon, so the static
plan itself is NOT pinned by this loader; see oracle/dmp_oracle.py header).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("SNAC_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "Env", "1D"))


def use_root(path: str) -> None:
    """Import the reference from ``path`` instead (call before the first load_class / load_multiprocess)."""
    global REF_ROOT
    REF_ROOT = path


def _install_stubs():
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")

        class Env:                       # gym.Env: no behaviour used by the reference
            pass

        class Wrapper:                   # gym.Wrapper.__init__(env) stores env
            def __init__(self, env):
                self.env = env

        gym.Env, gym.Wrapper = Env, Wrapper
        gym.spaces = types.ModuleType("gym.spaces")

        class Discrete:                  # gym.spaces.Discrete(n): the *_MCTS classes only store it
            def __init__(self, n):
                self.n = n

        gym.spaces.Discrete = Discrete
        sys.modules["gym"] = gym
        sys.modules["gym.spaces"] = gym.spaces
    if "matplotlib" not in sys.modules:
        from oracle.dmp_oracle import _point_in_polygon, _polygon_vertices

        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        patches = types.ModuleType("matplotlib.patches")

        class CirclePolygon:             # restated polygon (PARITY UNPINNED for this class)
            def __init__(self, xy, radius=5, resolution=20, **kw):
                self.r = radius
                # centre is (12.5, 12.5) at every reference call site
                assert float(xy[0]) == 12.5 and float(xy[1]) == 12.5
                self.vx, self.vy = _polygon_vertices(radius, resolution)

            def contains_point(self, pt, radius=None):
                if self.r <= 0:
                    return False
                return _point_in_polygon(float(pt[0]), float(pt[1]), self.vx, self.vy)

        patches.CirclePolygon = CirclePolygon
        mpl.pyplot, mpl.patches = plt, patches
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
        sys.modules["matplotlib.patches"] = patches
        m3d = types.ModuleType("mpl_toolkits")
        m3d2 = types.ModuleType("mpl_toolkits.mplot3d")
        m3d2.Axes3D = object
        sys.modules.setdefault("mpl_toolkits", m3d)
        sys.modules.setdefault("mpl_toolkits.mplot3d", m3d2)


_MODULES = {
    # key: (subdir, module, class)
    ("1D", "static"): ("1D", "DMP_Env_1D_static", "deep_mobile_printing_1d1r"),
    ("1D", "dynamic"): ("1D", "DMP_Env_1D_dynamic_usedata_plan", "deep_mobile_printing_1d1r"),
    ("2D", "static"): ("2D", "DMP_Env_2D_static", "deep_mobile_printing_2d1r"),
    ("2D", "dynamic"): ("2D", "DMP_Env_2D_dynamic_usedata_plan", "deep_mobile_printing_2d1r"),
    ("3D", "static"): ("3D", "DMP_simulator_3d_static_circle", "deep_mobile_printing_3d1r"),
    ("3D", "dynamic"): ("3D", "DMP_simulator_3d_dynamic_triangle_usedata", "deep_mobile_printing_3d1r"),
    # observation-format variants used by the representation-learning baselines
    ("1D", "lnet"): ("1D", "DMP_Env_1D_static_Lnet", "deep_mobile_printing_1d1r"),
    ("2D", "lnet"): ("2D", "DMP_Env_2D_static_Lnet", "deep_mobile_printing_2d1r"),
    ("3D", "lnet"): ("3D", "DMP_simulator_3d_static_circle_Lnet", "deep_mobile_printing_3d1r"),
    # classes that carry the random plan generators
    ("1D", "hindsight_dynamic"): ("1D", "DMP_Env_1D_dynamic_hindsight_replay", "deep_mobile_printing_1d1r_hindsight"),
    ("2D", "hindsight_dynamic"): ("2D", "DMP_Env_2D_dynamic_hindsight_replay_usedata", "deep_mobile_printing_2d1r_hindsight"),
    ("3D", "hindsight_dynamic"): ("3D", "DMP_simulator_3d_dynamic_triangle_hindsight_replay", "deep_mobile_printing_3d1r_hindsight"),
    ("1D", "hindsight_static"): ("1D", "DMP_Env_1D_static_hindsight_replay", "deep_mobile_printing_1d1r_hindsight"),
    ("2D", "hindsight_static"): ("2D", "DMP_Env_2D_static_hindsight_replay", "deep_mobile_printing_2d1r_hindsight"),
    ("3D", "hindsight_static"): ("3D", "DMP_simulator_3d_static_circle_hindsight_replay", "deep_mobile_printing_3d1r_hindsight"),
    # tree-search variants: state tuples + functional transition(state, action)
    ("1D", "mcts_static"): ("1D", "DMP_Env_1D_static_MCTS", "deep_mobile_printing_1d1r_MCTS"),
    ("1D", "mcts_dynamic"): ("1D", "DMP_Env_1D_dynamic_MCTS", "deep_mobile_printing_1d1r_MCTS_obs"),
    ("2D", "mcts_static"): ("2D", "DMP_ENV_2D_static_MCTS", "deep_mobile_printing_2d1r_MCTS"),
    ("2D", "mcts_dynamic"): ("2D", "DMP_ENV_2D_dynamic_MCTS", "deep_mobile_printing_2d1r"),
    ("3D", "mcts_static"): ("3D", "DMP_simulator_3d_static_circle_MCTS", "deep_mobile_printing_3d1r"),
    ("3D", "mcts_dynamic"): ("3D", "DMP_simulator_3d_dynamic_triangle_MCTS", "deep_mobile_printing_3d1r"),
}


def load_class(dim: str, kind: str):
    """Return the unmodified reference class for e.g. ("2D", "static")."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_stubs()
    sub, mod, cls = _MODULES[(dim, kind)]
    path = os.path.join(REF_ROOT, "Env", sub)
    if path not in sys.path:
        sys.path.append(path)
    return getattr(importlib.import_module(mod), cls)


def load_multiprocess():
    """Import the reference's multiprocess.py (VectorizedEnvWrapper)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_stubs()
    for sub in ("1D", "2D", "3D"):
        p = os.path.join(REF_ROOT, "Env", sub)
        if p not in sys.path:
            sys.path.append(p)
    # loaded by path under a private name: a third-party package called ``multiprocess`` may be installed
    import importlib.util
    spec = importlib.util.spec_from_file_location("snac_reference_multiprocess", os.path.join(REF_ROOT, "multiprocess.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def dataset_path(dim: str, density: str = "dense", split: str = "train") -> str:
    if dim == "1D":
        return os.path.join(REF_ROOT, "Env/1D/data_1d_dynamic_sin_envplan_500_%s.pkl" % split)
    d = dim.lower()
    return os.path.join(REF_ROOT, "Env/%s/data_%s_dynamic_%s_envplan_500_%s.pkl" % (dim, d, density, split))


def load_dataset(dim: str, density: str = "dense", split: str = "train"):
    import joblib
    return [np.asarray(p, dtype=np.float64) for p in joblib.load(dataset_path(dim, density, split))]
