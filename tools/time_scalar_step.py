#!/usr/bin/env python
"""Latency of the scalar drop-in classes (one env, reference-style loop: step, reset on done): microseconds per step(),
beside the unmodified reference classes on one host core where oracle/_ref is staged.

    python tools/time_scalar_step.py [steps]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 3000


def loop(env, A, n):
    rng = np.random.RandomState(1)
    acts = rng.randint(A, size=n)
    env.reset()
    t0 = time.perf_counter()
    for a in acts:
        _, _, d = env.step(int(a))
        if d:
            env.reset()
    return 1e6 * (time.perf_counter() - t0) / n


def main():
    import snac_b200 as S
    out = {}
    for name, cls, A in (("1D static", S.deep_mobile_printing_1d1r, 3), ("2D static", S.deep_mobile_printing_2d1r, 5),
                         ("3D static", S.deep_mobile_printing_3d1r, 8)):
        env = cls(plan_choose=0)
        loop(env, A, 200)
        out[name] = {"snac_b200_us_per_step": loop(env, A, N)}
    try:
        import bench
        if bench.reference_root():
            for name, wl, A in (("1D static", "1d_static_step", 3), ("2D static", "2d_static_dense", 5), ("3D static", "3d_static_dense", 8)):
                cls = bench._reference_class(wl)
                env = cls(plan_choose=bench.WORKLOADS[wl][2])
                loop(env, A, 200)
                out[name]["reference_us_per_step"] = loop(env, A, N)
    except Exception as e:
        out["reference"] = repr(e)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
