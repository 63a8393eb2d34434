#!/usr/bin/env python
"""Times the UNMODIFIED reference env classes (imported from /root/reference through oracle/refload.py's gym /
matplotlib stubs) beside the python oracle port, single core, in the build container -- to show how the `port`
CPU baseline of bench.py relates to the real reference.  Cannot run on the GPU box (no /root/reference there).

    python tools/time_reference_here.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dmp_oracle as O  # noqa: E402
from oracle import refload  # noqa: E402

N = 60000
for dim, A in (("1D", 3), ("2D", 5), ("3D", 8)):
    cls = refload.load_class(dim, "static")
    env = cls(plan_choose=0)
    env.reset()                                  # static 2D/3D: create_plan() runs 1352 contains_point calls per reset
    rng = np.random.RandomState(1)
    acts = rng.randint(A, size=N)
    t0 = time.perf_counter()
    resets = 0
    t_reset = 0.0
    for a in acts:
        _, _, d = env.step(int(a))
        if d:
            r0 = time.perf_counter()
            env.reset()
            t_reset += time.perf_counter() - r0
            resets += 1
    t_ref = time.perf_counter() - t0
    o = O.make_env(int(dim[0]), False, plan_choose=0)
    o.reset(0)
    sizes = rng.randint(1, 4, size=N)
    t0 = time.perf_counter()
    for a, s in zip(acts, sizes):
        _, _, d = o.step(int(a), int(s))
        if d:
            o.reset(0)
    t_port = time.perf_counter() - t0
    print("%s static: reference %.2f us/step (%.2f without its %d resets), oracle port %.2f us/step  [1 core]"
          % (dim, 1e6 * t_ref / N, 1e6 * (t_ref - t_reset) / N, resets, 1e6 * t_port / N))
