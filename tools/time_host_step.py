"""Where a host-buffer step of the compact kinds spends its time (1 GPU): the pieces of HostStepper.step timed one by one.
    python tools/time_host_step.py [--envs 1048576] [--kind bits]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from snac_b200.compat import HostStepper          # noqa: E402
from snac_b200.vecenv import BatchedDMPEnv        # noqa: E402


def ev_time(fn, reps=50):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))


def wall(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e6)
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--kind", default="bits")
    ap.add_argument("--dim", type=int, default=2)
    a = ap.parse_args()
    n = a.envs
    out = {}
    for mapped in (False, True, "out"):
        env = BatchedDMPEnv(a.dim, num_envs=n, obs_dtype=a.kind, auto_reset=True)
        env.reset()
        hs = HostStepper(env, mapped=mapped)
        acts = np.random.randint(0, env.action_dim, size=n).astype(np.uint8)
        tag = {False: "staged", True: "mapped", "out": "mapped_out"}[mapped]
        out[tag + " step(actions) wall us"] = wall(lambda: hs.step(acts))
        out[tag + " step(actions_buffer) wall us"] = wall(lambda: hs.step(hs.actions_buffer))
        st, io = env._st, hs._ios[0][0]
        from snac_b200 import _lib as L
        import ctypes as C
        s = torch.cuda.current_stream().cuda_stream

        def kern():
            L.lib.dmp_rollout(C.byref(st), C.byref(io), 1, s)
        out[tag + " kernel alone (events) us"] = ev_time(kern)
        if mapped is False:
            out["numpy copy of actions us"] = wall(lambda: np.copyto(hs.actions_buffer, acts))
            out["H2D actions (events) us"] = ev_time(lambda: hs._in_dev[:n].copy_(hs._in_pin[:n], non_blocking=True))
            out["D2H result (events) us"] = ev_time(lambda: hs._res_pin[0].copy_(hs._res_dev, non_blocking=True))
            out["empty sync wall us"] = wall(lambda: torch.cuda.current_stream().synchronize())
            out["result bytes"] = hs.d2h_bytes
    for k, v in out.items():
        print("%-44s %10.1f" % (k, v))


if __name__ == "__main__":
    main()
