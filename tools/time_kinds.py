"""Kernel-only throughput of one rollout launch per observation kind (device buffers, CUDA events, median of 30):
    python tools/time_kinds.py [--dim 2] [--envs 1048576] [--steps 20]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from snac_b200.vecenv import BatchedDMPEnv        # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=2)
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    kinds = [torch.float32, torch.int16, "record"] + (["bits"] if a.dim > 1 else [])
    for kind in kinds:
        env = BatchedDMPEnv(a.dim, num_envs=a.envs, obs_dtype=kind, auto_reset=True)
        env.reset()
        K, n = a.steps, a.envs
        outs = [(torch.empty((K, n, env.obs_row), dtype=env.obs_dtype, device=env.device),
                 torch.empty((K, n), dtype=torch.float32, device=env.device),
                 torch.empty((K, n), dtype=torch.uint8, device=env.device)) for _ in range(4)]      # > L2 in total for f32
        for i in range(8):
            env.rollout(K, out=outs[i % 4])
        ts = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(30):
            e0.record()
            env.rollout(K, out=outs[i % 4])
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        row = env.obs_row * torch.empty(0, dtype=env.obs_dtype).element_size()
        print("%dD %-8s %3d B/env-step  %8.3f ms per %d-step launch  %7.2f G env-steps/s  %6.0f GB/s of results"
              % (a.dim, str(kind).replace("torch.", ""), row, ms, K, n * K / ms / 1e6, n * K * (row + 5) / ms / 1e6))
        env.check_errors()


if __name__ == "__main__":
    main()
