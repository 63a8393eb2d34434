#!/usr/bin/env python
"""Throughput of the device-resident acting loop (snac_b200/policy_loop.py) on one GPU.

    python tools/bench_policy_loop.py [--envs N] [--horizon T] [--reps R]

Prints one JSON line per policy: env-steps/s of `DeviceRollout.collect()` (CUDA-graph replay of T x (policy, dmp_step)),
timed with CUDA events after two warm-up collects.  The Q-network is the reference's critic
(script/DQN/2d/DQN_2d_static.py:78-98: 52 -> 64 -> 128 -> 128 -> 1, scored for all 5 actions per env)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from snac_b200.policy_loop import DeviceRollout, EpsilonGreedy, QSAAdapter, RandomPolicy  # noqa: E402
from snac_b200.vecenv import BatchedDMPEnv  # noqa: E402


class QNet(torch.nn.Module):
    def __init__(self, d):
        super().__init__()
        self.f = torch.nn.Sequential(torch.nn.Linear(d + 1, 64), torch.nn.ReLU(), torch.nn.Linear(64, 128), torch.nn.ReLU(),
                                     torch.nn.Linear(128, 128), torch.nn.ReLU(), torch.nn.Linear(128, 1))

    def forward(self, s, a):
        return self.f(torch.cat((s, a), dim=1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=2)
    ap.add_argument("--envs", type=int, default=1048576)
    ap.add_argument("--horizon", type=int, default=32)
    ap.add_argument("--reps", type=int, default=8)
    args = ap.parse_args()
    torch.manual_seed(0)
    A = {1: 3, 2: 5, 3: 8}[args.dim]
    D = {1: 7, 2: 51, 3: 51}[args.dim]
    q = QSAAdapter(QNet(D).cuda(), A)
    window_sum = lambda o: torch.remainder((o[:, :-2].sum(1) + o[:, -1]).to(torch.int64), A).to(torch.uint8)
    policies = [("kernel Philox actions (no policy)", None), ("torch.randint", RandomPolicy(A)),
                ("window-sum mod A (3 elementwise/reduction kernels)", window_sum),
                ("epsilon-greedy over the reference Q(s,a) MLP, fp32, all %d actions per env" % A, EpsilonGreedy(q, A, 0.2))]
    for name, pol in policies:
        n = args.envs if pol is None or not isinstance(pol, EpsilonGreedy) else min(args.envs, 262144)
        env = BatchedDMPEnv(args.dim, plan_choose=0, num_envs=n, auto_reset=True, reset_obs=True)
        env.reset()
        loop = DeviceRollout(env, pol, horizon=args.horizon)
        loop.collect(); loop.collect()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            loop.collect()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        env.check_errors()
        print(json.dumps({"policy": name, "dim": args.dim, "envs": n, "horizon": loop.T, "collects": args.reps,
                          "env_steps_per_s": n * loop.T * args.reps / (ms * 1e-3), "ms_per_step": ms / (loop.T * args.reps)}), flush=True)
        del loop, env


if __name__ == "__main__":
    main()
