#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched SNAC simulator (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

A "step" is one pass of the hot path over the whole batch: every env of the vector env advances by
one step (dmp_step: action in, observation/reward/done out).  Workload = BASELINE.json configs[2]:
2D static dense plan, random-action rollouts, 1,048,576 envs in total, sharded over the ranks
(one independent slice per GPU, no data-path collective; one NCCL all-reduce of the 4-double
episode-statistics vector after the timed region).

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every key.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (dim, dynamic, plan_choose, density, default total envs, B_alg bytes per env-step (f32 obs, step mode))
    "1d_dynamic": (1, True, 0, "dense", 65536, 67),
    "1d_static_step": (1, False, 2, None, 65536, 64),
    "2d_static_dense": (2, False, 0, None, 1048576, 262),
    "2d_static_sparse": (2, False, 1, None, 1048576, 262),
    "2d_dynamic_dense": (2, True, 0, "dense", 1048576, 268),
    "3d_static_dense": (3, False, 0, None, 262144, 330),
    "3d_dynamic_dense": (3, True, 0, "dense", 262144, 336),
}
METRIC = "env-steps/sec"
L2_BYTES = 126 * 1024 * 1024


def load_plans_fixture(dim, density, split="train"):
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", "plans_packed.npz"))
    if dim == 1:
        return z["1d_%s" % split].astype(np.float64)
    a = z["%dd_%s_%s" % (dim, density, split)]
    bits = np.unpackbits(a, axis=1)[:, :400].reshape(len(a), 20, 20).astype(np.float64)
    out = np.zeros((len(a), 26, 26))
    out[:, 3:23, 3:23] = bits * (6.0 if dim == 3 else 1.0)
    return out


# ------------------------------------------------------------------------------------------------
# clocks sampler (NVML) -- runs during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port (python restatement of the reference algorithm), one process per core
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    wl, n_envs, n_steps, seed = args
    import numpy as np
    from oracle import dmp_oracle as O
    dim, dynamic, plan_choose, density, _, _ = WORKLOADS[wl]
    plans = load_plans_fixture(dim, density) if dynamic else None
    envs = [O.make_env(dim, dynamic, plan_choose=plan_choose, plans=plans) for _ in range(n_envs)]
    rng = np.random.RandomState(seed)
    for e in envs:
        e.reset(int(rng.randint(len(e.plans))))
    A = O.SPEC[dim]["actions"]
    t0 = time.perf_counter()
    for _ in range(n_steps):
        acts = rng.randint(A, size=n_envs)
        sizes = rng.randint(1, 4, size=n_envs)
        for i, e in enumerate(envs):
            _, _, d = e.step(int(acts[i]), int(sizes[i]))
            if d:
                e.reset(int(rng.randint(len(e.plans))))
    return n_envs * n_steps, time.perf_counter() - t0


def cpu_port_throughput(wl, n_envs_per_proc, n_steps, procs=None):
    """Aggregate env-steps/s of the oracle port over `procs` processes (default: all usable cores)."""
    import multiprocessing as mp
    cores = procs or len(os.sched_getaffinity(0))
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, [(wl, n_envs_per_proc, n_steps, 1000 + i) for i in range(cores)])
        wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    return total / max(max(r[1] for r in res), 1e-9), cores, total, wall


def cpu_c_port_throughput(wl, n_envs=65536, K=2000):
    """The same algorithm as a compiled, OpenMP-parallel C port (oracle/dmp_oracle.c) on all host cores:
    a much stronger CPU baseline than the python port; observations are materialised like on the GPU."""
    import numpy as np
    from oracle.c_oracle import COracleBatch
    dim, dynamic, plan_choose, density, _, _ = WORKLOADS[wl]
    plans = load_plans_fixture(dim, density) if dynamic else None
    cb = COracleBatch(dim, dynamic, n_envs, plan_choose, plans)
    rng = np.random.RandomState(5)
    cb.reset(rng.randint(cb.n_plans, size=n_envs).astype(np.int32) if dynamic else None)
    from oracle import dmp_oracle as O
    A = O.SPEC[dim]["actions"]
    acts = rng.randint(A, size=(K, n_envs)).astype(np.uint8)
    sizes = rng.randint(1, 4, size=(K, n_envs)).astype(np.uint8)
    nxt = rng.randint(cb.n_plans, size=(K, n_envs)).astype(np.int32) if dynamic else None
    cb.rollout(acts[:20], sizes[:20], None if nxt is None else nxt[:20], want_obs=False)     # warm-up
    t0 = time.perf_counter()
    cb.rollout(acts, sizes, nxt, want_obs=False)
    dt = time.perf_counter() - t0
    return {"value": n_envs * K / dt, "unit": "env-steps/s", "cores": len(os.sched_getaffinity(0)), "kind": "port-c",
            "sample": "%d envs x %d steps of %s, C/OpenMP oracle port, reset on done, no obs copy-out (%.2f s)" % (n_envs, K, wl, dt)}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port, kind "port": the reference is Python
    and cannot travel to the GPU box), all host cores, same metric/config; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    cores = len(os.sched_getaffinity(0))
    # bounded sample: about 1.5 M env-steps per core in total (~20-30 s of python per core)
    per_step = max(1, min(64, int(1.5e6 / max(args.steps + args.warmup, 1))))
    cpu_port_throughput(wl, per_step, max(args.warmup, 1), cores)                   # warm-up (untimed)
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(wl, per_step, args.steps, 2000 + i) for i in range(cores)])
    t = max(r[1] for r in res)
    total = sum(r[0] for r in res)
    value = total / t
    sample = "%d procs x %d envs x %d steps of %s (python oracle port of the reference algorithm, reset on done)" % (
        cores, per_step, args.steps, wl)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl, "envs_per_step": cores * per_step},
            "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    try:
        line["cpu_baseline_c"] = cpu_c_port_throughput(wl)
    except Exception as e:                                  # gcc missing etc.: report, do not fail the arm
        line["cpu_baseline_c"] = {"unavailable": repr(e)}
    print(json.dumps(line), flush=True)


def plan_launches(K, KL, G):
    """How K timed steps are executed with KL steps per launch and G launches per CUDA graph:
    (replays of the main graph, full launches in the tail graph, steps of the tail's short last launch or 0).
    n_replay * G * KL + n_tail * KL + rem == K exactly."""
    full, rem = divmod(K, KL)
    n_replay, n_tail = divmod(full, G)
    return n_replay, n_tail, rem


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from snac_b200.vecenv import BatchedDMPEnv

    wl = args.workload
    dim, dynamic, plan_choose, density, default_envs, b_alg_step = WORKLOADS[wl]
    state_bytes = {1: 72, 2: 64, 3: 1216}[dim]
    total_envs = args.envs or default_envs
    from snac_b200.sharding import shard_bounds
    env_base, n = shard_bounds(total_envs, rank, world)  # strong scaling: the batch is sharded
    env = BatchedDMPEnv(dim, dynamic=dynamic, plan_choose=plan_choose,
                        plans=load_plans_fixture(dim, density) if dynamic else None,
                        num_envs=n, device=dev, auto_reset=True, env_base=env_base,
                        obs_dtype=torch.float32, action_dist=args.action_dist)
    D, A = env.obs_dim, env.action_dim
    K, W = args.steps, args.warmup

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    RA = 64
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    act_pool = None
    if args.actions == "buffer":
        act_pool = torch.randint(0, A, (RA, n), dtype=torch.uint8, device=dev, generator=g)
    env.reset()
    torch.cuda.synchronize()
    stream = torch.cuda.Stream(device=dev)

    def timed_run(KL, K, W):
        """Time K vector steps executed as launches of KL steps each (KL = 1: step mode, one dmp_step per
        vector step; KL > 1: rollout mode, dmp_rollout keeps the state on chip for KL steps).  Returns
        (ms, steps actually timed, warm-up steps, launches, ring description, clocks)."""
        # rollout storage: a ring of [KL, n, D] observation buffers larger than 2 x L2, so stores reach HBM
        obs_bytes = KL * n * D * 4
        R = min(256, max(2, -(-2 * L2_BYTES // obs_bytes)))
        obs_ring = torch.empty((R, KL, n, D), dtype=torch.float32, device=dev)
        rew_ring = torch.empty((R, KL, n), dtype=torch.float32, device=dev)
        done_ring = torch.empty((R, KL, n), dtype=torch.uint8, device=dev)
        acts = None
        if act_pool is not None:
            acts = [act_pool[(torch.arange(KL) + 7 * j) % RA].contiguous() for j in range(8)]
        G = max(16 if KL > 1 else 32, R + (R & 1))            # launches per graph (even: t_dev slots alternate)

        def launch(i, kl=KL):
            a = None if acts is None else acts[i % 8][:kl]
            env.rollout(kl, actions=a, out=(obs_ring[i % R][:kl], rew_ring[i % R][:kl], done_ring[i % R][:kl]),
                        use_device_t=True, t_slot=i & 1)

        # EXACTLY K steps are timed: n_replay replays of a graph of G full launches, then one replay of a tail graph
        # holding the remaining full launches and (when KL does not divide K) one shorter launch
        n_replay, n_tail, rem = plan_launches(K, KL, G)
        with torch.cuda.stream(stream):
            for i in range(4):
                launch(i)                                    # sets func attributes before capture
            stream.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                for i in range(G):
                    launch(i)
            tail = None
            if n_tail or rem:
                tail = torch.cuda.CUDAGraph()
                with torch.cuda.graph(tail, stream=stream):
                    for i in range(n_tail):
                        launch(i)
                    if rem:
                        launch(n_tail, rem)
            per_replay = G * KL
            n_replay_w = max(1, -(-W // per_replay))
            for _ in range(n_replay_w):
                graph.replay()
            stream.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with ClockSampler(local) as clk:
                e0.record(stream)
                for _ in range(n_replay):
                    graph.replay()
                if tail is not None:
                    tail.replay()
                e1.record(stream)
                stream.synchronize()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            ms = e0.elapsed_time(e1)
            n_launch = n_replay * G + n_tail + (1 if rem else 0)
            if (n_tail + (1 if rem else 0)) & 1:             # keep the two device step-counter slots alternating
                launch(1)
                stream.synchronize()
        tm = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ring = "obs ring of %d x %.1f MB (> 2 x L2) + state %.1f MB" % (R, obs_bytes / 1e6, n * state_bytes / 1e6)
        del graph, tail, obs_ring, rew_ring, done_ring
        return float(tm.item()), K, n_replay_w * per_replay, n_launch, ring, clk.summary()

    def b_alg_of(KL, mean_len=None):
        """algorithmic bytes per env-step with fp32 observations.
        step mode: the table value (1D 64/67, 2D 262/268, 3D 330/336 + E); rollout mode:
        D*4 + 5 + (2*S + 8 + 2*grid_bytes)/K + E.  The action byte is dropped when actions are drawn in-kernel;
        E = 1600/L (IoU read + map clear per episode of mean length L) is counted for 3D only."""
        act = 0 if act_pool is None else 1
        E = (1600.0 / mean_len) if (dim == 3 and mean_len) else 0.0
        if KL == 1:
            return b_alg_step - 1 + act + E
        S, grid = {1: (5, 60), 2: (6, 50), 3: (6, 800)}[dim]
        return act + D * 4 + 5 + (2 * S + 8 + 2 * grid) / KL + E

    def b_layout_of(KL, mean_len=None):
        """The same formula with the bytes of THIS repo's 3D layout (byte maps: window 49 x 1 B + 1 B brick, 400 B maps,
        E = 800/L): what a kernel on this layout has to move at least.  1D/2D: identical to b_alg_of."""
        if dim != 3:
            return b_alg_of(KL, mean_len)
        act = 0 if act_pool is None else 1
        E = (800.0 / mean_len) if mean_len else 0.0
        if KL == 1:
            return b_alg_step - 50 - 1 + act + E
        return act + D * 4 + 5 + (2 * 6 + 8 + 2 * 400) / KL + E

    KL = 1 if args.mode == "step" else args.rollout_k
    ms, K_eff, W_eff, launches, ring, clocks = timed_run(KL, K, W)
    other = None
    if not args.single_mode:
        oKL = args.rollout_k if KL == 1 else 1
        oms, oK, _, olaunch, oring, _ = timed_run(oKL, max(K // 4, 256), max(W // 4, 64))
        other = {"mode": "rollout K=%d" % oKL if oKL > 1 else "step", "ms_per_step": oms / oK,
                 "value": total_envs * oK / (oms * 1e-3), "steps": oK, "gpu_launches": olaunch,
                 }
    stats = env.stats(allreduce=True).cpu().numpy()          # NCCL all-reduce of the 4-double stats vector
    env.check_errors()
    mean_len = float(stats[3] / stats[2]) if stats[2] > 0 else None
    b_alg = b_alg_of(KL, mean_len)
    if other is not None:
        ob = b_alg_of(1 if KL > 1 else args.rollout_k, mean_len)
        other["bytes_per_env_step"] = ob
        other["roofline_frac"] = (other["value"] / world) * ob / 1e9 / peak
        other["roofline_frac_this_layout"] = (other["value"] / world) * b_layout_of(1 if KL > 1 else args.rollout_k, mean_len) / 1e9 / peak
    value = total_envs * K_eff / (ms * 1e-3)
    per_gpu_steps_s = n * K_eff / (ms * 1e-3)
    achieved = per_gpu_steps_s * b_alg / 1e9
    b_lay = b_layout_of(KL, mean_len)

    # ---- e2e: the reference-facing call with HOST buffers (actions in, obs/reward/done out) -------
    from snac_b200.compat import HostStepper
    if args.no_e2e:                                          # kernel A/B runs only; the default run always measures e2e
        args.e2e_steps, args.no_e2e_i16 = 4, True
    hs = HostStepper(env)
    Ke = max(4, min(K_eff, args.e2e_steps))
    rng = np.random.RandomState(99 + rank)
    host_actions = [rng.randint(0, A, size=n).astype(np.uint8) for _ in range(4)]
    for i in range(3):
        hs.step(host_actions[i % 4])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(Ke):
        o, r, d = hs.step(host_actions[i % 4])
    torch.cuda.synchronize()
    te = time.perf_counter() - t0
    tt = torch.tensor([te], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    te = float(tt.item())
    e2e = {"value": total_envs * Ke / te, "unit": "env-steps/s", "steps": Ke,
           "h2d_bytes_per_step": int(hs.h2d_bytes) * world, "d2h_bytes_per_step": int(hs.d2h_bytes) * world,
           "api": "HostStepper.step(actions: np.uint8[N]) -> (obs f32[N,%d], reward f32[N], done bool[N]) numpy, pinned" % D}

    # the same call with int16 observations (every raw observation value is a small integer, so i16 is exact;
    # SURVEY.md 8(d) lists it as the compact obs_t): half the D2H bytes of the PCIe-bound f32 call
    e2e_i16 = None
    if not dynamic and not args.no_e2e_i16:
        env16 = BatchedDMPEnv(dim, dynamic=dynamic, plan_choose=plan_choose, plans=None, num_envs=n, device=dev,
                              auto_reset=True, env_base=env_base, obs_dtype=torch.int16)
        env16.reset()
        hs16 = HostStepper(env16)
        for i in range(3):
            hs16.step(host_actions[i % 4])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(Ke):
            hs16.step(host_actions[i % 4])
        torch.cuda.synchronize()
        t16 = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t16, op=dist.ReduceOp.MAX)
        e2e_i16 = {"value": total_envs * Ke / float(t16.item()), "unit": "env-steps/s", "steps": Ke,
                   "h2d_bytes_per_step": int(hs16.h2d_bytes) * world, "d2h_bytes_per_step": int(hs16.d2h_bytes) * world,
                   "api": "HostStepper.step on an env built with obs_dtype=int16"}
        del hs16, env16

    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("%s|%d|K=%d" % (wl, n, KL))
    except Exception:
        pass
    line = {"metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K_eff, "warmup": W_eff,
            "ms_per_step": ms / K_eff, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32 bit-grid / i32 counters, f32 observations", "data": "synthetic",
            "config": {"workload": wl, "total_envs": total_envs, "envs_per_gpu": n, "mode": ("step: one dmp_step launch per vector step" if KL == 1 else "rollout: dmp_rollout, %d steps per launch, every step's obs/reward/done materialised" % KL) + ", CUDA-graph replay",
                       "actions": (("uniform over %d actions, " % A) if args.action_dist == "uniform" else
                                   "the reference's p = [.2, .2, .2, .2, .05, .05, .05, .05] (Env/3D/DMP_simulator_3d_static_circle.py:361-362), ") + ("pool of %d pre-generated vectors in HBM" % RA if act_pool is not None
                                                                          else "Philox4x32-10 in-kernel, counter (global env id, step)"),
                       "step_size": "Philox4x32-10 in-kernel", "auto_reset": True,
                       "l2": ring,
                       "parallelism": "env-sharded x%d, NCCL all-reduce of episode stats only" % world},
            "clocks": clocks,
            "e2e": e2e, "e2e_i16": e2e_i16, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_note": "bytes per launch from the committed ncu capture (profiles/), not measured live; algorithmic bytes per launch = %d" % int(b_alg * n * KL),
                         "peak_source": peak_src, "kernel": ("k3d_step_bytes<float> (K=1)" if (dim == 3 and KL == 1) else
                                    "k%dd%s_rollout<float> (K=%d)" % (dim, "_cache" if dim == 3 else "", KL)),
                         "bytes_per_env_step": b_alg, "envs_per_launch": n,
                         "bytes_per_env_step_this_layout": b_lay, "frac_this_layout": per_gpu_steps_s * b_lay / 1e9 / peak,
                         "layout_note": None if dim != 3 else "frac uses SURVEY 8(d)'s canonical u16 maps (800 B, window 98 B, E = 1600/L); this repo's 3D state is byte maps (400 B, window 49 B, E = 800/L): frac_this_layout is the fraction by those bytes"},
            "episode_stats": {"mean_episode_length": mean_len, "mean_iou": (float(stats[1] / stats[2]) if stats[2] > 0 else None),
                              "mean_return": (float(stats[0] / stats[2]) if stats[2] > 0 else None),
                              "sum_return": float(stats[0]), "sum_iou": float(stats[1]), "episodes": float(stats[2]),
                              "steps": float(stats[3])}}
    if other is not None:
        line["other_mode"] = other
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, total, wall = cpu_port_throughput(wl, 64, 20000)
        line["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
                                "sample": "%d procs x 64 envs x 20000 steps of %s, python oracle port, reset on done (%.1f s wall)" % (cores, wl, wall)}
        try:
            line["cpu_baseline_c"] = cpu_c_port_throughput(wl)
        except Exception as e:
            line["cpu_baseline_c"] = {"unavailable": repr(e)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16384)
    ap.add_argument("--warmup", type=int, default=1024)
    ap.add_argument("--actions", default="philox", choices=["philox", "buffer"])
    ap.add_argument("--mode", default="rollout", choices=["rollout", "step"],
                    help="rollout: K steps per launch (dmp_rollout); step: one launch per vector step (dmp_step)")
    ap.add_argument("--rollout-k", type=int, default=0,
                    help="steps per dmp_rollout launch (default: 16; 64 for the 1D workloads, whose 65 536-env launches "
                         "are otherwise dominated by launch latency)")
    ap.add_argument("--single-mode", action="store_true", help="skip the secondary measurement of the other mode")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="2d_static_dense", choices=sorted(WORKLOADS))
    ap.add_argument("--action-dist", default="uniform", choices=["uniform", "ref3d"],
                    help="in-kernel action distribution: uniform over the env's actions, or (3D) the reference's own "
                         "p = [.2, .2, .2, .2, .05, .05, .05, .05] (SURVEY.md 8(d) cfg 5)")
    ap.add_argument("--envs", type=int, default=0, help="total envs over all GPUs (default: the BASELINE config's)")
    ap.add_argument("--e2e-steps", type=int, default=48)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e-i16", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="kernel A/B runs: 4 e2e steps only, no int16 e2e")
    args = ap.parse_args()
    if not args.rollout_k:
        args.rollout_k = 64 if WORKLOADS[args.workload][0] == 1 else 16
    if args.action_dist == "ref3d" and (WORKLOADS[args.workload][0] != 3 or args.actions != "philox"):
        ap.error("--action-dist ref3d needs a 3D workload and in-kernel (philox) actions")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
