#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched SNAC simulator (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU classes (oracle/_ref), host cores

A "step" is one pass of the hot path over the whole batch: every env of the vector env advances by one step (action in,
observation / reward / done out).  Workload = BASELINE.json configs[2]: 2D static dense plan, random-action rollouts,
1,048,576 envs in total, sharded over the ranks (one independent slice per GPU, no data-path collective; one NCCL
all-reduce of the 4-double episode-statistics vector, timed separately).

Timed region = EXACTLY --steps steps: one dmp_rollout(K = steps) launch when steps <= 64 (else launches of --rollout-k
steps), inside a CUDA graph whose first and last nodes are the two timing events, so no host latency sits between the
events.  The region is repeated (`repeats` in the line, >= 30) and the MEDIAN region time is reported, max over ranks.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every key.
"""
from __future__ import annotations

import argparse
import contextlib
import io
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
KL_MAX = 64                      # most steps one dmp_rollout launch of the timed region carries


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
# clocks sampler (NVML) -- runs during the timed regions
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index, period_s=0.002):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.period = period_s
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
            self._stop.wait(self.period)

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
# CPU arms.  All of this is BASELINE infrastructure: it executes oracle/ (the staged reference classes, or the oracle
# port where no archive was staged) on the host cores and never touches snac_b200/.
# ------------------------------------------------------------------------------------------------
_REF_ROOT = None


def reference_root():
    """Directory holding the unmodified reference hot-path files, unpacked from oracle/_ref (None: nothing staged)."""
    global _REF_ROOT
    if _REF_ROOT is None:
        try:
            from oracle import refload, stage_ref
            if not stage_ref.staged() and stage_ref.source_available():
                stage_ref.stage()                             # build container: (re)create the archive
            if stage_ref.staged():
                _REF_ROOT = stage_ref.unpack()
                refload.use_root(_REF_ROOT)
        except Exception as e:                                # pragma: no cover - reported in the line
            sys.stderr.write("reference archive unusable: %r\n" % (e,))
            _REF_ROOT = False
        if _REF_ROOT is None:
            _REF_ROOT = False
    return _REF_ROOT or None


def _reference_class(wl, cache_static_plan=True):
    """The unmodified reference class of a workload.  Static 2D/3D classes re-run create_plan() (1 352 point-in-polygon
    tests, ~5 ms) in every reset(); with cache_static_plan the METHOD is memoised at run time (the source is untouched) so
    that the step loop, not the plan generator, is what the arm times -- the uncached reset cost is reported beside it."""
    from oracle import refload
    dim, dynamic, plan_choose, density, _, _ = WORKLOADS[wl]
    cls = refload.load_class("%dD" % dim, "dynamic" if dynamic else "static")
    if not dynamic and cache_static_plan and not getattr(cls, "_snac_bench_cached", False):
        orig, memo = cls.create_plan, {}

        def create_plan(self):
            if self.plan_choose not in memo:
                memo[self.plan_choose] = orig(self)
            plan, area = memo[self.plan_choose]
            return plan.copy(), area

        cls.create_plan, cls._snac_bench_orig_create_plan, cls._snac_bench_cached = create_plan, orig, True
    return cls


def _ref_worker(args):
    """n_envs independent reference env objects, random actions over the env's full action set, reset on done."""
    wl, n_envs, n_steps, n_warm, seed = args
    import numpy as np
    from oracle import refload
    dim, dynamic, plan_choose, density, _, _ = WORKLOADS[wl]
    cls = _reference_class(wl)
    np.random.seed(seed)
    if dynamic:
        path = refload.dataset_path("%dD" % dim, density, "train")
        envs = [cls(data_path=path, random_choose_paln=True) for _ in range(n_envs)]
    else:
        envs = [cls(plan_choose=plan_choose) for _ in range(n_envs)]
    for e in envs:
        e.reset()
    A = envs[0].action_dim
    rng = np.random.RandomState(seed + 1)
    t_run = 0.0
    for t in range(n_warm + n_steps):
        acts = rng.randint(A, size=n_envs)
        t0 = time.perf_counter()
        for i, e in enumerate(envs):
            _, _, d = e.step(int(acts[i]))
            if d:
                e.reset()
        if t >= n_warm:
            t_run += time.perf_counter() - t0
    t_reset = None
    if not dynamic:                                          # what one reset() costs with the reference's own create_plan()
        orig = cls._snac_bench_orig_create_plan
        t0 = time.perf_counter()
        orig(envs[0])
        t_reset = time.perf_counter() - t0
    return n_envs * n_steps, t_run, t_reset


def _port_worker(args):
    wl, n_envs, n_steps, n_warm, seed = args
    import numpy as np
    from oracle import dmp_oracle as O
    dim, dynamic, plan_choose, density, _, _ = WORKLOADS[wl]
    plans = load_plans_fixture(dim, density) if dynamic else None
    envs = [O.make_env(dim, dynamic, plan_choose=plan_choose, plans=plans) for _ in range(n_envs)]
    rng = np.random.RandomState(seed)
    for e in envs:
        e.reset(int(rng.randint(len(e.plans))))
    A = O.SPEC[dim]["actions"]
    t_run = 0.0
    for t in range(n_warm + n_steps):
        acts = rng.randint(A, size=n_envs)
        sizes = rng.randint(1, 4, size=n_envs)
        t0 = time.perf_counter()
        for i, e in enumerate(envs):
            _, _, d = e.step(int(acts[i]), int(sizes[i]))
            if d:
                e.reset(int(rng.randint(len(e.plans))))
        if t >= n_warm:
            t_run += time.perf_counter() - t0
    return n_envs * n_steps, t_run, None


def cpu_throughput(wl, envs_per_proc, n_steps, n_warm, kind=None, procs=None):
    """Aggregate env-steps/s over one process per usable core.  kind "reference": the unmodified reference classes from
    oracle/_ref; "port": the python oracle port; None: the reference when staged.  Each "step" advances every env of
    every process once; throughput = all env-steps / the slowest process' stepping time."""
    import multiprocessing as mp
    if kind is None:
        kind = "reference" if reference_root() else "port"
    if kind == "reference" and not reference_root():
        raise RuntimeError("no staged reference archive (oracle/_ref)")
    cores = procs or len(os.sched_getaffinity(0))
    worker = _ref_worker if kind == "reference" else _port_worker
    if kind == "reference":
        _reference_class(wl)                                 # import once in the parent; the forked workers inherit it
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(worker, [(wl, envs_per_proc, n_steps, n_warm, 2000 + 17 * i) for i in range(cores)])
    wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    t = max(r[1] for r in res)
    resets = [r[2] for r in res if r[2] is not None]
    return {"value": total / max(t, 1e-9), "unit": "env-steps/s", "cores": cores, "kind": kind,
            "seconds_stepping": t, "seconds_wall": wall, "env_steps": total,
            "uncached_reset_ms": (1e3 * statistics.median(resets)) if resets else None,
            "sample": "%d procs x %d envs x %d steps of %s (%s, random actions over the full action set, reset on done%s)" % (
                cores, envs_per_proc, n_steps, wl,
                "UNMODIFIED reference classes from oracle/_ref through gym/matplotlib stubs" if kind == "reference"
                else "python oracle port of the reference algorithm",
                "; static plan memoised at run time, the reference's own reset() re-runs create_plan()" if resets else "")}


def cpu_c_port_throughput(wl, n_envs=65536, K=2000):
    """The same algorithm as a compiled, OpenMP-parallel C port (oracle/dmp_oracle.c) on all host cores:
    a much stronger CPU baseline than the python classes; no observation copy-out."""
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


def config1_cpu():
    """BASELINE config 1 on the host: `multiprocess.py --env 1DStatic --plan_type 2 --num_envs 5`, verbatim (one shared
    env object behind the five slots, actions 0..2, no reset on done: 750 iterations = 3 750 env.step calls) and fixed
    (five independent env objects, reset on done).  One core, like the reference."""
    import numpy as np
    out = {}
    kind = "reference" if reference_root() else "port"
    out["kind"] = kind
    if kind == "reference":
        from oracle import refload
        mod = refload.load_multiprocess()
        ns = argparse.Namespace(env="1DStatic", plan_type=2, num_envs=5)
        np.random.seed(7)
        sink = io.StringIO()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(sink):
            mod.main(ns)
        dt = time.perf_counter() - t0
        out["verbatim"] = {"env_steps_per_s": 3750 / dt, "seconds": dt, "printed": sink.getvalue().split(),
                           "what": "reference multiprocess.main() as is (shared env object, no reset on done)"}
        cls = refload.load_class("1D", "static")
        envs = [cls(plan_choose=2) for _ in range(5)]
        step = lambda e, a: e.step(a)
        reset = lambda e: e.reset()
    else:
        from oracle import dmp_oracle as O
        envs = [O.make_env(1, False, plan_choose=2) for _ in range(5)]
        rs = np.random.RandomState(3)
        step = lambda e, a: e.step(a, int(rs.randint(1, 4)))
        reset = lambda e: e.reset(0)
    np.random.seed(7)
    for e in envs:
        reset(e)
    t0 = time.perf_counter()
    for t in range(750):
        acts = np.random.randint(3, size=5)
        for e, a in zip(envs, acts):
            _, _, d = step(e, int(a))
            if d:
                reset(e)
    dt = time.perf_counter() - t0
    out["fixed"] = {"env_steps_per_s": 3750 / dt, "seconds": dt,
                    "what": "five independent env objects, reset on done, same action stream"}
    return out


def config1_ours(dev):
    """The drop-in for config 1: snac_b200.multiprocess's main() with the same flags (five device-resident envs, one
    launch + one D2H copy per vector step).  Five envs cannot fill a GPU: this is a latency number."""
    import numpy as np
    import torch
    from snac_b200 import compat
    np.random.seed(7)
    sink = io.StringIO()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(sink):
        compat.main(["--env", "1DStatic", "--plan_type", "2", "--num_envs", "5"])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"env_steps_per_s": 3750 / dt, "seconds": dt, "printed": sink.getvalue().split(),
            "what": "python -m snac_b200.multiprocess --env 1DStatic --plan_type 2 --num_envs 5 (main() timed in-process: "
                    "env construction + reset + 750 VectorizedEnvWrapper.step calls with numpy in / numpy out)"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores -- the UNMODIFIED env classes
    staged in oracle/_ref (kind "reference"); the python oracle port only where no archive was staged (kind "port").
    Same metric / unit / workload as our arm; a step advances a bounded batch (cores x envs_per_proc envs) once.  Rank 0
    only; the other ranks exit without work."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))     # torchrun pins it to 1; the C port uses OpenMP
    wl = args.workload
    dim, dynamic, plan_choose, density, default_envs, _ = WORKLOADS[wl]
    per_proc = args.ref_envs_per_proc or (256 if dynamic else 2048)
    steps, warm = max(args.steps, 1), max(args.warmup, 0)
    r = cpu_throughput(wl, per_proc, steps, warm)
    batch = r["cores"] * per_proc
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "env-steps/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": 1e3 * r["seconds_stepping"] / steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl, "total_envs": args.envs or default_envs, "envs_per_step_sampled": batch,
                       "note": "CPU throughput does not depend on the batch size (a python loop over env objects): the arm "
                               "steps a bounded sample of %d envs per step instead of the GPU arm's %d" % (batch, args.envs or default_envs),
                       "actions": "uniform over the env's actions (numpy)", "step_size": "the reference's own np.random.randint(1, 4)",
                       "auto_reset": "reset() on done"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "uncached_reset_ms")},
            "e2e": {"value": r["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if r["kind"] == "reference":
        try:
            p = cpu_throughput(wl, 64, max(steps, 200), 10, kind="port")
            line["cpu_baseline_port"] = {k: p[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:
            line["cpu_baseline_port"] = {"unavailable": repr(e)}
    try:
        line["cpu_baseline_c"] = cpu_c_port_throughput(wl)
    except Exception as e:                                  # gcc missing etc.: report, do not fail the arm
        line["cpu_baseline_c"] = {"unavailable": repr(e)}
    print(json.dumps(line), flush=True)


def plan_launches(K, KL):
    """Steps per launch of a K-step timed region with at most KL steps per launch: full launches, then one shorter one.
    sum(result) == K exactly."""
    full, rem = divmod(K, KL)
    return [KL] * full + ([rem] if rem else [])


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, self.world))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.dist = dist
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        from snac_b200.compat import bind_to_gpu_numa_node
        self.cpus = bind_to_gpu_numa_node(self.dev)          # pinned staging buffers land on the GPU's NUMA node
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            self.peak, self.peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        else:
            self.peak, self.peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        self.stream = torch.cuda.Stream(device=self.dev)

    # -------------------------------------------------------------------------------------------
    def make_env(self, wl, total_envs, obs_dtype=None, action_dist="uniform"):
        import torch
        from snac_b200.sharding import shard_bounds
        from snac_b200.vecenv import BatchedDMPEnv
        dim, dynamic, plan_choose, density, _, _ = WORKLOADS[wl]
        env_base, n = shard_bounds(total_envs, self.rank, self.world)   # strong scaling: the batch is sharded
        env = BatchedDMPEnv(dim, dynamic=dynamic, plan_choose=plan_choose,
                            plans=load_plans_fixture(dim, density) if dynamic else None,
                            num_envs=n, device=self.dev, auto_reset=True, env_base=env_base,
                            obs_dtype=torch.float32 if obs_dtype is None else obs_dtype, action_dist=action_dist)
        env.reset()
        return env

    def max_over_ranks(self, x):
        import torch
        if self.world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def barrier(self):
        import torch
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    # -------------------------------------------------------------------------------------------
    def time_regions(self, env, K, KL, W, act_pool=None, min_repeats=30, target_s=0.25, max_repeats=1500):
        """Time a region of EXACTLY K vector steps, executed as launches of at most KL steps (KL = 1: step mode, one
        dmp_step per vector step; KL > 1: rollout mode, the state stays on chip inside a launch), `repeats` times.
        Each region is one CUDA graph: [event e0] -> launches -> [event e1]; the events are graph nodes, so what lies
        between them is device time only.  Observations go to a ring of [K, n, D] buffers larger than 2 x L2 in total
        (stores reach HBM).  Warm-up: exactly W steps, untimed, through the same launches.  Returns a dict."""
        import torch
        n, D = env.num_envs, env.obs_row
        dev, stream = self.dev, self.stream
        esz = torch.empty(0, dtype=env.obs_dtype).element_size()
        launches = plan_launches(K, KL)
        kb = max(launches)                                   # steps per buffer = the longest launch
        obs_bytes = kb * n * D * esz
        nbuf = max(2, -(-2 * L2_BYTES // obs_bytes))         # launches cycle through > 2 x L2 of observation buffers
        n_graphs = 2                                         # even: the two device step-counter slots alternate
        nbuf = max(2, min(nbuf + (nbuf & 1), n_graphs * len(launches)))     # no more buffers than launches that use them
        bufs = [(torch.empty((kb, n, D), dtype=env.obs_dtype, device=dev), torch.empty((kb, n), dtype=torch.float32, device=dev),
                 torch.empty((kb, n), dtype=torch.uint8, device=dev)) for _ in range(nbuf)]
        RA = 0 if act_pool is None else act_pool.shape[0]
        acts = None
        if act_pool is not None:
            acts = [act_pool[(torch.arange(kb) + 7 * j) % RA].contiguous() for j in range(8)]
        count = [0]                                          # launches issued so far (slot parity of the device counter)

        def launch(kl):
            i = count[0]
            o, r, d = bufs[i % nbuf]
            a = None if acts is None else acts[i % 8][:kl]
            env.rollout(kl, actions=a, out=(o[:kl], r[:kl], d[:kl]), use_device_t=True, t_slot=i & 1)
            count[0] += 1

        def region():
            for kl in launches:
                launch(kl)

        t_start = env.t
        with torch.cuda.stream(stream):
            env._t_dev.fill_(t_start)
            # ---- warm-up: exactly W steps (also sets the kernels' function attributes before any capture) ----------
            done_w = 0
            while done_w < W:
                kl = min(kb, W - done_w)
                launch(kl)
                done_w += kl
            if count[0] & 1:                                 # keep the slot parity of the graphs independent of W
                env._t_dev[0] = env._t_dev[1]
                count[0] += 1
            stream.synchronize()
            base = count[0]
            graphs, events, in_graph = [], [], True
            try:
                if self.args.timing == "stream":
                    raise RuntimeError("--timing stream")
                for g in range(n_graphs):
                    count[0] = base + g * len(launches)
                    e0 = torch.cuda.Event(enable_timing=True, external=True)
                    e1 = torch.cuda.Event(enable_timing=True, external=True)
                    gr = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gr, stream=stream):
                        e0.record(stream)
                        region()
                        e1.record(stream)
                    graphs.append(gr)
                    events.append((e0, e1))
            except Exception as ex:                          # event nodes unsupported: events around the replay instead
                sys.stderr.write("in-graph timing events unavailable (%r): stream events after a spin kernel\n" % (ex,))
                in_graph, graphs, events = False, [], []
                for g in range(n_graphs):
                    count[0] = base + g * len(launches)
                    gr = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gr, stream=stream):
                        region()
                    graphs.append(gr)
                    events.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))

            def one(r):
                g = r % n_graphs
                e0, e1 = events[g]
                if in_graph:
                    graphs[g].replay()
                else:
                    torch.cuda._sleep(200000)               # the host enqueues e0 / graph / e1 while this spins
                    e0.record(stream)
                    graphs[g].replay()
                    e1.record(stream)
                stream.synchronize()
                return e0.elapsed_time(e1)

            self.barrier()
            probe = [one(r) for r in range(n_graphs)]        # untimed: first replays upload the graphs
            est = max(statistics.median(probe), 1e-3)
            R = int(min(max_repeats, max(min_repeats, target_s * 1e3 / est)))
            R += R & 1                                       # both graphs equally often: the step-counter slots stay in phase
            self.barrier()
            with ClockSampler(self.local) as clk:
                ms = [one(r) for r in range(R)]
            self.barrier()
        steps_run = W + (n_graphs + R) * K
        env._st.t = t_start + steps_run                      # host-side step counter follows the device's
        med = self.max_over_ranks(statistics.median(ms))
        res = {"ms": med, "ms_min": min(ms), "ms_max": max(ms), "ms_mean": statistics.fmean(ms), "repeats": R, "steps": K,
               "warmup": W, "launches_per_region": len(launches), "steps_per_launch": launches,
               "timing": "CUDA events as first / last node of each region's graph" if in_graph else "stream events after a spin kernel",
               "ring": "launches cycle through %d x [%d, %d, %d] obs buffers = %d x %.1f MB (> 2 x L2 in total)" % (nbuf, kb, n, D, nbuf, obs_bytes / 1e6),
               "clocks": clk.summary()}
        del graphs, bufs, acts
        return res

    # -------------------------------------------------------------------------------------------
    @staticmethod
    def b_alg(dim, b_alg_step, D, KL, philox_actions, mean_len=None):
        """algorithmic bytes per env-step with fp32 observations.
        step mode: the table value (1D 64/67, 2D 262/268, 3D 330/336 + E); rollout mode with KL steps per launch:
        D*4 + 5 + (2*S + 8 + 2*grid_bytes)/KL + E.  The action byte is dropped when actions are drawn in-kernel;
        E = 1600/L (IoU read + map clear per episode of mean length L) is counted for 3D only."""
        act = 0 if philox_actions else 1
        E = (1600.0 / mean_len) if (dim == 3 and mean_len) else 0.0
        if KL == 1:
            return b_alg_step - 1 + act + E
        S, grid = {1: (5, 60), 2: (6, 50), 3: (6, 800)}[dim]
        return act + D * 4 + 5 + (2 * S + 8 + 2 * grid) / KL + E

    @staticmethod
    def b_layout(dim, b_alg_step, D, KL, philox_actions, mean_len=None):
        """The same formula with the bytes of THIS repo's 3D layout (nibble maps: window 49 x 0.5 B + the 1 B store that
        carries the brick instead of 49 x 2 + 2, 208 B maps instead of 800 B, E = 416/L): what a kernel on this layout has
        to move at least.  1D/2D: identical to b_alg."""
        if dim != 3:
            return Bench.b_alg(dim, b_alg_step, D, KL, philox_actions, mean_len)
        act = 0 if philox_actions else 1
        E = (416.0 / mean_len) if mean_len else 0.0
        if KL == 1:
            return b_alg_step - 100 + 25.5 - 1 + act + E
        return act + D * 4 + 5 + (2 * 6 + 8 + 2 * 208) / KL + E

    def kernel_result(self, wl, total_envs, env, r, KL_eff, philox_actions, mean_len):
        """value / roofline of one time_regions() result (KL_eff: steps per launch the byte formula is evaluated at)."""
        dim, _, _, _, _, b_alg_step = WORKLOADS[wl]
        n = env.num_envs
        per_gpu = n * r["steps"] / (r["ms"] * 1e-3)
        ba = self.b_alg(dim, b_alg_step, env.obs_dim, KL_eff, philox_actions, mean_len)
        bl = self.b_layout(dim, b_alg_step, env.obs_dim, KL_eff, philox_actions, mean_len)
        return {"value": total_envs * r["steps"] / (r["ms"] * 1e-3), "ms_per_step": r["ms"] / r["steps"],
                "bytes_per_env_step": ba, "bytes_per_env_step_this_layout": bl,
                "achieved_gbs": per_gpu * ba / 1e9, "frac": per_gpu * ba / 1e9 / self.peak,
                "frac_this_layout": per_gpu * bl / 1e9 / self.peak}

    def episode_stats(self, env, timed=False):
        """NCCL all-reduce of the 4-double statistics vector (the path's only collective).  timed: (stats, ms)."""
        import torch
        if not timed:
            return env.stats(allreduce=True).cpu().numpy()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        env.stats(allreduce=True)                            # warm-up (communicator set-up)
        self.barrier()
        e0.record()
        s = env.stats(allreduce=True)
        e1.record()
        torch.cuda.synchronize()
        return s.cpu().numpy(), self.max_over_ranks(e0.elapsed_time(e1))

    # -------------------------------------------------------------------------------------------
    def host_loop(self, step_fn, Ke, total_envs, batches=5):
        """Wall-clock env-steps/s of Ke synchronous host-buffer steps (median of `batches` batches, max over ranks)."""
        import torch
        for i in range(3):
            step_fn(i)
        ts = []
        for b in range(batches):
            self.barrier()
            t0 = time.perf_counter()
            for i in range(Ke):
                step_fn(i)
            torch.cuda.synchronize()
            ts.append(self.max_over_ranks(time.perf_counter() - t0))
        return total_envs * Ke / statistics.median(ts)

    def pcie_probe(self, mb=256, reps=4):
        """Device-to-host copy bandwidth with ALL ranks copying at the same time (pinned host buffers): what the host-buffer
        numbers are bound by.  Returns GB/s per GPU (slowest rank) and for the whole job."""
        import torch
        n = mb << 20
        d = torch.empty(n, dtype=torch.uint8, device=self.dev)
        h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        h.copy_(d, non_blocking=True)
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            h.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
        t = self.max_over_ranks(time.perf_counter() - t0)
        per = reps * n / t / 1e9
        return {"d2h_gbs_per_gpu": per, "d2h_gbs_all_gpus": per * self.world,
                "what": "%d ranks copy %d x %d MB device -> pinned host at the same time (slowest rank's time)" % (self.world, reps, mb)}

    def e2e_suite(self, wl, total_envs, K):
        """The reference-facing calls with HOST buffers: numpy actions in, numpy results out, H2D + D2H inside the timed
        region.  `bits` = bit-packed step records (2D 16 B / 3D 32 B per env: one buffer, one D2H copy, the smallest
        host-facing kind and the line's `e2e`); `record` = byte records (56 B / 16 B per env); f32 / i16 = [N, D] observations + reward + done (one fused D2H copy); wrapper = the drop-in
        VectorizedEnvWrapper.step of multiprocess.py (float64 (N,1,D) observations, numpy-RNG step sizes)."""
        import numpy as np
        import torch
        from snac_b200.compat import HostStepper
        args = self.args
        dim, dynamic, plan_choose, density, _, _ = WORKLOADS[wl]
        Ke = max(4, min(K, args.e2e_steps))
        rng = np.random.RandomState(99 + self.rank)
        out = {}
        kinds = [("record", "record", False), ("record_mapped", "record", True)]
        if dim > 1:                                          # bit records: 16 B (2D) / 32 B (3D) per env-step
            kinds += [("bits", "bits", None), ("bits_staged", "bits", False)]
        if not args.no_e2e:
            kinds += [("f32", torch.float32, False), ("i16", torch.int16, False)]
        for name, dt, mapped in kinds:
            env = self.make_env(wl, total_envs, obs_dtype=dt)
            n, A = env.num_envs, env.action_dim
            host_actions = [rng.randint(0, A, size=n).astype(np.uint8) for _ in range(4)]
            hs = HostStepper(env, mapped=mapped)
            v = self.host_loop(lambda i: hs.step(host_actions[i % 4]), Ke, total_envs)
            env.check_errors()
            api = {"record": "HostStepper.step(actions: np.uint8[N]) -> records: structured numpy array [N] (%d B each: u8 window + 1 [49] | flags | u16 count_brick | u16 count_step | i8 reward | bool done), a view of the pinned buffer the one D2H copy filled" % (hs.d2h_bytes // n) if dim > 1 else
                             "HostStepper.step(actions: np.uint8[N]) -> records: structured numpy array [N] (16 B each: i16 window[5] | u16 count_brick | u16 count_step | i8 reward | bool done)",
                   "bits": "HostStepper.step(actions: np.uint8[N]) -> bit records: np.uint8 [N, %d] (DMP_OBS_BITS: 49 x %d-bit window code | 12-bit count_brick | 12-bit count_step | 3-bit reward code | done | saturated), a view of the pinned buffer the one D2H copy filled; snac_b200.vecenv.unpack_bits / unpack_records_device expand them to the [N, 51] rows" % (hs.d2h_bytes // n, 2 if dim == 2 else 4),
                   "f32": "HostStepper.step(actions) -> (obs f32[N,%d], reward f32[N], done bool[N]) numpy views of one pinned buffer" % env.obs_dim,
                   "i16": "HostStepper.step(actions) -> (obs i16[N,%d], reward f32[N], done bool[N])" % env.obs_dim}[name.split("_")[0]]
            if hs.mapped:
                api += "; mapped=True: no staging copies, the kernel reads the actions from and writes the records into mapped pinned host memory"
            out[name] = {"value": v, "unit": "env-steps/s", "steps": Ke, "h2d_bytes_per_step": int(hs.h2d_bytes) * self.world,
                         "d2h_bytes_per_step": int(hs.d2h_bytes) * self.world, "api": api}
            del hs, env
        if not args.no_e2e:
            from snac_b200 import compat
            cls = {1: compat.deep_mobile_printing_1d1r, 2: compat.deep_mobile_printing_2d1r, 3: compat.deep_mobile_printing_3d1r}[dim]
            if not dynamic:
                from snac_b200.sharding import shard_bounds
                _, n = shard_bounds(total_envs, self.rank, self.world)
                proto = cls(plan_choose=plan_choose, device=self.dev)
                for name, kw in (("wrapper", {}), ("wrapper_f32_philox", {"obs_dtype": torch.float32, "step_size_rng": "philox"})):
                    vw = compat.VectorizedEnvWrapper(proto, num_envs=n, auto_reset=True, **kw)
                    vw.reset()
                    acts = [rng.randint(0, proto.action_dim, size=n) for _ in range(4)]
                    Kw = max(4, min(Ke, 12))
                    v = self.host_loop(lambda i: vw.step(acts[i % 4]), Kw, total_envs, batches=2)
                    out[name] = {"value": v, "unit": "env-steps/s", "steps": Kw, "h2d_bytes_per_step": int(vw.h2d_bytes) * self.world,
                                 "d2h_bytes_per_step": int(vw.d2h_bytes) * self.world,
                                 "api": "VectorizedEnvWrapper(env, N).step(actions) -> ((N,1,%d) %s, (N,) float64, (N,) bool) numpy, %s step sizes (multiprocess.py:24-32)"
                                        % (proto.state_dim, "float64" if not kw else "float32",
                                           "np.random.randint(1,4,size=N)" if not kw else "in-kernel Philox")}
                    del vw
        return out

    # -------------------------------------------------------------------------------------------
    def secondary(self, wl, total_envs, action_dist="uniform", K=None):
        """One short kernel-only measurement of another BASELINE configuration (rollout and step mode)."""
        dim = WORKLOADS[wl][0]
        env = self.make_env(wl, total_envs, action_dist=action_dist)
        K = K or (64 if dim == 1 else 20)
        out = {"total_envs": total_envs, "envs_per_gpu": env.num_envs, "action_dist": action_dist}
        # warm up long enough that episodes end and auto-resets are part of the steady state (3D: ~23-step episodes)
        W = {1: 128, 2: 40, 3: 200}[dim]
        r = self.time_regions(env, K, min(K, KL_MAX), W, min_repeats=30, target_s=0.12)
        s = self.time_regions(env, 64, 1, 16, min_repeats=30, target_s=0.12)
        stats, ar_ms = self.episode_stats(env, timed=True)
        env.check_errors()
        mean_len = float(stats[3] / stats[2]) if stats[2] > 0 else None
        kr = self.kernel_result(wl, total_envs, env, r, min(K, KL_MAX), True, mean_len)
        ks = self.kernel_result(wl, total_envs, env, s, 1, True, mean_len)
        out["rollout"] = dict(kr, steps=K, repeats=r["repeats"], clocks=r["clocks"])
        out["step"] = dict(ks, steps=64, repeats=s["repeats"])
        out["mean_episode_length"] = mean_len
        out["mean_iou"] = float(stats[1] / stats[2]) if stats[2] > 0 else None
        out["stats_allreduce_ms"] = ar_ms
        del env
        return out

    # -------------------------------------------------------------------------------------------
    def run(self):
        import torch
        args = self.args
        wl = args.workload
        dim, dynamic, plan_choose, density, default_envs, b_alg_step = WORKLOADS[wl]
        state_bytes = {1: 72, 2: 64, 3: 416}[dim]
        total_envs = args.envs or default_envs
        K, W = args.steps, max(args.warmup, 3)
        env = self.make_env(wl, total_envs, action_dist=args.action_dist)
        n, D, A = env.num_envs, env.obs_dim, env.action_dim
        act_pool = None
        if args.actions == "buffer":
            g = torch.Generator(device=self.dev)
            g.manual_seed(1234 + self.rank)
            act_pool = torch.randint(0, A, (64, n), dtype=torch.uint8, device=self.dev, generator=g)
        philox_actions = act_pool is None

        KL = 1 if args.mode == "step" else (args.rollout_k or min(K, KL_MAX))
        main = self.time_regions(env, K, KL, W, act_pool)
        other = None
        if not args.single_mode:
            oKL = 1 if KL > 1 else (args.rollout_k or min(K, KL_MAX))
            oK = K if oKL > 1 else max(K, 64)
            other = self.time_regions(env, oK, oKL, W, act_pool, target_s=0.15)
        stats, ar_ms = self.episode_stats(env, timed=True)
        env.check_errors()
        mean_len = float(stats[3] / stats[2]) if stats[2] > 0 else None
        KL_eff = 1 if KL == 1 else max(main["steps_per_launch"])
        kr = self.kernel_result(wl, total_envs, env, main, KL_eff, philox_actions, mean_len)

        e2e = self.e2e_suite(wl, total_envs, K)

        pcie = self.pcie_probe()

        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("%s|%d|K=%d" % (wl, n, KL_eff))
        except Exception:
            pass
        mode_txt = ("step: one dmp_step launch per vector step" if KL == 1 else
                    "rollout: dmp_rollout, %s steps per launch, every step's obs/reward/done materialised" % "+".join(map(str, main["steps_per_launch"])))
        line = {"metric": METRIC, "value": kr["value"], "unit": "env-steps/s", "n_gpus": self.world, "steps": K, "warmup": W,
                "ms_per_step": kr["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "u32 bit-grid / i32 counters, f32 observations" if dim == 2 else
                         ("u16 heights / i32 counters, f32 observations" if dim == 1 else "u4 heights (u16 escape) / i32 counters, f32 observations"),
                "data": "synthetic", "repeats": main["repeats"],
                "region_ms": {"median": main["ms"], "min": main["ms_min"], "max": main["ms_max"], "mean": main["ms_mean"]},
                "config": {"workload": wl, "total_envs": total_envs, "envs_per_gpu": n,
                           "mode": mode_txt + "; the %d-step region is one CUDA graph, timed %d times, median reported" % (K, main["repeats"]),
                           "timing": main["timing"],
                           "actions": (("uniform over %d actions, " % A) if args.action_dist == "uniform" else
                                       "the reference's p = [.2, .2, .2, .2, .05, .05, .05, .05] (Env/3D/DMP_simulator_3d_static_circle.py:361-362), ")
                                      + ("pool of 64 pre-generated vectors in HBM" if act_pool is not None else "Philox4x32-10 in-kernel, counter (global env id, step)"),
                           "step_size": "Philox4x32-10 in-kernel", "auto_reset": True,
                           "l2": main["ring"] + " + state %.1f MB" % (n * state_bytes / 1e6),
                           "parallelism": "env-sharded x%d, NCCL all-reduce of episode stats only (%.3f ms, outside the timed region)" % (self.world, ar_ms)},
                "clocks": main["clocks"],
                "e2e": e2e["bits" if dim > 1 else "record"], "e2e_bits_staged": e2e.get("bits_staged"),
                "e2e_record": e2e["record"], "e2e_record_mapped": e2e.get("record_mapped"), "e2e_f32": e2e.get("f32"), "e2e_i16": e2e.get("i16"),
                "e2e_wrapper": e2e.get("wrapper"), "e2e_wrapper_f32_philox": e2e.get("wrapper_f32_philox"),
                "pcie": pcie,
                "gpu_launches": main["launches_per_region"] * main["repeats"],
                "gpu_launches_per_region": main["launches_per_region"],
                "roofline": {"bound": "hbm", "achieved": kr["achieved_gbs"], "peak": self.peak, "unit": "GB/s", "frac": kr["frac"],
                             "traffic": traffic,
                             "traffic_note": "DRAM bytes per launch of this kernel at this size from the committed ncu --set full capture (profiles/traffic.json), not measured live; algorithmic bytes per launch = %d" % int(kr["bytes_per_env_step"] * n * KL_eff),
                             "peak_source": self.peak_src,
                             "kernel": ("k3d_step_bytes<float> (K=1)" if (dim == 3 and KL == 1) else
                                        "k%dd%s_rollout<float> (K=%d)" % (dim, "_cache" if dim == 3 else "", KL_eff)),
                             "bytes_per_env_step": kr["bytes_per_env_step"], "envs_per_launch": n,
                             "bytes_per_env_step_this_layout": kr["bytes_per_env_step_this_layout"], "frac_this_layout": kr["frac_this_layout"],
                             "layout_note": None if dim != 3 else "frac uses SURVEY 8(d)'s canonical u16 maps (800 B, window 98 B, E = 1600/L); this repo's 3D state is nibble maps (208 B, window 24.5 B, E = 416/L): frac_this_layout is the fraction by those bytes"},
                "episode_stats": {"mean_episode_length": mean_len, "mean_iou": (float(stats[1] / stats[2]) if stats[2] > 0 else None),
                                  "mean_return": (float(stats[0] / stats[2]) if stats[2] > 0 else None),
                                  "sum_return": float(stats[0]), "sum_iou": float(stats[1]), "episodes": float(stats[2]),
                                  "steps": float(stats[3]), "allreduce_ms": ar_ms}}
        if other is not None:
            oKL_eff = 1 if other["launches_per_region"] == other["steps"] else max(other["steps_per_launch"])
            ko = self.kernel_result(wl, total_envs, env, other, oKL_eff, philox_actions, mean_len)
            line["other_mode"] = {"mode": "rollout K=%d" % oKL_eff if oKL_eff > 1 else "step", "ms_per_step": ko["ms_per_step"],
                                  "value": ko["value"], "steps": other["steps"], "repeats": other["repeats"],
                                  "gpu_launches": other["launches_per_region"] * other["repeats"],
                                  "bytes_per_env_step": ko["bytes_per_env_step"], "roofline_frac": ko["frac"],
                                  "roofline_frac_this_layout": ko["frac_this_layout"]}
        del env

        # ---- the other BASELINE configurations, one short measurement each (kernel only) -----------------------
        if not args.no_workloads:
            wls = {}
            for name, envs, dist_ in (("1d_dynamic", 65536, "uniform"), ("2d_dynamic_dense", 1048576, "uniform"),
                                      ("3d_static_dense", 262144, "uniform"), ("3d_static_dense", 262144, "ref3d"),
                                      ("3d_dynamic_dense", 262144, "uniform"), ("3d_dynamic_dense", 262144, "ref3d")):
                key = name if dist_ == "uniform" else name + "|ref3d"
                if name == wl and dist_ == args.action_dist:
                    continue
                try:
                    wls[key] = self.secondary(name, envs, dist_)
                except Exception as e:                       # a failing side measurement must not lose the headline
                    wls[key] = {"failed": repr(e)}
            line["workloads"] = wls

        if self.rank == 0 and self.world == 1 and not args.no_cpu_baseline:
            dynamic_wl = WORKLOADS[wl][1]
            try:
                c = cpu_throughput(wl, 128 if dynamic_wl else 1024, 25, 5)
                line["cpu_baseline"] = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample", "uncached_reset_ms")}
            except Exception as e:
                line["cpu_baseline"] = {"unavailable": repr(e)}
            try:
                line["cpu_baseline_c"] = cpu_c_port_throughput(wl)
            except Exception as e:
                line["cpu_baseline_c"] = {"unavailable": repr(e)}
            try:
                line["config1"] = {"cpu": config1_cpu(), "ours": config1_ours(self.dev),
                                   "what": "BASELINE configs[0]: 1D static plan_choose=2, random actions 0..2, num_envs=5, 750 iterations"}
            except Exception as e:
                line["config1"] = {"failed": repr(e)}
        if self.rank == 0:
            print(json.dumps(line), flush=True)
        if self.world > 1:
            self.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--actions", default="philox", choices=["philox", "buffer"])
    ap.add_argument("--mode", default="rollout", choices=["rollout", "step"],
                    help="rollout: dmp_rollout launches (the whole region in one launch when --steps <= 64); "
                         "step: one launch per vector step (dmp_step)")
    ap.add_argument("--rollout-k", type=int, default=0, help="steps per dmp_rollout launch (default: min(--steps, 64))")
    ap.add_argument("--single-mode", action="store_true", help="skip the secondary measurement of the other mode")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="2d_static_dense", choices=sorted(WORKLOADS))
    ap.add_argument("--action-dist", default="uniform", choices=["uniform", "ref3d"],
                    help="in-kernel action distribution: uniform over the env's actions, or (3D) the reference's own "
                         "p = [.2, .2, .2, .2, .05, .05, .05, .05] (SURVEY.md 8(d) cfg 5)")
    ap.add_argument("--envs", type=int, default=0, help="total envs over all GPUs (default: the BASELINE config's)")
    ap.add_argument("--e2e-steps", type=int, default=48)
    ap.add_argument("--ref-envs-per-proc", type=int, default=0, help="--impl reference: env objects per host process")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-workloads", action="store_true", help="skip the short measurements of the other BASELINE configs")
    ap.add_argument("--no-e2e", action="store_true", help="kernel A/B runs: only the record-kind e2e")
    ap.add_argument("--timing", default="graph", choices=["graph", "stream"],
                    help="graph: the timing events are the first / last node of each region's CUDA graph (default); stream: "
                         "stream events around the graph replay, behind a spin kernel that hides the host's launch latency "
                         "(the fallback where event nodes cannot be captured)")
    args = ap.parse_args()
    if args.action_dist == "ref3d" and (WORKLOADS[args.workload][0] != 3 or args.actions != "philox"):
        ap.error("--action-dist ref3d needs a 3D workload and in-kernel (philox) actions")
    if args.steps < 1:
        ap.error("--steps must be >= 1")
    if args.impl == "reference":
        run_reference(args)
    else:
        Bench(args).run()


if __name__ == "__main__":
    main()
