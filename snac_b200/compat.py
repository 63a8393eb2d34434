"""Reference-facing surface: the same class names, constructor arguments, methods, attributes and
return types as the reference's env classes and ``multiprocess.py``, backed by the CUDA vector env.

    deep_mobile_printing_1d1r(plan_choose)            Env/1D/DMP_Env_1D_static.py:6
    deep_mobile_printing_1d1r_dynamic(data_path, random_choose_paln)   Env/1D/DMP_Env_1D_dynamic_usedata_plan.py:7
    deep_mobile_printing_2d1r / _2d1r_dynamic         Env/2D/DMP_Env_2D_static.py:6, ..._dynamic_usedata_plan.py:6
    deep_mobile_printing_3d1r / _3d1r_dynamic         Env/3D/DMP_simulator_3d_static_circle.py:7, ..._dynamic_triangle_usedata.py:6
    VectorizedEnvWrapper(env, num_envs)               multiprocess.py:15-32

Stochastic draws follow the reference exactly: ``step`` draws ``np.random.randint(1, 4)`` and a random
``reset`` draws ``np.random.randint(0, len(dataset))`` from the process-global numpy RNG, in the same
order as the reference, and injects them into the kernel -- so ``np.random.seed(s)`` reproduces the
reference's trajectories bit for bit.  ``step(action, step_size)`` (the *_hindsight_replay form) is
accepted as well.

These scalar classes exist for drop-in compatibility; one env per launch is latency-bound on a GPU.
Throughput comes from ``VectorizedEnvWrapper`` / ``BatchedDMPEnv`` with thousands of envs.
"""
from __future__ import annotations

import argparse
from typing import Optional

import numpy as np
import torch

import ctypes as C

from . import _lib as L
from .vecenv import BatchedDMPEnv, generate_plans, load_plan_dataset


def bind_to_gpu_numa_node(device) -> Optional[list]:
    """Pin this process to the CPUs of the NUMA node the GPU hangs off (NVML's ideal CPU affinity), so that pinned staging
    buffers allocated afterwards are node-local and PCIe copies do not cross the socket interconnect.  Matters when one
    process per GPU drives host-buffer steps on a multi-socket box (bench.py --gpus N); returns the CPU list or None."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        idx = torch.device(device).index
        h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device() if idx is None else idx)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w in range(words) for b in range(64) if (mask[w] >> b) & 1]
        cpus = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None


def _align16(x: int) -> int:
    return (x + 15) & ~15


class HostStepper:
    """Host-buffer step for a BatchedDMPEnv: numpy actions in, numpy results out.

    One step = ONE host-to-device copy (actions, + injected step sizes when given), one kernel launch, ONE device-to-host
    copy of the step's whole result, one stream synchronisation.  The result lives in a single device buffer mirrored by a
    single pinned host buffer, ``[observations | rewards | done flags]`` back to back; the arrays returned are views of
    the pinned buffer (valid until the step after next: two host buffers alternate).

    * numeric obs kinds: ``step(actions) -> (obs [N, D], reward f32 [N], done bool [N])``
    * ``obs_dtype="record"`` envs: ``step(actions) -> records``, a structured array (``record_dtype``) whose fields ``win``,
      ``count_brick``, ``count_step``, ``reward``, ``done`` are views into the one buffer that was copied -- 56 B per env
      instead of 209 B: the PCIe link carries 3.7x fewer bytes per step.
    * ``obs_dtype="bits"`` envs (2D / 3D): ``step(actions) -> uint8 [N, 16 | 32]`` bit-packed records (DMP_OBS_BITS): 13x /
      6.5x fewer bytes than float32 rows; ``unpack_bits`` (host) / ``unpack_records_device`` expand them.

    ``mapped=True`` drops both copies: the pinned host buffers are mapped into the device's address space (unified
    addressing), the kernel reads the actions from host memory and writes its results straight into the pinned result
    buffer -- one launch and one synchronisation per step, and the transfer overlaps the kernel.  ``mapped="out"`` maps only
    the result buffer (the actions still go through one staged copy: a kernel reading one byte per env over the link issues
    32-byte requests).  Default: mapped for small batches (a step is then bound by call latency, not bytes) and for the 16 B
    records -- 2D bit records, 1D step records -- (one 128-bit store per env, 512 contiguous bytes per warp: the kernel's own
    writes keep the link as busy as a copy would, and nothing waits for the kernel to finish first -- 1 M 2D envs: 403 us per
    step against 441 us staged; 65 536 1D envs: 46 against 56 us),
    staged copies otherwise (56 B records and wider rows: the copy engine moves them ~15 % faster than the SMs' writes do;
    3D bit records leave as two 128-bit stores per env, half-filled sectors that the link carries badly: 262 144 envs
    0.51 G env-steps/s mapped against 1.22 G staged).
    """
    MAPPED_MAX_BYTES = 1 << 18       # default switch-over: results of at most 256 KB per step go through mapped memory

    def __init__(self, env: BatchedDMPEnv, buffers: int = 2, mapped: Optional[bool] = None):
        self.env = env
        n, row = env.num_envs, env.obs_row
        dev = env.device
        esz = torch.empty(0, dtype=env.obs_dtype).element_size()
        self._obs_bytes = n * row * esz
        if env.records:
            self._off_rew = self._off_done = None
            total = self._obs_bytes
        else:
            self._off_rew = _align16(self._obs_bytes)
            self._off_done = self._off_rew + _align16(4 * n)
            total = self._off_done + n
        self._total = total
        if mapped is None:
            mapped = total <= self.MAPPED_MAX_BYTES or (env.records and env.obs_row == 16)     # one 128-bit store per env
        self.mapped = mapped if mapped == "out" else bool(mapped)
        self._map_in, self._map_out = self.mapped is True, bool(self.mapped)
        self._res_pin = [torch.empty(total, dtype=torch.uint8, pin_memory=True) for _ in range(max(1, int(buffers)))]
        self._views = [self._host_views(b.numpy()) for b in self._res_pin]
        # what the kernel writes: the device result buffer, or (mapped) each pinned buffer itself
        self._res_dev = None if self._map_out else torch.empty(total, dtype=torch.uint8, device=dev)
        self._outs = [self._out_tensors(b) for b in ((self._res_pin) if self._map_out else [self._res_dev])]
        self._in_pin = torch.empty(2 * n, dtype=torch.uint8, pin_memory=True)          # [actions | step sizes]
        self._in_dev = torch.empty(2 * n, dtype=torch.uint8, device=dev)
        self._in_np = self._in_pin.numpy()
        self.actions_buffer = self._in_np[:n]               # write actions here to skip one host copy
        self.step_sizes_buffer = self._in_np[n:]
        self._i = 0
        self.h2d_bytes = n                                  # 2 n when step sizes are injected
        self.d2h_bytes = total
        # the C-ABI call itself, prepared once per result buffer and per "step sizes injected?": a step of a compact kind is
        # a few hundred microseconds, of which the argument checks / tensor views of the generic path were a quarter
        src = self._in_pin if self._map_in else self._in_dev
        self._ios = []
        for o, r, d in self._outs:
            pair = []
            for with_sizes in (False, True):
                io = L.DmpIO()
                io.actions = src.data_ptr()
                io.step_sizes = src.data_ptr() + n if with_sizes else None
                io.next_plan = None
                io.obs, io.reward, io.done = o.data_ptr(), (None if r is None else r.data_ptr()), (None if d is None else d.data_ptr())
                io.obs_kind, io.flags = env._kind_of(env.obs_dtype), env._flags(1)
                pair.append(io)
            self._ios.append(pair)
        self._dev_index = env.device.index

    def _out_tensors(self, buf: torch.Tensor):
        env, n = self.env, self.env.num_envs
        obs = buf[:self._obs_bytes].view(env.obs_dtype).view(1, n, env.obs_row)
        if env.records:
            return obs, None, None
        return (obs, buf[self._off_rew:self._off_rew + 4 * n].view(torch.float32).view(1, n),
                buf[self._off_done:self._off_done + n].view(1, n))

    def _host_views(self, buf: np.ndarray):
        env, n = self.env, self.env.num_envs
        if env.records:
            from .vecenv import BITS, record_dtype
            if env.record_kind == BITS:                     # bit records have no field view: uint8 [N, 16 | 32] (unpack_bits)
                return buf[:self._obs_bytes].reshape(n, env.obs_row)
            return buf[:self._obs_bytes].view(record_dtype(env.dim))
        npdt = {torch.float32: np.float32, torch.float64: np.float64, torch.int16: np.int16}[env.obs_dtype]
        return (buf[:self._obs_bytes].view(npdt).reshape(n, env.obs_row),
                buf[self._off_rew:self._off_rew + 4 * n].view(np.float32),
                buf[self._off_done:self._off_done + n].view(np.bool_))

    def _step_on_device(self, j: int, with_sizes: bool, nin: int):
        env = self.env
        stream = torch.cuda.current_stream(env.device)
        if not self._map_in:
            self._in_dev[:nin].copy_(self._in_pin[:nin], non_blocking=True)
        st = env._st
        st.t_dev = None
        rc = L.lib.dmp_rollout(C.byref(st), C.byref(self._ios[j if self._map_out else 0][1 if with_sizes else 0]), 1, stream.cuda_stream)
        if rc:
            L.check(rc, "dmp_rollout")
        st.t = st.t + 1
        if not self._map_out:
            self._res_pin[j].copy_(self._res_dev, non_blocking=True)
        stream.synchronize()

    def step(self, actions, step_sizes=None):
        env, n = self.env, self.env.num_envs
        if env._needs_initial_reset:
            raise RuntimeError("call reset() before step()")
        if actions is not self.actions_buffer:
            self.actions_buffer[:] = actions
        if step_sizes is not None and step_sizes is not self.step_sizes_buffer:
            self.step_sizes_buffer[:] = step_sizes
        nin = n if step_sizes is None else 2 * n
        j = self._i
        self._i = (j + 1) % len(self._res_pin)
        if torch.cuda.current_device() == self._dev_index:
            self._step_on_device(j, step_sizes is not None, nin)
        else:
            with torch.cuda.device(env.device):
                self._step_on_device(j, step_sizes is not None, nin)
        return self._views[j]


# --------------------------------------------------------------------------------------------------
# scalar drop-in classes
# --------------------------------------------------------------------------------------------------
class _ScalarDMP:
    """One reference-style env (N = 1 on the device)."""
    _dim = 0
    _dynamic = False
    _lnet = False              # *_Lnet observation-format variants

    def _setup(self, plan_choose=0, data_path=None, random_choose_paln=True, device="cuda", plans=None):
        self.plan_choose = plan_choose
        self.random_choose_paln = random_choose_paln
        self.index_for_non_random = 0
        self.index_random = None
        if self._dynamic:
            if plans is None:
                plans = load_plan_dataset(data_path, self._dim)
            self.plan_dataset = [np.asarray(p, dtype=np.float64) for p in plans]
            self.plan_dataset_len = len(self.plan_dataset)
        self._device = device
        self._plans_arg = None if not self._dynamic else np.asarray(self.plan_dataset)
        self._env: Optional[BatchedDMPEnv] = None
        self._plan_cache = None            # content of self.plan as the device last saw it (hindsight overwrites)
        self._plan_row = 0
        self.step_size = 1
        self.count_step = 0
        self.total_brick = 0
        self.plan = None
        self.input_plan = None
        self.one_hot = None
        self.environment_memory = None
        self.position_memory = None
        self.observation = None
        self.brick_memory = None
        self._set_count_brick(None)
        self._dataset_idx = None
        # constants (reference __init__ blocks)
        probe = {1: (30, 20, 2, 3, 7, 750, 750), 2: (20, 20, 3, 5, 51, 600, 600), 3: (20, 20, 3, 8, 51, 1300, 1000)}[self._dim]
        self.plan_width, self.plan_height, self.HALF_WINDOW_SIZE, self.action_dim, self.state_dim = probe[:5]
        self.total_step = probe[6] if self._dynamic else probe[5]
        self.environment_width = self.plan_width + 2 * self.HALF_WINDOW_SIZE
        if self._dim == 1:
            self.environment_height = 100
            self.wall = np.ones((1, 2)) * (-1)
        else:
            self.environment_height = self.plan_height + 2 * self.HALF_WINDOW_SIZE
        if self._dim == 3:
            self.plan_length, self.z, self.blank_size, self.check, self.start = 10, 6, 2, [], None
            self.environment_length = self.plan_length + 2 * self.HALF_WINDOW_SIZE

    # 1D base classes spell the counter 'conut_brick' (Env/1D/DMP_Env_1D_static.py:14); expose both
    def _set_count_brick(self, v):
        self.count_brick = v
        self.conut_brick = v

    def _ensure(self):
        if self._env is None:
            self._env = BatchedDMPEnv(self._dim, dynamic=self._dynamic, plan_choose=self.plan_choose,
                                      plans=self._plans_arg, num_envs=1, device=self._device,
                                      obs_dtype=torch.float64, random_choose_paln=self.random_choose_paln,
                                      dynamic_rules=True if (self._lnet and self._dim == 3) else None)
            self._dense_plans = self._env.plans_dense()
            # one spare plan row: a plan written by the caller (hindsight relabelling) is packed there
            tab, tot = self._env.plan_table().clone(), self._env.plan_totals().clone()
            self._env._install_plans(torch.cat([tab, torch.zeros_like(tab[:1])]), torch.cat([tot, tot[:1]]))
            self._scratch_row = self._env.n_plans - 1
            self._prepare_step()

    # ---- one env per step is bound by call latency, not by bytes: the step and the state export that refreshes the
    # reference's attributes run back to back on mapped pinned host buffers (the kernels read the action from and write
    # observation / reward / done / grid / scalars into host memory) behind ONE stream synchronisation
    def _prepare_step(self):
        env = self._env
        lay = env._lay
        cells = lay.grid_rows * lay.grid_cols
        self._pin = dict(inp=torch.zeros(2, dtype=torch.uint8, pin_memory=True),
                         obs=torch.zeros(lay.obs_dim, dtype=torch.float64, pin_memory=True),
                         rew=torch.zeros(1, dtype=torch.float32, pin_memory=True),
                         done=torch.zeros(1, dtype=torch.uint8, pin_memory=True),
                         grid=torch.zeros(cells, dtype=torch.int32, pin_memory=True),
                         sc=torch.zeros(8, dtype=torch.int32, pin_memory=True))
        self._pin_np = {k: v.numpy() for k, v in self._pin.items()}
        self._grid_shape = (1, lay.grid_cols) if self._dim == 1 else (lay.grid_rows, lay.grid_cols)
        io = L.DmpIO()
        io.actions, io.step_sizes, io.next_plan = self._pin["inp"].data_ptr(), self._pin["inp"].data_ptr() + 1, None
        io.obs, io.reward, io.done = self._pin["obs"].data_ptr(), self._pin["rew"].data_ptr(), self._pin["done"].data_ptr()
        io.obs_kind, io.flags = L.OBS_F64, env._flags(1)
        self._step_io = io
        self._dev_index = env.device.index

    def _launch_step_and_export(self):
        env = self._env
        stream = torch.cuda.current_stream(env.device)
        env._st.t_dev = None
        rc = L.lib.dmp_rollout(C.byref(env._st), C.byref(self._step_io), 1, stream.cuda_stream)
        if rc:
            L.check(rc, "dmp_rollout")
        env._st.t = env._st.t + 1
        rc = L.lib.dmp_export_state(C.byref(env._st), self._pin["grid"].data_ptr(), self._pin["sc"].data_ptr(), None,
                                    stream.cuda_stream)
        if rc:
            L.check(rc, "dmp_export_state")
        stream.synchronize()

    def _step_device(self, a, step_size):
        """One step of the env + a fresh export of its state, through the prepared call.  Returns (obs, reward, done, pos)."""
        inp = self._pin_np["inp"]
        inp[0], inp[1] = a, step_size
        if torch.cuda.current_device() == self._dev_index:
            self._launch_step_and_export()
        else:
            with torch.cuda.device(self._env.device):
                self._launch_step_and_export()
        sc = self._pin_np["sc"]
        self.environment_memory = self._pin_np["grid"].reshape(self._grid_shape).astype(np.float64)
        if self._lnet and self._dim == 2:                    # frame value 2 (Env/2D/DMP_Env_2D_static_Lnet.py:61-64)
            self.environment_memory[self.environment_memory == -1] = 2
        self._set_count_brick(int(sc[2]))
        self.count_step = int(sc[3])
        pos = int(sc[0]) if self._dim == 1 else [int(sc[0]), int(sc[1])]
        return (self._pin_np["obs"].reshape(1, -1).copy(), float(self._pin_np["rew"][0]), bool(self._pin_np["done"][0]), pos)

    # ---- hindsight relabelling: the reference's learners overwrite ``env.plan`` between reset() and step(), by
    # assignment (script/DRQN_hindsight/1d/DRQN_hindsight_1D_dynamic.py:255) or in place
    # (script/DRQN_hindsight/2d/DRQN_hindsight_2D_dynamic.py:274-276).  Both are noticed at the next step().
    def _pack_plan_row(self, plan, row, budget=None):
        env = self._env
        shape = (1, 30) if self._dim == 1 else (1, 26, 26)
        raw = torch.as_tensor(np.ascontiguousarray(np.asarray(plan, dtype=np.float64).reshape(shape)), device=env.device)
        rb = env._lay.plan_row_bytes
        tot = torch.zeros(1, dtype=torch.int32, device=env.device)
        with torch.cuda.device(env.device):
            L.check(L.lib.dmp_plans_pack(self._dim, raw.data_ptr(), 1, env._plans.data_ptr() + row * rb, tot.data_ptr(),
                                         env._stream()), "dmp_plans_pack")
        env._plan_total[row] = tot[0] if budget is None else int(np.ceil(budget))

    def _point_env_at_row(self, row):
        sc = self._env.export_state()["scalars"]
        sc[0, 4] = row
        self._env.import_state(scalars=sc)
        self._plan_row = row

    def _sync_plan(self):
        if self.plan is None or self._plan_cache is None:
            return
        cur = np.asarray(self.plan, dtype=np.float64)
        if cur.shape == self._plan_cache.shape and np.array_equal(cur, self._plan_cache):
            return
        if cur.shape != self._plan_cache.shape:
            raise ValueError("plan must keep its shape %s" % (self._plan_cache.shape,))
        # the episode goes on against the new plan with the budget reset() fixed (the reference does not recompute it)
        self._pack_plan_row(cur, self._scratch_row, budget=self.total_brick)
        self._point_env_at_row(self._scratch_row)
        if self._dynamic and self._dataset_idx is not None and np.shares_memory(self.plan, self._dense_plans[self._dataset_idx]):
            # in-place edit of a dataset entry: it persists in the reference's plan_dataset, so it persists here
            self._pack_plan_row(cur, self._dataset_idx)
        self._plan_cache = cur.copy()

    def _sync_attrs(self):
        st = self._env.export_state()
        sc = st["scalars"][0].cpu().numpy()
        self.environment_memory = st["grid"][0].cpu().numpy().astype(np.float64)
        if self._lnet and self._dim == 2:                    # frame value 2 (Env/2D/DMP_Env_2D_static_Lnet.py:61-64)
            self.environment_memory[self.environment_memory == -1] = 2
        self._set_count_brick(int(sc[2]))
        self.count_step = int(sc[3])
        pos = int(sc[0]) if self._dim == 1 else [int(sc[0]), int(sc[1])]
        return pos

    def _choose_plan(self) -> int:
        """Plan-table row for the next episode, consuming the global numpy RNG exactly like the reference's reset()."""
        idx = 0
        if self._dynamic:
            if self.random_choose_paln:                      # e.g. Env/2D/DMP_Env_2D_dynamic_usedata_plan.py:35-38
                self.index_random = int(np.random.randint(0, self.plan_dataset_len))
                idx = self.index_random
            else:                                            # :39-44
                idx = self.index_for_non_random
                self.index_for_non_random += 1
                if self.index_for_non_random == self.plan_dataset_len:
                    self.index_for_non_random = 0
        return idx

    def reset(self):
        self._ensure()
        idx = self._choose_plan()
        self._dataset_idx = idx if self._dynamic else None
        self._plan_row = idx
        obs = self._env.reset(plan_idx=[idx]).cpu().numpy().reshape(1, -1)
        # dataset classes alias the dataset entry (as the reference does); static ones get a fresh array per reset()
        self.plan = self._dense_plans[idx] if self._dynamic else self._dense_plans[idx].copy()
        tb = float(self._env.plan_totals()[idx].item())
        self.total_brick = tb
        if self._dim != 1:
            h = self.HALF_WINDOW_SIZE
            self.input_plan = self.plan[h:h + self.plan_height, h:h + self.plan_width]
        if self._dim == 1 and not self._dynamic:
            self.one_hot = None
        self._plan_cache = np.asarray(self.plan, dtype=np.float64).copy()
        pos = self._sync_attrs()
        self.position_memory = [pos]
        self.brick_memory = [[-1, -1]] if self._dim == 1 else None
        if self._dim == 3:
            self.check, self.step_size = [], 1
        return self._format_obs(obs, pos, reset=True)

    def step(self, action, step_size=None):
        if self._env is None or self.position_memory is None:
            raise AttributeError("call reset() before step()")
        a = int(action)
        self._sync_plan()
        self.step_size = int(np.random.randint(1, 4)) if step_size is None else int(step_size)
        if self._dim != 3 and not (0 <= a < self.action_dim):
            # the reference increments count_step, draws, then fails on the unbound 'position'
            self._step_device(255, self.step_size)
            self._env._err.zero_()
            raise UnboundLocalError("local variable 'position' referenced before assignment")
        if self._dim == 3 and a < 0:
            raise ValueError("negative actions are not supported by the device path")
        obs, r, d, pos = self._step_device(min(a, 255), self.step_size)
        self.position_memory.append(pos)
        if self._dim == 1:
            self.brick_memory.append([pos, self.environment_memory[0, pos]] if a == 2 else [-1, -1])
        if a >= self.action_dim:                             # 3D: an action past the table is an unbuilt brick, not an error
            self._env._err.zero_()
        return self._format_obs(obs, pos, reset=False), self._reward_type(a, r, d), d

    def _reward_type(self, a, r, d):
        """The reference returns the python int 0 on moves (and on 2D drops above the plan) and floats
        elsewhere; 3D always returns floats."""
        if self._dim == 3:
            return r
        drop = self.action_dim - 1
        if a != drop:
            return 0
        if self._dim == 2 and r == 0.0 and not (self.count_brick >= self.total_brick):
            return 0
        return r

    def _normalised(self, obs):
        o = obs.copy()
        o[0, -2] = self.count_brick / self.total_brick
        o[0, -1] = self.count_step / self.total_step
        return o

    def _format_obs(self, obs, pos, reset):
        if self._lnet:
            if self._dim == 1:                               # Env/1D/DMP_Env_1D_static_Lnet.py:83 (+position)
                return np.hstack((obs, np.array([[pos]], dtype=np.float64)))
            if self._dim == 2:
                obs = obs.copy()
                w = obs[:, :49]
                w[w == -1] = 2
            return [self._normalised(obs), pos]              # Env/2D/DMP_Env_2D_static_Lnet.py:75-76
        if not self._dynamic:
            return obs
        if self._dim == 1:                                   # Env/1D/DMP_Env_1D_dynamic_usedata_plan.py:66-70, :114-118
            out = [obs, self._normalised(obs), self.plan]
            if reset:
                out.append(pos)
            return out
        return [self._normalised(obs), self.input_plan, pos]  # Env/2D/..._usedata_plan.py:64-66

    def iou(self):
        return float(self._env.iou().item())

    def render(self, *a, **k):
        raise NotImplementedError("render() needs matplotlib and is outside the simulator hot path")


class deep_mobile_printing_1d1r(_ScalarDMP):
    _dim, _dynamic = 1, False

    def __init__(self, plan_choose=0, device="cuda"):
        self._setup(plan_choose=plan_choose, device=device)


class deep_mobile_printing_1d1r_dynamic(_ScalarDMP):
    _dim, _dynamic = 1, True

    def __init__(self, data_path=None, random_choose_paln=True, device="cuda", plans=None):
        self._setup(data_path=data_path, random_choose_paln=random_choose_paln, device=device, plans=plans)


class deep_mobile_printing_2d1r(_ScalarDMP):
    _dim, _dynamic = 2, False

    def __init__(self, plan_choose=0, device="cuda"):
        self._setup(plan_choose=plan_choose, device=device)


class deep_mobile_printing_2d1r_dynamic(_ScalarDMP):
    _dim, _dynamic = 2, True

    def __init__(self, data_path=None, random_choose_paln=True, device="cuda", plans=None):
        self._setup(data_path=data_path, random_choose_paln=random_choose_paln, device=device, plans=plans)


class deep_mobile_printing_3d1r(_ScalarDMP):
    _dim, _dynamic = 3, False

    def __init__(self, plan_choose=1, device="cuda"):          # 3D default is 1 (…static_circle.py:8)
        self._setup(plan_choose=plan_choose, device=device)


class deep_mobile_printing_3d1r_dynamic(_ScalarDMP):
    _dim, _dynamic = 3, True

    def __init__(self, data_path=None, random_choose_paln=True, device="cuda", plans=None):
        self._setup(data_path=data_path, random_choose_paln=random_choose_paln, device=device, plans=plans)


class deep_mobile_printing_1d1r_Lnet(deep_mobile_printing_1d1r):
    """Env/1D/DMP_Env_1D_static_Lnet.py: observation (1,8) = window, count_brick, count_step, position."""
    _lnet = True


class deep_mobile_printing_2d1r_Lnet(deep_mobile_printing_2d1r):
    """Env/2D/DMP_Env_2D_static_Lnet.py: frame value 2, normalised counters, returns [obs, position]."""
    _lnet = True


class deep_mobile_printing_3d1r_Lnet(deep_mobile_printing_3d1r):
    """Env/3D/DMP_simulator_3d_static_circle_Lnet.py: static plan, the dynamic class's termination rules
    (-100 when boxed in), normalised counters, returns [obs, position]."""
    _lnet = True


# --------------------------------------------------------------------------------------------------
# *_hindsight_replay classes (explicit step_size).  The reference gives all six the same class
# name per dimension and tells them apart by module; here the dataset / generator ones keep the reference name and the
# static ones get a ``_static`` suffix.
# --------------------------------------------------------------------------------------------------
class deep_mobile_printing_1d1r_hindsight(_ScalarDMP):
    """Env/1D/DMP_Env_1D_dynamic_hindsight_replay.py: a fresh random sinusoid from ``create_plan()`` at every reset()
    (generated on the device from the reference's three numpy draws); observations are ``[raw (1,7), plan (30,)]``."""
    _dim, _dynamic = 1, True

    def __init__(self, device="cuda"):
        self._setup(device=device, plans=np.full((1, 30), 20.0), random_choose_paln=False)

    def create_plan(self):
        k_1 = np.random.uniform(3, 12)                       # :32-34, same draws in the same order
        k_2 = np.random.randint(1, 4)
        phase = np.random.uniform(-1, 1) * np.pi
        self.one_hot = [k_1, k_2, phase]
        table, _, _ = generate_plans(1, 1, draws=np.array([[k_1, k_2, phase]]), device=self._device)
        y = table[0, :30].cpu().numpy().astype(np.float64)
        return y, sum(y)

    def _choose_plan(self) -> int:
        self.one_hot = None
        y, _ = self.create_plan()
        self._dense_plans[0] = y
        self._pack_plan_row(y, 0)
        return 0

    def _format_obs(self, obs, pos, reset):
        return [obs, self.plan]                              # :69-70, :96-109


class deep_mobile_printing_2d1r_hindsight(_ScalarDMP):
    """Env/2D/DMP_Env_2D_dynamic_hindsight_replay_usedata.py: dataset plans like the ``usedata_plan`` class, but reset()
    first runs ``create_plan()`` (a random triangle that is then discarded -- it still consumes the numpy RNG) and
    observations carry the raw counters: ``[raw (1,51), input_plan, position]``."""
    _dim, _dynamic = 2, True

    def __init__(self, data_path=None, random_choose_paln=True, device="cuda", plans=None, plan_choose=None):
        self._setup(data_path=data_path, random_choose_paln=random_choose_paln, device=device, plans=plans)
        if plan_choose is not None:                          # the reference infers it from the file name (:31-32)
            self.plan_choose = plan_choose
        elif data_path is not None and "sparse" in data_path:
            self.plan_choose = 1

    def create_plan(self):
        if self.plan_choose not in (0, 1):
            raise ValueError(' 0: Dense triangle, 1: Sparse triangle')
        while True:                                          # :41-57: redraw until the area exceeds 50 / 20
            x = np.random.randint(0, self.plan_width, size=3)
            y = np.random.randint(0, self.plan_height, size=3)
            try:
                table, _, _ = generate_plans(2, 1, self.plan_choose, draws=np.concatenate([x, y])[None, None, :],
                                             device=self._device)
            except ValueError:
                continue
            bits = np.unpackbits(table[0, :52].cpu().numpy(), bitorder="little")[:400].reshape(20, 20)
            plan = np.zeros((26, 26))
            plan[3:23, 3:23] = bits
            return plan, sum(sum(plan))

    def _choose_plan(self) -> int:
        self.create_plan()
        return _ScalarDMP._choose_plan(self)

    def _format_obs(self, obs, pos, reset):
        return [obs, self.input_plan, pos]


class deep_mobile_printing_3d1r_hindsight(_ScalarDMP):
    """Env/3D/DMP_simulator_3d_dynamic_triangle_hindsight_replay.py: the dataset class with an explicit step size;
    reset() returns ``[obs, input_plan]`` (no position, :70-72)."""
    _dim, _dynamic = 3, True

    def __init__(self, data_path=None, random_choose_paln=True, device="cuda", plans=None):
        self._setup(data_path=data_path, random_choose_paln=random_choose_paln, device=device, plans=plans)

    def _format_obs(self, obs, pos, reset):
        if reset:
            return [obs, self.input_plan]
        return [self._normalised(obs), self.input_plan, pos]


class deep_mobile_printing_1d1r_hindsight_static(deep_mobile_printing_1d1r):
    """Env/1D/DMP_Env_1D_static_hindsight_replay.py: the static class with ``step(action, step_size)``."""


class deep_mobile_printing_2d1r_hindsight_static(deep_mobile_printing_2d1r):
    """Env/2D/DMP_Env_2D_static_hindsight_replay.py."""


class deep_mobile_printing_3d1r_hindsight_static(deep_mobile_printing_3d1r):
    """Env/3D/DMP_simulator_3d_static_circle_hindsight_replay.py."""


# --------------------------------------------------------------------------------------------------
# tree-search variants (Env/*/...MCTS*.py): reset() -> (state, obs), step(a) -> (state, obs, reward,
# done), and the functional transition(state, a) that script/MCTS/utils/uct.py expands tree nodes with.
# state = (position, environment_memory, count_brick, count_step); observations are always the raw (1, D) row.
# --------------------------------------------------------------------------------------------------
class _MCTSMixin:
    _transition_copies = False      # only the 1D dataset class copies the caller's array (Env/1D/DMP_Env_1D_dynamic_MCTS.py:86)

    def _format_obs(self, obs, pos, reset):
        return obs

    def _snapshot(self):
        pos = self.position_memory[-1]
        pos = pos if self._dim == 1 else list(pos)
        self.state = (pos, self.environment_memory.copy(), self.count_brick, self.count_step)
        return self.state

    def reset(self):
        self.state = None
        if not hasattr(self, "action_space"):
            self.action_space = _Space(n=self.action_dim)
        obs = super().reset()
        return self._snapshot(), obs

    def step(self, action, step_size=None):
        obs, r, d = super().step(action, step_size)
        return self._snapshot(), obs, r, d

    def equality_operator(self, o1, o2):
        return bool(np.array_equal(o1, o2))

    def iou_MCTS(self, environment_memory):
        """IoU of a tree node's grid against the current plan (Env/1D/DMP_Env_1D_static_MCTS.py:250-264)."""
        tw = self._twin(1)
        sc = torch.zeros((1, 8), dtype=torch.int32, device=tw.device)
        sc[0, 0], sc[0, 1], sc[0, 4] = self.HALF_WINDOW_SIZE, (self.HALF_WINDOW_SIZE if self._dim != 1 else 0), self._plan_row
        tw.import_state(grid=self._grid_to_device(environment_memory, tw), scalars=sc)
        return float(tw.iou().item())

    # ---- expansion ------------------------------------------------------------------------------
    def _twin(self, n):
        """A second device env of n slots that shares this env's plan table: tree nodes are stepped there, so
        expanding never disturbs the episode the env itself is in."""
        if self._env is None:
            raise AttributeError("call reset() before transition()")
        tw = getattr(self, "_twins", {}).get(n)
        if tw is None:
            if not hasattr(self, "_twins"):
                self._twins = {}
            src = self._env
            tw = BatchedDMPEnv(self._dim, dynamic=self._dynamic, plan_choose=self.plan_choose,
                               plans=(src.plan_table(), src.plan_totals()) if self._dynamic else None, num_envs=n,
                               device=self._device, obs_dtype=torch.float64, dynamic_rules=src.dynamic_rules)
            if not self._dynamic:
                tw._install_plans(src.plan_table(), src.plan_totals())       # incl. the hindsight scratch row
            self._twins[n] = tw
        return tw

    def _grid_to_device(self, grids, tw):
        g = np.asarray(grids, dtype=np.float64)
        if self._lnet and self._dim == 2:
            g = np.where(g == 2, -1, g)
        return torch.as_tensor(np.ascontiguousarray(g).astype(np.int32), device=tw.device)

    def transition_batch(self, states, actions, step_sizes=None):
        """``transition`` for many (state, action) pairs in ONE kernel launch (leaf-parallel expansion).  Step sizes
        default to one ``np.random.randint(1, 4)`` draw per pair, in order -- the stream a loop over transition()
        would consume.  Returns a list of (state', observation, reward, done)."""
        n = len(states)
        acts = [int(a) for a in actions]
        if len(acts) != n:
            raise ValueError("one action per state")
        if step_sizes is None:
            step_sizes = [int(np.random.randint(1, 4)) for _ in range(n)]
        self._sync_plan()
        tw = self._twin(n)
        pos = np.asarray([[s[0], 0] if self._dim == 1 else list(s[0]) for s in states], dtype=np.int32)
        grids = np.stack([np.asarray(s[1], dtype=np.float64).reshape(tw._lay.grid_rows, tw._lay.grid_cols) for s in states])
        a_dev = [min(a, 255) if a >= 0 else 255 for a in acts]
        if self._dim == 3 and any(a < 0 for a in acts):
            raise ValueError("negative actions are not supported by the device path")
        npos, ngrid, ncb, ncs, obs, rew, done = tw.transition_dense(
            pos[:, 0] if self._dim == 1 else pos, self._grid_to_device(grids, tw),
            [int(s[2]) for s in states], [int(s[3]) for s in states], a_dev, step_sizes,
            plan_idx=[self._plan_row] * n)
        tw._err.zero_()                                      # out-of-range actions are no-op steps here, not errors
        npos, ncb, ncs = npos.cpu().numpy(), ncb.cpu().numpy(), ncs.cpu().numpy()
        ngrid = ngrid.cpu().numpy().astype(np.float64)
        obs, rew, done = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()
        out = []
        for i, st in enumerate(states):
            g_in = st[1]
            g_new = ngrid[i].reshape(np.shape(g_in))
            if self._transition_copies or not isinstance(g_in, np.ndarray):
                g_out = g_new
            else:                                            # the reference mutates the caller's array in place
                g_in[...] = g_new
                g_out = g_in
            p = int(npos[i]) if self._dim == 1 else [int(npos[i, 0]), int(npos[i, 1])]
            cb, r, d = int(ncb[i]), float(rew[i]), bool(done[i])
            if self._dim == 2:                               # 2D: int 0 on moves and on drops above the plan, like step()
                a = acts[i]
                if a != 4 or (r == 0.0 and not cb >= self.total_brick):
                    r = 0
            out.append(((p, g_out, cb, int(ncs[i])), obs[i].reshape(1, -1).copy(), r, d))
        return out

    def transition(self, state, action, is_model_dynamic=True):
        """(state, action) -> (state', observation, reward, done), e.g. Env/2D/DMP_ENV_2D_static_MCTS.py:110-169,
        Env/3D/DMP_simulator_3d_static_circle_MCTS.py:215-289.  Draws its step size from the global numpy stream."""
        step_size = int(np.random.randint(1, 4))
        return self.transition_batch([state], [action], [step_size])[0]


class deep_mobile_printing_1d1r_MCTS(_MCTSMixin, deep_mobile_printing_1d1r):
    """Env/1D/DMP_Env_1D_static_MCTS.py."""


class deep_mobile_printing_1d1r_MCTS_obs(_MCTSMixin, deep_mobile_printing_1d1r_dynamic):
    """Env/1D/DMP_Env_1D_dynamic_MCTS.py (dataset plans; transition() leaves the caller's grid untouched)."""
    _transition_copies = True


class deep_mobile_printing_2d1r_MCTS(_MCTSMixin, deep_mobile_printing_2d1r):
    """Env/2D/DMP_ENV_2D_static_MCTS.py."""


class deep_mobile_printing_2d1r_MCTS_dynamic(_MCTSMixin, deep_mobile_printing_2d1r_dynamic):
    """Env/2D/DMP_ENV_2D_dynamic_MCTS.py (class ``deep_mobile_printing_2d1r`` of that module)."""


class deep_mobile_printing_3d1r_MCTS(_MCTSMixin, deep_mobile_printing_3d1r):
    """Env/3D/DMP_simulator_3d_static_circle_MCTS.py (class ``deep_mobile_printing_3d1r`` of that module)."""


class deep_mobile_printing_3d1r_MCTS_dynamic(_MCTSMixin, deep_mobile_printing_3d1r_dynamic):
    """Env/3D/DMP_simulator_3d_dynamic_triangle_MCTS.py (class ``deep_mobile_printing_3d1r`` of that module)."""


# --------------------------------------------------------------------------------------------------
# learner-side copies of the envs (script/SAC/environments/*.py, script/PPO/*/DMP_*.py): the same simulators with a
# flat (D,) observation, and -- for the stable-baselines PPO scripts -- gym spaces and a 4-tuple step()
# --------------------------------------------------------------------------------------------------
class _Space:
    """Minimal stand-in for gym.spaces.{Discrete,Box} (gym is not a dependency of this package)."""

    def __init__(self, n=None, low=None, high=None, dtype=np.int64):
        self.n, self.low, self.high, self.dtype = n, low, high, dtype
        self.shape = () if low is None else np.asarray(low).shape

    def sample(self):
        if self.n is not None:
            return int(np.random.randint(self.n))
        return np.random.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        if self.n is not None:
            return 0 <= int(x) < self.n
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))


class FlatObsEnv:
    """``FlatObsEnv(env)``: reset() -> (D,) ; step(a) -> ((D,), reward, done) like script/SAC/environments/DMP_Env_1D_static.py:72-118;
    ``FlatObsEnv(env, gym_api=True)``: step(a) -> ((D,), reward, done, {}) plus ``action_space`` / ``observation_space``
    like script/PPO/2d_static/DMP_Env_2D_static.py:32-34, :144-161.  Everything else is forwarded to the wrapped env."""

    def __init__(self, env: _ScalarDMP, gym_api: bool = False):
        self.env, self.gym_api = env, bool(gym_api)
        win = env.state_dim - 2
        self.action_space = _Space(n=env.action_dim)
        self.observation_space = _Space(low=np.array([-1] * win + [0, 0]),
                                        high=np.array([99] * win + [env.total_step, env.total_step]))

    def __getattr__(self, name):
        return getattr(self.env, name)

    @staticmethod
    def _flat(o):
        return np.asarray(o[0] if isinstance(o, list) else o).reshape(-1)

    def reset(self):
        return self._flat(self.env.reset())

    def step(self, action, step_size=None):
        o, r, d = self.env.step(action, step_size)
        return (self._flat(o), r, d, {}) if self.gym_api else (self._flat(o), r, d)


# --------------------------------------------------------------------------------------------------
# multiprocess.py replacement
# --------------------------------------------------------------------------------------------------
class VectorizedEnvWrapper:
    """``VectorizedEnvWrapper(env, num_envs)`` of multiprocess.py:15-32 with num_envs INDEPENDENT,
    device-resident copies of ``env`` (the reference repeats one shared object, quirk Q1 -- not reproduced).
    reset() -> (N,1,D) float64; step(actions) -> ((N,1,D) float64, (N,) float64, (N,) bool), all numpy.

    step_size_rng  "numpy" (default): per-env step sizes are drawn as np.random.randint(1, 4, size=N), the same
                   global-RNG stream the reference consumes when it loops over N env objects (bit-exact drop-in; the
                   draw costs ~6 ns per env on the host); "philox": the kernels' counter-based stream (throughput).
    obs_dtype      torch.float64 (default, the reference's dtype), float32 or int16: dtype of the returned observations.
    shards         the batch is held as this many independent device shards (contiguous env ranges) that a step runs as a
                   pipeline: while the copy engine brings shard c's results to the host, the host draws shard c+1's step
                   sizes (in env order, so the numpy stream is consumed exactly as by one draw of N) and launches it.
                   Default: one shard per 131 072 envs (at most 16) with numpy step sizes, else 1.  Results do not depend
                   on the number of shards (Philox streams are keyed by the global env index).
    The arrays returned by step() are zero-copy views of pinned staging buffers (two alternate: an array stays valid until
    the step after next; copy what must live longer).  Rewards are converted to float64 on the device."""
    PIPELINE_SHARD_ENVS = 1 << 17

    def __init__(self, env_: _ScalarDMP, num_envs: int = 1, obs_dtype=torch.float64, auto_reset: bool = False,
                 step_size_rng: str = "numpy", mapped: Optional[bool] = None, shards: Optional[int] = None):
        if step_size_rng not in ("numpy", "philox"):
            raise ValueError('step_size_rng must be "numpy" or "philox"')
        self.env = env_
        self.num_envs = num_envs
        self.envs = [env_ for _ in range(num_envs)]
        self._proto = env_
        self.total_step = env_.total_step
        self.step_size_rng = step_size_rng
        n = num_envs
        esz = torch.empty(0, dtype=obs_dtype).element_size()
        lay = L.DmpLayout()
        L.check(L.lib.dmp_layout(env_._dim, n, C.byref(lay)), "dmp_layout")
        D = lay.obs_dim
        # one device result buffer [obs | reward f64 | done] mirrored by two pinned host buffers
        self._obs_bytes = n * D * esz
        self._off_rew = _align16(self._obs_bytes)
        self._off_done = self._off_rew + 8 * n
        total = self._off_done + n
        # small batches (the reference's own --num_envs 3 / 5): no staging copies at all -- the kernel reads the actions from
        # and writes its results into mapped pinned host memory; a step is one launch and one synchronisation
        self.mapped = (total <= HostStepper.MAPPED_MAX_BYTES) if mapped is None else bool(mapped)
        if shards is None:
            shards = 1 if (step_size_rng != "numpy" or self.mapped) else min(16, n // self.PIPELINE_SHARD_ENVS)
        self.n_shards = max(1, min(int(shards), n))
        if self.mapped and self.n_shards != 1:
            raise ValueError("mapped host buffers are for small batches: one shard")
        from .sharding import shard_bounds
        self._bounds = [shard_bounds(n, c, self.n_shards) for c in range(self.n_shards)]
        self.vecs = []
        for first, cnt in self._bounds:
            plans = env_._plans_arg
            if self.vecs and env_._dynamic:                    # one packed plan table, shared by every shard
                plans = (self.vecs[0].plan_table(), self.vecs[0].plan_totals())
            self.vecs.append(BatchedDMPEnv(env_._dim, dynamic=env_._dynamic, plan_choose=env_.plan_choose, plans=plans,
                                           num_envs=cnt, device=env_._device, obs_dtype=obs_dtype, env_base=first,
                                           random_choose_paln=env_.random_choose_paln, auto_reset=auto_reset,
                                           normalise=False))
        self.vec = self.vecs[0]
        dev = self.vec.device
        self._D, self._esz = D, esz
        self._res_dev = torch.empty(total, dtype=torch.uint8, device=dev)
        self._obs_dev = self._res_dev[:self._obs_bytes].view(obs_dtype).view(1, n, D)
        self._rew64_dev = self._res_dev[self._off_rew:self._off_done].view(torch.float64)
        self._done_dev = self._res_dev[self._off_done:].view(1, n)
        self._rew32_dev = torch.empty((1, n), dtype=torch.float32, device=dev)
        npdt = {torch.float32: np.float32, torch.float64: np.float64, torch.int16: np.int16}[obs_dtype]
        self._res_pin, self._views = [], []
        for _ in range(2):
            b = torch.empty(total, dtype=torch.uint8, pin_memory=True)
            a = b.numpy()
            self._res_pin.append(b)
            self._views.append((a[:self._obs_bytes].view(npdt).reshape(n, 1, D),
                                a[self._off_rew:self._off_done].view(np.float64),
                                a[self._off_done:].view(np.bool_)))
        self._in_pin = torch.empty(2 * n, dtype=torch.uint8, pin_memory=True)
        self._in_dev = torch.empty(2 * n, dtype=torch.uint8, device=dev)
        self._in_np = self._in_pin.numpy()
        self._i = 0
        self.h2d_bytes = 2 * n if step_size_rng == "numpy" else n
        self.d2h_bytes = total
        self._dev_index = dev.index
        from .vecenv import _TORCH_OBS
        kind = _TORCH_OBS[obs_dtype]
        numpy_rng = step_size_rng == "numpy"
        if self.mapped:
            self._rew32_pin = torch.empty((1, n), dtype=torch.float32, pin_memory=True)
            self._rew32_np = self._rew32_pin.numpy()[0]
            self._outs = [(b[:self._obs_bytes].view(obs_dtype).view(1, n, D), self._rew32_pin, b[self._off_done:].view(1, n))
                          for b in self._res_pin]
            # the call itself, prepared once: a step of a small batch is bound by host latency, not by bytes
            self._ios = []
            for o, r, d in self._outs:
                io = L.DmpIO()
                io.actions = self._in_pin.data_ptr()
                io.step_sizes = self._in_pin.data_ptr() + n if numpy_rng else None
                io.next_plan = None
                io.obs, io.reward, io.done = o.data_ptr(), r.data_ptr(), d.data_ptr()
                io.obs_kind, io.flags = kind, self.vec._flags(1)
                self._ios.append(io)
        else:
            # staged path: per shard, the prepared call on its slice of the shared buffers and the copies around it
            self._stages = []
            for (first, cnt), vec in zip(self._bounds, self.vecs):
                io = L.DmpIO()
                io.actions = self._in_dev.data_ptr() + first
                io.step_sizes = self._in_dev.data_ptr() + n + first if numpy_rng else None
                io.next_plan = None
                io.obs = self._res_dev.data_ptr() + first * D * esz
                io.reward = self._rew32_dev.data_ptr() + 4 * first
                io.done = self._res_dev.data_ptr() + self._off_done + first
                io.obs_kind, io.flags = kind, vec._flags(1)
                sl = slice(first, first + cnt)
                ob = slice(first * D * esz, (first + cnt) * D * esz)
                rw = slice(self._off_rew + 8 * first, self._off_rew + 8 * (first + cnt))
                dn = slice(self._off_done + first, self._off_done + first + cnt)
                h2d = [(self._in_dev[sl], self._in_pin[sl])]
                if numpy_rng:
                    h2d.append((self._in_dev[n + first:n + first + cnt], self._in_pin[n + first:n + first + cnt]))
                d2h = [[(b[ob], self._res_dev[ob]), (b[rw], self._res_dev[rw]), (b[dn], self._res_dev[dn])] for b in self._res_pin]
                self._stages.append((vec, io, sl, h2d, (self._rew64_dev[sl], self._rew32_dev[0, sl]), d2h))

    def _draw_plans(self, n):
        if not self._proto._dynamic:
            return None
        if self._proto.random_choose_paln:
            return np.random.randint(0, self._proto.plan_dataset_len, size=n).astype(np.int32)
        return None

    def reset(self):
        p = self._draw_plans(self.num_envs)
        obs = [vec.reset(plan_idx=None if p is None else p[f:f + c]) for (f, c), vec in zip(self._bounds, self.vecs)]
        obs = obs[0] if len(obs) == 1 else torch.cat(obs)
        return obs.cpu().numpy().reshape(self.num_envs, 1, -1)

    def reset_at(self, env_index):
        p = self._draw_plans(1)
        c = next(i for i, (f, k) in enumerate(self._bounds) if f <= env_index < f + k)
        vec, local = self.vecs[c], env_index - self._bounds[c][0]
        m = np.zeros(vec.num_envs, np.uint8)
        m[local] = 1
        pi = None
        if p is not None:
            pi = np.zeros(vec.num_envs, np.int32)
            pi[local] = p[0]
        obs = vec.reset(mask=m, plan_idx=pi)
        return obs[local].cpu().numpy().reshape(1, -1)

    def _step_prepared(self, j):
        vec = self.vec
        stream = torch.cuda.current_stream(vec.device)
        vec._st.t_dev = None
        rc = L.lib.dmp_rollout(C.byref(vec._st), C.byref(self._ios[j]), 1, stream.cuda_stream)
        if rc:
            L.check(rc, "dmp_rollout")
        vec._st.t = vec._st.t + 1
        stream.synchronize()

    def _step_staged(self, j, actions):
        """The shard pipeline: everything is enqueued on one stream; the host only waits at the end, so drawing shard c+1's
        step sizes overlaps shard c's kernel and device-to-host copies."""
        numpy_rng = self.step_size_rng == "numpy"
        n = self.num_envs
        stream = torch.cuda.current_stream(self.vec.device)
        sp = stream.cuda_stream
        for vec, io, sl, h2d, rew, d2h in self._stages:
            if numpy_rng:
                self._in_np[n + sl.start:n + sl.stop] = np.random.randint(1, 4, size=sl.stop - sl.start)
            self._in_np[sl] = actions[sl]
            for dst, src in h2d:
                dst.copy_(src, non_blocking=True)
            vec._st.t_dev = None
            rc = L.lib.dmp_rollout(C.byref(vec._st), C.byref(io), 1, sp)
            if rc:
                L.check(rc, "dmp_rollout")
            vec._st.t = vec._st.t + 1
            rew[0].copy_(rew[1])                                # float64 rewards like the reference's np.asarray(...)
            for dst, src in d2h[j]:
                dst.copy_(src, non_blocking=True)
        stream.synchronize()

    def step(self, actions):
        n = self.num_envs
        if self.vec._needs_initial_reset:
            raise RuntimeError("call reset() before step()")
        j = self._i
        out = self._views[j]
        self._i ^= 1
        if self.mapped:
            if self.step_size_rng == "numpy":
                self._in_np[n:] = np.random.randint(1, 4, size=n)
            self._in_np[:n] = actions
            if torch.cuda.current_device() == self._dev_index:
                self._step_prepared(j)
            else:
                with torch.cuda.device(self.vec.device):
                    self._step_prepared(j)
            out[1][:] = self._rew32_np                          # float64 rewards like the reference's np.asarray(...)
            return out
        actions = np.asarray(actions)
        if actions.shape != (n,):
            raise ValueError("actions must have shape (%d,)" % n)
        if torch.cuda.current_device() == self._dev_index:
            self._step_staged(j, actions)
        else:
            with torch.cuda.device(self.vec.device):
                self._step_staged(j, actions)
        return out


def main(args=None):
    """CLI of multiprocess.py:34-97: --env {1DStatic,...,3DDynamic} --plan_type INT --num_envs INT."""
    parser = argparse.ArgumentParser()
    parser.add_argument('--env', type=str, default=None,
                        help='Environment Name: {1DStatic, 1DDynamic,2DStatic, 2DDynamic, 3DStatic, 3DDynamic}')
    parser.add_argument('--plan_type', type=int, default=None, help='type of shapes')
    parser.add_argument('--num_envs', type=int, default=3, help='Number of environments')
    parser.add_argument('--data_root', type=str, default='./Env', help='directory holding the reference .pkl plan datasets')
    args = parser.parse_args(args)
    if args.env is None:
        print("please choose an environment in the list: {1DStatic, 1DDynamic,2DStatic, 2DDynamic, 3DStatic, 3DDynamic} ")
        return None
    need_plan = {"1DStatic": "{0: sin, 1:Gaussian, 2: Step}", "2DStatic": "{0: Dense, 1: Sparse}",
                 "2DDynamic": "{0: Dense, 1: Sparse}", "3DStatic": "{0: Dense, 1: Sparse}", "3DDynamic": "{0: Dense, 1: Sparse}"}
    if args.env in need_plan and args.plan_type is None:
        print("please choose a shape from list: " + need_plan[args.env])
        return None
    dens = "sparse" if args.plan_type == 1 else "dense"
    if args.env == "1DStatic":
        env_input = deep_mobile_printing_1d1r(plan_choose=args.plan_type)
    elif args.env == "1DDynamic":
        env_input = deep_mobile_printing_1d1r_dynamic(data_path=args.data_root + "/1D/data_1d_dynamic_sin_envplan_500_train.pkl")
    elif args.env == "2DStatic":
        env_input = deep_mobile_printing_2d1r(plan_choose=args.plan_type)
    elif args.env == "2DDynamic":
        env_input = deep_mobile_printing_2d1r_dynamic(data_path=args.data_root + "/2D/data_2d_dynamic_" + dens + "_envplan_500_train.pkl")
    elif args.env == "3DStatic":
        env_input = deep_mobile_printing_3d1r(plan_choose=args.plan_type)
    elif args.env == "3DDynamic":      # the reference loads the 2D dataset here (quirk Q5); we load the 3D one
        env_input = deep_mobile_printing_3d1r_dynamic(data_path=args.data_root + "/3D/data_3d_dynamic_" + dens + "_envplan_500_train.pkl")
    else:
        raise SystemExit("unknown --env %s" % args.env)
    env = VectorizedEnvWrapper(env_input, num_envs=args.num_envs)
    T = env.total_step
    observations = env.reset()
    for t in range(T):
        actions = np.random.randint(3, size=args.num_envs)     # multiprocess.py:83 (actions 0..2 only, quirk Q2)
        observations, rewards, dones = env.step(actions)
    print(observations.shape)
    print(rewards.shape)
    print(dones.shape)
    return observations, rewards, dones


if __name__ == '__main__':
    main()
