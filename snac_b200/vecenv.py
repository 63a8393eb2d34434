"""Device-resident vector environment for SNAC's mobile-construction tasks.

``BatchedDMPEnv`` is the batched (tensor) form of the reference's gym-style surface
``reset() -> obs`` / ``step(action) -> (obs, reward, done)`` (e.g. Env/2D/DMP_Env_2D_static.py:54,95)
and replaces ``multiprocess.VectorizedEnvWrapper`` (multiprocess.py:15-32).  All state lives in
PyTorch CUDA tensors laid out as described in include/dmp.h; every method is a thin wrapper over one
C-ABI call into libdmp.so.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib as L

_TORCH_OBS = {torch.float32: L.OBS_F32, torch.float64: L.OBS_F64, torch.int16: L.OBS_I16, torch.uint8: L.OBS_REC}
RECORD = "record"      # obs_dtype=RECORD: packed step records (DMP_OBS_REC, include/dmp.h), uint8 [..., rec_bytes]
BITS = "bits"          # obs_dtype=BITS: bit-packed step records (DMP_OBS_BITS, 2D 16 B / 3D 32 B), uint8 [..., bits_bytes]
_NP_OBS = {torch.float32: np.float32, torch.float64: np.float64, torch.int16: np.int16}


def record_dtype(dim: int) -> np.dtype:
    """numpy structured dtype of one DMP_OBS_REC step record (include/dmp.h): view a uint8 [..., rec_bytes] buffer with it.
    2D / 3D (56 B): ``win`` u8[49] = window value + 1 (0 is the -1 frame), ``flags``, ``count_brick``, ``count_step``,
    ``reward`` (i8), ``done`` (bool).  1D (16 B): ``win`` i16[5] raw heights, ``count_brick``, ``count_step``, ``reward``,
    ``done``."""
    if dim == 1:
        return np.dtype({"names": ["win", "count_brick", "count_step", "reward", "done"],
                         "formats": [("<i2", (5,)), "<u2", "<u2", "i1", "?"], "offsets": [0, 10, 12, 14, 15], "itemsize": 16})
    return np.dtype({"names": ["win", "flags", "count_brick", "count_step", "reward", "done"],
                     "formats": [("u1", (49,)), "u1", "<u2", "<u2", "i1", "?"], "offsets": [0, 49, 50, 52, 54, 55],
                     "itemsize": 56})


def bits_bytes(dim: int) -> int:
    """Size of one DMP_OBS_BITS record (include/dmp.h): 16 B in 2D, 32 B in 3D (1D has no bit records)."""
    if dim not in (2, 3):
        raise ValueError("bit records exist for 2D and 3D envs (the 1D step record is 16 B already)")
    return 16 if dim == 2 else 32


def unpack_bits(rec, dim: int, dtype=np.float64):
    """Bit-packed step records (DMP_OBS_BITS) -> (obs [..., 51], reward float32 [...], done bool [...], saturated bool
    [...]) on the host with numpy.  ``rec``: uint8 array [..., 16 | 32] (numpy or torch).  ``saturated`` marks records whose
    observation is not exact (a counter above 4 095, or in 3D a window cell of height >= 14)."""
    if torch.is_tensor(rec):
        rec = rec.cpu().numpy()
    rb = bits_bytes(dim)
    if rec.dtype != np.uint8 or rec.shape[-1] != rb:
        raise ValueError("expected a uint8 array [..., %d]" % rb)
    w = np.ascontiguousarray(rec).view("<u4")                       # [..., 4 | 8]
    cw = 2 if dim == 2 else 4
    obs = np.empty(rec.shape[:-1] + (51,), dtype=dtype)
    for j in range(49):
        b = j * cw
        obs[..., j] = ((w[..., b >> 5] >> np.uint32(b & 31)) & np.uint32((1 << cw) - 1)).astype(np.int64) - 1
    tr = (w[..., 3] >> np.uint32(2)) if dim == 2 else w[..., 7]
    obs[..., 49] = tr & np.uint32(0xFFF)
    obs[..., 50] = (tr >> np.uint32(12)) & np.uint32(0xFFF)
    reward = np.asarray(L.BITS_REWARDS, np.float32)[(tr >> np.uint32(24)) & np.uint32(7)]
    return obs, reward, ((tr >> np.uint32(27)) & np.uint32(1)).astype(bool), ((tr >> np.uint32(28)) & np.uint32(1)).astype(bool)


def unpack_records_device(rec: torch.Tensor, dim: int, kind: str = RECORD, dtype: torch.dtype = torch.float32):
    """The same expansion ON THE DEVICE (``dmp_records_unpack``): a uint8 CUDA tensor [..., record bytes] of step records
    (``kind`` "record" or "bits") -> (obs [..., D] of ``dtype``, reward float32 [...], done bool [...], saturated bool
    [...]) CUDA tensors.  For learners that keep a replay buffer of compact records on the host and expand each sampled
    minibatch after uploading it."""
    if not (torch.is_tensor(rec) and rec.is_cuda and rec.dtype == torch.uint8):
        raise ValueError("rec must be a uint8 CUDA tensor")
    rk = {RECORD: L.OBS_REC, BITS: L.OBS_BITS}[kind]
    rb = bits_bytes(dim) if kind == BITS else record_dtype(dim).itemsize
    if rec.shape[-1] != rb:
        raise ValueError("expected records of %d bytes" % rb)
    rec = rec.contiguous()
    shape = tuple(rec.shape[:-1])
    n = int(np.prod(shape)) if shape else 1
    D = 7 if dim == 1 else 51
    dev = rec.device
    obs = torch.empty(shape + (D,), dtype=dtype, device=dev)
    reward = torch.empty(shape, dtype=torch.float32, device=dev)
    done = torch.empty(shape, dtype=torch.uint8, device=dev)
    sat = torch.empty(shape, dtype=torch.uint8, device=dev)
    if n:
        with torch.cuda.device(dev):
            L.check(L.lib.dmp_records_unpack(dim, rk, rec.data_ptr(), n, obs.data_ptr(), _TORCH_OBS[dtype], reward.data_ptr(),
                                             done.data_ptr(), sat.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                    "dmp_records_unpack")
    return obs, reward, done.view(torch.bool), sat.view(torch.bool)


def unpack_records(rec, dim: int, dtype=np.float64):
    """Step records -> (obs [..., D], reward float32 [...], done bool [...], saturated bool [...]) numpy arrays in the
    layout of the other observation kinds (raw counters).  ``rec``: uint8 array [..., rec_bytes] (numpy or torch) or an
    array of record_dtype(dim); a uint8 array whose last axis has the size of a bit record (2D 16 / 3D 32) goes through
    unpack_bits()."""
    if torch.is_tensor(rec):
        rec = rec.cpu().numpy()
    if not rec.dtype.names and dim != 1 and rec.shape[-1] == bits_bytes(dim):
        return unpack_bits(rec, dim, dtype)
    r = rec if rec.dtype.names else np.ascontiguousarray(rec).view(record_dtype(dim))[..., 0]
    win = r["win"].astype(dtype) - (0 if dim == 1 else 1)
    obs = np.concatenate([win, r["count_brick"][..., None].astype(dtype), r["count_step"][..., None].astype(dtype)], axis=-1)
    sat = np.zeros(r.shape, bool) if dim == 1 else (r["flags"] & L.REC_SATURATED) != 0
    return obs, r["reward"].astype(np.float32), r["done"].copy(), sat


# reference constants exposed as attributes (Env/*/…__init__)
_SPEC = {
    1: dict(plan_width=30, plan_height=20, HALF_WINDOW_SIZE=2, action_dim=3, state_dim=7),
    2: dict(plan_width=20, plan_height=20, HALF_WINDOW_SIZE=3, action_dim=5, state_dim=51),
    3: dict(plan_width=20, plan_height=20, HALF_WINDOW_SIZE=3, action_dim=8, state_dim=51, z=6, plan_length=10),
}


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def load_plan_dataset(data_path: str, dim: int, key: Optional[str] = None) -> np.ndarray:
    """Load a plan dataset in the reference's format: a joblib ``.pkl`` list of float64 arrays
    (Env/*/data_*_envplan_500_*.pkl), or this repo's compact ``.npz`` re-encoding (``key`` selects
    the array, see tests/golden/make_golden.py:pack_plans).  Returns float64 [n,30] / [n,26,26]."""
    if data_path.endswith(".npz"):
        z = np.load(data_path)
        if key is None:
            raise ValueError("npz plan files need key=, one of %s" % list(z.keys()))
        a = z[key]
        if dim == 1:
            return a.astype(np.float64)
        bits = np.unpackbits(a, axis=1)[:, :400].reshape(len(a), 20, 20).astype(np.float64)
        out = np.zeros((len(a), 26, 26))
        out[:, 3:23, 3:23] = bits * (6.0 if dim == 3 else 1.0)
        return out
    import joblib
    return np.asarray(joblib.load(data_path), dtype=np.float64)


def generate_plans(dim: int, n_plans: int, plan_choose: int = 0, *, seed: int = L.SEED_DEFAULT, first_id: int = 0,
                   draws=None, max_attempts: int = 256, device="cuda"):
    """On-device random plan generator (``dmp_plans_generate``): the reference's ``create_plan`` of the generator
    classes -- 1D random sinusoid (Env/1D/DMP_Env_1D_dynamic_hindsight_replay.py:29-42), 2D/3D random dense / sparse
    triangles (Env/2D/DMP_Env_2D_dynamic_hindsight_replay_usedata.py:37-59) -- for plan ids
    ``first_id .. first_id + n_plans - 1``.

    draws  None: counter-based Philox stream keyed by (seed, plan id); otherwise the reference's numpy draws:
           1D float64 [n,3] = (k_1, k_2, phase); 2D/3D int32 [n,A,6] = (x0,x1,x2,y0,y1,y2) of each of A attempts.
    Returns (table uint8 [n, row_bytes], totals int32 [n], aux) with aux = the 1D parameters used (float64 [n,3],
    the reference's ``one_hot``) or the number of attempts per plan (int32 [n]).  Raises ValueError for a bad
    plan_choose (like the reference) or if an injected vertex is outside the 20x20 grid / every attempt was rejected."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("snac_b200 has no CPU path; device must be a CUDA device")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    if dim != 1 and plan_choose not in (0, 1):
        raise ValueError(' 0: Dense triangle, 1: Sparse triangle')
    lay = L.DmpLayout()
    L.check(L.lib.dmp_layout(dim, 1, C.byref(lay)), "dmp_layout")
    n = int(n_plans)
    with torch.cuda.device(dev):
        table = torch.zeros((n, lay.plan_row_bytes), dtype=torch.uint8, device=dev)
        totals = torch.zeros(n, dtype=torch.int32, device=dev)
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        d = None
        if dim == 1:
            aux = torch.zeros((n, 3), dtype=torch.float64, device=dev)
            if draws is not None:
                d = torch.as_tensor(np.ascontiguousarray(draws, dtype=np.float64), device=dev).reshape(n, 3).contiguous()
        else:
            aux = torch.zeros(n, dtype=torch.int32, device=dev)
            if draws is not None:
                d = torch.as_tensor(np.ascontiguousarray(draws, dtype=np.int32), device=dev)
                if d.dim() == 2:
                    d = d[:, None, :]
                if d.dim() != 3 or d.shape[0] != n or d.shape[2] != 6:
                    raise ValueError("draws must have shape [n_plans, attempts, 6]")
                d = d.contiguous()
                max_attempts = int(d.shape[1])
        L.check(L.lib.dmp_plans_generate(dim, int(plan_choose), C.c_uint64(int(seed)), int(first_id), n, _ptr(d),
                                         int(max_attempts), table.data_ptr(), totals.data_ptr(), aux.data_ptr(),
                                         err.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "dmp_plans_generate")
        if int(err.item()):
            raise ValueError("plan generation failed: a vertex outside 0..19 or no accepted triangle in %d attempts" % max_attempts)
    return table, totals, aux


class BatchedDMPEnv:
    """N independent DMP environments on one GPU.

    dim               1, 2 or 3
    plan_choose       static plan selector (1D: 0 sine / 1 Gaussian / 2 step; 2D, 3D: 0 dense / 1 sparse)
    plans             dynamic envs: float64 array of plans in the reference's format ([n,30] or [n,26,26]), or the
                      string "generate": n_plans random plans from the on-device generators (generate_plans();
                      plan_choose selects dense / sparse triangles), or a (table, totals) pair of device tensors
    n_plans           size of the generated plan table (plans="generate" only; at most 65 535)
    random_choose_paln  (sic, reference spelling) True: random plan per reset; False: sequential with wrap
    auto_reset        fold finished episodes into the per-env statistics and reset them inside step()
    reset_obs         with auto_reset: the observation returned for an env whose episode ends in a step is the one its reset
                      returns (the next episode's first policy input, as gym's vector envs do) instead of the terminal
                      observation; reward / done are the finished episode's (DMP_F_RESET_OBS).  Default: False
    obs_dtype         torch.float32 (default), torch.float64 (the reference's dtype), torch.int16, or "record": packed step
                      records (window + counters + reward + done in one uint8 [N, rec_bytes] buffer: the compact
                      host-facing kind, see record_dtype() / unpack_records()), or "bits" (2D / 3D): the same content
                      bit-packed into 16 B / 32 B per env (DMP_OBS_BITS; unpack_bits() / unpack_records_device())
    normalise         emit the dynamic classes' normalised counter columns (default: False = raw counters)
    env_base          global index of env 0 (multi-GPU sharding; keeps Philox streams shard-independent)
    tuning            names of kernel tuning switches (snac_b200._lib.TUNING_FLAGS: "no_pdl", "tile_ldst", "generic",
                      "rollout_k1", "no_l2_hint"), decided once here and passed to the library as DmpIO.flags bits;
                      default: the SNAC_B200_TUNING environment variable (comma separated), read at construction
    dynamic_rules     3D only: use the dataset classes' termination rules (re-check after placement, -100 when
                      boxed in) independently of where the plan comes from -- the static *_Lnet class does that
                      (Env/3D/DMP_simulator_3d_static_circle_Lnet.py:210-236).  Default: same as `dynamic`.
    """

    def __init__(self, dim: int, *, dynamic: bool = False, plan_choose: int = 0,
                 plans: Optional[np.ndarray] = None, num_envs: int = 1, device="cuda",
                 random_choose_paln: bool = True, auto_reset: bool = False,
                 obs_dtype: torch.dtype = torch.float32, normalise: bool = False,
                 seed: int = L.SEED_DEFAULT, env_base: int = 0, action_dist: str = "uniform",
                 total_step: Optional[int] = None, dynamic_rules: Optional[bool] = None, n_plans: int = 4096,
                 plan_id_base: int = 0, tuning: Optional[Sequence[str]] = None, reset_obs: bool = False):
        if dim not in (1, 2, 3):
            raise ValueError("dim must be 1, 2 or 3")
        self.records = isinstance(obs_dtype, str) and obs_dtype in (RECORD, BITS)
        self.record_kind = obs_dtype if self.records else None
        if self.records:
            if obs_dtype == BITS and dim == 1:
                raise ValueError('obs_dtype="bits" exists for 2D and 3D envs; the 1D step record ("record") is 16 B already')
            obs_dtype = torch.uint8
        elif obs_dtype not in (torch.float32, torch.float64, torch.int16):
            raise ValueError('obs_dtype must be float32, float64, int16, "record" or "bits"')
        if normalise and obs_dtype in (torch.int16, torch.uint8):
            raise ValueError("normalised counters need a floating obs_dtype")
        if tuning is None:
            tuning = [x for x in os.environ.get("SNAC_B200_TUNING", "").split(",") if x]
        if os.environ.get("SNAC_B200_L2_HINTS", "1") == "0":
            tuning = list(tuning) + ["no_l2_hint"]
        unknown = [x for x in tuning if x not in L.TUNING_FLAGS]
        if unknown:
            raise ValueError("unknown tuning switches %s (known: %s)" % (unknown, sorted(L.TUNING_FLAGS)))
        self.tuning = tuple(tuning)
        self._tuning_flags = 0
        for x in self.tuning:
            self._tuning_flags |= L.TUNING_FLAGS[x]
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("snac_b200 has no CPU path; device must be a CUDA device")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.dim, self.dynamic = dim, bool(dynamic)
        self.l2_hints = not (self._tuning_flags & L.F_NO_L2_HINT)
        self.num_envs = int(num_envs)
        self.auto_reset = bool(auto_reset)
        self.reset_obs = bool(reset_obs)
        if self.reset_obs and not self.auto_reset:
            raise ValueError("reset_obs needs auto_reset=True")
        self.obs_dtype, self.normalise = obs_dtype, bool(normalise)
        for k, v in _SPEC[dim].items():
            setattr(self, k, v)
        lay = L.DmpLayout()
        L.check(L.lib.dmp_layout(dim, self.num_envs, C.byref(lay)), "dmp_layout")
        self._lay = lay
        self.total_step = int(total_step) if total_step is not None else (
            lay.total_step_dynamic if self.dynamic else lay.total_step_static)
        self.environment_width = self.plan_width + 2 * self.HALF_WINDOW_SIZE
        self.plan_choose = plan_choose

        dev, n = self.device, self.num_envs
        with torch.cuda.device(dev):
            u8 = dict(dtype=torch.uint8, device=dev)
            self._cells = torch.zeros(lay.cells_bytes, **u8)
            self._aux = torch.zeros(max(lay.aux_bytes, 16), **u8)
            self._err = torch.zeros(1, dtype=torch.int32, device=dev)
            self._ep_cnt = torch.zeros(n, dtype=torch.int32, device=dev)
            self._ep_len = torch.zeros(n, dtype=torch.int32, device=dev)
            self._ep_ret = torch.zeros(n, dtype=torch.float64, device=dev)
            self._ep_iou = torch.zeros(n, dtype=torch.float64, device=dev)
            self._t_dev = torch.zeros(2, dtype=torch.int64, device=dev)
            self._stats = torch.zeros(4, dtype=torch.float64, device=dev)
            self._stats_scratch = torch.zeros(int(L.lib.dmp_stats_scratch_bytes(n)), **u8)
            # ---- plan table -----------------------------------------------------------------
            if self.dynamic and (isinstance(plans, str) or isinstance(plans, tuple)):
                if isinstance(plans, str):
                    if plans != "generate":
                        raise ValueError('plans must be an array, a (table, totals) pair or "generate"')
                    table, totals, self.plan_aux = generate_plans(dim, n_plans, plan_choose, seed=seed,
                                                                  first_id=plan_id_base, device=dev)
                else:
                    table, totals = plans
                self._plans_raw = None
                self._install_plans(table, totals)
            elif self.dynamic:
                if plans is None:
                    raise ValueError("dynamic envs need plans= (see load_plan_dataset)")
                raw = torch.as_tensor(np.ascontiguousarray(plans, dtype=np.float64), device=dev)
                want = (30,) if dim == 1 else (26, 26)
                if tuple(raw.shape[1:]) != want:
                    raise ValueError("plans must have shape [n,%s]" % ",".join(map(str, want)))
                self.n_plans = int(raw.shape[0])
                self._plans = torch.zeros(self.n_plans * lay.plan_row_bytes, **u8)
                self._plan_total = torch.zeros(self.n_plans, dtype=torch.int32, device=dev)
                L.check(L.lib.dmp_plans_pack(dim, raw.data_ptr(), self.n_plans, self._plans.data_ptr(),
                                             self._plan_total.data_ptr(), self._stream()), "dmp_plans_pack")
                self._plans_raw = raw
            else:
                self.n_plans = 1
                self._plans = torch.zeros(lay.plan_row_bytes, **u8)
                self._plan_total = torch.zeros(1, dtype=torch.int32, device=dev)
                rc = L.lib.dmp_plan_static(dim, int(plan_choose), self._plans.data_ptr(),
                                           self._plan_total.data_ptr(), self._stream())
                if rc == L.EINVAL:       # the reference raises at reset(); we raise at construction
                    raise ValueError('0: Sin, 1: Gaussian, 2: Step' if dim == 1 else '0: Dense circle, 1: Sparse circle')
                L.check(rc, "dmp_plan_static")
                self._plans_raw = None
            # ---- output buffers (reused by step(); rollout() allocates [K,...] on demand) -----
            D = self.obs_row
            self._obs = torch.zeros((n, D), dtype=obs_dtype, device=dev)
            self._reward = torch.zeros(n, dtype=torch.float32, device=dev)
            self._done = torch.zeros(n, dtype=torch.uint8, device=dev)

        st = L.DmpState()
        self.dynamic_rules = self.dynamic if dynamic_rules is None else bool(dynamic_rules)
        st.dim, st.dynamic, st.n_plans, st.total_step = dim, int(self.dynamic_rules), self.n_plans, self.total_step
        st.plan_mode = (L.PLAN_KEEP if not self.dynamic else
                        (L.PLAN_PHILOX if random_choose_paln else L.PLAN_SEQUENTIAL))
        if action_dist == "ref3d" and dim != 3:
            raise ValueError('action_dist="ref3d" is the 3D envs\' own action distribution '
                             '(Env/3D/DMP_simulator_3d_static_circle.py:361-362); 1D/2D envs draw uniformly')
        st.action_dist = {"uniform": L.ACT_UNIFORM, "ref3d": L.ACT_REF3D}[action_dist]
        st.n_envs, st.env_base, st.seed, st.t = n, int(env_base), int(seed), 0
        st.t_dev = None
        st.cells, st.aux = self._cells.data_ptr(), self._aux.data_ptr()
        st.plans, st.plan_total = self._plans.data_ptr(), self._plan_total.data_ptr()
        st.ep_cnt, st.ep_len = self._ep_cnt.data_ptr(), self._ep_len.data_ptr()
        st.ep_ret, st.ep_iou = self._ep_ret.data_ptr(), self._ep_iou.data_ptr()
        st.err = self._err.data_ptr()
        self._st = st
        self.random_choose_paln = bool(random_choose_paln)
        self._needs_initial_reset = True
        import torch.cuda.nvtx as _nvtx
        self._nvtx = _nvtx

    # ------------------------------------------------------------------------------------------
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    @property
    def obs_dim(self) -> int:
        return self._lay.obs_dim

    @property
    def obs_row(self) -> int:
        """Elements of obs_dtype per env in an observation buffer: obs_dim, or the record size in bytes."""
        if self.records:
            return self._lay.bits_bytes if self.record_kind == BITS else self._lay.rec_bytes
        return self._lay.obs_dim

    def _kind_of(self, dtype) -> int:
        """DMP_OBS_* of an observation buffer of this env with the given torch dtype."""
        if dtype == torch.uint8:
            return L.OBS_BITS if self.record_kind == BITS else L.OBS_REC
        return _TORCH_OBS[dtype]

    @property
    def t(self) -> int:
        return int(self._st.t)

    def _flags(self, K: int = 1) -> int:
        # L2 evict_last hints for the state pay off when the state can actually stay resident: measured on
        # B200 (profiles/README.md) they help K>1 rollouts and shards <= ~40 MB, and cost ~4 % on a 67 MB
        # state stepped one launch at a time (the two L2 partitions hold less than 126 MB of distinct lines;
        # fractional evict_last / evict_first policies that keep a half or a quarter of the lines were slower still).
        hint = self.l2_hints and (K > 1 or self._cells.numel() <= 40 * 1024 * 1024)
        return ((L.F_AUTORESET if self.auto_reset else 0) | (L.F_NORMALISE if self.normalise else 0)
                | (L.F_RESET_OBS if self.reset_obs else 0)
                | (self._tuning_flags & ~L.F_NO_L2_HINT) | (0 if hint else L.F_NO_L2_HINT))

    def _u8(self, x, shape, what) -> torch.Tensor:
        if not torch.is_tensor(x):
            x = torch.as_tensor(np.asarray(x), device=self.device)
        if x.device != self.device and not (x.device.type == "cpu" and x.is_pinned()):
            x = x.to(self.device, non_blocking=True)        # pinned host tensors are mapped: the kernel reads them in place
        if x.dtype != torch.uint8:
            x = x.to(torch.uint8)
        if tuple(x.shape) != tuple(shape):
            raise ValueError("%s must have shape %s, got %s" % (what, tuple(shape), tuple(x.shape)))
        return x.contiguous()

    def _i32(self, x, shape, what) -> torch.Tensor:
        if not torch.is_tensor(x):
            x = torch.as_tensor(np.asarray(x), device=self.device)
        x = x.to(device=self.device, dtype=torch.int32)
        if tuple(x.shape) != tuple(shape):
            raise ValueError("%s must have shape %s, got %s" % (what, tuple(shape), tuple(x.shape)))
        return x.contiguous()

    # ------------------------------------------------------------------------------------------
    def reset(self, mask=None, plan_idx=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Reset all envs (or those with mask != 0).  Returns the observation buffer [N, D]
        (rows of envs that were not reset keep their previous content).
        Dynamic envs: ``plan_idx`` (int32 [N]) injects the reference's ``index_random`` draw
        (Env/2D/DMP_Env_2D_dynamic_usedata_plan.py:36); default is a Philox draw, or -- with
        random_choose_paln=False -- plan 0 on an env's first reset and +1 (wrapping) afterwards (in the kernel)."""
        n = self.num_envs
        m = None if mask is None else self._u8(mask, (n,), "mask")
        p = None if plan_idx is None else self._i32(plan_idx, (n,), "plan_idx")
        obs = self._obs if out is None else out
        self._nvtx.range_push("dmp_reset")
        with torch.cuda.device(self.device):
            L.check(L.lib.dmp_reset(C.byref(self._st), _ptr(m), _ptr(p), C.c_uint64(L.T_INIT - self._st.t),
                                    obs.data_ptr(), self._kind_of(obs.dtype), self._stream()), "dmp_reset")
        self._nvtx.range_pop()
        if mask is None:
            self._needs_initial_reset = False
        self._keep = (m, p)
        return obs

    def reset_at(self, env_index: int) -> torch.Tensor:
        """multiprocess.py:22-23 -- reset one env, return its observation row [D]."""
        m = torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device)
        m[env_index] = 1
        return self.reset(mask=m)[env_index]

    def step(self, actions, step_sizes=None, next_plan=None):
        """One step of every env.  ``actions`` uint8 [N] (None: Philox synthetic actions).
        ``step_sizes`` uint8 [N] in {1,2,3} injects the reference's ``np.random.randint(1, 4)``
        draw (the *_hindsight_replay ``step(action, step_size)`` form); None = Philox.
        Returns (obs [N,D] (records: uint8 [N, rec_bytes]), reward f32 [N], done bool [N]) -- views of internal
        buffers that the next call overwrites."""
        obs, rew, done = self.rollout(1, actions=None if actions is None else self._u8(actions, (self.num_envs,), "actions")[None],
                                      step_sizes=None if step_sizes is None else self._u8(step_sizes, (self.num_envs,), "step_sizes")[None],
                                      next_plan=None if next_plan is None else self._i32(next_plan, (self.num_envs,), "next_plan")[None],
                                      out=(self._obs[None], self._reward[None], self._done[None]))
        return obs[0], rew[0], done[0]

    def rollout(self, K: int, actions=None, step_sizes=None, next_plan=None, out=None,
                materialise_obs: bool = True, use_device_t: bool = False, t_slot: int = 0):
        """Advance every env K steps in ONE kernel launch (state stays on chip between steps).
        actions / step_sizes: uint8 [K,N] or None (Philox).  Returns (obs [K,N,D], reward [K,N], done [K,N])."""
        n, D = self.num_envs, self.obs_row
        if self._needs_initial_reset:
            raise RuntimeError("call reset() before step()/rollout()")
        a = None if actions is None else self._u8(actions, (K, n), "actions")
        s = None if step_sizes is None else self._u8(step_sizes, (K, n), "step_sizes")
        p = None if next_plan is None else self._i32(next_plan, (K, n), "next_plan")
        if out is None:
            obs = torch.empty((K, n, D), dtype=self.obs_dtype, device=self.device) if materialise_obs else None
            rew = torch.empty((K, n), dtype=torch.float32, device=self.device)
            done = torch.empty((K, n), dtype=torch.uint8, device=self.device)
        else:
            obs, rew, done = out
        io = L.DmpIO()
        io.actions, io.step_sizes, io.next_plan = _ptr(a), _ptr(s), _ptr(p)
        io.obs, io.reward, io.done = _ptr(obs), _ptr(rew), _ptr(done)
        io.obs_kind = self._kind_of(self.obs_dtype if obs is None else obs.dtype)
        io.flags = self._flags(K) | (L.F_TSLOT1 if (use_device_t and t_slot) else 0)
        self._st.t_dev = self._t_dev.data_ptr() if use_device_t else None
        self._nvtx.range_push("dmp_step" if K == 1 else "dmp_rollout")
        with torch.cuda.device(self.device):
            L.check(L.lib.dmp_rollout(C.byref(self._st), C.byref(io), int(K), self._stream()), "dmp_rollout")
        self._nvtx.range_pop()
        if not use_device_t:
            self._st.t = self._st.t + K
        self._keep_io = (a, s, p)
        return obs, rew, (None if done is None else done.view(torch.bool))

    STAGES = ("move", "deposit", "observe", "reward", "done_reset")

    def step_staged(self, actions, step_sizes=None, next_plan=None, stages=STAGES):
        """The same step as step(), executed as the five standalone stage kernels
        (a) move, (b) deposit, (c) observe, (d) reward, (e) done/reset -- one launch each
        (include/dmp.h "stage kernels").  For unit parity and per-stage timing; the fused
        step()/rollout() is the production path."""
        n = self.num_envs
        if self._needs_initial_reset:
            raise RuntimeError("call reset() before step_staged()")
        a = None if actions is None else self._u8(actions, (n,), "actions")
        s = None if step_sizes is None else self._u8(step_sizes, (n,), "step_sizes")
        p = None if next_plan is None else self._i32(next_plan, (n,), "next_plan")
        if not hasattr(self, "_stage_scratch"):
            self._stage_scratch = torch.zeros((n, 4), dtype=torch.int32, device=self.device)
        io = L.DmpIO()
        io.actions, io.step_sizes, io.next_plan = _ptr(a), _ptr(s), _ptr(p)
        io.obs, io.reward, io.done = self._obs.data_ptr(), self._reward.data_ptr(), self._done.data_ptr()
        io.obs_kind, io.flags = self._kind_of(self.obs_dtype), self._flags()
        self._st.t_dev = None
        fns = {"move": L.lib.dmp_stage_move, "deposit": L.lib.dmp_stage_deposit, "observe": L.lib.dmp_stage_observe,
               "reward": L.lib.dmp_stage_reward, "done_reset": L.lib.dmp_stage_done_reset}
        with torch.cuda.device(self.device):
            for name in stages:
                if name == "done_reset" and not self.auto_reset:
                    continue
                L.check(fns[name](C.byref(self._st), C.byref(io), self._stage_scratch.data_ptr(), self._stream()),
                        "dmp_stage_" + name)
        if "move" in stages:
            self._st.t = self._st.t + 1
        return self._obs, self._reward, self._done.view(torch.bool)

    # ------------------------------------------------------------------------------------------
    def iou(self) -> torch.Tensor:
        """Per-env IoU of the current grid vs. its plan, float64 [N]
        (Env/1D/DMP_Env_1D_static.py:138-151, Env/2D/DMP_Env_2D_static.py:169-175,
        Env/3D/DMP_simulator_3d_static_circle.py:257-276)."""
        out = torch.empty(self.num_envs, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            L.check(L.lib.dmp_iou(C.byref(self._st), out.data_ptr(), self._stream()), "dmp_iou")
        return out

    def stats(self, allreduce: bool = False) -> torch.Tensor:
        """float64 [4] = (sum of episode returns, sum of final IoUs, episodes, steps) over the finished
        episodes of this shard; ``allreduce=True`` sums it over all ranks (NCCL over NVLink)."""
        with torch.cuda.device(self.device):
            L.check(L.lib.dmp_stats_reduce(C.byref(self._st), self._stats.data_ptr(),
                                           self._stats_scratch.data_ptr(), self._stream()), "dmp_stats_reduce")
        if allreduce:
            from .sharding import allreduce_stats
            allreduce_stats(self._stats)
        return self._stats

    def clear_stats(self) -> None:
        with torch.cuda.device(self.device):
            L.check(L.lib.dmp_stats_clear(C.byref(self._st), self._stream()), "dmp_stats_clear")

    def episode_stats(self):
        """Per-env (episodes, total length, sum of returns, sum of IoUs) tensors."""
        return self._ep_cnt, self._ep_len, self._ep_ret, self._ep_iou

    def errors(self) -> int:
        """OR of DMP_ERR_* bits latched by the kernels so far (synchronises)."""
        return int(self._err.item())

    def check_errors(self) -> None:
        e = self.errors()
        if e & L.ERR_ACTION:
            raise UnboundLocalError("an action outside the env's action set was stepped "
                                    "(the reference leaves 'position' unbound, Env/1D/DMP_Env_1D_static.py:130-133)")
        if e & L.ERR_OVERFLOW:
            raise OverflowError("count_step or count_brick of an env passed 65 535 (envs stepped on after done without a "
                                "reset): the packed counters hold 16 bits and have saturated")
        if e:
            raise ValueError("libdmp latched error bits 0x%x" % e)

    # ------------------------------------------------------------------------------------------
    def export_state(self) -> dict:
        """Dense copy of the state in the reference's own shapes (device tensors):
        grid int32 [N,34] / [N,26,26] (environment_memory, -1 frame), scalars int32 [N,8]
        (pos_row|pos, pos_col, count_brick, count_step, plan_idx, total_brick, 0, 0), ret f32 [N]."""
        n, lay = self.num_envs, self._lay
        grid = torch.empty((n, lay.grid_rows, lay.grid_cols), dtype=torch.int32, device=self.device)
        sc = torch.empty((n, 8), dtype=torch.int32, device=self.device)
        ret = torch.empty(n, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            L.check(L.lib.dmp_export_state(C.byref(self._st), grid.data_ptr(), sc.data_ptr(), ret.data_ptr(),
                                           self._stream()), "dmp_export_state")
        if self.dim == 1:
            grid = grid.reshape(n, 1, lay.grid_cols)
        return dict(grid=grid, scalars=sc, ret=ret)

    def import_state(self, grid=None, scalars=None, ret=None) -> None:
        """Inverse of export_state (any subset)."""
        n, lay = self.num_envs, self._lay
        g = None if grid is None else self._i32(torch.as_tensor(grid).reshape(n, lay.grid_rows, lay.grid_cols),
                                                (n, lay.grid_rows, lay.grid_cols), "grid")
        s = None if scalars is None else self._i32(scalars, (n, 8), "scalars")
        r = None if ret is None else torch.as_tensor(ret, dtype=torch.float32, device=self.device).contiguous()
        with torch.cuda.device(self.device):
            L.check(L.lib.dmp_import_state(C.byref(self._st), _ptr(g), _ptr(s), _ptr(r), self._stream()),
                    "dmp_import_state")
        self._needs_initial_reset = False

    def get_state(self) -> dict:
        """Raw SoA snapshot (cheap clone of the packed tensors) -- the MCTS variants' state tuple
        (Env/1D/DMP_Env_1D_static_MCTS.py:87) in batched form."""
        return dict(cells=self._cells.clone(), aux=self._aux.clone(), t=int(self._st.t),
                    ep=(self._ep_cnt.clone(), self._ep_len.clone(), self._ep_ret.clone(), self._ep_iou.clone()))

    def set_state(self, state: dict) -> None:
        self._cells.copy_(state["cells"])
        self._aux.copy_(state["aux"])
        self._st.t = int(state["t"])
        for dst, src in zip((self._ep_cnt, self._ep_len, self._ep_ret, self._ep_iou), state["ep"]):
            dst.copy_(src)
        self._needs_initial_reset = False

    def transition(self, state: dict, actions, step_sizes=None):
        """Functional step of the MCTS variants, batched: (state, action) -> (state', obs, reward, done)
        (Env/1D/DMP_Env_1D_static_MCTS.py:94-145, Env/2D/DMP_ENV_2D_static_MCTS.py:110-169).  `state` is a
        get_state() snapshot; the env's own state is replaced by it, stepped, and the successor returned."""
        self.set_state(state)
        obs, rew, done = self.step(actions, step_sizes)
        return self.get_state(), obs.clone(), rew.clone(), done.clone()

    def transition_dense(self, position, grid, count_brick, count_step, actions, step_sizes=None, plan_idx=None):
        """The MCTS variants' ``transition(state, action)`` for N tree nodes in ONE launch, with the state in the
        reference's own tuple format (Env/2D/DMP_ENV_2D_static_MCTS.py:110-169,
        Env/3D/DMP_simulator_3d_static_circle_MCTS.py:215-289): ``position`` [N] (1D) or [N,2], ``grid`` =
        environment_memory incl. its -1 frame ([N,1,34] / [N,26,26]), ``count_brick`` / ``count_step`` [N].
        ``plan_idx`` [N] selects each node's plan row (default: the rows the envs already point at).
        Returns (position', grid', count_brick', count_step', obs, reward, done) as device tensors; this env's
        state is replaced by the successor states (episode statistics are not touched: use auto_reset=False)."""
        n = self.num_envs
        if self.auto_reset:
            raise RuntimeError("transition_dense() needs an env built with auto_reset=False")
        pos = torch.as_tensor(position, device=self.device).to(torch.int32).reshape(n, -1)
        sc = torch.zeros((n, 8), dtype=torch.int32, device=self.device)
        sc[:, 0] = pos[:, 0]
        if self.dim != 1:
            sc[:, 1] = pos[:, 1]
        sc[:, 2] = torch.as_tensor(count_brick, device=self.device).to(torch.int32).reshape(n)
        sc[:, 3] = torch.as_tensor(count_step, device=self.device).to(torch.int32).reshape(n)
        if plan_idx is None:
            sc[:, 4] = self.export_state()["scalars"][:, 4] if not self._needs_initial_reset else 0
        else:
            sc[:, 4] = torch.as_tensor(plan_idx, device=self.device).to(torch.int32).reshape(n)
        g = torch.as_tensor(grid, device=self.device).to(torch.int32)
        self.import_state(grid=g, scalars=sc, ret=torch.zeros(n, dtype=torch.float32, device=self.device))
        obs, rew, done = self.step(actions, step_sizes)
        st = self.export_state()
        s2 = st["scalars"]
        npos = s2[:, 0].clone() if self.dim == 1 else s2[:, 0:2].clone()
        return npos, st["grid"], s2[:, 2].clone(), s2[:, 3].clone(), obs.clone(), rew.clone(), done.clone()

    # plan sources --------------------------------------------------------------------------------
    def _install_plans(self, table: torch.Tensor, totals: torch.Tensor) -> None:
        table = table.to(self.device).contiguous().view(torch.uint8).reshape(-1, self._lay.plan_row_bytes)
        totals = totals.to(device=self.device, dtype=torch.int32).contiguous()
        if table.shape[0] != totals.shape[0] or not 1 <= table.shape[0] <= 65535:
            raise ValueError("plan table needs 1..65535 rows and one budget per row")
        self.n_plans = int(table.shape[0])
        self._plans, self._plan_total = table.reshape(-1), totals
        if hasattr(self, "_st"):
            self._st.n_plans = self.n_plans
            self._st.plans, self._st.plan_total = self._plans.data_ptr(), self._plan_total.data_ptr()

    def set_plan_table(self, table: torch.Tensor, totals: torch.Tensor) -> None:
        """Swap the plan table (packed rows as produced by generate_plans() / hindsight_plans() / plan_table()).
        Envs keep their plan index; call reset() (or pass plan_idx) afterwards."""
        if not self.dynamic:
            raise RuntimeError("static envs have a fixed plan; build the env with dynamic=True")
        self._install_plans(table, totals)

    def regenerate_plans(self, first_id: int, n_plans: Optional[int] = None, draws=None) -> None:
        """Fresh random plans for ids first_id.. (removes the reference's 500-plan dataset limit)."""
        table, totals, self.plan_aux = generate_plans(self.dim, n_plans or self.n_plans, self.plan_choose,
                                                      seed=int(self._st.seed), first_id=first_id, draws=draws,
                                                      device=self.device)
        self.set_plan_table(table, totals)

    def hindsight_plans(self):
        """Hindsight relabelling: (table [N,row_bytes], totals [N]) where row i is what env i has built so far --
        ``env_hindsight.plan = env.environment_memory[...]`` of script/DRQN_hindsight/1d/DRQN_hindsight_1D_static.py:242-245.
        Feed it to a second env with set_plan_table() and reset(plan_idx=arange(N)) to replay the episode against it."""
        table = torch.empty((self.num_envs, self._lay.plan_row_bytes), dtype=torch.uint8, device=self.device)
        totals = torch.empty(self.num_envs, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            L.check(L.lib.dmp_plans_from_state(C.byref(self._st), table.data_ptr(), totals.data_ptr(), self._stream()),
                    "dmp_plans_from_state")
        return table, totals

    # plan views ----------------------------------------------------------------------------------
    def plan_table(self) -> torch.Tensor:
        """The packed plan table as uploaded (uint8 view; layout in include/dmp.h)."""
        return self._plans.view(self.n_plans, -1)

    def plan_totals(self) -> torch.Tensor:
        return self._plan_total

    def plans_dense(self) -> np.ndarray:
        """Plans in the reference's array format (float64 [n_plans,30] or [n_plans,26,26])."""
        tab = self.plan_table().cpu().numpy()
        if self.dim == 1:
            return tab[:, :30].astype(np.float64)
        out = np.zeros((self.n_plans, 26, 26))
        if self.dim == 2:
            bits = np.unpackbits(tab[:, :52].copy(), axis=1, bitorder="little")[:, :400]
            out[:, 3:23, 3:23] = bits.reshape(-1, 20, 20)
        else:
            out[:, 3:23, 3:23] = tab[:, :400].reshape(-1, 20, 20)
        return out
