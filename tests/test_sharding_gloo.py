"""Multi-rank host logic on CPU (gloo, world_size 2): shard bounds, shard-independent Philox streams,
and the statistics all-reduce == a single-process run over all envs.  The per-env arithmetic is done by
the C oracle here (no GPU in this container); the GPU twin is test_parity_gpu.py::test_shard_independence_2d."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_plans
from oracle import dmp_oracle as O
from oracle import philox

SEED = 0x534E4143


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run_shard(dim, dynamic, base, count, K, plans):
    """Oracle rollout of envs [base, base+count) with the kernels' Philox streams; returns stats[4] and obs."""
    from oracle.c_oracle import COracleBatch
    cb = COracleBatch(dim, dynamic, count, 0, plans)
    ids = np.arange(base, base + count)
    p0 = philox.reset_draw(SEED, ids, 0, cb.n_plans) if dynamic else None
    cb.reset(p0)
    A = O.SPEC[dim]["actions"]
    acts = np.zeros((K, count), np.uint8)
    sizes = np.zeros((K, count), np.uint8)
    nxt = np.zeros((K, count), np.int32)
    for k in range(K):
        s, a, p = philox.draws(SEED, ids, k, A, cb.n_plans)
        acts[k], sizes[k], nxt[k] = a, s, p
    obs, rew, done, err = cb.rollout(acts, sizes, nxt if dynamic else None)
    assert err == 0
    stats = np.array([cb.ep_ret.sum(), cb.ep_iou.sum(), cb.ep_cnt.sum(), cb.ep_len.sum()], np.float64)
    return stats, obs


def _worker(rank, world, port, dim, dynamic, total, K, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from snac_b200.sharding import allreduce_stats, shard_bounds
        plans = load_plans(dim, "dense", "val") if dynamic else None
        base, count = shard_bounds(total, rank, world)
        stats, obs = _run_shard(dim, dynamic, base, count, K, plans)
        t = torch.from_numpy(stats.copy())
        allreduce_stats(t)
        np.save(os.path.join(out_dir, "obs_%d.npy" % rank), obs)
        np.save(os.path.join(out_dir, "stats_%d.npy" % rank), t.numpy())
        np.save(os.path.join(out_dir, "bounds_%d.npy" % rank), np.array([base, count]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dim,dynamic,total,K", [(2, False, 37, 650), (3, True, 21, 120), (1, True, 10, 800)])
def test_two_rank_sharded_run_equals_single_run(tmp_path, dim, dynamic, total, K):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, dim, dynamic, total, K, str(tmp_path)), nprocs=world, join=True)
    plans = load_plans(dim, "dense", "val") if dynamic else None
    full_stats, full_obs = _run_shard(dim, dynamic, 0, total, K, plans)
    b0, b1 = np.load(tmp_path / "bounds_0.npy"), np.load(tmp_path / "bounds_1.npy")
    assert b0[0] == 0 and b1[0] == b0[1] and b0[1] + b1[1] == total          # contiguous cover
    obs = np.concatenate([np.load(tmp_path / "obs_0.npy"), np.load(tmp_path / "obs_1.npy")], axis=1)
    assert np.array_equal(obs, full_obs)                                      # env-for-env identical
    s0, s1 = np.load(tmp_path / "stats_0.npy"), np.load(tmp_path / "stats_1.npy")
    assert np.array_equal(s0, s1)                                             # every rank holds the sum
    assert s0[0] == full_stats[0] and s0[2] == full_stats[2] and s0[3] == full_stats[3]
    assert s0[2] > 0
    assert abs(s0[1] - full_stats[1]) <= 1e-9 * max(1.0, abs(full_stats[1]))


def test_shard_bounds_cover_and_balance():
    from snac_b200.sharding import shard_bounds
    for total in (0, 1, 5, 1048576, 262144, 1000003):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (b0, c0), (b1, _) in zip(spans, spans[1:]):
                assert b0 + c0 == b1
            counts = [c for _, c in spans]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def test_allreduce_stats_is_noop_without_process_group():
    from snac_b200.sharding import allreduce_stats
    t = torch.tensor([1.0, 2.0, 3.0, 4.0], dtype=torch.float64)
    assert torch.equal(allreduce_stats(t.clone()), t)
    with pytest.raises(ValueError):
        allreduce_stats(torch.zeros(3, dtype=torch.float64))
