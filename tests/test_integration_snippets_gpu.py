"""The code shown in INTEGRATION.md runs as written (host-buffer records, tensor API with reset_obs, acting loop)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_host_buffer_record_snippet():
    from snac_b200 import BatchedDMPEnv, HostStepper
    from snac_b200.vecenv import record_dtype, unpack_records
    env = BatchedDMPEnv(2, plan_choose=0, num_envs=1 << 12, auto_reset=True, obs_dtype="record")
    env.reset()
    hs = HostStepper(env)
    actions_u8 = np.random.RandomState(0).randint(0, 5, size=1 << 12).astype(np.uint8)
    rec = hs.step(actions_u8)
    assert rec.dtype == record_dtype(2) and rec.shape == (1 << 12,) and rec.dtype.itemsize == 56
    assert rec["win"].shape == (1 << 12, 49) and rec["win"].dtype == np.uint8 and rec["win"].max() <= 2
    obs, reward, done, saturated = unpack_records(rec, dim=2)
    assert obs.shape == (1 << 12, 51) and obs.dtype == np.float64 and set(np.unique(obs[:, :49])) <= {-1.0, 0.0, 1.0}
    assert (obs[:, 50] == 1).all() and not done.any() and not saturated.any() and reward.dtype == np.float32


def test_host_buffer_bit_record_snippet():
    from snac_b200 import BatchedDMPEnv, HostStepper
    from snac_b200.vecenv import unpack_bits, unpack_records_device
    n = 1 << 13
    env = BatchedDMPEnv(2, plan_choose=0, num_envs=n, auto_reset=True, obs_dtype="bits")
    env.reset()
    hs = HostStepper(env)
    actions_u8 = np.random.RandomState(0).randint(0, 5, size=n).astype(np.uint8)
    rec = hs.step(actions_u8)
    assert rec.dtype == np.uint8 and rec.shape == (n, 16) and hs.mapped
    obs, reward, done, saturated = unpack_bits(rec[:4096], dim=2)
    assert obs.shape == (4096, 51) and obs.dtype == np.float64 and set(np.unique(obs[:, :49])) <= {-1.0, 0.0, 1.0}
    assert (obs[:, 50] == 1).all() and not done.any() and not saturated.any() and reward.dtype == np.float32
    batch = torch.as_tensor(rec[:4096]).cuda()
    obs_t, reward_t, done_t, sat_t = unpack_records_device(batch, 2, "bits", torch.float32)
    assert np.array_equal(obs_t.cpu().numpy(), obs.astype(np.float32)) and np.array_equal(reward_t.cpu().numpy(), reward)
    assert not done_t.any() and not sat_t.any()


def test_tensor_api_and_acting_loop_snippet():
    from snac_b200 import BatchedDMPEnv, DeviceRollout, RandomPolicy
    env = BatchedDMPEnv(2, plan_choose=0, num_envs=1 << 12, device="cuda:0", auto_reset=True, reset_obs=True, total_step=9)
    obs = env.reset()
    assert obs.shape == (1 << 12, 51) and obs.dtype == torch.float32
    obs, reward, done = env.rollout(9)
    assert obs.shape == (9, 1 << 12, 51) and bool(done[8].all())
    reset_row = env.reset()[0]
    env.reset()
    obs, reward, done = env.rollout(9)
    assert torch.equal(obs[8], reset_row.expand_as(obs[8]))          # finished envs return their reset observation
    loop = DeviceRollout(env, RandomPolicy(5), horizon=18)
    traj = loop.collect()
    torch.cuda.synchronize()
    assert traj["next_obs"].data_ptr() == traj["obs"][1:].data_ptr() or traj["next_obs"].shape == traj["obs"].shape
    assert traj["done"].sum() == 2 * (1 << 12)
    st = env.stats(allreduce=False)
    assert st[2].item() >= 3 * (1 << 12)
    env.check_errors()
