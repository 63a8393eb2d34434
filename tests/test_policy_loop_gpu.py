"""Device-resident acting loop (snac_b200/policy_loop.py): the CUDA-graph replay of
obs -> policy -> dmp_step equals stepping the oracle with the same policy evaluated on the CPU."""
import numpy as np
import pytest
import torch

from oracle import dmp_oracle as O
from oracle import philox
from oracle_batch import OracleBatch

pytestmark = pytest.mark.gpu
SEED = 0x534E4143


def int_policy(A):
    """A deterministic policy whose arithmetic is exact in every float type: (sum of the window + step counter) mod A."""
    def f(obs):
        w = obs[:, :-2].sum(dim=1) + obs[:, -1]
        return torch.remainder(w.to(torch.int64), A).to(torch.uint8)
    return f


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("graph", [True, False])
@pytest.mark.parametrize("reset_obs", [True, False], ids=["kernel_reset_obs", "patched"])
def test_loop_matches_oracle(dim, graph, reset_obs):
    """reset_obs=True: the kernel writes the reset observation for finished envs (next_obs IS obs[1:]); False: next_obs
    keeps the terminal observation and obs[t + 1] is patched.  Either way the policy acts on the reset observation after
    an episode ends."""
    from snac_b200.policy_loop import DeviceRollout
    from snac_b200.vecenv import BatchedDMPEnv
    n, T, A = 80, 30, O.SPEC[dim]["actions"]
    env = BatchedDMPEnv(dim, plan_choose=0, num_envs=n, auto_reset=True, obs_dtype=torch.float32, seed=SEED, env_base=5,
                        reset_obs=reset_obs, total_step=None if dim == 3 else 25)
    ob = OracleBatch(dim, False, n, 0)
    if dim != 3:
        for e in ob.envs:
            e.total_step = 25                                 # short episodes: every env is reset inside each collect()
    o = env.reset()
    cur = ob.reset()
    assert np.array_equal(o.cpu().numpy().astype(np.float64), cur)
    reset_row = cur[0].copy()                                 # reset() returns the same row for every env
    assert (cur == reset_row).all()
    loop = DeviceRollout(env, int_policy(A), horizon=T, graph=graph)
    assert np.array_equal(loop.reset_obs.cpu().numpy().astype(np.float64), reset_row)
    t, n_done = 0, 0
    for rep in range(3):                                     # replays continue where the previous one stopped
        traj = loop.collect()
        torch.cuda.synchronize()
        for k in range(T):
            a = (cur[:, :-2].sum(1) + cur[:, -1]).astype(np.int64) % A
            s, _, _ = philox.draws(SEED, np.arange(5, 5 + n), t, A)
            assert np.array_equal(traj["obs"][k].cpu().numpy().astype(np.float64), cur), (rep, k)
            assert np.array_equal(traj["actions"][k].cpu().numpy(), a), (rep, k)
            nxt, r, d = ob.step(a, s)
            # next_obs is the terminal observation the step returned; on done the oracle batch has reset the env, and the
            # NEXT policy input of that env is the reset observation -- ``state = env.reset()`` in the reference's loops
            cur = np.where(d[:, None], reset_row[None, :], nxt)
            assert np.array_equal(traj["next_obs"][k].cpu().numpy().astype(np.float64), cur if reset_obs else nxt), (rep, k)
            assert np.array_equal(traj["reward"][k].cpu().numpy(), r)
            assert np.array_equal(traj["done"][k].cpu().numpy(), d)
            n_done += int(d.sum())
            t += 1
    assert env.t == 3 * T
    assert n_done > (n if dim != 3 else 20)                  # episodes finished all over the batch: the reset rows were exercised
    g_ref, sc_ref = ob.export()
    st = env.export_state()
    assert np.array_equal(st["grid"].cpu().numpy().reshape(g_ref.shape), g_ref)
    assert np.array_equal(st["scalars"].cpu().numpy()[:, :4], sc_ref[:, :4])
    env.check_errors()


def test_epsilon_greedy_qsa_graph_equals_eager():
    """The reference's critic form Q(s, a) scored for all actions at once; graph replay == eager loop."""
    from snac_b200.policy_loop import DeviceRollout, EpsilonGreedy, QSAAdapter
    from snac_b200.vecenv import BatchedDMPEnv

    class QNet(torch.nn.Module):                             # the layer sizes of script/DQN/2d/DQN_2d_static.py:78-98
        def __init__(self):
            super().__init__()
            self.f = torch.nn.Sequential(torch.nn.Linear(52, 64), torch.nn.ReLU(), torch.nn.Linear(64, 128), torch.nn.ReLU(),
                                         torch.nn.Linear(128, 128), torch.nn.ReLU(), torch.nn.Linear(128, 1))

        def forward(self, s, a):
            # a state-dependent phase on top of the MLP, so that the greedy action of a random-init net varies
            return self.f(torch.cat((s, a), dim=1)) + torch.sin(0.37 * s.sum(dim=1, keepdim=True) + 1.3 * a)

    torch.manual_seed(0)
    q = QSAAdapter(QNet().cuda(), 5)
    outs = []
    for graph in (True, False):
        env = BatchedDMPEnv(2, plan_choose=0, num_envs=256, auto_reset=True, seed=SEED)
        env.reset()
        loop = DeviceRollout(env, EpsilonGreedy(q, 5, epsilon=0.0), horizon=16, graph=graph)
        a = {k: v.clone() for k, v in loop.collect().items()}
        b = {k: v.clone() for k, v in loop.collect().items()}
        torch.cuda.synchronize()
        outs.append((a, b))
        assert a["actions"].max() <= 4 and len(torch.unique(a["actions"])) > 1
    for x, y in zip(outs[0], outs[1]):
        for k in x:
            assert torch.equal(x[k], y[k]), k


def test_random_policy_and_kernel_actions_run():
    from snac_b200.policy_loop import DeviceRollout, RandomPolicy
    from snac_b200.vecenv import BatchedDMPEnv
    env = BatchedDMPEnv(3, plan_choose=0, num_envs=500, auto_reset=True, seed=SEED)
    env.reset()
    for pol in (RandomPolicy(8), None):
        loop = DeviceRollout(env, pol, horizon=40)
        tr = loop.collect()
        tr = loop.collect()
        torch.cuda.synchronize()
        assert tr["obs"].shape == (40, 500, 51) and tr["done"].dtype == torch.bool
    assert env.stats()[2].item() > 0                           # episodes finished and were reset inside the graph
    env.check_errors()
