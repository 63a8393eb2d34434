"""Drive N independent oracle envs through the same streams as the CUDA vector env (test helper)."""
import numpy as np

from oracle import dmp_oracle as O
from oracle import philox


class OracleBatch:
    """N oracle envs with the vector env's auto-reset + statistics semantics (include/dmp.h)."""

    def __init__(self, dim, dynamic, n, plan_choose=0, plans=None, sequential=False):
        self.dim, self.dynamic, self.n = dim, dynamic, n
        self.envs = [O.make_env(dim, dynamic, plan_choose=plan_choose, plans=plans) for _ in range(n)]
        self.n_plans = len(self.envs[0].plans)
        self.sequential = sequential
        self.D = self.envs[0].D
        self.ret = np.zeros(n)
        self.ep_cnt = np.zeros(n, np.int64)
        self.ep_len = np.zeros(n, np.int64)
        self.ep_ret = np.zeros(n)
        self.ep_iou = np.zeros(n)

    def reset(self, plan_idx=None):
        obs = np.zeros((self.n, self.D))
        for i, e in enumerate(self.envs):
            obs[i] = e.reset(0 if plan_idx is None else int(plan_idx[i]))[0]
        self.ret[:] = 0
        return obs

    def step(self, actions, step_sizes, next_plan=None, auto_reset=True, normalise=False, reset_obs=False):
        """reset_obs: the row of an env that finishes (and is reset) in this step is the observation its reset returns
        (DMP_F_RESET_OBS) instead of the terminal observation."""
        n = self.n
        obs = np.zeros((n, self.D))
        rew = np.zeros(n, np.float32)
        done = np.zeros(n, bool)
        for i, e in enumerate(self.envs):
            o, r, d = e.step(int(actions[i]), int(step_sizes[i]))
            obs[i] = (e.obs_normalised() if normalise else o)[0]
            rew[i], done[i] = r, d
            self.ret[i] += r
            if d and auto_reset:
                self.ep_cnt[i] += 1
                self.ep_len[i] += e.count_step
                self.ep_ret[i] += self.ret[i]
                self.ep_iou[i] += e.iou()
                self.ret[i] = 0
                if next_plan is not None:
                    p = int(next_plan[i])
                elif not self.dynamic:
                    p = 0
                elif self.sequential:
                    p = (e.plan_idx + 1) % self.n_plans
                else:
                    raise ValueError("random plan mode needs next_plan")
                o = e.reset(p)
                if reset_obs:
                    obs[i] = (e.obs_normalised() if normalise else o)[0]
        return obs, rew, done

    def export(self):
        grids = np.stack([e.grid.astype(np.int32) for e in self.envs])
        sc = np.zeros((self.n, 8), np.int32)
        for i, e in enumerate(self.envs):
            pos = [e.pos, 0] if self.dim == 1 else e.pos
            sc[i, :6] = [pos[0], pos[1], e.count_brick, e.count_step, e.plan_idx, int(np.ceil(e.total_brick))]
        return grids, sc

    def iou(self):
        return np.array([e.iou() for e in self.envs])


def philox_rollout(batch, K, seed, env_base, t0, n_actions, ref3d=False, auto_reset=True, normalise=False,
                   actions=None, step_sizes=None, reset_obs=False):
    """K steps with the kernels' Philox streams (oracle/philox.py).  Returns obs [K,N,D], reward, done."""
    n = batch.n
    ids = np.arange(env_base, env_base + n)
    obs = np.zeros((K, n, batch.D))
    rew = np.zeros((K, n), np.float32)
    done = np.zeros((K, n), bool)
    used_a = np.zeros((K, n), np.uint8)
    used_s = np.zeros((K, n), np.uint8)
    for k in range(K):
        s, a, p = philox.draws(seed, ids, t0 + k, n_actions, batch.n_plans, ref3d)
        if actions is not None:
            a = actions[k]
        if step_sizes is not None:
            s = step_sizes[k]
        used_a[k], used_s[k] = a, s
        nxt = p if (batch.dynamic and not batch.sequential) else None
        obs[k], rew[k], done[k] = batch.step(a, s, next_plan=nxt, auto_reset=auto_reset, normalise=normalise,
                                             reset_obs=reset_obs)
    return obs, rew, done, used_a, used_s
