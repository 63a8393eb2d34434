"""Device-resident acting loop: observation -> policy -> action -> step, with nothing leaving the GPU.

The reference's learners run ``action = agent.choose_action(state); state_next, r, done = env.step(action)`` one env
and one Python iteration at a time (script/DQN/2d/DQN_2d_static.py:185-207, ``choose_action`` :108-123).  Here the
same loop runs for N envs at once: the observation tensor written by ``dmp_step`` is the policy's input, the policy's
uint8 action tensor is ``dmp_step``'s input, and T iterations are captured ONCE in a CUDA graph and replayed -- no
host round trip, no per-step launch overhead beyond the graph's kernel nodes.  The per-step draws
(``np.random.randint(1, 4)``) come from the kernels' Philox stream; the step counter lives on the device
(``DmpState.t_dev``) so a replayed graph keeps advancing it.

    loop = DeviceRollout(env, EpsilonGreedy(qnet, n_actions=env.action_dim, epsilon=0.2), horizon=64)
    traj = loop.collect()          # dict of [T, N, ...] tensors (views of static buffers, overwritten by the next call)

Policies are callables ``policy(obs [N, D]) -> actions uint8 [N]`` made of graph-capturable torch ops.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from .vecenv import BatchedDMPEnv


class RandomPolicy:
    """Uniform random actions (multiprocess.py:83 draws ``np.random.randint(3, size=N)`` on the host)."""

    def __init__(self, n_actions: int, generator: Optional[torch.Generator] = None):
        self.n_actions, self.generator = int(n_actions), generator

    def __call__(self, obs: torch.Tensor) -> torch.Tensor:
        return torch.randint(0, self.n_actions, (obs.shape[0],), device=obs.device, dtype=torch.uint8,
                             generator=self.generator)


class QSAAdapter(torch.nn.Module):
    """Wraps a reference-style critic ``Q(s, a) -> [B, 1]`` (``Q_NET.forward(s, a)`` concatenates the action to the
    state, script/DQN/2d/DQN_2d_static.py:78-98) into ``q(obs) -> [N, A]`` by scoring all A actions in one batch
    instead of the reference's ``for item in range(Action_dim)`` loop (:116-120)."""

    def __init__(self, q_sa: torch.nn.Module, n_actions: int):
        super().__init__()
        self.q_sa, self.n_actions = q_sa, int(n_actions)

    def forward(self, obs: torch.Tensor) -> torch.Tensor:
        n, A = obs.shape[0], self.n_actions
        s = obs.unsqueeze(1).expand(n, A, obs.shape[1]).reshape(n * A, -1)
        a = torch.arange(A, device=obs.device, dtype=obs.dtype).repeat(n).unsqueeze(1)
        return self.q_sa(s, a).reshape(n, A)


class EpsilonGreedy:
    """``choose_action`` of the reference's DQN agents, batched: with probability epsilon a uniform random action,
    else argmax_a Q(obs)[a].  ``q`` maps obs [N, D] (float) to Q-values [N, A]."""

    def __init__(self, q: Callable[[torch.Tensor], torch.Tensor], n_actions: int, epsilon: float = 0.0,
                 generator: Optional[torch.Generator] = None):
        self.q, self.n_actions, self.generator = q, int(n_actions), generator
        self.epsilon = float(epsilon)

    @torch.no_grad()
    def __call__(self, obs: torch.Tensor) -> torch.Tensor:
        greedy = self.q(obs.float() if obs.dtype != torch.float32 else obs).argmax(dim=1)
        if self.epsilon <= 0.0:
            return greedy.to(torch.uint8)
        n = obs.shape[0]
        explore = torch.rand(n, device=obs.device, generator=self.generator) <= self.epsilon
        rnd = torch.randint(0, self.n_actions, (n,), device=obs.device, generator=self.generator)
        return torch.where(explore, rnd, greedy).to(torch.uint8)


def reset_observation(dim: int, dtype: torch.dtype, device) -> torch.Tensor:
    """The observation every reset() returns (raw or normalised counters: both are 0): the window at the start position
    over an empty grid -- 1D [-1, -1, 0, 0, 0 | 0, 0] (Env/1D/DMP_Env_1D_static.py:81-83), 2D/3D rows and columns 0..2 of
    the 7x7 window on the -1 frame (Env/2D/DMP_Env_2D_static.py:60-76)."""
    if dim == 1:
        o = torch.tensor([-1, -1, 0, 0, 0, 0, 0])
    else:
        w = torch.zeros((7, 7), dtype=torch.int64)
        w[:3, :] = -1
        w[:, :3] = -1
        o = torch.cat([w.reshape(-1), torch.zeros(2, dtype=torch.int64)])
    return o.to(device=device, dtype=dtype)


class DeviceRollout:
    """T steps of ``obs -> policy -> dmp_step`` for every env, captured in one CUDA graph.

    env        a BatchedDMPEnv that has been reset.  auto_reset=True keeps finished envs going; the policy input after an
               episode ends is the reset observation, like ``state = env.reset()`` in the reference's loops.  Build the
               env with reset_obs=True for the fast form: the kernel writes that observation itself, ``next_obs`` is
               ``obs[1:]`` (gym's vector-env convention: the terminal observation is not materialised) and the loop
               is nothing but policy + step.  With reset_obs=False ``next_obs[t]`` keeps the terminal observation and
               ``obs[t + 1]`` is patched with the reset observation by one extra pass over the batch per step
    policy     callable obs [N, D] -> uint8 actions [N] (see above); None = the kernels' own Philox actions
    horizon    steps per collect() (rounded up to an even number: the device step counter alternates two slots)
    graph      False runs the same loop eagerly (debugging / policies that cannot be captured)
    """

    def __init__(self, env: BatchedDMPEnv, policy: Optional[Callable] = None, horizon: int = 32, graph: bool = True):
        if env._needs_initial_reset:
            raise RuntimeError("reset() the env before building a DeviceRollout")
        self.env, self.policy = env, policy
        self.T = int(horizon) + (int(horizon) & 1)
        if env.records:
            raise ValueError("DeviceRollout feeds observations to a torch policy: build the env with a numeric obs_dtype")
        n, D, dev = env.num_envs, env.obs_dim, env.device
        self.obs = torch.zeros((self.T + 1, n, D), dtype=env.obs_dtype, device=dev)      # obs[t] is the input of step t
        # what step t returned.  With auto_reset the kernel returns the TERMINAL observation of a finished episode and
        # resets the env in the same launch; the next policy input of that env is then the reset observation, like
        # ``state = env.reset()`` in the reference's loops (script/DQN/2d/DQN_2d_static.py:186-206)
        self._patch = env.auto_reset and not env.reset_obs
        self.next_obs = torch.zeros((self.T, n, D), dtype=env.obs_dtype, device=dev) if self._patch else self.obs[1:]
        self.reset_obs = reset_observation(env.dim, env.obs_dtype, dev)                  # constant [D] row
        self._cur = torch.zeros((n, D), dtype=env.obs_dtype, device=dev)                 # observation carried between collects
        self.actions = torch.zeros((self.T, n), dtype=torch.uint8, device=dev)
        self.reward = torch.zeros((self.T, n), dtype=torch.float32, device=dev)
        self.done = torch.zeros((self.T, n), dtype=torch.uint8, device=dev)
        self._graph = None
        self._use_graph = bool(graph)
        self._stream = torch.cuda.Stream(device=dev)
        self._primed = False

    def set_initial_obs(self, obs: torch.Tensor) -> None:
        """Observation the first step's policy call sees (normally what env.reset() returned)."""
        self._cur.copy_(obs)
        self._primed = True

    def _body(self) -> None:
        env = self.env
        self.obs[0].copy_(self._cur)
        for t in range(self.T):
            a = None
            if self.policy is not None:
                self.actions[t].copy_(self.policy(self.obs[t]))
                a = self.actions[t:t + 1]
            env.rollout(1, actions=a, out=(self.next_obs[t:t + 1], self.reward[t:t + 1], self.done[t:t + 1]),
                        use_device_t=True, t_slot=t & 1)
            if self._patch:                                  # envs reset in this launch continue from the reset observation
                torch.where(self.done[t].view(torch.bool)[:, None], self.reset_obs, self.next_obs[t], out=self.obs[t + 1])
        self._cur.copy_(self.obs[self.T])                    # next collect() continues from the last observation

    def collect(self) -> dict:
        env = self.env
        if not self._primed:
            self.set_initial_obs(env._obs)                   # reset() / step() leave the current observation there
        cur = torch.cuda.current_stream(env.device)
        self._stream.wait_stream(cur)
        with torch.cuda.stream(self._stream):
            if self._graph is None:
                if self._use_graph:
                    # warm-up pass outside capture (lazy initialisation, cuBLAS workspaces); it must not count as
                    # experience, so the env is put back exactly where it was before capturing
                    first, snapshot = self._cur.clone(), env.get_state()
                    env._t_dev.fill_(env.t)
                    self._body()
                    env.set_state(snapshot)
                    self._cur.copy_(first)
                    self._stream.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=self._stream):
                        self._body()
                    self._graph = g
                else:
                    self._graph = False
            env._t_dev.fill_(env.t)                          # hand the host-side step counter to the device
            if self._graph:
                self._graph.replay()
            else:
                self._body()
        cur.wait_stream(self._stream)
        env._st.t = env._st.t + self.T
        return dict(obs=self.obs[:self.T], next_obs=self.next_obs, actions=self.actions, reward=self.reward,
                    done=self.done.view(torch.bool))
