"""Single-environment facade with the shape of the reference's ``SnakeGymEnv`` / ``Snake`` pair.

``SnakeGymEnv(robot, args)`` (reference ``SnakeGymEnv.py:4-103``) and ``Snake(pybullet_client,
urdf_root, args)`` (``snake.py:12-32``) keep their constructor signatures and the attributes callers
touch (``alpha/beta/gamma/mode/_gaitSelection``, ``robot.numMotors``, ``robot.START_POSITION``,
``robot.calculateEnergy``, ``observation_space``/``action_space``), but the physics is one
environment of the CUDA batch.

Semantics.  The batch implements what the *vector wrapper* makes of the env (``ppo/multiprocessing_env.py:11-16``):
on ``done`` the returned observation is the post-reset one and the next reward measures progress from the reset pose.
The reference class used *without* the wrapper (``ppo/utils.py:71-83``, ``ars/train.py:47-59``, ``a3c/agent.py:99-125``)
behaves differently: ``step`` returns the TERMINAL observation of the dead episode (``SnakeGymEnv.py:35-42``
computes it before the in-step reset) and, because ``self._observation`` keeps that observation, the first reward after a
``done`` uses the dead episode's x as x_prev.  ``SnakeGymEnv(..., raw_semantics=True)`` (the default, = the reference class)
reproduces exactly that on top of the batch; ``raw_semantics=False`` gives the wrapper's behaviour.
"""
from __future__ import annotations

import numpy as np

from .vec_env import SnakeVecEnv


class Snake:
    """Parameter carrier mirroring ``snake.Snake``; the pybullet client argument is accepted and ignored."""

    def __init__(self, pybullet_client=None, urdf_root=None, args=None):
        self.numMotors = 16
        self._pybulletClient = pybullet_client
        self._urdf = urdf_root
        self._timeStep = 1 / 100.0
        self.START_POSITION = [0, 0, 0]
        self.FRICTION_VALUES = [1, 0.1, 0.01]
        self.MAX_TORQUE = np.inf
        self.args = args
        self.mode = getattr(args, "mode", "train") if args is not None else "train"
        self._gaitSelection = getattr(args, "gaitSelection", 1) if args is not None else 1
        self.SCALING_FACTOR = np.pi / (getattr(args, "scaling_factor", 6) * 1.0) if args is not None else np.pi / 6

    def calculateEnergy(self, observation):  # snake.py:336-341
        n = self.numMotors
        return float(np.sum(np.asarray(observation[n:2 * n]) * np.asarray(observation[2 * n:3 * n]) * self._timeStep))


class RawSingleEnv:
    """The unwrapped ``SnakeGymEnv.step`` semantics on top of a one-environment batch with wrapper semantics.

    ``vec`` needs ``step_traced(action[1,8]) -> (obs[1,56], rew[1], done[1], last_internal_obs or None)`` and
    ``observe() -> obs[1,56]``.  Kept free of CUDA so that the logic is testable against the reference's own Python."""

    def __init__(self, vec, alpha):
        self.vec, self.alpha = vec, float(alpha)
        self.x_prev_shift = 0.0  # x_prev of the reference minus the batch's x_prev (= x at step start) for the coming step

    def reset(self):
        self.x_prev_shift = 0.0  # SnakeGymEnv.reset overwrites self._observation (SnakeGymEnv.py:28-31)

    def step(self, a):
        before = self.vec.observe()[0]
        obs, r, d, last = self.vec.step_traced(a)
        obs, r, d = np.array(obs[0], dtype=np.float64), float(r[0]), bool(d[0])
        r -= self.alpha * self.x_prev_shift          # reward = alpha * (x - x_prev_reference) + ...  (SnakeGymEnv.py:91)
        self.x_prev_shift = 0.0
        if d:
            terminal = np.array(last if last is not None else before, dtype=np.float64)   # observation before the in-step reset
            self.x_prev_shift = terminal[48] - obs[48]    # self._observation = terminal observation (SnakeGymEnv.py:42)
            obs = terminal
        return obs, r, d


class _BatchOfOne:
    def __init__(self, vec):
        self.vec = vec

    def observe(self):
        return self.vec.observe().cpu().numpy()

    def step_traced(self, a):
        mode = self.vec.mode
        self.vec.mode = "test"
        try:
            obs, r, d, infos = self.vec.step(a)
        finally:
            self.vec.mode = mode
        io = infos[0]["internal_observations"]
        self.last_info = infos[0]
        return obs, r, d, (io[-1] if len(io) else None)


class SnakeGymEnv:
    def __init__(self, robot=None, args=None, device=None, raw_semantics=True):
        self.robot = robot if robot is not None else Snake(None, None, args)
        if args is None:
            args = getattr(self.robot, "args", None)
        self.alpha = getattr(args, "alpha", 1) if args is not None else 1
        self.beta = getattr(args, "beta", 0.01) if args is not None else 0.01
        self.gamma = getattr(args, "gamma", 0.1) if args is not None else 0.1
        self.mode = getattr(args, "mode", "train") if args is not None else "train"
        self._gaitSelection = getattr(args, "gaitSelection", 1) if args is not None else 1
        self._action_bound = 1
        urdf = getattr(self.robot, "_urdf", None)
        import os
        self._vec = SnakeVecEnv(num_envs=1, args=args, device=device, urdf_path=urdf if (urdf and os.path.exists(urdf)) else None,
                                mode=self.mode)
        self.observation_space = self._vec.observation_space
        self.action_space = self._vec.action_space
        self._observation = None
        self._one = _BatchOfOne(self._vec)
        self._raw = RawSingleEnv(self._one, self.alpha) if raw_semantics else None

    def reset(self, hardReset=False):  # SnakeGymEnv.py:28-31
        self._observation = self._vec.reset(hard=bool(hardReset))[0]
        if self._raw is not None:
            self._raw.reset()
        return self._observation

    def step(self, action):
        a = np.asarray(action, np.float32)
        if isinstance(action, (list, np.ndarray)):  # checkBound clips in place (SnakeGymEnv.py:82-88)
            for i in range(len(action)):
                if action[i] < -1 or action[i] > 1:
                    action[i] = np.clip(action[i], -1, 1)
        if self._raw is not None:  # the reference class as its single-env callers see it (terminal observation, Q8 x_prev)
            self._raw.alpha = float(self.alpha)
            ob, r, d = self._raw.step(a[None, :])
            self._observation = ob
            return ob, r, d, (self._one.last_info if self.mode == "test" else {})  # the reference reads self.mode per step (SnakeGymEnv.py:43)
        self._vec.mode = self.mode
        obs, r, d, infos = self._vec.step(a[None, :])
        self._observation = obs[0]
        return obs[0], float(r[0]), bool(d[0]), infos[0]  # {} in train mode, the per-tick stream in test mode (SnakeGymEnv.py:43-46)

    def render(self):  # SnakeGymEnv.py:52-58: train mode; test mode would need a display (out of scope)
        return np.array([])

    def close(self):
        self._vec.close()
