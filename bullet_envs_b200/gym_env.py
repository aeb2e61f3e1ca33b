"""Single-environment facade with the shape of the reference's ``SnakeGymEnv`` / ``Snake`` pair.

``SnakeGymEnv(robot, args)`` (reference ``SnakeGymEnv.py:4-103``) and ``Snake(pybullet_client,
urdf_root, args)`` (``snake.py:12-32``) keep their constructor signatures and the attributes callers
touch (``alpha/beta/gamma/mode/_gaitSelection``, ``robot.numMotors``, ``robot.START_POSITION``,
``robot.calculateEnergy``, ``observation_space``/``action_space``), but the physics is one
environment of the CUDA batch.  Semantics are those of the vector-wrapper path: after a
``done`` the environment is already reset and the returned observation is the post-reset one.
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


class SnakeGymEnv:
    def __init__(self, robot=None, args=None, device=None):
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

    def reset(self, hardReset=False):  # SnakeGymEnv.py:28-31
        self._observation = self._vec.reset(hard=bool(hardReset))[0]
        return self._observation

    def step(self, action):
        a = np.asarray(action, np.float32)
        if isinstance(action, (list, np.ndarray)):  # checkBound clips in place (SnakeGymEnv.py:82-88)
            for i in range(len(action)):
                if action[i] < -1 or action[i] > 1:
                    action[i] = np.clip(action[i], -1, 1)
        self._vec.mode = self.mode  # the reference reads self.mode at every step (SnakeGymEnv.py:43)
        obs, r, d, infos = self._vec.step(a[None, :])
        self._observation = obs[0]
        return obs[0], float(r[0]), bool(d[0]), infos[0]  # {} in train mode, the per-tick stream in test mode (SnakeGymEnv.py:43-46)

    def render(self):  # SnakeGymEnv.py:52-58: train mode; test mode would need a display (out of scope)
        return np.array([])

    def close(self):
        self._vec.close()
