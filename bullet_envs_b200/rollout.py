"""Device-resident rollout storage for the PPO caller (second half).

The reference's rollout (``ppo/train.py:112-140``) appends per-step tensors to Python lists, moves rewards and
masks host->device every step and then runs ``compute_gae`` (``ppo/agent.py:14-22``) over those lists.  Here the
buffers are preallocated ``[T, N, ...]`` tensors on the env's device -- the step kernel writes observations,
rewards and done flags straight into their slices -- and the GAE recursion is a T-step scan over ``[N]`` vectors
on the same device.  torch is used for storage and elementwise arithmetic only.
"""
from __future__ import annotations


class RolloutBuffer:
    """Preallocated PPO rollout buffers; slice ``t`` is what ``SnakeVecEnv.step(..., out=buf.out(t))`` fills."""

    def __init__(self, num_steps, num_envs, obs_dim=56, act_dim=8, device=None):
        import torch
        kw = dict(device=device, dtype=torch.float32)
        self.T, self.N = int(num_steps), int(num_envs)
        self.obs = torch.zeros((self.T + 1, self.N, obs_dim), **kw)   # obs[t] = state the policy saw at step t
        self.actions = torch.zeros((self.T, self.N, act_dim), **kw)
        self.log_probs = torch.zeros((self.T, self.N, act_dim), **kw)
        self.values = torch.zeros((self.T, self.N), **kw)
        self.rewards = torch.zeros((self.T, self.N), **kw)
        self.dones = torch.zeros((self.T, self.N), device=device, dtype=torch.uint8)

    def out(self, t):
        """(obs, reward, done) slices for step ``t`` -- contiguous views the kernel writes in place."""
        return self.obs[t + 1], self.rewards[t], self.dones[t]

    def masks(self):
        """``1 - done`` (``ppo/train.py:134``) as float, [T, N]."""
        return 1.0 - self.dones.float()

    def roll(self):
        """start the next rollout from the last observation (``state = next_state``, ``ppo/train.py:139``)"""
        self.obs[0].copy_(self.obs[self.T])

    def gae(self, next_value, gamma=0.99, tau=0.95):
        """``compute_gae`` (``ppo/agent.py:14-22``) over the whole buffer as ONE kernel launch (C-ABI ``snk_gae``) on the current
        torch stream: returns ``(returns, advantages)``, both [T, N] (``advantage = returns - values``, ``ppo/train.py:178``).
        ``next_value`` [N] = V of ``obs[T]`` (``ppo/train.py:170-171``).  CUDA buffers only -- there is no CPU fallback."""
        import ctypes
        import torch
        from . import _abi
        if not self.rewards.is_cuda:
            raise RuntimeError("RolloutBuffer.gae runs the CUDA kernel snk_gae: the buffer must live on a CUDA device")
        lib = _abi.load_library()
        nv = next_value.detach().reshape(self.N).to(dtype=torch.float32).contiguous()
        values = self.values.detach()
        returns = torch.empty_like(self.rewards)
        adv = torch.empty_like(self.rewards)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        dev = self.rewards.device
        _abi.check(lib.snk_gae(dev.index if dev.index is not None else torch.cuda.current_device(), p(self.rewards), p(self.dones), p(values), p(nv),
                               float(gamma), float(tau), p(returns), p(adv), self.T, self.N,
                               ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), lib)
        return returns, adv

    def flat(self):
        """[T*N, ...] views in the layout ``ppo_update`` consumes (``torch.cat(states)`` etc., ``ppo/train.py:174-180``)."""
        T, N = self.T, self.N
        return (self.obs[:T].reshape(T * N, -1), self.actions.reshape(T * N, -1), self.log_probs.reshape(T * N, -1),
                self.values.reshape(T * N, 1))


def compute_gae(next_value, rewards, masks, values, gamma=0.99, tau=0.95):
    """Plain-tensor statement of the recursion (host logic; T small eager steps) -- the product path for a device-resident rollout is
    :meth:`RolloutBuffer.gae`, one kernel.  Generalised advantage estimation, the recursion of ``ppo/agent.py:14-22`` on [T, N] tensors:
    delta_t = r_t + gamma V_{t+1} m_t - V_t ;  gae_t = delta_t + gamma tau m_t gae_{t+1} ;  return_t = gae_t + V_t.
    ``next_value`` [N] is V of the state after the last step.  Returns the [T, N] returns tensor."""
    import torch
    T = rewards.shape[0]
    returns = torch.empty_like(rewards)
    gae = torch.zeros_like(rewards[0])
    nxt = next_value.reshape(rewards[0].shape)
    for t in range(T - 1, -1, -1):
        delta = rewards[t] + gamma * nxt * masks[t] - values[t]
        gae = delta + gamma * tau * masks[t] * gae
        returns[t] = gae + values[t]
        nxt = values[t]
    return returns
