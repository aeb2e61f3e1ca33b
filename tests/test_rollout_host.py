"""Host logic of the PPO rollout storage: the tensor GAE equals the reference's list recursion."""
import numpy as np
import torch

from bullet_envs_b200.rollout import RolloutBuffer, compute_gae


def reference_gae(next_value, rewards, masks, values, gamma=0.99, tau=0.95):
    # restatement of ppo/agent.py:14-22 on lists of [N,1] tensors, as ppo/train.py:170-173 calls it
    values = values + [next_value]
    gae = 0
    returns = []
    for step in reversed(range(len(rewards))):
        delta = rewards[step] + gamma * values[step + 1] * masks[step] - values[step]
        gae = delta + gamma * tau * masks[step] * gae
        returns.insert(0, gae + values[step])
    return returns


def test_gae_matches_the_reference_recursion():
    g = torch.Generator().manual_seed(0)
    T, N = 20, 37
    rewards = torch.randn((T, N), generator=g, dtype=torch.float64)
    values = torch.randn((T, N), generator=g, dtype=torch.float64)
    nxt = torch.randn((N,), generator=g, dtype=torch.float64)
    masks = (torch.rand((T, N), generator=g) > 0.15).double()
    ours = compute_gae(nxt, rewards, masks, values)
    ref = reference_gae(nxt[:, None], [r[:, None] for r in rewards], [m[:, None] for m in masks], [v[:, None] for v in values])
    assert torch.allclose(ours, torch.cat(ref, 1).T, atol=1e-12, rtol=0)


def test_rollout_buffer_views():
    buf = RolloutBuffer(5, 12, device="cpu")
    o, r, d = buf.out(2)
    assert o.is_contiguous() and r.is_contiguous() and d.is_contiguous() and d.dtype == torch.uint8
    o.fill_(3.0); r.fill_(-1.0); d.fill_(1)
    assert float(buf.obs[3].min()) == 3.0 and float(buf.rewards[2].max()) == -1.0 and float(buf.masks()[2].max()) == 0.0
    buf.obs[5].fill_(7.0); buf.roll()
    assert float(buf.obs[0].min()) == 7.0
    s, a, lp, v = buf.flat()
    assert s.shape == (60, 56) and a.shape == (60, 8) and lp.shape == (60, 8) and v.shape == (60, 1)
    assert np.shares_memory(s.numpy(), buf.obs.numpy())
