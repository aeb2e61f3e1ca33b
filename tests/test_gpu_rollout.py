"""snk_rollout_linear (n env-steps per launch with a linear policy per environment):
bit-exact against the same rollout driven step by step through snk_step, and against the CPU oracle's twin."""
import numpy as np
import pytest

from bullet_envs_b200 import default_params
from oracle.oracle_py import Oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (there is no CPU fallback)")
    return torch


def test_fused_rollout_equals_the_stepwise_loop(torch):
    """Weights with a single non-zero per action row make the policy arithmetic exact (one product), so the fused
    rollout and `for t: env.step(policy(obs))` must agree bit for bit: returns, final state, observation trace."""
    from bullet_envs_b200 import SnakeVecEnv
    n, T = 1024, 6
    g = torch.Generator().manual_seed(2)
    noise = torch.rand((T, n, 56), generator=g).cuda()
    cols = torch.tensor([33, 1, 17, 49, 9, 55, 3, 20])
    gain = (torch.rand((n, 8), generator=g) * 4 - 2)
    W = torch.zeros((n, 8, 56)); W[:, torch.arange(8), cols] = gain
    W, gain, cols = W.cuda(), gain.cuda(), cols.cuda()
    fused = SnakeVecEnv(num_envs=n, device=0); fused.reset(as_torch=True)
    ret, trace = fused.rollout_linear(W, T, noise=noise, trace=True)
    step = SnakeVecEnv(num_envs=n, device=0)
    obs = step.reset(as_torch=True)
    acc = torch.zeros(n, device="cuda")
    for t in range(T):
        x = obs + noise[t]
        assert torch.equal(x, trace[t])
        obs, rew, done, _ = step.step(gain * x[:, cols])
        acc += rew
    assert torch.equal(acc, ret)
    assert torch.equal(fused.get_state(), step.get_state())
    assert fused.counters()["ticks"] > n * T                  # the rollout really moved
    fused.close(); step.close()


def test_fused_rollout_vs_oracle(torch):
    from bullet_envs_b200 import SnakeVecEnv
    n, T = 192, 3
    rng = np.random.default_rng(8)
    W = (rng.normal(size=(n, 8, 56)) * 0.05).astype(np.float32)
    noise = rng.uniform(0, 1, (T, n, 56)).astype(np.float32)
    mean = rng.uniform(0.3, 0.7, 56).astype(np.float32); inv_std = rng.uniform(1.0, 3.0, 56).astype(np.float32)
    env = SnakeVecEnv(num_envs=n, device=0); env.reset(as_torch=True)
    ret, tr = env.rollout_linear(W, T, mean=mean, inv_std=inv_std, noise=noise, trace=True)
    o = Oracle(n); o.reset()
    oret, otr = o.rollout_linear(W, T, mean=mean, inv_std=inv_std, noise=noise, trace=True)
    assert np.abs(tr[0].cpu().numpy() - otr[0]).max() < 1e-6   # first policy input: reset observation + noise
    c, oc = env.counters(), o.counters()
    assert abs(c["ticks"] - oc["ticks"]) <= 0.02 * oc["ticks"] and abs(c["dones"] - oc["dones"]) <= 0.05 * oc["dones"] + 3
    d = np.abs(ret.cpu().numpy() - oret)
    assert np.median(d) < 5e-3 and np.isfinite(ret.cpu().numpy()).all()
    env.close()


def test_rollout_argument_errors(torch):
    from bullet_envs_b200 import SnakeVecEnv
    env = SnakeVecEnv(num_envs=4, device=0, params=default_params(motor_solver=0))
    with pytest.raises(RuntimeError, match="exact motor solver"):
        env.rollout_linear(np.zeros((4, 8, 56), np.float32), 2)
    env.close()


def test_gae_kernel_matches_the_reference_recursion(torch):
    """snk_gae (one launch) against compute_gae of ppo/agent.py:14-22 restated on lists of [N,1] tensors, as ppo/train.py:170-173
    calls it -- in fp32 on the same device (same operation order: bit-exact up to fma contraction, asserted to 1e-5) and in fp64."""
    from bullet_envs_b200.rollout import RolloutBuffer
    g = torch.Generator().manual_seed(0)
    for T, N in ((20, 65536), (7, 1001), (1, 5)):
        buf = RolloutBuffer(T, N, device="cuda")
        buf.rewards.copy_(torch.randn((T, N), generator=g)); buf.values.copy_(torch.randn((T, N), generator=g))
        buf.dones.copy_((torch.rand((T, N), generator=g) < 0.15).to(torch.uint8))
        nxt = torch.randn((N,), generator=g).cuda()
        ret, adv = buf.gae(nxt, gamma=0.99, tau=0.95)

        def reference(dtype):   # ppo/agent.py:14-22
            values = [v.to(dtype)[:, None] for v in buf.values] + [nxt.to(dtype)[:, None]]
            rewards = [r.to(dtype)[:, None] for r in buf.rewards]; masks = [1 - d.to(dtype)[:, None] for d in buf.dones]
            gae = 0; returns = []
            for step in reversed(range(len(rewards))):
                delta = rewards[step] + 0.99 * values[step + 1] * masks[step] - values[step]
                gae = delta + 0.99 * 0.95 * masks[step] * gae
                returns.insert(0, gae + values[step])
            return torch.cat(returns, 1).T
        assert (ret - reference(torch.float32)).abs().max() < 1e-5
        assert (ret.double() - reference(torch.float64)).abs().max() < 1e-4
        assert torch.equal(adv, ret - buf.values) or (adv - (ret - buf.values)).abs().max() < 1e-5
