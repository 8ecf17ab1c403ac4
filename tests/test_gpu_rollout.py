"""GPU (-m gpu): the "next" row — rollout collector + GAE kernel (through the C-ABI) against a plain restatement."""
import numpy as np
import pytest
import torch

from test_multi_rank_cpu import gae_numpy

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("T,B,N", [(128, 512, 8), (7, 33, 3), (1, 1, 1)])
def test_gae_kernel_matches_restatement(T, B, N):
    from sigmarl_b200.rollout import RolloutBuffer, compute_gae
    rng = np.random.default_rng(T)
    buf = RolloutBuffer(T, B, N, 4, "cuda:0")
    host = {k: rng.standard_normal((T, B, N)).astype(np.float32) for k in ("reward", "value", "next_value")}
    done = rng.random((T, B)) < 0.1
    for k, v in host.items():
        getattr(buf, k).copy_(torch.from_numpy(v))
    buf.done.copy_(torch.from_numpy(done.astype(np.uint8)))
    adv, tgt = compute_gae(buf, 0.99, 0.9)
    want_a, want_t = gae_numpy(host["reward"], host["value"], host["next_value"], done, 0.99, 0.9)
    assert np.max(np.abs(adv.cpu().numpy() - want_a)) <= 1e-5
    assert np.max(np.abs(tgt.cpu().numpy() - want_t)) <= 1e-5


def test_collect_fills_buffers_and_keeps_reset_semantics():
    from sigmarl_b200 import EnvConfig, RoadTrafficEnv
    from sigmarl_b200.rollout import RolloutBuffer, collect, compute_gae, all_gather_advantages
    env = RoadTrafficEnv(EnvConfig(scenario_type="cpm_entire", n_agents=8), num_envs=256, device="cuda:0", seed=1)
    env.reset()
    T = 16
    buf = RolloutBuffer(T, env.B, env.N, env.D, env.device)
    g = torch.Generator(device="cuda").manual_seed(0)
    ur = torch.tensor([1.0, 31 * np.pi / 180], device="cuda")
    policy = lambda obs: (torch.rand(env.B, env.N, 2, device="cuda", generator=g) * 2 - 1) * ur  # noqa: E731
    value = lambda obs: obs[..., 0] * 0.5 + obs[..., 7]  # noqa: E731  any deterministic function of obs
    first_obs = env.obs.clone()
    collect(env, policy, buf, value_fn=value)
    assert torch.equal(buf.obs[0], first_obs)
    assert torch.isfinite(buf.obs).all() and torch.isfinite(buf.reward).all()
    assert buf.done.sum() > 0
    # the observation stored at t+1 of a reset env is the post-reset one: its own speed entry equals |v| of the new pose
    adv, tgt = compute_gae(buf)
    a_all, _ = all_gather_advantages(buf)
    assert a_all.shape == (1, T, env.B, env.N) and torch.equal(a_all[0], adv)
    want_a, _ = gae_numpy(buf.reward.cpu().numpy(), buf.value.cpu().numpy(), buf.next_value.cpu().numpy(),
                          buf.done.cpu().numpy().astype(bool), 0.99, 0.9)
    assert np.max(np.abs(adv.cpu().numpy() - want_a)) <= 1e-5


def test_in_place_collect_equals_copying_collect():
    """Binding the env's outputs to the rollout buffer (no per-step copies) must fill exactly the same buffers."""
    from sigmarl_b200 import EnvConfig, RoadTrafficEnv
    from sigmarl_b200.rollout import RolloutBuffer, collect
    bufs = []
    for in_place in (True, False):
        env = RoadTrafficEnv(EnvConfig(scenario_type="cpm_mixed", n_agents=4), num_envs=512, device="cuda:0", seed=5)
        env.reset()
        buf = RolloutBuffer(12, env.B, env.N, env.D, env.device)
        g = torch.Generator(device="cuda").manual_seed(3)
        ur = torch.tensor([1.0, 31 * np.pi / 180], device="cuda")
        # 1.5 x the action range: the kernel clamps its own copy, the buffer must keep the RAW policy output
        policy = lambda obs: (torch.rand(env.B, env.N, 2, device="cuda", generator=g) * 2 - 1) * ur * 1.5  # noqa: E731
        value = lambda obs: obs[..., 0] - obs[..., 7]  # noqa: E731
        collect(env, policy, buf, value_fn=value, in_place=in_place)
        bufs.append((buf, env.obs.clone(), env.reward.clone(), env.done.clone(), env.pose.clone()))
    a, b = bufs
    for name in ("obs", "action", "reward", "done", "value", "next_value"):
        assert torch.equal(getattr(a[0], name), getattr(b[0], name)), name
    for x, y in zip(a[1:], b[1:]):
        assert torch.equal(x, y)
    assert a[0].done.sum() > 0
    assert float(a[0].action[..., 0].abs().max()) > 1.2        # raw (unclamped) actions were stored ...
    assert float(env.action[..., 0].abs().max()) <= 1.0        # ... while the env stepped on the clamped copy


@pytest.mark.parametrize("T,B,N", [(32, 257, 5), (32, 257, 8)])
def test_fused_gae_allgather_on_one_rank_equals_gae(T, B, N):
    """sgb_gae_allgather with world == 1 (the only peer is this rank's own buffer) leaves exactly what sgb_gae leaves;
    N = 5: one column per thread, N = 8: four columns per thread with 16-byte loads / stores."""
    from sigmarl_b200.rollout import RolloutBuffer, compute_gae, gae_allgather
    g = torch.Generator(device="cuda").manual_seed(4)
    buf = RolloutBuffer(T, B, N, 4, "cuda:0")
    for name in ("reward", "value", "next_value"):
        getattr(buf, name).copy_(torch.randn(T, B, N, device="cuda", generator=g))
    buf.done.copy_((torch.rand(T, B, device="cuda", generator=g) < 0.1).to(torch.uint8))
    a, t = compute_gae(buf, 0.99, 0.9)
    a, t = a.clone(), t.clone()
    buf.adv_all.zero_(); buf.target_all.zero_()
    a_all, t_all = gae_allgather(buf, 0.99, 0.9)
    assert torch.equal(a_all[0], a) and torch.equal(t_all[0], t)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs of one node (peer memory)")
@pytest.mark.parametrize("multicast", ["0", "1"])
def test_fused_gae_allgather_across_two_gpus_equals_gae_plus_nccl_all_gather(multicast):
    """Two ranks, symmetric-memory gather buffers: the fused kernel (peer stores / NVSwitch multicast stores) leaves in
    EVERY rank's buffers bit for bit what sgb_gae + NCCL all_gather_into_tensor leave (tests/tools/fused_gather_check.py)."""
    import os
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(repo, "tests", "tools", "fused_gather_check.py"), multicast]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=repo)
    assert out.returncode == 0 and out.stdout.count("FUSED-GATHER-OK") == 2, out.stdout[-2000:] + out.stderr[-4000:]
