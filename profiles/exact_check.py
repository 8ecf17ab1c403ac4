"""One-off exactness sweep: pruned kernel vs cfg.exhaustive=1 (every segment / pair through the full interX
predicate, no gates, no votes) at the headline shape, all output buffers compared bitwise every step.
    python profiles/exact_check.py [B] [T] [scenario] [N]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sigmarl_b200 import EnvConfig, RoadTrafficEnv
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 30
sc = sys.argv[3] if len(sys.argv) > 3 else "cpm_entire"
N = int(sys.argv[4]) if len(sys.argv) > 4 else 8
envs = [RoadTrafficEnv(EnvConfig(scenario_type=sc, n_agents=N, rew_method="ttc_sparse", exhaustive=ex,
                                 threshold_near_other_agents_c2c_low=0.1635), num_envs=B, device="cuda:0", seed=11,
                       info=True) for ex in (False, True)]
for e in envs:
    e.reset()
ur = torch.tensor([1.0, 31 * np.pi / 180], device="cuda")
g = torch.Generator(device="cuda").manual_seed(3)
names = ("pose", "aux", "carry", "obs", "reward", "done", "agent_flags", "collide_with", "step_count", "info")
n_lane = n_a2a = 0
for t in range(T):
    act = (torch.rand(B, N, 2, device="cuda", generator=g) * 2 - 1) * ur
    if t % 3 == 2:      # calmer driving every third step: long episodes reach other parts of the paths
        act[..., 1] *= 0.1
    for e in envs:
        e.step(act)
    torch.cuda.synchronize()
    for n in names:
        a, b = getattr(envs[0], n), getattr(envs[1], n)
        assert torch.equal(a, b), f"step {t}: {n} differs in {int((a != b).sum())} entries"
    n_lane += int(((envs[0].agent_flags & 2) != 0).sum()); n_a2a += int(((envs[0].agent_flags & 1) != 0).sum())
    for e in envs:
        e.reset_done(write_obs=True)
    assert torch.equal(envs[0].pose, envs[1].pose) and torch.equal(envs[0].obs, envs[1].obs)
print(f"exact_check ok: {sc} B={B} N={N} T={T}: {B * N * T} agent-steps bit-identical "
      f"(lane hits {n_lane}, agent hits {n_a2a})")
