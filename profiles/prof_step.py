"""Tiny driver for ncu: a few fused steps (+ device resets) of the headline workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sigmarl_b200 import EnvConfig, RoadTrafficEnv
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env = RoadTrafficEnv(EnvConfig(scenario_type="cpm_entire", n_agents=8), num_envs=B, device="cuda:0", seed=0)
env.reset()
ur = torch.tensor([1.0, 31 * np.pi / 180], device="cuda")
g = torch.Generator(device="cuda").manual_seed(0)
for t in range(6):
    env.step((torch.rand(B, 8, 2, device="cuda", generator=g) * 2 - 1) * ur)
    env.reset_done(write_obs=True)
torch.cuda.synchronize()
print("done", float(env.done.float().mean()))
