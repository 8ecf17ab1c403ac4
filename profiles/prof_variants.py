"""Driver for the ncu captures of the other instantiations: a few steps (+ device resets) of
   python profiles/prof_variants.py <scenario> <n_agents> <B> [layout k=v,...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sigmarl_b200 import EnvConfig, RoadTrafficEnv
sc, N, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
layout = {k: bool(int(v)) for k, v in (kv.split("=") for kv in (sys.argv[4].split(",") if len(sys.argv) > 4 and sys.argv[4] else []))}
env = RoadTrafficEnv(EnvConfig(scenario_type=sc, n_agents=N, **layout), num_envs=B, device="cuda:0", seed=0)
env.reset()
ur = torch.tensor([1.0, 31 * np.pi / 180], device="cuda")
g = torch.Generator(device="cuda").manual_seed(0)
for t in range(6):
    env.step((torch.rand(B, N, 2, device="cuda", generator=g) * 2 - 1) * ur)
    env.reset_done(write_obs=True)
torch.cuda.synchronize()
print("done", float(env.done.float().mean()))
