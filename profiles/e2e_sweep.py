"""e2e micro-bench: ms per sgb_step_reset_host call at the headline shape for several chunk sizes of the host pipeline.
   python profiles/e2e_sweep.py   (SGB_HOST_CHUNK_WAVES is read per call)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sigmarl_b200 import EnvConfig, RoadTrafficEnv
B, N = 65536, 8
env = RoadTrafficEnv(EnvConfig(scenario_type="cpm_entire", n_agents=N), num_envs=B, device="cuda:0", seed=0)
env.reset()
ur = torch.tensor([1.0, 31 * np.pi / 180])
acts = [((torch.rand(B, N, 2) * 2 - 1) * ur).contiguous().pin_memory() for _ in range(2)]
for waves in (1, 2, 3, 4, 7, 14):
    os.environ["SGB_HOST_CHUNK_WAVES"] = str(waves)
    for i in range(3):
        env.step_host(acts[i % 2], reset_done=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(10):
        env.step_host(acts[i % 2], reset_done=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print(f"waves/chunk {waves:2d} ({-(-B // (waves * 4736))} chunks): {1e3 * dt:.3f} ms per step -> {B * N / dt / 1e6:.1f} M agent-steps/s e2e, "
          f"D2H {(B * N * 33 * 4 + B) / dt / 1e9:.1f} GB/s")
