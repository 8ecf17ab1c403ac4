"""Step-kernel micro-bench: ms per launch of env_step_kernel at the headline shape (CUDA events, L2 flushed).
   SGB_LIBRARY=<variant.so> python profiles/kbench.py [B] [iters] [scenario] [n_agents]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sigmarl_b200 import EnvConfig, RoadTrafficEnv
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
K = int(sys.argv[2]) if len(sys.argv) > 2 else 60
SC = sys.argv[3] if len(sys.argv) > 3 else "cpm_entire"
NA = int(sys.argv[4]) if len(sys.argv) > 4 else 8
# KB_LAYOUT="is_ego_view=0,is_obs_steering=1": observation-layout switches (0 / 1) for the flag-driven writer
LAYOUT = {k: bool(int(v)) for k, v in (kv.split("=") for kv in os.environ.get("KB_LAYOUT", "").split(",") if kv)}
REW = os.environ.get("KB_REW", "distance")            # reward method
GENTLE = bool(int(os.environ.get("KB_GENTLE", "0")))  # 1: path-following actions (pure pursuit): long episodes
env = RoadTrafficEnv(EnvConfig(scenario_type=SC, n_agents=NA, rew_method=REW, **LAYOUT), num_envs=B, device="cuda:0", seed=0)
env.reset()
ur = torch.tensor([1.0, 31 * np.pi / 180], device="cuda")
g = torch.Generator(device="cuda").manual_seed(0)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
ts, tr = [], []
chk = 0.0
WARM = 5 if not GENTLE else 60      # path following: let the population spread before timing
for t in range(K + WARM):
    if GENTLE:
        # pure pursuit on the 2nd short-term reference point (ego frame, obs[3:5]) + a little steering noise: agents
        # follow their paths, episodes run to the time limit, the population spreads along the whole map
        a = torch.rand(B, NA, 2, device="cuda", generator=g)
        steer = torch.clamp(1.5 * torch.atan2(env.obs[..., 4], env.obs[..., 3]) + (a[..., 1] * 2 - 1) * 0.03, -float(ur[1]), float(ur[1]))
        env.action[..., 0] = 0.5 + 0.3 * a[..., 0]; env.action[..., 1] = steer
    else:
        env.action.copy_((torch.rand(B, NA, 2, device="cuda", generator=g) * 2 - 1) * ur)
    flush.zero_()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(); env.step(None); e[1].record(); env.reset_done(write_obs=bool(int(os.environ.get('KB_WRITE_OBS', '0')))); e[2].record()
    torch.cuda.synchronize()
    if t >= WARM:
        ts.append(e[0].elapsed_time(e[1])); tr.append(e[1].elapsed_time(e[2]))
    chk += float(env.reward.double().sum()) + float(env.obs.double().sum())
print(f"{os.environ.get('SGB_LIBRARY', 'default')[-28:]:28s} {SC} B={B} N={NA} D={env.D} rew={REW} {'gentle' if GENTLE else 'uniform'} {os.environ.get('KB_LAYOUT', '')} done-rate {float(env.done.float().mean()):.2f} step {np.mean(ts):.4f} ms (min {np.min(ts):.4f})  reset+refresh {np.mean(tr):.4f} ms  "
      f"-> {B * NA / (np.mean(ts) + np.mean(tr)) / 1e3:.1f} M agent-steps/s   checksum {chk:.6f}")
