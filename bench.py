#!/usr/bin/env python
"""bench.py — agent-steps/s of the road-traffic environment step on the CPM map (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--envs B] [--agents N]

One "step" = one pass of the hot path over one batch: the fused step kernel (dynamics -> collisions ->
reward -> observation -> done) followed by the masked device reset/respawn of finished envs, exactly what a
rollout executes per environment step.  N>1: launched by torchrun, one rank per GPU; envs are sharded by
index with no data-path collective -> weak scaling, `value` = all ranks' agent-steps / max time.

--impl reference: the reference's CPU implementation of the same path.  The reference is pure Python (it
cannot be compiled into oracle/_ref), so this arm times the C oracle port (oracle/sigmarl_oracle.c, pinned
bit-exactly to the reference's golden vectors) on all host threads, on a bounded sample of the workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

ALGO_BYTES = lambda D: 65 + 4 * D  # noqa: E731  SURVEY.md §8(d): algorithmic HBM bytes per agent-step
METRIC = "agent-steps/sec (num_envs x n_agents / step_time), CPM map"


def ncu_traffic():
    """DRAM bytes per launch of the fused step kernel from the committed `ncu --set full` capture (profiles/)."""
    p = os.path.join(REPO, "profiles", "ncu_step_kernel_r1.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"])
        except Exception:
            return None
    return None


def ncu_issue():
    """Issue-slot utilisation / active lanes of the fused step kernel from the committed ncu capture: the bound that
    actually limits this kernel (DESIGN.md roofline section); static, for context next to the HBM fraction."""
    p = os.path.join(REPO, "profiles", "ncu_full_r1.json")
    try:
        ks = json.load(open(p))
        k = max(ks, key=lambda d: float(d["gpu__time_duration.sum"].split()[0]))
        busy = float(k["smsp__issue_active.avg.pct_of_peak_sustained_active"].split()[0])
        lanes = float(k["smsp__thread_inst_executed_per_inst_executed.ratio"].split()[0])
        winst = float(k["smsp__inst_executed.sum"].split()[0])
        return {"issue_slots_busy_pct": busy, "active_lanes_per_warp_inst": lanes, "warp_inst_per_launch": winst,
                # share of the SMs' lane-issue capacity (4 schedulers x 32 lanes per clock) doing useful work: the
                # "fp32 / ALU roofline" reading SURVEY.md §8d asks for next to the HBM fraction
                "lane_issue_frac": busy / 100.0 * lanes / 32.0,
                "thread_inst_per_agent_step": winst * lanes / (65536 * 8),   # the capture's shape: 65536 envs x 8 agents
                "source": "profiles/ncu_full_r1.json"}
    except Exception:
        return None


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe): one
    `nvidia-smi -lms 20` process streams samples while the timed loop runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.rows = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.proc.stdout.readline()      # first sample = the sampler is up before the timed region starts
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return
        time.sleep(0.03)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=3)
        except Exception:
            self.proc.kill()
            out = ""
        self.rows = [[c.strip() for c in line.split(",")] for line in out.strip().splitlines() if line.strip()]

    def summary(self):
        num = lambda v: v.replace(".", "", 1).isdigit()  # noqa: E731
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and num(r[0])]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and num(r[1])]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_baseline_sample(n_envs, n_agents, steps, threads, seed=0):
    """Time the oracle port on `threads` host threads: `steps` steps of `n_envs` envs (+ resets of done envs)."""
    from oracle import oracle as O
    w = O.OracleWorld("cpm_entire", n_envs, n_agents, mode="params", rew_method="distance")
    for b in range(n_envs):
        assert w.reset_env(b) == 0
    rng = np.random.default_rng(seed)
    ur = np.asarray([1.0, 31 * np.pi / 180], np.float32)
    acts = [((rng.random((n_envs, n_agents, 2), np.float32) * 2 - 1) * ur).astype(np.float32) for _ in range(steps + 1)]
    w.step(acts[0], n_threads=threads)  # warm-up
    t0 = time.perf_counter()
    for k in range(steps):
        _, _, done, _ = w.step(acts[k + 1], n_threads=threads)
        for b in np.where(done)[0]:
            w.reset_env(int(b))
    dt = time.perf_counter() - t0
    return n_envs * n_agents * steps / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_envs = args.ref_envs
    # warm-up steps, then K timed steps of the bounded sample
    v, dt = cpu_baseline_sample(n_envs, args.agents, args.steps, threads)
    D = 10 + 11 * min(2, args.agents - 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "agent-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"cpm_entire num_envs={args.envs} n_agents={args.agents} per GPU (reference arm: bounded sample of {n_envs} envs)",
                   "obs_dim": D, "rew_method": "distance", "dt": 0.1},
        "cpu_baseline": {"value": v, "unit": "agent-steps/s", "cores": threads, "kind": "port",
                         "sample": f"{n_envs} envs x {args.agents} agents x {args.steps} steps, oracle/sigmarl_oracle.c on {threads} pthreads "
                                   f"(reference is pure Python: 2-3e3 agent-steps/s measured under the import shim, BASELINE.md)"},
        "e2e": {"value": v, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from sigmarl_b200 import EnvConfig, RoadTrafficEnv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, N, K, W = args.envs, args.agents, args.steps, max(3, args.warmup)
    cfg = EnvConfig(scenario_type="cpm_entire", n_agents=N, mode="params", rew_method="distance")
    env = RoadTrafficEnv(cfg, num_envs=B, device=dev, seed=args.seed, env_offset=rank * B)
    env.reset()
    D = env.D
    ur = torch.tensor([1.0, 31 * np.pi / 180], device=dev)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)

    def new_action():
        # SURVEY.md §8d action distribution (i): U(-1,1)^2 * [v_max, delta_max]
        return (torch.rand(B, N, 2, device=dev, generator=gen) * 2 - 1) * ur

    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm ----------------
    for _ in range(W):
        env.step(new_action())
        env.reset_done(write_obs=True)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    launches0 = env.launches
    done_rate = 0.0
    t_wall0 = time.perf_counter()
    for k in range(K):
        act = new_action()
        env.action.copy_(act)
        flush.zero_()                      # evict state/obs from L2 between timed iterations
        ev[k][0].record()
        env.step(None)
        ev[k][1].record()
        env.reset_done(write_obs=True)
        ev[k][2].record()
        done_rate += float(env.done.float().mean()) / K
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = env.launches - launches0
    t_step = sum(e[0].elapsed_time(e[1]) for e in ev) * 1e-3      # fused step kernel only
    t_total = sum(e[0].elapsed_time(e[2]) for e in ev) * 1e-3     # + masked reset/respawn + refresh

    # ---------------- end-to-end arm: host buffers through sgb_step_host ----------------
    h_act = [((torch.rand(B, N, 2) * 2 - 1) * ur.cpu()).contiguous().pin_memory() for _ in range(2)]
    for i in range(2):
        env.step_host(h_act[i % 2])
        env.reset_done(write_obs=True)
    barrier()
    Ke = max(3, min(K, 10))
    t0 = time.perf_counter()
    for k in range(Ke):
        h_obs, h_rew, h_done = env.step_host(h_act[k % 2])
        env.reset_done(write_obs=True)
    barrier()
    t_e2e = time.perf_counter() - t0
    sampler.stop()

    times = torch.tensor([t_total, t_step, t_e2e / Ke * K], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_total, t_step, t_e2e_k = [float(x) for x in times]
    agent_steps = B * N * K * world
    value = agent_steps / t_total
    hbm_peak, peak_src = peaks()
    per_gpu_step_rate = B * N * K / t_step
    achieved = ALGO_BYTES(D) * per_gpu_step_rate / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": 1e3 * t_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"cpm_entire num_envs={B} n_agents={N} per GPU (BASELINE configs[2] shape), fused step + masked device reset/respawn with fresh observations for reset envs",
                   "obs_dim": D, "rew_method": "distance", "dt": 0.1, "l2": "flushed (512 MiB write) between timed iterations",
                   "actions": "U(-1,1)^2*[1.0, 31deg]", "done_rate_per_step": round(done_rate, 4),
                   "map_smem_bytes": env.map_bytes},
        "gpu_launches": launches,
        "e2e": {"value": agent_steps / t_e2e_k, "unit": "agent-steps/s",
                "h2d_bytes_per_step": B * N * 2 * 4, "d2h_bytes_per_step": B * N * D * 4 + B * N * 4 + B,
                "note": "sgb_step_host: pinned host action in, obs/reward/done out, plus device reset"},
        "roofline": {"bound": "hbm", "kernel": "env_step_kernel (fused step)", "achieved": achieved, "peak": hbm_peak,
                     "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": ncu_traffic(), "peak_source": peak_src,
                     "algorithmic_bytes_per_agent_step": ALGO_BYTES(D), "kernel_ms": 1e3 * t_step / K,
                     "issue_bound": ncu_issue(),
                     "note": "path is fp32-ALU/shared-memory bound, not HBM bound (DESIGN.md roofline section)"},
        "clocks": sampler.summary(),
        "wall_s_timed_region": t_wall,
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, dt = cpu_baseline_sample(args.ref_envs, N, 4, threads)
            line["cpu_baseline"] = {"value": v, "unit": "agent-steps/s", "cores": threads, "kind": "port",
                                    "sample": f"{args.ref_envs} envs x {N} agents x 4 steps (+resets), C oracle port on {threads} pthreads"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_rollout(args):
    """--workload rollout: BASELINE configs[4] shape per GPU (CPM map, T=128 rollout + GAE + the all-gather of the
    advantage / value-target buffers at PPO-update time).  The policy / critic networks are
    dense NN work outside the path: actions and values are pre-generated on the device."""
    import torch
    import torch.distributed as dist
    from sigmarl_b200 import EnvConfig, RoadTrafficEnv
    from sigmarl_b200.rollout import RolloutBuffer, all_gather_advantages, collect, compute_gae

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, N, T, K, W = args.envs, args.agents, args.horizon, args.steps, max(1, min(args.warmup, 2))
    env = RoadTrafficEnv(EnvConfig(scenario_type="cpm_entire", n_agents=N, mode="params", rew_method="distance"),
                         num_envs=B, device=dev, seed=args.seed, env_offset=rank * B)
    env.reset()
    buf = RolloutBuffer(T, B, N, env.D, dev, world=world, rank=rank)
    ur = torch.tensor([1.0, 31 * np.pi / 180], device=dev)
    gen = torch.Generator(device=dev).manual_seed(99 + rank)
    acts = (torch.rand(T, B, N, 2, device=dev, generator=gen) * 2 - 1) * ur
    buf.value.copy_(torch.rand(T, B, N, device=dev, generator=gen))
    buf.next_value.copy_(torch.rand(T, B, N, device=dev, generator=gen))
    step = {"t": 0}

    def policy(_obs):
        a = acts[step["t"] % T]
        step["t"] += 1
        return a

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one():
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        collect(env, policy, buf)
        ev[1].record()
        compute_gae(buf, 0.99, 0.9)
        ev[2].record()
        all_gather_advantages(buf)
        ev[3].record()
        return ev

    for _ in range(W):
        one()
    barrier()
    l0 = env.launches
    evs = [one() for _ in range(K)]
    barrier()
    t = torch.tensor([sum(e[0].elapsed_time(e[3]) for e in evs), sum(e[0].elapsed_time(e[1]) for e in evs),
                      sum(e[1].elapsed_time(e[2]) for e in evs), sum(e[2].elapsed_time(e[3]) for e in evs)],
                     device=dev, dtype=torch.float64) * 1e-3
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_all, t_col, t_gae, t_ag = [float(x) for x in t]
    if rank == 0:
        print(json.dumps({
            "metric": METRIC + " — full rollout + GAE + all-gather", "value": world * B * N * T * K / t_all,
            "unit": "agent-steps/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": 1e3 * t_all / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"cpm_entire num_envs={B} n_agents={N} per GPU, T={T} rollout (step + masked reset with fresh "
                                   f"obs + [T,B,N,*] buffer writes) + GAE kernel + NCCL all-gather of advantage/value target",
                       "policy": "pre-generated actions/values (NN forward is outside the path)"},
            "breakdown_ms": {"collect": 1e3 * t_col / K, "gae": 1e3 * t_gae / K, "all_gather": 1e3 * t_ag / K},
            "gathered_bytes_per_rank": 2 * T * B * N * 4, "gpu_launches": env.launches - l0 + K}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=65536, help="envs per GPU")
    ap.add_argument("--agents", type=int, default=8)
    ap.add_argument("--ref-envs", type=int, default=4096, help="bounded sample for the CPU arm")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="step", choices=["step", "rollout"],
                    help="step = the headline metric (default); rollout = T-step rollout + GAE + all-gather (extra line)")
    ap.add_argument("--horizon", type=int, default=128, help="rollout length T (max_steps, config.json)")
    args = ap.parse_args()
    if args.workload == "rollout" and args.impl == "ours":
        run_rollout(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
