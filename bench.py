#!/usr/bin/env python
"""bench.py — agent-steps/s of the road-traffic environment step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--scenario cpm_entire] [--envs B] [--agents N] [--rew-method distance] [--actions uniform|gentle]
                    [--no-rollout] [--no-cpu-baseline]

One "step" = one pass of the hot path over one batch: the fused step kernel (dynamics -> collisions ->
reward -> observation -> done) followed by the masked device reset/respawn of finished envs, exactly what a
rollout executes per environment step.  N>1: launched by torchrun, one rank per GPU; envs are sharded by
index with no data-path collective -> weak scaling, `value` = all ranks' agent-steps / max time.

The JSON line also carries
  e2e      the same metric through the host-buffer entry point (sgb_step_reset_host: pinned host actions in;
           reward / done and the observation to act on next — post-reset for finished envs — out), next to the
           measured ceiling of the box's concurrent D2H copies (`d2h_ceiling_gbs`, `frac_of_d2h_ceiling`);
  rollout  BASELINE configs[4] per-GPU shape (32768 envs x 8 agents, T = 128): rollout collection into [T,B,N,*]
           buffers + GAE kernel + the design's ONE collective, the NCCL all-gather of advantage / value target,
           with its own timing (`breakdown_ms.all_gather`);
  roofline both readings: algorithmic HBM bytes / kernel time against the measured HBM peak, and the issue-slot
           reading from the committed ncu capture of this round (profiles/ncu_full_r2.json, commit id inside).

--scenario / --agents select other maps: BASELINE configs[3] is `--scenario on_ramp_2_multilane --agents 12 --envs 8192`
and `--scenario roundabout_2 --agents 12 --envs 8192` on 4 GPUs (8192 envs per GPU).

--impl reference: the reference's CPU implementation of the same path.  The reference is pure Python (it
cannot be compiled into oracle/_ref), so this arm times the C oracle port (oracle/sigmarl_oracle.c, pinned
bit-exactly to the reference's golden vectors) on all host threads, on a bounded sample of the workload.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

ALGO_BYTES = lambda D: 65 + 4 * D  # noqa: E731  SURVEY.md §8(d): algorithmic HBM bytes per agent-step
METRIC = "agent-steps/sec (num_envs x n_agents / step_time)"
NCU_ROUND = "r2"
# measured in the build container with the UNMODIFIED Python reference behind the import shim
PY_REFERENCE_NOTE = ("the Python reference itself: 3035 agent-steps/s step-only at cpm_entire B=4096 N=8, 1437 with resets at "
                     "B=1024 (8 vCPU, torch CPU; BASELINE.md §2) — it cannot travel to the GPU box")


def ncu_capture():
    """Static facts of the committed `ncu --set full` capture of this round's step kernel: DRAM bytes per launch and
    the issue-slot reading (the bound that actually limits the kernel), with the commit the capture was taken at."""
    p = os.path.join(REPO, "profiles", f"ncu_full_{NCU_ROUND}.json")
    try:
        d = json.load(open(p))
        k = d["kernels"]["env_step_kernel<4,0,0>"]
        busy, lanes, winst = k["issue_slots_busy_pct"], k["active_lanes_per_warp_inst"], k["warp_inst_per_launch"]
        shape = d["shape_agents"]
        return {"traffic": k["dram_bytes_read"] + k["dram_bytes_write"],
                "issue_bound": {"issue_slots_busy_pct": busy, "active_lanes_per_warp_inst": lanes,
                                "warp_inst_per_launch": winst,
                                # share of the SMs' lane-issue capacity (4 schedulers x 32 lanes per clock) doing useful
                                # work: the fp32 / ALU reading of the roofline SURVEY.md §8d asks for next to HBM
                                "lane_issue_frac": busy / 100.0 * lanes / 32.0,
                                "thread_inst_per_agent_step": winst * lanes / shape,
                                "kernel_us_under_ncu": k["duration_us"],
                                "source": f"profiles/ncu_full_{NCU_ROUND}.json", "commit": d.get("commit")}}
    except Exception:
        return {"traffic": None, "issue_bound": None}


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


def bind_to_gpu_numa(index):
    """Run this rank on the CPUs next to its GPU BEFORE pinned host buffers are allocated: pinned pages then come from
    that NUMA node, so N ranks copying at once do not all go through one socket's memory controller."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def action_sampler(kind, B, N, dev, gen, env=None):
    import torch
    ur = torch.tensor([1.0, 31 * np.pi / 180], device=dev)
    if kind == "uniform" or env is None:       # SURVEY.md §8d distribution (i): U(-1,1)^2 * [v_max, delta_max]
        return lambda: (torch.rand(B, N, 2, device=dev, generator=gen) * 2 - 1) * ur, "U(-1,1)^2*[1.0, 31deg]"

    # (ii) "gentle": pure pursuit on the 2nd short-term reference point of the observation (ego frame, obs[3:5]) with a
    # little steering noise, speeds 0.5-0.8 m/s — agents follow their paths, episodes run long, the population spreads
    def pursuit():
        a = torch.rand(B, N, 2, device=dev, generator=gen)
        o = env.obs
        steer = torch.clamp(1.5 * torch.atan2(o[..., 4], o[..., 3]) + (a[..., 1] * 2 - 1) * 0.03, -float(ur[1]), float(ur[1]))
        return torch.stack([0.5 + 0.3 * a[..., 0], steer], -1)
    return pursuit, "gentle: pure pursuit on the short-term reference path, v in U(0.5,0.8)"


def cpu_baseline_sample(scenario, n_envs, n_agents, steps, threads, rew_method="distance", seed=0):
    """Time the oracle port on `threads` host threads: `steps` steps of `n_envs` envs (+ resets of done envs)."""
    from oracle import oracle as O
    w = O.OracleWorld(scenario, n_envs, n_agents, mode="params", rew_method=rew_method)
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


def workload_name(args):
    return f"{args.scenario} num_envs={args.envs} n_agents={args.agents} per GPU"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_envs = args.ref_envs
    # warm-up steps, then K timed steps of the bounded sample
    v, dt = cpu_baseline_sample(args.scenario, n_envs, args.agents, args.steps, threads, args.rew_method)
    D = 10 + 11 * min(2, args.agents - 1)
    line = {
        "impl": "reference", "metric": METRIC + f", {args.scenario} map", "value": v, "unit": "agent-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args) + f" (reference arm: bounded sample of {n_envs} envs)",
                   "obs_dim": D, "rew_method": args.rew_method, "dt": 0.1},
        "cpu_baseline": {"value": v, "unit": "agent-steps/s", "cores": threads, "kind": "port",
                         "sample": f"{n_envs} envs x {args.agents} agents x {args.steps} steps (+resets), oracle/sigmarl_oracle.c on "
                                   f"{threads} pthreads; " + PY_REFERENCE_NOTE},
        "e2e": {"value": v, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def measure_d2h_ceiling(nbytes, dev, barrier, reps=8):
    """What the box gives N ranks copying device -> pinned host at once: plain cudaMemcpyAsync of the e2e arm's
    per-step byte count, all ranks started together, GB/s of THIS rank (the caller takes the minimum over ranks)."""
    import torch
    src = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dst = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dst.copy_(src, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    return nbytes * reps / dt / 1e9


def rollout_record(args, dev, world, rank, barrier):
    """BASELINE configs[4] per-GPU shape: T-step rollout into [T,B,N,*] buffers (step + masked reset with fresh
    observations, written in place) + GAE kernel + the NCCL all-gather of advantage / value target at PPO-update time
    (mappo_cavs.py:357-378, helper_training.py:686-788).  The policy / critic networks are dense NN work outside the
    path: actions and values are pre-generated on the device."""
    import torch
    import torch.distributed as dist
    from sigmarl_b200 import EnvConfig, RoadTrafficEnv
    from sigmarl_b200.rollout import RolloutBuffer, all_gather_advantages, collect, compute_gae, gae_allgather

    B, N, T = args.rollout_envs, args.agents, args.horizon
    env = RoadTrafficEnv(EnvConfig(scenario_type=args.scenario, n_agents=N, mode="params", rew_method=args.rew_method),
                         num_envs=B, device=dev, seed=args.seed, env_offset=rank * B)
    env.reset()
    # gather buffers in symmetric (peer-mapped) memory: GAE and the all-gather are ONE kernel (sgb_gae_allgather)
    fused = world > 1 and not args.nccl_gather
    fused_note = None
    if fused:
        # symmetric memory needs peer mappings between all GPUs of the node; if the box refuses them, every rank falls
        # back to sgb_gae + NCCL all-gather together (still GPU kernels + NCCL; the record says which one ran)
        try:
            buf = RolloutBuffer(T, B, N, env.D, dev, world=world, rank=rank, symmetric=True)
            ok = torch.ones(1, device=dev)
        except Exception as e:      # noqa: BLE001
            fused_note = f"symmetric memory unavailable ({type(e).__name__}: {str(e)[:120]})"
            ok = torch.zeros(1, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if not bool(ok):
            fused = False
            fused_note = fused_note or "symmetric memory unavailable on another rank"
    if not fused:
        buf = RolloutBuffer(T, B, N, env.D, dev, world=world, rank=rank)
    gen = torch.Generator(device=dev).manual_seed(99 + rank)
    sample, _ = action_sampler(args.actions, B, N, dev, gen)
    acts = torch.stack([sample() for _ in range(T)])
    buf.value.copy_(torch.rand(T, B, N, device=dev, generator=gen))
    buf.next_value.copy_(torch.rand(T, B, N, device=dev, generator=gen))
    step = {"t": 0}

    def policy(_obs):
        a = acts[step["t"] % T]
        step["t"] += 1
        return a

    def one(mode):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        collect(env, policy, buf)
        ev[1].record()
        if mode == "fused":
            ev[2].record()
            gae_allgather(buf, 0.99, 0.9, multicast=args.multicast)
        else:
            compute_gae(buf, 0.99, 0.9)
            ev[2].record()
            all_gather_advantages(buf)
        ev[3].record()
        return ev

    one("nccl")
    check = None
    if fused:
        # the fused kernel must leave exactly what GAE + NCCL all-gather leave
        torch.cuda.synchronize()
        want = (buf.adv_all.clone(), buf.target_all.clone())
        buf.adv_all.zero_(); buf.target_all.zero_()
        gae_allgather(buf, 0.99, 0.9, multicast=args.multicast)
        torch.cuda.synchronize()
        same = torch.tensor([int(torch.equal(want[0], buf.adv_all) and torch.equal(want[1], buf.target_all))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        check = bool(int(same))
        if not check:
            raise SystemExit("sgb_gae_allgather differs from sgb_gae + NCCL all_gather")
        del want
        one("fused")
    barrier()
    K = args.rollouts
    l0 = env.launches

    def timed(mode):
        evs = [one(mode) for _ in range(K)]
        barrier()
        t = torch.tensor([sum(e[0].elapsed_time(e[3]) for e in evs), sum(e[0].elapsed_time(e[1]) for e in evs),
                          sum(e[1].elapsed_time(e[2]) for e in evs), sum(e[2].elapsed_time(e[3]) for e in evs)],
                         device=dev, dtype=torch.float64) * 1e-3
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    t_all, t_col, t_gae, t_ag = timed("fused" if fused else "nccl")
    launches = env.launches - l0 + K
    gathered = 2 * T * B * N * 4 * world            # advantage + value target of every rank, received by each rank
    rec = {"value": world * B * N * T * K / t_all, "unit": "agent-steps/s",
           "workload": f"{args.scenario} num_envs={B} n_agents={N} per GPU, T={T} (BASELINE configs[4] per-GPU shape): rollout + GAE "
                       f"+ all-gather of advantage / value target",
           "rollouts_timed": K, "ms_per_rollout": 1e3 * t_all / K,
           "breakdown_ms": {"collect": 1e3 * t_col / K, "gae": 1e3 * t_gae / K, "all_gather": 1e3 * t_ag / K},
           "all_gather_share": t_ag / t_all, "gathered_bytes_per_rank": gathered,
           "all_gather_gbs_per_rank": (gathered * (world - 1) / world) / (t_ag / K) / 1e9 if world > 1 and t_ag > 0 else None,
           "collective": ("sgb_gae_allgather: GAE fused with the all-gather — one kernel stores every value into all ranks' "
                          "[world, T, B, N] buffers over NVLink peer memory (torch symmetric memory; "
                          + ("NVSwitch multicast stores" if args.multicast else "one store per peer")
                          + "), barriers before / after; breakdown_ms.gae is 0 and all_gather is the fused kernel") if fused
                         else ("NCCL all_gather_into_tensor, in place ([world, T, B, N] buffers; GAE writes this rank's slot)" if world > 1
                               else "none at 1 GPU (the gather is the identity)"),
           "gpu_launches": launches,
           "policy": "pre-generated actions / values (the NN forward is outside the path)",
           "gae_checked_against": "numpy restatement of the TorchRL recurrence (TorchRL is not installable here)"}
    if fused_note:
        rec["fused_gather_note"] = fused_note
    if fused:
        n_all, n_col, n_gae, n_ag = timed("nccl")
        rec["fused_equals_gae_plus_nccl_all_gather"] = check
        rec["nccl_two_step_ms"] = {"gae": 1e3 * n_gae / K, "all_gather": 1e3 * n_ag / K,
                                   "note": "same rollouts with sgb_gae + NCCL all_gather_into_tensor, for comparison"}
    env.close()
    del env, buf, acts
    torch.cuda.empty_cache()
    return rec


def run_ours(args):
    import torch
    import torch.distributed as dist
    from sigmarl_b200 import EnvConfig, RoadTrafficEnv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    numa_cpus = bind_to_gpu_numa(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, N, K, W = args.envs, args.agents, args.steps, max(3, args.warmup)
    cfg = EnvConfig(scenario_type=args.scenario, n_agents=N, mode="params", rew_method=args.rew_method)
    env = RoadTrafficEnv(cfg, num_envs=B, device=dev, seed=args.seed, env_offset=rank * B)
    env.reset()
    D = env.D
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    new_action, action_desc = action_sampler(args.actions, B, N, dev, gen, env)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm ----------------
    for _ in range(W if args.actions == "uniform" else max(W, 60)):     # path following: let the population spread first
        env.step(new_action())
        env.reset_done(write_obs=True)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    launches0 = env.launches
    done_acc = torch.zeros((), device=dev)
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
        done_acc += env.done.float().mean()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    done_rate = float(done_acc) / K
    launches = env.launches - launches0
    t_step = sum(e[0].elapsed_time(e[1]) for e in ev) * 1e-3      # fused step kernel only
    t_total = sum(e[0].elapsed_time(e[2]) for e in ev) * 1e-3     # + masked reset/respawn + refresh

    # ---------------- end-to-end arm: host buffers through sgb_step_reset_host ----------------
    ur = torch.tensor([1.0, 31 * np.pi / 180])
    h_act = [((torch.rand(B, N, 2) * 2 - 1) * ur).contiguous().pin_memory() for _ in range(2)]
    for i in range(2):
        env.step_host(h_act[i % 2], reset_done=True)
    barrier()
    Ke = max(3, min(K, 10))
    t0 = time.perf_counter()
    for k in range(Ke):
        h_obs, h_rew, h_done = env.step_host(h_act[k % 2], reset_done=True)
    barrier()
    t_e2e = time.perf_counter() - t0
    sampler.stop()
    d2h_bytes = B * N * D * 4 + B * N * 4 + B
    ceiling = measure_d2h_ceiling(d2h_bytes, dev, barrier)

    times = torch.tensor([t_total, t_step, t_e2e / Ke * K, -ceiling], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_total, t_step, t_e2e_k, neg_ceiling = [float(x) for x in times]
    ceiling = -neg_ceiling                                   # slowest rank's D2H rate with all ranks copying
    agent_steps = B * N * K * world
    value = agent_steps / t_total
    hbm_peak, peak_src = peaks()
    per_gpu_step_rate = B * N * K / t_step
    achieved = ALGO_BYTES(D) * per_gpu_step_rate / 1e9
    e2e_step_s = t_e2e_k / K
    cap = ncu_capture() if (args.scenario, B, N) == ("cpm_entire", 65536, 8) else {"traffic": None, "issue_bound": None}
    line = {
        "metric": METRIC + f", {args.scenario} map", "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": 1e3 * t_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args) + ", fused step + masked device reset/respawn with fresh observations for reset envs",
                   "obs_dim": D, "rew_method": args.rew_method, "dt": 0.1, "l2": "flushed (512 MiB write) between timed iterations",
                   "actions": action_desc, "done_rate_per_step": round(done_rate, 4), "map_smem_bytes": env.map_bytes},
        "gpu_launches": launches,
        "e2e": {"value": agent_steps / t_e2e_k, "unit": "agent-steps/s",
                "h2d_bytes_per_step": B * N * 2 * 4, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": 1e3 * e2e_step_s,
                "d2h_gbs_per_gpu": d2h_bytes / e2e_step_s / 1e9, "d2h_ceiling_gbs": ceiling,
                "frac_of_d2h_ceiling": d2h_bytes / e2e_step_s / 1e9 / ceiling,
                "pinned_buffers_numa_bound_cpus": numa_cpus,
                "note": "sgb_step_reset_host: pinned host action in; reward / done of the step and the observation to act on "
                        "next (post-reset for finished envs) out; ceiling = plain cudaMemcpyAsync D2H of the same bytes with all "
                        "ranks copying at once (slowest rank)"},
        "roofline": {"bound": "hbm", "kernel": "sgb::env_step_kernel (fused step)", "achieved": achieved, "peak": hbm_peak,
                     "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": cap["traffic"], "peak_source": peak_src,
                     "algorithmic_bytes_per_agent_step": ALGO_BYTES(D), "kernel_ms": 1e3 * t_step / K,
                     "issue_bound": cap["issue_bound"],
                     "note": "the path is issue / fp32-ALU bound, not HBM bound (DESIGN.md §3.2): the HBM fraction is reported "
                             "because BASELINE.json asks for it, issue_bound.lane_issue_frac is the limiting resource"},
        "clocks": sampler.summary(),
        "wall_s_timed_region": t_wall,
    }
    env.close()
    del env, flush
    torch.cuda.empty_cache()
    if not args.no_rollout:
        line["rollout"] = rollout_record(args, dev, world, rank, barrier)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, dt = cpu_baseline_sample(args.scenario, args.ref_envs, N, 4, threads, args.rew_method)
            line["cpu_baseline"] = {"value": v, "unit": "agent-steps/s", "cores": threads, "kind": "port",
                                    "sample": f"{args.ref_envs} envs x {N} agents x 4 steps (+resets), C oracle port on {threads} "
                                              f"pthreads; " + PY_REFERENCE_NOTE}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scenario", default="cpm_entire", help="map (sigmarl/constants.py SCENARIOS), default the CPM map")
    ap.add_argument("--envs", type=int, default=65536, help="envs per GPU")
    ap.add_argument("--agents", type=int, default=8)
    ap.add_argument("--rew-method", default="distance", choices=["distance", "ttc", "sparse", "distance_sparse", "ttc_sparse"])
    ap.add_argument("--actions", default="uniform", choices=["uniform", "gentle"])
    ap.add_argument("--ref-envs", type=int, default=4096, help="bounded sample for the CPU arm")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nccl-gather", action="store_true", help="rollout record: sgb_gae + NCCL all-gather instead of the fused kernel")
    ap.add_argument("--multicast", type=lambda v: {"auto": None, "1": True, "0": False}[v], default=None,
                    help="fused GAE + all-gather: 1 = NVSwitch multicast stores, 0 / auto = one store per peer (default)")
    ap.add_argument("--no-rollout", action="store_true", help="skip the rollout + GAE + all-gather sub-record")
    ap.add_argument("--rollout-envs", type=int, default=32768, help="envs per GPU of the rollout sub-record")
    ap.add_argument("--rollouts", type=int, default=3, help="timed rollouts of the sub-record")
    ap.add_argument("--horizon", type=int, default=128, help="rollout length T (max_steps, config.json)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
