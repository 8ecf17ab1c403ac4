"""GPU (-m gpu): parity of the CUDA path, called through the C-ABI (ctypes), against
  (1) golden vectors from the unmodified reference (tests/golden, teacher-forced),
  (2) the CPU oracle on seeded random states at sizes the oracle finishes in seconds,
  (3) size-independent properties at BASELINE.json's full sizes.
Tolerances (BASELINE.json north_star): 1e-5 abs for fp32 state / obs / reward, bit-exact masks & indices.
"""
import os

import numpy as np
import pytest
import torch

from conftest import golden_files
from test_oracle_golden import fresh_from_reset

pytestmark = pytest.mark.gpu
TOL = 1e-5
UR = np.asarray([1.0, 31 * np.pi / 180], np.float32)


OBS_FLAG_NAMES = ["is_ego_view", "is_observe_vertices", "is_obs_steering", "is_observe_ref_path_other_agents",
                  "is_observe_distance_to_agents", "is_observe_distance_to_center_line",
                  "is_observe_distance_to_boundaries", "is_apply_mask"]


def _obs_flags_of_golden(g):
    """Observation-layout parameters the reference run was configured with (older fixtures: the defaults)."""
    return {n: bool(g["cfg_" + n]) for n in OBS_FLAG_NAMES if ("cfg_" + n) in g.files}


def _env_from_golden(g, B=None, **over):
    from sigmarl_b200 import EnvConfig, RoadTrafficEnv
    over = {**_obs_flags_of_golden(g), **over}
    for name in ("reset_agent_fixed_duration", "is_use_mtv_distance"):       # fixtures of tests/golden/next/
        if ("cfg_" + name) in g.files:
            over.setdefault(name, g["cfg_" + name].item())
    cfg = EnvConfig(
        scenario_type=str(g["cfg_scenario_type"]), n_agents=int(g["cfg_N"]), mode=str(g["cfg_mode"]),
        dt=float(g["cfg_dt"]), max_steps=int(g["cfg_max_steps"]), rew_method=str(g["cfg_rew_method"]),
        n_nearing_agents_observed=int(g["cfg_n_nearing_agents_observed"]),
        reward_progress=float(g["cfg_reward_progress"]),
        threshold_near_boundary_high=float(g["cfg_near_boundary_high"]),
        threshold_near_boundary_low=float(g["cfg_near_boundary_low"]),
        threshold_near_other_agents_c2c_high=float(g["cfg_near_other_agents_high"]),
        threshold_near_other_agents_c2c_low=float(g["cfg_near_other_agents_low"]),
        ttc_low=float(g["cfg_ttc_low"]), ttc_high=float(g["cfg_ttc_high"]),
        penalty_near_boundary=float(g["cfg_penalty_near_boundary"]),
        penalty_near_other_agents=float(g["cfg_penalty_near_other_agents"]),
        is_testing_mode=bool(g["cfg_is_testing_mode"]), **over)
    return RoadTrafficEnv(cfg, num_envs=B or int(g["cfg_B"]), device="cuda:0", debug=True, info=True)


def _coll_matrix(env):
    cw = env.collide_with.cpu().numpy().astype(np.uint32)
    N = env.N
    return ((cw[..., None] >> np.arange(N, dtype=np.uint32)) & 1).astype(bool)


def _close(name, got, want, ctx):
    err = float(np.max(np.abs(np.asarray(got, np.float64) - np.asarray(want, np.float64)))) if np.size(got) else 0.0
    assert err <= TOL, f"{ctx} {name}: max abs err {err:.3e}"
    return err


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: os.path.basename(p)[:-4])
@pytest.mark.parametrize("exhaustive", [False, True], ids=["pruned", "exhaustive"])
def test_cuda_matches_reference_goldens(path, exhaustive):
    g = np.load(path)
    env = _env_from_golden(g, exhaustive=exhaustive)
    assert env.D == g["obs"].shape[-1]
    assert env.map.max_ref_path_points == int(g["cfg_max_ref_path_points"])
    T, N = int(g["cfg_T"]), env.N
    for t in range(T):
        ctx = f"{os.path.basename(path)} t={t}"
        gp = env.map.global_path(g["pre_scenario_id"][t], g["pre_path_id"][t])
        env.set_state(g["pre_pos"][t], g["pre_rot"][t], g["pre_speed"][t], g["pre_steering"][t], gp,
                      step_count=g["pre_step"][t])
        if not env.config.is_observe_distance_to_boundaries:
            env.set_pose_history(fresh_from_reset(g, t))
        obs, rew, done = env.step(torch.as_tensor(g["action"][t]).cuda())
        torch.cuda.synchronize()
        _close("pos", env.pos.cpu(), g["post_pos"][t], ctx)
        _close("rot", env.rot.cpu(), g["post_rot"][t], ctx)
        _close("speed", env.speed.cpu(), g["post_speed"][t], ctx)
        _close("steering", env.steering.cpu(), g["post_steering"][t], ctx)
        _close("vel", env.vel.cpu(), g["post_vel"][t], ctx)
        _close("sideslip", env.sideslip_angle.cpu(), g["post_sideslip"][t], ctx)
        _close("obs", obs.cpu(), g["obs"][t], ctx)
        _close("reward", rew.cpu(), g["reward"][t], ctx)
        dbg = env.dbg.cpu().numpy()
        _close("d_ref", dbg[..., 0], g["d_ref"][t], ctx)
        _close("d_left_cg", dbg[..., 2], g["d_left"][t][..., 0], ctx)
        _close("d_right_cg", dbg[..., 7], g["d_right"][t][..., 0], ctx)
        if N > 1:  # agent 0's stored vertex distances are one step stale in the reference
            _close("d_left_v", dbg[:, 1:, 3:7], g["d_left"][t][:, 1:, 1:], ctx)
            _close("d_right_v", dbg[:, 1:, 8:12], g["d_right"][t][:, 1:, 1:], ctx)
        _close("d_bound", dbg[..., 12], g["d_bound"][t], ctx)
        assert np.array_equal(dbg[..., 1].view(np.int32), g["idx_ref"][t]), f"{ctx} idx_ref"
        fl = env.agent_flags.cpu().numpy()
        assert np.array_equal(done.cpu().numpy().astype(bool), g["done"][t]), f"{ctx} done"
        assert np.array_equal((fl & 2) != 0, g["col_lane"][t]), f"{ctx} col_lane"
        assert np.array_equal((fl & 4) != 0, g["col_entry"][t]), f"{ctx} col_entry"
        assert np.array_equal((fl & 8) != 0, g["col_exit"][t]), f"{ctx} col_exit"
        assert np.array_equal(_coll_matrix(env), g["col_agents"][t]), f"{ctx} col_agents"
        assert np.array_equal((fl & 1) != 0, g["col_agents"][t].any(-1)), f"{ctx} any col_agents"
        # respawn requests = entry/exit crossers of not-done envs (road_traffic.py:1462-1472); in testing mode
        # every colliding or leaving agent of a not-done env, on every map (:1435-1447)
        if bool(g["cfg_is_testing_mode"]):
            req = ((fl & 15) != 0) & ~g["done"][t][:, None]
        else:
            req = ((fl & 12) != 0) & ~g["done"][t][:, None] & (str(g["cfg_scenario_type"]) != "cpm_entire")
        assert np.array_equal(req, g["respawn_mask"][t]), f"{ctx} respawn"
        assert np.array_equal(env.step_count.cpu().numpy(), g["pre_step"][t] + 1), f"{ctx} step_count"
        # the info block the kernel writes == what the reference's info(agent) returned in that step
        blk = env.info.cpu().numpy()
        _close("info.ref", blk[..., 0:6], g["info_ref"][t], ctx)
        _close("info.distance_ref", blk[..., 6], g["info_distance_ref"][t], ctx)
        _close("info.distance_left_b", blk[..., 7], g["info_distance_left_b"][t], ctx)
        _close("info.distance_right_b", blk[..., 8], g["info_distance_right_b"][t], ctx)
        _close("info.rew_near_other_agents", blk[..., 9], g["info_rew_near_other_agents"][t], ctx)
        _close("info.rew_collide_other_agents", blk[..., 10], g["info_rew_collide_other_agents"][t], ctx)
        _close("info.rew_collide_lane", blk[..., 11], g["info_rew_collide_lane"][t], ctx)
        _close("info.rew_reach_goal", blk[..., 12], g["info_rew_reach_goal"][t], ctx)
        _close("info.rew_total", blk[..., 13], g["info_rew_total"][t], ctx)
        # evaluation counters accumulate over the run (:998-1002, :1029-1035)
        assert np.array_equal(env.task_tries.cpu().numpy(), g["num_task_tries"][t]), f"{ctx} num_task_tries"
        assert np.array_equal(env.task_success.cpu().numpy(), g["task_success_times"][t]), f"{ctx} task_success_times"


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: os.path.basename(p)[:-4])
def test_cuda_reset_obs_matches_reference(path):
    """place + refresh(write_obs) reproduces the reference's post-reset state and its all-fresh observation."""
    g = np.load(path)
    env = _env_from_golden(g)
    n = 0
    for t in range(int(g["cfg_T"])):
        m = g["reset_mask"][t]
        if not m.any():
            continue
        gp = env.map.global_path(g["reset_scenario_id"][t], g["reset_path_id"][t])
        amask = np.repeat(m[:, None], env.N, axis=1)
        env.place(gp, g["reset_point_id"][t], g["reset_speed"][t], agent_mask=amask)
        obs = env.refresh(env_mask=torch.as_tensor(m), write_obs=True)
        torch.cuda.synchronize()
        assert np.array_equal(env.pos.cpu().numpy()[m], g["reset_pos"][t][m])
        assert np.array_equal(env.rot.cpu().numpy()[m], g["reset_rot"][t][m])
        _close("reset_vel", env.vel.cpu().numpy()[m], g["reset_vel"][t][m], f"t={t}")
        _close("reset_obs", obs.cpu().numpy()[m], g["reset_obs"][t][m], f"{os.path.basename(path)} t={t}")
        n += int(m.sum())
    assert n > 0


def _oracle_for(env, O, rew_method, mode):
    over = {}
    if env.config.is_use_mtv_distance:     # MTV thresholds, road_traffic.py:264-270, 632-648
        over.update(use_mtv=True, na_low=env.config.threshold_near_other_agents_MTV_low,
                    na_high=env.config.threshold_near_other_agents_MTV_high)
    if env.config.reset_agent_fixed_duration:
        over.update(fixed_duration=env.config.reset_agent_fixed_duration)
    return O.OracleWorld(env.config.scenario_type, env.B, env.N, mode=mode, rew_method=rew_method,
                         n_nearing_agents_observed=env.config.n_nearing_agents_observed,
                         max_steps=env.config.max_steps,
                         obs_flags=O.obs_flags_from(lambda n: getattr(env.config, n)), **over)


@pytest.mark.parametrize("scenario,N,rew,mode,B", [
    ("cpm_entire", 8, "distance", "params", 1024),
    ("cpm_entire", 8, "ttc_sparse", "kwargs", 512),
    ("cpm_mixed", 8, "distance_sparse", "params", 512),
    ("intersection_1", 2, "distance", "kwargs", 256),
    ("on_ramp_2_multilane", 12, "ttc", "kwargs", 256),
    ("roundabout_2", 12, "sparse", "params", 256),
    ("cpm_entire", 8, "ttc", "params", 256),           # run with n_nearing_agents_observed = 5 below (k > 2 path)
    ("cpm_entire", 15, "ttc_sparse", "params", 128),       # the reference's default n_agents on this map (G = 2)
    ("cpm_entire", 18, "distance_sparse", "params", 96),   # N > 16: one lane per agent (G = 1 instantiation)
    ("cpm_entire", 1, "distance", "params", 64),           # single agent: nobody to observe, k = 0
])
def test_cuda_matches_oracle_free_running_with_device_resets(oracle_mod, scenario, N, rew, mode, B):
    """GPU drives (device resets included); every step the oracle is teacher-forced from the GPU's pre-step state."""
    k_obs = 5 if (scenario, rew) == ("cpm_entire", "ttc") else 2   # 5: neighbours beyond the two kept in registers
    env = _free_run(oracle_mod, scenario, N, rew, mode, B, k_obs)
    assert env.D == 10 + 11 * min(k_obs, N - 1)


@pytest.mark.parametrize("scenario,N,rew,mode,B,k_obs,flags", [
    # every layout flag at once, ego view; G = 4 / 2 / 1 instantiations of the flag-driven writer
    ("cpm_entire", 8, "distance", "params", 512, 2,
     dict(is_observe_vertices=False, is_obs_steering=True, is_observe_ref_path_other_agents=True,
          is_observe_distance_to_agents=False, is_observe_distance_to_center_line=False)),
    ("on_ramp_2_multilane", 12, "ttc", "kwargs", 256, 3, dict(is_obs_steering=True, is_observe_ref_path_other_agents=True)),
    ("cpm_entire", 18, "distance_sparse", "params", 64, 4, dict(is_observe_vertices=False, is_obs_steering=True)),
    # bird view: default fields, and everything switched
    ("cpm_entire", 8, "ttc_sparse", "params", 512, 2, dict(is_ego_view=False)),
    ("cpm_mixed", 6, "distance", "params", 256, 5,
     dict(is_ego_view=False, is_observe_vertices=False, is_obs_steering=True, is_observe_ref_path_other_agents=True,
          is_observe_distance_to_agents=False, is_observe_distance_to_center_line=False)),
    ("intersection_1", 3, "distance", "kwargs", 256, 2, dict(is_ego_view=False, is_observe_ref_path_other_agents=True)),
    # boundary points instead of boundary distances (exact argmin scans on both boundaries, reset / step history bit)
    ("cpm_entire", 8, "distance", "params", 512, 2, dict(is_observe_distance_to_boundaries=False)),
    ("cpm_mixed", 6, "ttc", "params", 256, 2,
     dict(is_observe_distance_to_boundaries=False, is_ego_view=False, is_obs_steering=True)),
    ("on_ramp_2_multilane", 12, "distance", "kwargs", 128, 2,
     dict(is_observe_distance_to_boundaries=False, is_observe_vertices=False)),
    ("cpm_entire", 18, "sparse", "params", 64, 2, dict(is_observe_distance_to_boundaries=False)),
])
def test_cuda_observation_layouts_match_oracle(oracle_mod, scenario, N, rew, mode, B, k_obs, flags):
    """Non-default observation layouts (observation_provider_rt.py:594-925; SGB_OBS_*), GPU vs the oracle that is
    pinned on the reference's own outputs for these layouts (tests/golden/obsvar_*.npz)."""
    env = _free_run(oracle_mod, scenario, N, rew, mode, B, k_obs, **flags)
    assert env.D == env.config.obs_dim(N) and env.config.obs_flags() != 0


@pytest.mark.parametrize("layout", [dict(), dict(is_ego_view=False, is_obs_steering=True)], ids=["default", "birdview"])
def test_observation_noise_is_uniform_additive_and_reproducible(layout):
    """is_obs_noise (observation_provider_rt.py:611-617): obs + level * U[0,1) on every element, fresh draws on every
    call (torch.rand_like).  The reference draws from torch's global generator, the library from a counter-based device
    generator keyed by (seed, API-call counter, GLOBAL env index, agent, column) — distribution-equivalent: noisy - clean
    must lie in [0, level), look uniform, differ between columns / agents / envs / steps (also when the state does NOT
    change, and between envs in identical states), be reproducible, and not depend on how envs are sharded."""
    from sigmarl_b200 import EnvConfig, RoadTrafficEnv
    B, N, level = 4096, 8, 0.05
    mk = lambda noise, seed=0, nenv=B, off=0: RoadTrafficEnv(  # noqa: E731
        EnvConfig(scenario_type="cpm_entire", n_agents=N, is_obs_noise=noise, obs_noise_seed=seed, **layout),
        num_envs=nenv, device="cuda:0", seed=3, env_offset=off)
    clean, noisy, noisy2, other_seed = mk(False), mk(True), mk(True), mk(True, seed=1)
    half = mk(True, nenv=B // 2, off=B // 2)        # the second shard of the same batch
    assert noisy.D == clean.D
    g = torch.Generator(device="cuda").manual_seed(1)
    ur = torch.as_tensor(UR).cuda()
    for e in (clean, noisy, noisy2, other_seed, half):
        e.reset()
    prev = None
    for t in range(4):
        act = (torch.rand(B, N, 2, device="cuda", generator=g) * 2 - 1) * ur
        half.pose.copy_(clean.pose[B // 2:]); half.aux.copy_(clean.aux[B // 2:]); half.carry.copy_(clean.carry[B // 2:])
        half.path_id.copy_(clean.path_id[B // 2:]); half.step_count.copy_(clean.step_count[B // 2:])
        for e in (clean, noisy, noisy2, other_seed):
            e.step(act)
        half.step(act[B // 2:])
        d = (noisy.obs - clean.obs).double()
        # same state trajectory (noise touches the observation only)
        assert torch.equal(noisy.pose, clean.pose) and torch.equal(noisy.reward, clean.reward)
        assert float(d.min()) >= -1e-6 and float(d.max()) < level + 1e-6
        u = d / level
        n = u.numel()
        assert abs(float(u.mean()) - 0.5) < 4 / (12 * n) ** 0.5 + 1e-4          # 4 sigma of a uniform mean (+ fp32 rounding)
        assert abs(float(u.var()) - 1 / 12) < 2e-3
        cols = u.reshape(-1, noisy.D)
        assert float((cols.mean(0) - 0.5).abs().max()) < 0.02                   # every column on its own
        c = torch.corrcoef(cols[:, :6].T)
        assert float((c - torch.eye(6, device="cuda", dtype=c.dtype)).abs().max()) < 0.03   # columns uncorrelated
        assert torch.equal(noisy.obs, noisy2.obs)                               # reproducible
        assert not torch.equal(noisy.obs, other_seed.obs)                       # seed matters
        assert torch.equal(half.obs, noisy.obs[B // 2:])                        # sharding does not
        if prev is not None:
            assert float(((d - prev).abs() > 1e-4).double().mean()) > 0.95      # fresh draws every step
        prev = d
        for e in (clean, noisy, noisy2, other_seed, half):
            e.reset_done(write_obs=False)
        # resets use the same RNG stream in all four envs: the trajectories stay identical
        assert torch.equal(noisy.pose, clean.pose)
    # the observation right after a reset is noisy as well (get_observation is the same call)
    a, b = clean.reset(), noisy.reset()
    dd = (b - a).double()
    assert float(dd.min()) >= -1e-6 and float(dd.max()) < level + 1e-6 and float(dd.mean()) > 0.4 * level
    # The noise is NOT a function of the state: (1) envs put into the very same state draw different noise, (2) a
    # refresh of an unchanged state draws new noise (a deterministic policy must not see B copies of one rollout).
    for e in (clean, noisy):
        for name in ("pose", "aux", "carry", "path_id"):
            t_ = getattr(e, name)
            t_.copy_(t_[:1].expand_as(t_).clone())
    o_clean = clean.refresh(write_obs=True).clone()
    o1 = noisy.refresh(write_obs=True).clone()
    o2 = noisy.refresh(write_obs=True).clone()
    assert torch.equal(o_clean, o_clean[:1].expand_as(o_clean))                # identical states, identical clean rows
    n1, n2 = (o1 - o_clean).double() / level, (o2 - o_clean).double() / level
    assert float((n1[1:] - n1[:1]).abs().mean()) > 0.25                         # |U - U'| has mean 1/3
    assert float((n1 - n2).abs().mean()) > 0.25
    cc = torch.corrcoef(torch.stack([n1[0].flatten(), n1[1].flatten(), n2[0].flatten()]))
    assert float((cc - torch.eye(3, device="cuda", dtype=cc.dtype)).abs().max()) < 0.2


# closest-index ties per configuration of the oracle-vs-GPU free runs, as measured on a B200 (see _free_run)
TIES_MEASURED = {
    # every other configuration of the free runs: 0 (round 2, B200, 3.0e6 agent-steps in all)
    "on_ramp_2_multilane|N=12|ttc|kwargs|k=2|": 1,                                                     # 1.1e-5 per agent-step
    "on_ramp_2_multilane|N=12|ttc|kwargs|k=3|is_obs_steering=1,is_observe_ref_path_other_agents=1": 1,
}


def _free_run(oracle_mod, scenario, N, rew, mode, B, k_obs, **flags):
    from sigmarl_b200 import EnvConfig, RoadTrafficEnv
    O = oracle_mod
    env = RoadTrafficEnv(EnvConfig(scenario_type=scenario, n_agents=N, mode=mode, rew_method=rew,
                                   n_nearing_agents_observed=k_obs, **flags),
                         num_envs=B, device="cuda:0", seed=7, debug=True)
    w = _oracle_for(env, O, rew, mode)
    assert env.D == w.D
    env.reset()
    rng = np.random.default_rng(0)
    n_done = n_lane = n_a2a = n_ties = 0
    for t in range(30):
        torch.cuda.synchronize()
        w.set_state(env.pos.cpu().numpy(), env.rot.cpu().numpy(), env.speed.cpu().numpy(),
                    env.steering.cpu().numpy(), env.path_id.cpu().numpy())
        w.step_count[:] = env.step_count.cpu().numpy()
        w.near_fresh[:] = env.pose_from_reset.cpu().numpy()
        # the carried (stale) values on the GPU come from its own history, the oracle's from a refresh
        if t % 3 == 0:
            act = ((rng.random((B, N, 2), np.float32) * 2 - 1) * UR).astype(np.float32)
        else:  # gentler actions: longer episodes, more agent-agent encounters
            act = np.stack([0.3 + 0.5 * rng.random((B, N), np.float32),
                            (rng.random((B, N), np.float32) * 2 - 1) * 0.15], -1).astype(np.float32)
        obs, rew_, done = env.step(torch.as_tensor(act).cuda())
        o_obs, o_rew, o_done, o_resp = w.step(act, n_threads=8)
        torch.cuda.synchronize()
        ctx = f"{scenario} t={t}"
        _close("pos", env.pos.cpu(), w.pos, ctx)
        _close("rot", env.rot.cpu(), w.rot, ctx)
        _close("vel", env.vel.cpu(), w.vel, ctx)
        # Closest-point index: bit-identical whenever the post-step position is bit-identical.  CUDA libm and glibc
        # differ by <= 1 ulp in sin/cos/tan/atan, so a position may differ by 1 ulp; an agent whose closest point is
        # a polyline vertex has two segments tied to ~1e-9 m and the 1-ulp shift can flip the argmin (rate ~1e-5 per
        # agent-step).  Exactly those certified ties are tolerated (and excluded from the obs comparison below).
        dbg = env.dbg.cpu().numpy()
        gi = dbg[..., 1].view(np.int32)
        tie = gi != w.idx_ref
        if tie.any():
            same_pos = (env.pos.cpu().numpy() == w.pos).all(-1)
            assert not (tie & same_pos).any(), f"{ctx} idx_ref differs at a bit-identical position"
            assert np.all(np.abs(dbg[..., 0][tie] - w.d_ref[tie]) <= 1e-6), f"{ctx} idx_ref differs without a tie"
            assert np.abs(gi - w.idx_ref)[tie].max() == 1 and tie.sum() <= 4, f"{ctx} too many / non-adjacent ties"
            n_ties += int(tie.sum())
        # (a layout that shows other agents' reference paths exposes a neighbour's tie as well: skip the whole env)
        ok_rows = ~(tie | (tie.any(-1, keepdims=True) & env.config.is_observe_ref_path_other_agents))
        if not env.config.is_observe_distance_to_boundaries:
            # the same kind of tie on a BOUNDARY's closest index: only agent 0 scans at the post-step position (which
            # may differ from the oracle's by an ulp; agents >= 1 scan at the bit-identical pre-step position)
            bad = (np.abs(obs.cpu().numpy() - o_obs).max(-1) > TOL) & ok_rows
            assert not bad[:, 1:].any(), f"{ctx} boundary points of an agent >= 1 differ"
            assert bad.sum() <= 1, f"{ctx} {int(bad.sum())} agent-0 rows differ in one step"
            n_ties += int(bad.sum())
            ok_rows &= ~bad
        if int(env.cfg.obs_flags) & 256:
            # lanelet-relation mask: the current lanelet is an argmin over lanelet points at the post-step position, which
            # may differ from the oracle's by an ulp (libm); an agent within ~1e-7 m of the border between two lanelets'
            # nearest-point cells can come out on the other lanelet (expected < 1e-5 per agent-step)
            bad = (np.abs(obs.cpu().numpy() - o_obs).max(-1) > TOL) & ok_rows
            # (one agent on the other lanelet changes the rows of everybody who observes it: at most one env per step)
            assert bad.any(-1).sum() <= 1, f"{ctx} rows differ in {int(bad.any(-1).sum())} envs in one step"
            n_ties += int(bad.any(-1).sum())
            ok_rows &= ~bad
        _close("obs", obs.cpu().numpy()[ok_rows], o_obs[ok_rows], ctx)
        _close("reward", rew_.cpu(), o_rew, ctx)
        fl = env.agent_flags.cpu().numpy()
        assert np.array_equal(done.cpu().numpy().astype(bool), o_done), f"{ctx} done"
        assert np.array_equal((fl & 2) != 0, w.col_lane.astype(bool)), f"{ctx} col_lane"
        assert np.array_equal(_coll_matrix(env), w.col_agents.astype(bool)), f"{ctx} col_agents"
        assert np.array_equal((fl & 4) != 0, w.col_entry.astype(bool)), f"{ctx} col_entry"
        assert np.array_equal((fl & 8) != 0, w.col_exit.astype(bool)), f"{ctx} col_exit"
        n_done += int(o_done.sum()); n_lane += int(w.col_lane.sum()); n_a2a += int(w.col_agents.sum())
        env.reset_done()
    assert n_done > 0 and n_lane > 0
    # Tie budget: certified closest-index ties between CUDA libm and glibc poses.  The counts are
    # deterministic (fixed seeds); the measured ones are listed in TIES_MEASURED and the run may use three times that
    # (plus two), far below the 1e-4 per agent-step this used to tolerate; no configuration may exceed 2e-5 per agent-step.
    key = f"{scenario}|N={N}|{rew}|{mode}|k={k_obs}|" + ",".join(f"{k}={int(v)}" for k, v in sorted(flags.items()))
    print(f"TIES {key} -> {n_ties} in {B * N * 30} agent-steps ({n_ties / (B * N * 30):.2e})")
    assert n_ties <= 3 * TIES_MEASURED.get(key, 0) + 2, f"{key}: {n_ties} ties, measured before: {TIES_MEASURED.get(key, 0)}"
    assert n_ties <= max(2, 2e-5 * B * N * 30), f"{key}: tie rate {n_ties / (B * N * 30):.2e} > 2e-5"
    env.n_ties = n_ties
    return env


def test_pruned_equals_exhaustive_bitwise_at_c2_size():
    """BASELINE configs[1] size (B=8192, N=8): the pruned search must not change a single bit."""
    from sigmarl_b200 import EnvConfig, RoadTrafficEnv
    B, N = 8192, 8
    envs = [RoadTrafficEnv(EnvConfig(scenario_type="cpm_entire", n_agents=N, rew_method="ttc_sparse",
                                     threshold_near_other_agents_c2c_low=0.1635, exhaustive=ex),
                           num_envs=B, device="cuda:0", seed=11, debug=True) for ex in (False, True)]
    for e in envs:
        e.reset()
    g = torch.Generator(device="cuda").manual_seed(3)
    ur = torch.as_tensor(UR).cuda()
    for t in range(12):
        act = (torch.rand(B, N, 2, device="cuda", generator=g) * 2 - 1) * ur
        if t % 2:
            act[..., 0] = act[..., 0].abs() * 0.6 + 0.2
            act[..., 1] *= 0.2
        for e in envs:
            e.step(act)
        for name in ("pose", "aux", "carry", "obs", "reward", "done", "agent_flags", "collide_with", "step_count", "dbg"):
            a, b = getattr(envs[0], name), getattr(envs[1], name)
            assert torch.equal(a, b), f"t={t} {name} differs between pruned and exhaustive search"
        for e in envs:
            e.reset_done()
        assert torch.equal(envs[0].pose, envs[1].pose)


def test_properties_at_full_size_and_sharding_invariance():
    """BASELINE configs[2] size (B=65536, N=8): invariants + env-sharded run == unsharded run, bit for bit."""
    from sigmarl_b200 import EnvConfig, RoadTrafficEnv
    B, N = 65536, 8
    cfg = EnvConfig(scenario_type="cpm_entire", n_agents=N, rew_method="distance")
    full = RoadTrafficEnv(cfg, num_envs=B, device="cuda:0", seed=5)
    halves = [RoadTrafficEnv(cfg, num_envs=B // 2, device="cuda:0", seed=5, env_offset=k * (B // 2)) for k in range(2)]
    full.reset()
    for h in halves:
        h.reset()
    # reset invariants (world_state_rt_sim.py:215-311)
    pos = full.pos
    d = (pos[:, :, None, :] - pos[:, None, :, :]).norm(dim=-1) + torch.eye(N, device="cuda") * 10
    assert float(d.min()) >= 0.3669, "agents spawned closer than reset_agent_min_distance"
    assert int(full.n_failed.item()) == 0 and int(full.step_count.abs().sum()) == 0
    g = torch.Generator(device="cuda").manual_seed(1)
    ur = torch.as_tensor(UR).cuda()
    dones = 0
    for t in range(10):
        act = (torch.rand(B, N, 2, device="cuda", generator=g) * 2 - 1) * ur
        obs, rew, done = full.step(act)
        for k, h in enumerate(halves):
            h.step(act[k * (B // 2):(k + 1) * (B // 2)])
        assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
        assert float(rew.min()) >= -1.0 and float(rew.max()) <= 1.0
        fl = full.agent_flags
        cw = full.collide_with
        # collision symmetry: bit j of agent i == bit i of agent j (world_state_rt_sim.py:389-393)
        m = ((cw.unsqueeze(-1) >> torch.arange(N, device="cuda", dtype=torch.int32)) & 1).bool()
        assert torch.equal(m, m.transpose(1, 2)) and not m.diagonal(dim1=1, dim2=2).any()
        assert torch.equal((fl & 1) != 0, m.any(-1))
        want_done = ((fl & 3) != 0).any(dim=1) | (full.step_count == cfg.max_steps - 1)
        assert torch.equal(done.bool(), want_done)
        for name in ("pose", "aux", "obs", "reward", "done", "agent_flags", "carry"):
            cat = torch.cat([getattr(h, name) for h in halves], dim=0)
            assert torch.equal(cat, getattr(full, name)), f"t={t} sharded {name} != unsharded"
        dones += int(done.sum())
        full.reset_done()
        for h in halves:
            h.reset_done()
        assert torch.equal(torch.cat([h.pose for h in halves], 0), full.pose), "sharded reset != unsharded reset"
    assert dones > 0


def test_step_is_idempotent_under_refresh():
    """Rebuilding the carried values from the pose (sgb_refresh) must reproduce what the step itself carried."""
    from sigmarl_b200 import EnvConfig, RoadTrafficEnv
    env = RoadTrafficEnv(EnvConfig(scenario_type="cpm_mixed", n_agents=6, rew_method="distance"), num_envs=2048,
                         device="cuda:0", seed=2)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(9)
    for t in range(6):
        act = torch.stack([torch.rand(env.B, env.N, device="cuda", generator=g) * 0.7,
                           (torch.rand(env.B, env.N, device="cuda", generator=g) - 0.5) * 0.3], -1)
        env.step(act)
        carried = env.carry.clone()
        aux = env.aux.clone()
        env.refresh()
        assert torch.equal(carried, env.carry)
        assert torch.equal(aux, env.aux)
        env.reset_done()


@pytest.mark.parametrize("scenario,N", [("cpm_entire", 8), ("cpm_mixed", 6), ("on_ramp_2_multilane", 12)])
def test_spawn_table_reset_equals_generic_refresh(scenario, N):
    """A device reset takes carry / boundary distances of the spawn pose from the table built at context creation.
    They must be bit-identical to what the generic polyline scan (sgb_refresh) derives from the same pose — carry,
    aux, the all-fresh observation and the info block of reset envs — and respawned agents of not-done envs must
    keep their step-time observation."""
    from sigmarl_b200 import EnvConfig, RoadTrafficEnv
    env = RoadTrafficEnv(EnvConfig(scenario_type=scenario, n_agents=N, rew_method="distance"), num_envs=1024,
                         device="cuda:0", seed=7, info=True)
    obs0 = env.reset().clone()
    carry0, aux0, info0 = env.carry.clone(), env.aux.clone(), env.info.clone()
    env.refresh(write_obs=True)                      # generic scan of the same poses
    assert torch.equal(carry0, env.carry) and torch.equal(aux0, env.aux)
    assert torch.equal(obs0, env.obs) and torch.equal(info0[..., :9], env.info[..., :9])
    g = torch.Generator(device="cuda").manual_seed(1)
    ur = torch.as_tensor(UR).cuda()
    n_reset = n_resp = 0
    for t in range(12 if scenario == "cpm_entire" else 70):
        if scenario == "cpm_entire":
            act = (torch.rand(env.B, N, 2, device="cuda", generator=g) * 2 - 1) * ur
        else:   # drive forward so that agents also leave through exit segments
            o = env.obs
            steer = torch.clamp(1.5 * torch.atan2(o[..., 4], o[..., 3]), -float(UR[1]), float(UR[1]))
            act = torch.stack([0.6 + 0.3 * torch.rand(env.B, N, device="cuda", generator=g), steer], -1)
        obs_step = env.step(act)[0].clone()
        done = env.done.bool().clone()
        resp = ((env.agent_flags & 12) != 0) & ~done[:, None] & bool(env.cfg.respawn_on_exit)
        env.reset_done(write_obs=True)
        obs_r, carry_r, aux_r = env.obs.clone(), env.carry.clone(), env.aux.clone()
        touched = done | resp.any(1)
        # envs the reset did not fully re-place keep their step-time observation, respawned agents included
        assert torch.equal(obs_r[~done], obs_step[~done])
        env.refresh(write_obs=True)                  # generic scan of everything
        assert torch.equal(carry_r, env.carry) and torch.equal(aux_r, env.aux)
        assert torch.equal(obs_r[done], env.obs[done])
        assert int(env.agent_flags[touched].sum()) == 0
        n_reset += int(done.sum()); n_resp += int(resp.sum())
    assert n_reset > 0 and (scenario != "cpm_mixed" or n_resp > 0)


def test_vmas_facade_drives_the_same_kernel():
    """The BaseScenario-shaped facade (make_world / world.step / reward / observation / done / reset_world_at)."""
    from sigmarl_b200 import EnvConfig, RoadTrafficEnv, make_env
    env = make_env(scenario_type="cpm_mixed", num_envs=64, device="cuda:0", n_agents=4, seed=3, max_steps=128,
                   is_obs_noise=False)     # (the facade's kwargs-mode default is the reference's: noise on)
    sc = env.scenario
    ref = RoadTrafficEnv(EnvConfig(scenario_type="cpm_mixed", n_agents=4, mode="kwargs"), num_envs=64, device="cuda:0", seed=3)
    ref.reset()
    assert torch.equal(ref.pose, sc.env.pose)
    assert sc.world.agents[0].state.pos.shape == (64, 2) and sc.world.agents[0].state.rot.shape == (64, 1)
    assert sc.world.agents[0].dynamics.needed_action_size == 2
    g = torch.Generator(device="cuda").manual_seed(0)
    for t in range(5):
        acts = [torch.stack([torch.rand(64, device="cuda", generator=g) * 0.8,
                             (torch.rand(64, device="cuda", generator=g) - 0.5) * 0.4], -1) for _ in range(4)]
        obs, rews, dones, infos = env.step(acts)
        r_obs, r_rew, r_done = ref.step(torch.stack(acts, 1))
        assert torch.equal(torch.stack(obs, 1), r_obs) and torch.equal(torch.stack(rews, 1), r_rew)
        assert torch.equal(dones, r_done.bool())
        assert infos[0]["pos"].shape == (64, 2)
        for e in torch.where(dones)[0].tolist():
            env.reset_at(e)
        ref.reset_done()
        # both paths reset the same envs; positions of non-reset envs must agree exactly
        keep = ~dones
        touched = ((ref.agent_flags & 12) != 0).any(1)
        assert torch.equal(sc.env.pose[keep & ~touched], ref.pose[keep & ~touched])
        ref.pose.copy_(sc.env.pose); ref.aux.copy_(sc.env.aux); ref.path_id.copy_(sc.env.path_id)
        ref.carry.copy_(sc.env.carry); ref.step_count.copy_(sc.env.step_count)


@pytest.mark.parametrize("B", [512, 2 * 9472 + 40])
def test_host_buffer_step_matches_device_step(B):
    """sgb_step_host (chunked copy / compute pipeline: whole kernel waves of 4736 envs at N = 8, at least two chunks from
    two waves on — the larger size runs as chunks of 14208 + 4776 envs on two streams) == sgb_step on every buffer."""
    from sigmarl_b200 import EnvConfig, RoadTrafficEnv
    a = RoadTrafficEnv(EnvConfig(scenario_type="cpm_entire", n_agents=8), num_envs=B, device="cuda:0", seed=4, info=True)
    b = RoadTrafficEnv(EnvConfig(scenario_type="cpm_entire", n_agents=8), num_envs=B, device="cuda:0", seed=4, info=True)
    a.reset(); b.reset()
    for _ in range(3):
        act = ((torch.rand(B, 8, 2) * 2 - 1) * torch.as_tensor(UR)).contiguous().pin_memory()
        h_obs, h_rew, h_done = a.step_host(act)
        obs, rew, done = b.step(act.cuda())
        torch.cuda.synchronize()
        assert torch.equal(h_obs, obs.cpu()) and torch.equal(h_rew, rew.cpu()) and torch.equal(h_done, done.cpu())
        for name in ("pose", "aux", "carry", "agent_flags", "collide_with", "step_count", "info", "task_tries", "task_success"):
            assert torch.equal(getattr(a, name), getattr(b, name)), name
        a.reset_done(); b.reset_done()


@pytest.mark.parametrize("B,scenario,N", [(300, "cpm_mixed", 4), (2 * 9472 + 40, "cpm_entire", 8)])
def test_host_buffer_step_with_reset_delivers_the_post_reset_observation(B, scenario, N):
    """sgb_step_reset_host = one collector iteration for a host-resident policy: step, then (chunk by chunk, inside the
    copy / compute pipeline) the masked device reset; the host gets reward / done of the step and the observation to act
    on NEXT — post-reset for the envs that finished, step-time for the rest.  Must equal sgb_step + sgb_reset(write_obs)
    on every buffer, with the same reset draws whatever the chunking."""
    from sigmarl_b200 import EnvConfig, RoadTrafficEnv
    mk = lambda: RoadTrafficEnv(EnvConfig(scenario_type=scenario, n_agents=N), num_envs=B, device="cuda:0", seed=4,  # noqa: E731
                                info=True, env_offset=1000)
    a, b = mk(), mk()
    a.reset(); b.reset()
    n_done = 0
    for _ in range(4):
        act = ((torch.rand(B, N, 2) * 2 - 1) * torch.as_tensor(UR)).contiguous().pin_memory()
        h_obs, h_rew, h_done = a.step_host(act, reset_done=True)
        _, rew, done = b.step(act.cuda())
        rew, done = rew.clone(), done.clone()
        b.reset_done(write_obs=True)
        torch.cuda.synchronize()
        assert torch.equal(h_rew, rew.cpu()) and torch.equal(h_done, done.cpu())
        assert torch.equal(h_obs, b.obs.cpu()), "host observation == step-time obs with the reset envs' rows refreshed"
        for name in ("pose", "aux", "carry", "path_id", "agent_flags", "collide_with", "step_count", "obs", "n_failed"):
            assert torch.equal(getattr(a, name), getattr(b, name)), name
        n_done += int(done.sum())
        d = done.bool().cpu()
        if d.any():     # the rows of finished envs really are fresh ones: own speed entry == |v| of the NEW pose / v_max
            assert torch.allclose(h_obs[d][..., 0], a.speed.cpu()[d], atol=1e-6)
    assert n_done > 0


RESET_FIXTURES = sorted(os.listdir(os.path.join(os.path.dirname(__file__), "golden", "resets"))) \
    if os.path.isdir(os.path.join(os.path.dirname(__file__), "golden", "resets")) else []


@pytest.mark.parametrize("fixture", RESET_FIXTURES, ids=lambda f: f[:-4])
def test_device_reset_draws_follow_the_references_distribution(fixture):
    """SURVEY.md §8f-3.  The device reset is distribution-equivalent, not stream-equivalent, to the reference's
    (world_state_rt_sim.py:215-358): path ~ U{paths of the env's set}, point ~ U[3, n/2), speed ~ U(0, v_max), agents placed
    one after the other with rejection on the distance to the ones before; on cpm_mixed one path set per env ~
    multinomial(cpm_scenario_probabilities).  Checked against (1) the law itself where it is known in closed form (first
    agent, speeds, set frequencies) and (2) what the UNMODIFIED reference drew in 1200-2400 full resets
    (oracle/gen_reset_draws.py -> tests/golden/resets/): two-sample tests on every agent's path / point / speed marginals
    and on the distances between agents, i.e. on the conditional acceptance.  Family-wise alpha = 1e-3 over all tests of a
    fixture (Bonferroni; seeds are fixed on both sides, so this is a deterministic regression check, not a flaky one)."""
    from scipy import stats
    from sigmarl_b200 import EnvConfig, RoadTrafficEnv
    ALPHA = 1e-3
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "resets", fixture))
    st, N = str(g["cfg_scenario_type"]), int(g["cfg_N"])
    probs = tuple(float(x) for x in g["cfg_probabilities"])
    B = 16384
    # (the reference's rejection loop is unbounded; the device reset gives up after max_reset_tries and reports it —
    # with enough tries nothing is reported and the two laws coincide)
    env = RoadTrafficEnv(EnvConfig(scenario_type=st, n_agents=N, cpm_scenario_probabilities=probs), num_envs=B,
                         device="cuda:0", seed=5, max_reset_tries=4096)
    m = env.map
    path, pos, speed, sid = [], [], [], []
    for _ in range(4):
        env.reset()
        torch.cuda.synchronize()
        assert int(env.n_failed) == 0
        path.append(env.path_id.cpu().numpy()); pos.append(env.pos.cpu().numpy()); speed.append(env.speed.cpu().numpy())
        sid.append(env.scenario_id.cpu().numpy())
    path, pos, speed, sid = (np.concatenate(x) for x in (path, pos, speed, sid))
    S = len(path)
    # point index of every spawn pose: the pose IS a centre point of its path
    point = np.full((S, N), -1, np.int64)
    for p in range(m.n_paths):
        c = m.center_xy[m.center_off[p]:m.center_off[p + 1]]
        sel = np.argwhere(path == p)
        if len(sel):
            d = np.abs(pos[sel[:, 0], sel[:, 1]][:, None, :] - c[None]).sum(-1)
            k = d.argmin(1)
            assert (d[np.arange(len(k)), k] == 0).all(), "a spawn pose is exactly a centre point"
            point[sel[:, 0], sel[:, 1]] = k
    n_c = m.n_center[path]
    assert (point >= 3).all() and (point < n_c // 2).all()
    set_of = m.set_of_path(path)
    assert (set_of == sid[:, None]).all(), "all agents of an env drive on paths of the env's set"
    lo = np.asarray([m.set_range[s][0] for s in m.set_names])
    n_set = np.asarray([m.set_range[s][1] - m.set_range[s][0] for s in m.set_names])
    # reference sample in the same terms (its path ids count within the set; scenario ids are 1-based on cpm_mixed)
    r_sid = np.maximum(g["scenario_id"].astype(np.int64) - 1, 0)
    r_path = lo[r_sid] + g["path_id"]
    r_point, r_speed, r_pos = g["point_id"].astype(np.int64), g["speed"], g["pos"]
    r_nc = m.n_center[r_path]
    assert (r_point >= 3).all() and (r_point < r_nc // 2).all()
    u_of = lambda pt, nc: (pt - 3 + 0.5) / (nc // 2 - 3)  # noqa: E731   point index as a fraction of its range

    def chi2_two_sample(a, b, n_bins):
        ca, cb = np.bincount(a, minlength=n_bins).astype(float), np.bincount(b, minlength=n_bins).astype(float)
        keep = (ca + cb) > 0
        return stats.chi2_contingency(np.stack([ca[keep], cb[keep]]))[1]

    # Every test's p-value is collected and the FAMILY is judged (Bonferroni): min p > ALPHA / number of tests.  (Seeds are
    # fixed, so one of ~60 uniformity tests landing at p = 2e-4 — seed 5, on-ramp, 48-point paths; the same counter-based
    # draws replayed in python give exactly that value, seeds 6 / 7 give 0.69 / 0.62 — is the multiple-testing effect.)
    pv = {}
    # (1) closed-form parts of the law
    if len(m.set_names) > 1 and sum(x > 0 for x in probs) > 1:
        want = np.asarray(probs) / sum(probs)
        nz = want > 0                                      # a set with probability 0 is never drawn, on either side
        n_sid, n_rsid = np.bincount(sid, minlength=len(want)), np.bincount(r_sid[:, 0], minlength=len(want))
        assert (n_sid[~nz] == 0).all() and (n_rsid[~nz] == 0).all()
        pv["path-set frequencies"] = stats.chisquare(n_sid[nz], want[nz] * S)[1]
        pv["path-set frequencies (reference)"] = stats.chisquare(n_rsid[nz], want[nz] * len(r_sid))[1]
    for k in range(len(m.set_names)):                      # first agent: always feasible -> uniform path, uniform point
        sel = sid == k
        if sel.sum() < 200:
            continue
        pv[f"set {k}, first agent: uniform path"] = stats.chisquare(np.bincount(path[sel, 0] - lo[k], minlength=n_set[k]))[1]
        pt, nc = point[sel, 0], n_c[sel, 0]
        for n in np.unique(nc):                                # uniform point in [3, n/2), per path length
            q = pt[nc == n] - 3
            if len(q) >= 500:
                pv[f"set {k}, first agent: uniform point on {n}-point paths"] = stats.chisquare(np.bincount(q, minlength=n // 2 - 3))[1]
    pv["speed ~ U(0, v_max)"] = stats.kstest(speed.ravel() / 1.0, "uniform")[1]
    d_min = np.sqrt(((pos[:, :, None] - pos[:, None]) ** 2).sum(-1) + np.eye(N) * 1e6).min()
    assert d_min >= 0.3669 and np.sqrt(((r_pos[:, :, None] - r_pos[:, None]) ** 2).sum(-1) + np.eye(N) * 1e6).min() >= 0.3669
    # (2) against the reference's own draws, agent by agent (later agents carry the rejection's conditioning)
    for a in range(N):
        pv[f"agent {a}: path marginal"] = chi2_two_sample(path[:, a], r_path[:, a], m.n_paths)
        pv[f"agent {a}: point"] = stats.ks_2samp(u_of(point[:, a], n_c[:, a]), u_of(r_point[:, a], r_nc[:, a]))[1]
        pv[f"agent {a}: speed"] = stats.ks_2samp(speed[:, a], r_speed[:, a])[1]
        pv[f"agent {a}: point index"] = chi2_two_sample(point[:, a], r_point[:, a], int(max(point.max(), r_point.max())) + 1)
    for a in range(1, N):                                  # conditional acceptance: distance to the agents placed before
        d_gpu = np.sqrt(((pos[:, a, None] - pos[:, :a]) ** 2).sum(-1)).min(-1)
        d_ref = np.sqrt(((r_pos[:, a, None] - r_pos[:, :a]) ** 2).sum(-1)).min(-1)
        pv[f"agent {a}: distance to the nearest agent placed before it"] = stats.ks_2samp(d_gpu, d_ref)[1]
    worst = min(pv, key=pv.get)
    print(f"RESET-LAW {fixture[:-4]}: {len(pv)} tests, min p = {pv[worst]:.3g} ({worst}), family bound {ALPHA / len(pv):.2g}")
    assert np.isfinite(list(pv.values())).all(), pv
    assert pv[worst] > ALPHA / len(pv), (worst, pv[worst])


def test_respawns_keep_the_path_set_of_their_env():
    """cpm_mixed with several weighted sets: a single-agent respawn draws from the set its env was given at the last
    full reset (world_state_rt_sim.py:327-330), a full reset draws a new set."""
    from sigmarl_b200 import EnvConfig, RoadTrafficEnv
    B, N = 4096, 2
    env = RoadTrafficEnv(EnvConfig(scenario_type="cpm_mixed", n_agents=N, cpm_scenario_probabilities=(0.2, 0.5, 0.3)),
                         num_envs=B, device="cuda:0", seed=9)
    env.reset()
    sid0 = env.scenario_id.clone()
    assert len(torch.unique(sid0)) == 3
    for _ in range(5):
        before = env.pose.clone()
        env.reset_masked(agent_mask=torch.ones(B, N, dtype=torch.bool))      # respawn every agent, no env reset
        torch.cuda.synchronize()
        assert torch.equal(env.scenario_id, sid0)
        assert np.array_equal(env.map.set_of_path(env.path_id.cpu().numpy()), np.repeat(sid0.cpu().numpy()[:, None], N, 1))
        assert float((env.pose != before).any(-1).float().mean()) > 0.8
    env.reset_masked(env_mask=torch.ones(B, dtype=torch.bool))
    assert not torch.equal(env.scenario_id, sid0)                             # new sets after a full reset
    # one step + device reset: the done envs draw new sets, the others keep theirs
    g = torch.Generator(device="cuda").manual_seed(0)
    env.step((torch.rand(B, N, 2, device="cuda", generator=g) * 2 - 1) * torch.as_tensor(UR).cuda())
    sid1, done = env.scenario_id.clone(), env.done.bool().clone()
    env.reset_done()
    assert torch.equal(env.scenario_id[~done], sid1[~done]) and bool((env.scenario_id[done] != sid1[done]).any())
    assert np.array_equal(env.map.set_of_path(env.path_id.cpu().numpy()), np.repeat(env.scenario_id.cpu().numpy()[:, None], N, 1))


def test_nan_flag_word_reports_non_finite_state():
    """Optional health word (sgb_buffers.nan_flags): the reference asserts that positions hold no NaN / inf
    (road_traffic.py:1245-1246); the step kernel sets bit 0 when a step produces a non-finite pose / reward."""
    from sigmarl_b200 import EnvConfig, RoadTrafficEnv
    env = RoadTrafficEnv(EnvConfig(scenario_type="cpm_entire", n_agents=8), num_envs=256, device="cuda:0", seed=1)
    env.reset()
    act = torch.zeros(256, 8, 2, device="cuda")
    act[..., 0] = 0.5
    env.step(act)
    assert int(env.nan_flags) == 0
    env.pose[17, 3, 0] = float("nan")
    env.step(act)
    assert int(env.nan_flags) & 1
    env.nan_flags.zero_()
    env.reset()
    env.step(act)
    assert int(env.nan_flags) == 0


def test_library_refuses_bad_arguments():
    import ctypes as C
    from sigmarl_b200 import lib
    L = lib.load_library()
    assert L.sgb_step(None, 1, 1, None, None) == -1
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "sigmarl_b200.h")).read()
    assert f"#define SGB_VERSION {L.sgb_version()}\n" in hdr
    # an observation-layout bit the library does not know is refused at context creation (SGB_ERR_UNSUPPORTED)
    from sigmarl_b200 import EnvConfig
    from sigmarl_b200.maps import MapLibrary
    m = MapLibrary("intersection_1")
    cfg = EnvConfig(scenario_type="intersection_1", n_agents=2).lower(m)
    cfg.obs_flags = 1 << 9
    ctx = C.c_void_p()
    desc = m.desc()
    assert L.sgb_create(C.byref(ctx), 0, C.byref(desc), C.byref(cfg)) == -5 and not ctx.value


@pytest.mark.parametrize("name", ["c1_intersection_B4_N2", "cpm_entire_B8_N8_distance", "cpm_mixed_B8_N4_gentle"])
def test_cuda_free_running_matches_reference(name):
    """BASELINE configs[0] literally: start from the reference's initial state and run FREE (state is never
    re-synced; only the reference's own reset / respawn draws are replayed through sgb_place), 1e-5 abs on
    state / obs / reward and exact masks over the whole trajectory — i.e. fp drift stays inside the tolerance."""
    path = os.path.join(os.path.dirname(__file__), "golden", name + ".npz")
    g = np.load(path)
    env = _env_from_golden(g)
    T, B, N = int(g["cfg_T"]), env.B, env.N
    gp0 = env.map.global_path(g["pre_scenario_id"][0], g["pre_path_id"][0])
    env.set_state(g["pre_pos"][0], g["pre_rot"][0], g["pre_speed"][0], g["pre_steering"][0], gp0, step_count=g["pre_step"][0])
    worst = 0.0
    for t in range(T):
        ctx = f"{name} free-run t={t}"
        obs, rew, done = env.step(torch.as_tensor(g["action"][t]).cuda())
        torch.cuda.synchronize()
        for nm, got, want in [("pos", env.pos, g["post_pos"][t]), ("rot", env.rot, g["post_rot"][t]),
                              ("speed", env.speed, g["post_speed"][t]), ("steering", env.steering, g["post_steering"][t]),
                              ("vel", env.vel, g["post_vel"][t]), ("obs", obs, g["obs"][t]), ("reward", rew, g["reward"][t])]:
            worst = max(worst, _close(nm, got.cpu(), want, ctx))
        fl = env.agent_flags.cpu().numpy()
        assert np.array_equal(done.cpu().numpy().astype(bool), g["done"][t]), f"{ctx} done"
        assert np.array_equal((fl & 2) != 0, g["col_lane"][t]) and np.array_equal(_coll_matrix(env), g["col_agents"][t]), ctx
        assert np.array_equal((fl & 8) != 0, g["col_exit"][t]) and np.array_equal((fl & 4) != 0, g["col_entry"][t]), ctx
        # replay the reference's respawn (inside done()) and env-reset draws
        resp, rst = g["respawn_mask"][t], g["reset_mask"][t]
        if resp.any():
            gp = env.map.global_path(g["pre_scenario_id"][t], g["respawn_path_id"][t])  # a respawn keeps the scenario
            env.place(gp, g["respawn_point_id"][t], g["respawn_speed"][t], agent_mask=resp)
            env.refresh(env_mask=torch.as_tensor(resp.any(1)))
        if rst.any():
            gp = env.map.global_path(g["reset_scenario_id"][t], g["reset_path_id"][t])
            env.place(gp, g["reset_point_id"][t], g["reset_speed"][t], agent_mask=np.repeat(rst[:, None], N, 1))
            env.refresh(env_mask=torch.as_tensor(rst))
            env.step_count[torch.as_tensor(rst).cuda()] = 0
    assert worst <= TOL


@pytest.mark.parametrize("scenario,N,B", [("on_ramp_2_multilane", 12, 8192), ("roundabout_2", 12, 8192), ("cpm_mixed", 4, 16384)])
def test_config4_maps_pruned_equals_exhaustive_at_scale(scenario, N, B):
    """BASELINE configs[3] per-GPU shape (on-ramp / roundabout, 8192 envs x 12 agents per GPU; G = 2 lanes per
    agent) plus cpm_mixed (2 envs per warp): pruned == exhaustive bit for bit, invariants hold, respawns happen."""
    from sigmarl_b200 import EnvConfig, RoadTrafficEnv
    envs = [RoadTrafficEnv(EnvConfig(scenario_type=scenario, n_agents=N, rew_method="ttc_sparse", mode="params",
                                     threshold_near_other_agents_c2c_low=0.1635, exhaustive=ex),
                           num_envs=B, device="cuda:0", seed=21) for ex in (False, True)]
    for e in envs:
        e.reset()
    # sequential rejection sampling without backtracking can dead-end at N = 12 (the reference would spin forever,
    # SURVEY.md §4); the bounded device reset reports those instead — they must stay rare
    assert int(envs[0].n_failed.item()) <= 0.002 * B * N
    g = torch.Generator(device="cuda").manual_seed(5)
    n_exit = 0
    for t in range(60):
        # pure pursuit on the 2nd short-term reference point of the ego-frame observation (obs[3:5]) so that agents
        # actually travel to their path ends; both envs hold identical observations, so the actions are identical
        o = envs[0].obs
        steer = torch.clamp(1.5 * torch.atan2(o[..., 4], o[..., 3]) + (torch.rand(B, N, device="cuda", generator=g) - 0.5) * 0.06,
                            -float(UR[1]), float(UR[1]))
        act = torch.stack([0.6 + 0.4 * torch.rand(B, N, device="cuda", generator=g), steer], -1)
        for e in envs:
            e.step(act)
        for name in ("pose", "aux", "carry", "obs", "reward", "done", "agent_flags", "collide_with"):
            assert torch.equal(getattr(envs[0], name), getattr(envs[1], name)), f"{scenario} t={t} {name}"
        n_exit += int(((envs[0].agent_flags & 8) != 0).sum())
        assert torch.isfinite(envs[0].obs).all()
        for e in envs:
            e.reset_done(write_obs=True)
        assert torch.equal(envs[0].pose, envs[1].pose) and torch.equal(envs[0].obs, envs[1].obs)
    if scenario == "cpm_mixed":   # short paths: agents must reach their path ends, i.e. the respawn branch ran
        assert n_exit > 0, "no agent ever reached its path end: the respawn branch was not exercised"


def test_facade_done_respawns_exit_crossers_like_reference_flow():
    """VMAS order through the facade: done() must respawn entry/exit crossers of not-done envs (road_traffic.py:1462-1472)
    and leave done envs to the caller's reset_at."""
    from sigmarl_b200 import make_env
    env = make_env(scenario_type="cpm_mixed", num_envs=512, device="cuda:0", n_agents=4, seed=9, max_steps=128, dt=0.1)
    sc = env.scenario
    g = torch.Generator(device="cuda").manual_seed(2)
    respawned = 0
    for t in range(80):
        o = sc.env.obs   # pure pursuit on the 2nd short-term point so that agents reach their path ends
        steer = torch.clamp(1.5 * torch.atan2(o[..., 4], o[..., 3]), -float(UR[1]), float(UR[1]))
        acts = [torch.stack([0.6 + 0.3 * torch.rand(512, device="cuda", generator=g), steer[:, i]], -1) for i in range(4)]
        for a, agent in zip(acts, env.agents):
            agent.action.u = a
        sc.world.step()
        flags = sc.env.agent_flags.clone()
        pose_before = sc.env.pose.clone()
        infos = [sc.info(a) for a in env.agents]
        dones = sc.done()
        crossing = ((flags & 12) != 0) & ~dones[:, None]
        moved = (sc.env.pose[..., :2] != pose_before[..., :2]).any(-1)
        assert torch.equal(moved, crossing), "exactly the entry/exit crossers of not-done envs are re-placed"
        assert torch.equal(torch.stack([i["is_reach_goal"] for i in infos], 1), (flags & 8) != 0)
        respawned += int(crossing.sum())
        for e in torch.where(dones)[0].tolist()[:8]:
            env.reset_at(e)
        sc.env.reset_done(write_obs=True)  # remaining done envs in one launch
    assert respawned > 0


class _Params:
    """Attribute bag standing in for the reference's ``Parameters`` object (helper_common.py:26-252)."""


def _facade_from_golden(g):
    from sigmarl_b200.scenario import ScenarioRoadTrafficB200
    sc = ScenarioRoadTrafficB200()
    kw = dict(
        scenario_type=str(g["cfg_scenario_type"]), n_agents=int(g["cfg_N"]), dt=float(g["cfg_dt"]),
        max_steps=int(g["cfg_max_steps"]), rew_method=str(g["cfg_rew_method"]),
        n_nearing_agents_observed=int(g["cfg_n_nearing_agents_observed"]),
        reward_progress=float(g["cfg_reward_progress"]),
        threshold_near_boundary_high=float(g["cfg_near_boundary_high"]),
        threshold_near_boundary_low=float(g["cfg_near_boundary_low"]),
        threshold_near_other_agents_c2c_high=float(g["cfg_near_other_agents_high"]),
        threshold_near_other_agents_c2c_low=float(g["cfg_near_other_agents_low"]),
        ttc_low=float(g["cfg_ttc_low"]), ttc_high=float(g["cfg_ttc_high"]),
        penalty_near_boundary=float(g["cfg_penalty_near_boundary"]),
        penalty_near_other_agents=float(g["cfg_penalty_near_other_agents"]),
        is_testing_mode=bool(g["cfg_is_testing_mode"]), is_obs_noise=False,    # the goldens were recorded without noise
        **_obs_flags_of_golden(g))
    for name in ("reset_agent_fixed_duration", "is_use_mtv_distance"):       # fixtures of tests/golden/next/
        if ("cfg_" + name) in g.files:
            kw[name] = g["cfg_" + name].item()
    if str(g["cfg_mode"]) == "params":     # mappo_cavs.py:168-169: scenario.parameters = parameters; make_world(...)
        p = _Params()
        for k, v in kw.items():
            setattr(p, k, v)
        sc.parameters = p
        world = sc.env_make_world(int(g["cfg_B"]), "cuda:0")
    else:
        world = sc.env_make_world(int(g["cfg_B"]), "cuda:0", **kw)
    return sc, world


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: os.path.basename(p)[:-4])
def test_facade_info_matches_reference_goldens(path):
    """ScenarioRoadTrafficB200.info(agent): every key of the reference's info dict (road_traffic.py:1547-1633),
    same shapes / dtypes, values within 1e-5 (bit-exact for masks and ids), teacher-forced from the goldens."""
    g = np.load(path)
    sc, world = _facade_from_golden(g)
    env = sc.env
    keys = sorted(k[5:] for k in g.files if k.startswith("info_"))
    assert len(keys) == 39
    for t in range(0, int(g["cfg_T"]), 3):
        ctx = f"{os.path.basename(path)} t={t}"
        gp = env.map.global_path(g["pre_scenario_id"][t], g["pre_path_id"][t])
        env.set_state(g["pre_pos"][t], g["pre_rot"][t], g["pre_speed"][t], g["pre_steering"][t], gp,
                      step_count=g["pre_step"][t])
        for i, agent in enumerate(world.agents):
            agent.action.u = torch.as_tensor(g["action"][t][:, i]).cuda()
        world.step()
        infos = [sc.info(agent) for agent in world.agents]
        assert sorted(infos[0].keys()) == keys, f"{ctx}: info keys differ from the reference's"
        # state entries of the fixtures are views the reference mutates when it respawns an agent inside done()
        keep = ~g["respawn_mask"][t]
        for k in keys:
            want = g["info_" + k][t]
            got = torch.stack([torch.as_tensor(d[k]).reshape(env.B, -1).squeeze(-1) for d in infos], dim=1).cpu().numpy()
            assert got.shape == want.shape, f"{ctx} info[{k}] shape {got.shape} vs {want.shape}"
            if want.dtype == np.bool_ or np.issubdtype(want.dtype, np.integer):
                assert got.dtype == want.dtype, f"{ctx} info[{k}] dtype {got.dtype} vs {want.dtype}"
                assert np.array_equal(got[keep], want[keep]), f"{ctx} info[{k}] not bit-exact"
            else:
                _close(f"info[{k}]", got[keep], want[keep], ctx)
        sc.done()


def test_facade_testing_mode_respawns_colliding_agents_and_never_ends_early():
    """is_testing_mode (road_traffic.py:1429-1447): an env is done only at the time limit; colliding / leaving
    agents of not-done envs are re-placed one by one inside done(); reward = progress + sparse terms (:1050-1055)."""
    from sigmarl_b200 import make_env
    B, N = 256, 6
    env = make_env(scenario_type="cpm_entire", num_envs=B, device="cuda:0", n_agents=N, seed=3, max_steps=24, dt=0.1,
                   is_testing_mode=True)
    sc = env.scenario
    gen = torch.Generator(device="cuda").manual_seed(5)
    ur = torch.as_tensor(UR).cuda()
    n_resp = 0
    for t in range(30):
        for agent in env.agents:
            agent.action.u = (torch.rand(B, 2, device="cuda", generator=gen) * 2 - 1) * ur
        sc.world.step()
        flags, pose_before = sc.env.agent_flags.clone(), sc.env.pose.clone()
        step = sc.env.step_count.clone()
        rew = sc.env.reward.clone()
        dones = sc.done()
        assert torch.equal(dones, step == 24 - 1), "testing mode: only the time limit ends an env"
        hit = ((flags & 15) != 0) & ~dones[:, None]
        moved = (sc.env.pose[..., :2] != pose_before[..., :2]).any(-1)
        assert torch.equal(moved, hit)
        # colliding agents carry the full collision penalty (reward clamps at -1)
        assert bool((rew[(flags & 3) != 0] <= -0.8).all())
        n_resp += int(hit.sum())
        sc.env.reset_done(write_obs=True)
    assert n_resp > 0 and int(sc.env.n_failed) == 0


def test_reset_world_at_single_agent_on_any_map_and_mode():
    """reset_world_at(env_index, agent_index) from outside the step (road_traffic.py:816-923) re-places exactly that
    agent — also on cpm_entire in training mode, where done() itself never respawns anybody — at a feasible point."""
    from sigmarl_b200 import make_env
    env = make_env(scenario_type="cpm_entire", num_envs=64, device="cuda:0", n_agents=8, seed=2, max_steps=128, dt=0.1)
    sc = env.scenario
    for (b, a) in [(3, 2), (0, 0), (63, 7)]:
        before, obs_before = sc.env.pose.clone(), sc.env.obs.clone()
        sc.reset_world_at(env_index=b, agent_index=a)
        moved = (sc.env.pose != before).any(-1)
        want = torch.zeros_like(moved)
        want[b, a] = True
        assert torch.equal(moved, want)
        assert torch.equal(sc.env.obs, obs_before), "a respawn keeps the step-time observation"
        d = (sc.env.pose[b, :, :2] - sc.env.pose[b, a, :2]).norm(dim=-1)
        d[a] = 1e9
        assert float(d.min()) >= 0.3669
        carried = sc.env.carry.clone()
        sc.env.refresh()
        assert torch.equal(carried, sc.env.carry), "carry of the respawned agent comes from the spawn table"
    before = sc.env.pose.clone()
    sc.reset_world_at(env_index=5)
    changed = (sc.env.pose != before).any(-1).any(-1)
    assert bool(changed[5]) and int(changed.sum()) == 1 and int(sc.env.step_count[5]) == 0
    assert int(sc.env.n_failed) == 0


def test_programmatic_dependent_launch_changes_nothing_but_overlap(monkeypatch):
    """The library chains its kernels with programmatic dependent launch (griddepcontrol) and keeps the
    reset's list length in two alternating counters instead of a memset between the kernels.  Same seeds with the chain
    switched off (SGB_NO_PDL=1, read at context creation): every buffer stays bit-identical through a mixed sequence of
    steps, masked resets with and without fresh observations, explicit respawns and masked refreshes."""
    from sigmarl_b200 import EnvConfig, RoadTrafficEnv
    B, N = 4096, 8
    envs = []
    for no_pdl in ("0", "1"):
        monkeypatch.setenv("SGB_NO_PDL", no_pdl)
        envs.append(RoadTrafficEnv(EnvConfig(scenario_type="cpm_entire", n_agents=N), num_envs=B, device="cuda:0", seed=11))
    monkeypatch.delenv("SGB_NO_PDL")
    gen = torch.Generator(device="cuda").manual_seed(5)
    for e in envs:
        e.reset()
    names = ("pose", "aux", "carry", "path_id", "agent_flags", "collide_with", "step_count", "obs", "reward", "done", "n_failed")
    for t in range(12):
        act = (torch.rand(B, N, 2, generator=gen, device="cuda") * 2 - 1) * torch.as_tensor(UR).cuda()
        mask = torch.rand(B, generator=gen, device="cuda") < 0.1
        amask = torch.rand(B, N, generator=gen, device="cuda") < 0.05
        for e in envs:
            e.step(act.clone())
            e.reset_done(write_obs=(t % 2 == 0))          # back-to-back kernels: step -> reset (-> refresh)
            if t % 3 == 0:
                e.refresh(env_mask=mask, write_obs=True)   # its own counter slot
            if t % 4 == 1:
                e.reset_masked(agent_mask=amask, write_obs=True)
            e.reset_done(write_obs=True)                   # a second reset right behind the first: alternating counters
        torch.cuda.synchronize()
        for name in names:
            assert torch.equal(getattr(envs[0], name), getattr(envs[1], name)), (t, name)
    assert int(envs[0].done.sum()) >= 0 and int(envs[0].n_failed) == 0


def test_sub_warp_reset_equals_warp_per_env_reset():
    """Large batches of up to 8 agents are reset four envs per warp (sub-warps of 8 lanes, reset_kernel<8>), small ones one
    env per warp (reset_kernel<32>).  The draws are keyed by (seed, epoch, global env, agent, try), not by lanes: a batch of
    20 480 envs (sub-warps) and the same envs as four shards of 5 120 (warp per env) must stay bit-identical through full
    resets, masked resets of done envs and explicit respawns — on a roomy map and on a crowded one (many retries)."""
    from sigmarl_b200 import EnvConfig, RoadTrafficEnv
    for scenario, N in (("cpm_entire", 8), ("on_ramp_2_multilane", 8)):
        B, S = 20480, 4
        cfg = EnvConfig(scenario_type=scenario, n_agents=N)
        full = RoadTrafficEnv(cfg, num_envs=B, device="cuda:0", seed=17, max_reset_tries=512)
        shards = [RoadTrafficEnv(cfg, num_envs=B // S, device="cuda:0", seed=17, env_offset=k * (B // S), max_reset_tries=512)
                  for k in range(S)]
        full.reset()
        for h in shards:
            h.reset()
        gen = torch.Generator(device="cuda").manual_seed(2)
        names = ("pose", "aux", "carry", "path_id", "agent_flags", "step_count", "obs")

        def same(tag):
            torch.cuda.synchronize()
            for name in names:
                cat = torch.cat([getattr(h, name) for h in shards])
                assert torch.equal(getattr(full, name), cat), (scenario, tag, name)
            assert int(full.n_failed) == sum(int(h.n_failed) for h in shards)

        same("reset")
        for t in range(6):
            act = (torch.rand(B, N, 2, generator=gen, device="cuda") * 2 - 1) * torch.as_tensor(UR).cuda()
            amask = torch.rand(B, N, generator=gen, device="cuda") < 0.1
            full.step(act)
            full.reset_done(write_obs=True)
            for k, h in enumerate(shards):
                sl = slice(k * (B // S), (k + 1) * (B // S))
                h.step(act[sl])
                h.reset_done(write_obs=True)
            same(f"step {t}")
            if t % 2 == 1:
                full.reset_masked(agent_mask=amask, write_obs=True)
                for k, h in enumerate(shards):
                    h.reset_masked(agent_mask=amask[k * (B // S):(k + 1) * (B // S)], write_obs=True)
                same(f"respawn {t}")
