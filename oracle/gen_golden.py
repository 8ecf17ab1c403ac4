#!/usr/bin/env python
"""TEST INFRASTRUCTURE — generate golden vectors from the UNMODIFIED reference.

Imports the reference's own hot-path modules from ``/root/reference`` behind the import
shim (``oracle/refshim/install_shims.py``), drives
``ScenarioRoadTraffic`` through the VMAS ``Environment.step`` call order with seeded
random actions and records, for every step, the pre-step state, the action and everything
the step produced.  The ``.npz`` fixtures land in ``tests/golden/`` and are committed,
because ``/root/reference`` cannot travel to the GPU box.

Re-run (here only):  ``python oracle/gen_golden.py``

Recorded per step t (arrays are [T, B, N, ...]):
  pre_*   : pos, rot, speed, steering, path_id, scenario_id, step   (state the step starts from,
            i.e. after the previous step's respawns / env resets)
  action  : [T,B,N,2]
  post_*  : pos, rot, speed, steering, vel, sideslip                 (helper_training.py:856-861)
  obs [T,B,N,D], reward [T,B,N], done [T,B]                          (road_traffic.py:925,1334,1368)
  col_agents [T,B,N,N], col_lane/col_entry/col_exit [T,B,N]          (world_state_rt_sim.py:379-424)
  d_ref, d_left[...,5], d_right[...,5], d_bound, d_agents[T,B,N,N], idx_ref,
  short_term [T,B,N,3,2], vertices [T,B,N,5,2]                       (world_state_rt.py:582-684)
  reset_mask [T,B]  : env was reset after step t (reset_at), followed by
  reset_obs [T,B,N,D], reset_pos/rot/speed/vel/path_id/point_id/scenario_id  (post-reset state of those envs)
  respawn_mask [T,B,N] : agent was respawned inside done() (road_traffic.py:1462-1472)
  info_<key> [T,B,N,...] : every entry of info(agent) (road_traffic.py:1547-1633), num_task_tries / task_success_times [T,B]
"""
import os
import sys

os.environ["CICD_TESTING"] = "true"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "refshim"))
import install_shims  # noqa: E402,F401

import numpy as np  # noqa: E402
import torch  # noqa: E402
from vmas.simulator.environment import Environment  # noqa: E402
from sigmarl.helper_common import Parameters  # noqa: E402
from sigmarl.scenarios.road_traffic import ScenarioRoadTraffic  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

# info() keys recorded in the fixtures; base/priority observations exist only with prioritized MARL
INFO_KEYS = ["pos", "pos_nom", "rot", "rot_nom", "vel", "vel_nom", "act_vel", "act_vel_nom", "act_steer",
             "act_steer_nom", "ref", "ref_nom", "distance_ref", "distance_ref_nom", "distance_left_b",
             "distance_left_b_nom", "distance_right_b", "distance_right_b_nom", "is_collision_with_agents",
             "is_collision_with_lanelets", "is_reach_goal", "ref_lanelet_ids", "path_id", "applied_action_vel",
             "applied_action_steer", "nominal_action_vel", "nominal_action_steer",
             "rew_progress", "rew_reach_goal", "rew_speed", "rew_centerline", "rew_near_other_agents",
             "rew_near_left_lane", "rew_near_right_lane", "rew_collide_other_agents", "rew_collide_lane",
             "rew_energy_acceleration", "rew_energy_steering", "rew_total"]

# name -> config.  mode "kwargs": scenario built from make_world(**kwargs) (dt=0.05 ...);
# mode "params": a Parameters object from sigmarl/config.json (dt=0.1 ...), as mappo_cavs.py:168 does.
CONFIGS = {
    # BASELINE.json configs[0]
    "c1_intersection_B4_N2": dict(st="intersection_1", B=4, N=2, T=100, mode="kwargs", seed=0),
    "cpm_entire_B8_N8_distance": dict(st="cpm_entire", B=8, N=8, T=50, mode="params", seed=1),
    "cpm_entire_B8_N8_sparse_kw": dict(st="cpm_entire", B=8, N=8, T=40, mode="kwargs", seed=2,
                                      extra=dict(rew_method="sparse")),
    "cpm_mixed_B8_N6_ttc_sparse": dict(st="cpm_mixed", B=8, N=6, T=60, mode="params", seed=3,
                                       extra=dict(rew_method="ttc_sparse",
                                                  threshold_near_other_agents_c2c_low=0.1635)),
    # cpm_mixed with the merge-in / merge-out path sets (cpm_scenario_probabilities != [1,0,0]; one scenario id per env,
    # world_state_rt_sim.py:313-358).  Those sets hold 4 short paths each (31 spawn points): at the reset distance of
    # 0.367 m at most 3 (merge-in) / 2 (merge-out) agents fit, so the reference's unbounded rejection sampling
    # (:232-311) only terminates reliably for N <= 2 (probed: N = 4 spins forever; N = 1 crashes in
    # observation_provider_rt.py:790).  N = 2 works and is recorded below.
    "sets_cpm_mixed_B4_N2_merge_in_gentle": dict(st="cpm_mixed", B=4, N=2, T=60, mode="params", seed=71, gentle=True,
                                                extra=dict(cpm_scenario_probabilities=[0.0, 1.0, 0.0])),
    "sets_cpm_mixed_B8_N2_all_sets": dict(st="cpm_mixed", B=8, N=2, T=60, mode="params", seed=72,
                                         extra=dict(cpm_scenario_probabilities=[0.3, 0.3, 0.4], rew_method="ttc_sparse")),
    "on_ramp_2_B8_N12": dict(st="on_ramp_2_multilane", B=8, N=12, T=40, mode="kwargs", seed=5),
    "roundabout_2_B8_N12_ttc": dict(st="roundabout_2", B=8, N=12, T=40, mode="params", seed=6,
                                   extra=dict(rew_method="ttc")),
    # forward-driving actions: long episodes, agents reach path ends -> exit crossings + respawns
    "cpm_mixed_B8_N4_gentle": dict(st="cpm_mixed", B=8, N=4, T=80, mode="params", seed=8, gentle=True,
                                  extra=dict(rew_method="distance_sparse")),
    # testing mode (road_traffic.py:1050-1055, 1429-1447): sparse reward, colliding agents respawned one by one
    "cpm_entire_B4_N4_testing": dict(st="cpm_entire", B=4, N=4, T=60, mode="params", seed=11,
                                    extra=dict(is_testing_mode=True), max_steps=40),
    "cpm_mixed_B4_N4_testing_gentle": dict(st="cpm_mixed", B=4, N=4, T=80, mode="params", seed=12, gentle=True,
                                          extra=dict(is_testing_mode=True, rew_method="ttc_sparse"), max_steps=48),
    "cpm_entire_B4_N3_k1": dict(st="cpm_entire", B=4, N=3, T=40, mode="params", seed=7,
                               extra=dict(n_nearing_agents_observed=1)),
    # non-default observation layouts (observation_provider_rt.py:594-925)
    "obsvar_cpm_entire_B4_N5_centres": dict(st="cpm_entire", B=4, N=5, T=40, mode="params", seed=21,
                                           extra=dict(is_observe_vertices=False, is_obs_steering=True,
                                                      is_observe_ref_path_other_agents=True,
                                                      is_observe_distance_to_agents=False,
                                                      is_observe_distance_to_center_line=False)),
    "obsvar_cpm_mixed_B4_N4_birdview_gentle": dict(st="cpm_mixed", B=4, N=4, T=60, mode="params", seed=22, gentle=True,
                                                  extra=dict(is_ego_view=False)),
    "obsvar_intersection_B4_N3_birdview_all": dict(st="intersection_1", B=4, N=3, T=40, mode="kwargs", seed=23,
                                                  extra=dict(is_ego_view=False, is_observe_vertices=False,
                                                             is_obs_steering=True,
                                                             is_observe_ref_path_other_agents=True)),
    "obsvar_roundabout_2_B4_N6_steer_refs": dict(st="roundabout_2", B=4, N=6, T=30, mode="params", seed=24,
                                                extra=dict(is_obs_steering=True, is_observe_ref_path_other_agents=True,
                                                           n_nearing_agents_observed=3)),
    # boundary points instead of boundary distances (world_state_rt.py:689-725): loop paths, open paths driven to
    # their ends (tail padding, respawns), bird view.  The closest boundary INDEX of an agent sitting exactly on a
    # centre point (every reset pose) is a near-tie between two segments that only an exact re-statement of
    # torch.norm's rounding resolves like the reference (orc_norm2); seed 31 holds two such agent-steps.
    "obsvar_cpm_entire_B4_N4_bpoints": dict(st="cpm_entire", B=4, N=4, T=40, mode="params", seed=25,
                                           extra=dict(is_observe_distance_to_boundaries=False)),
    "obsvar_cpm_mixed_B4_N3_bpoints_gentle": dict(st="cpm_mixed", B=4, N=3, T=100, mode="params", seed=31, gentle=True,
                                                 extra=dict(is_observe_distance_to_boundaries=False,
                                                            is_obs_steering=True)),
    "obsvar_on_ramp_2_B4_N6_bpoints_birdview": dict(st="on_ramp_2_multilane", B=4, N=6, T=40, mode="kwargs", seed=27,
                                                   extra=dict(is_observe_distance_to_boundaries=False,
                                                              is_ego_view=False, is_observe_vertices=False)),
    # reset_agent_fixed_duration (road_traffic.py:1388-1393): envs also end every 2 s (training mode) / 1 s of
    # simulated time (testing mode, dt 0.1); gentle driving so that the fixed-duration reset is what ends episodes
    "cpm_entire_B4_N3_fixed2s_gentle": dict(st="cpm_entire", B=4, N=3, T=70, mode="params", seed=41, gentle=True,
                                           extra=dict(reset_agent_fixed_duration=2)),
    "cpm_mixed_B4_N3_fixed1s_testing_gentle": dict(st="cpm_mixed", B=4, N=3, T=50, mode="params", seed=42, gentle=True,
                                                  extra=dict(reset_agent_fixed_duration=1, is_testing_mode=True),
                                                  max_steps=64),
    # MTV agent distance (is_use_mtv_distance; helper_scenario.py:1030-1138, world_state_rt_sim.py:360-396): mutual
    # distances from LAST step's rectangles, thresholds 0 / agent length, agents "collide" only at distance == 0
    "mtv_cpm_entire_B4_N6_distance": dict(st="cpm_entire", B=4, N=6, T=40, mode="params", seed=51,
                                         extra=dict(is_use_mtv_distance=True)),
    "mtv_cpm_mixed_B4_N5_ttc_sparse_gentle": dict(st="cpm_mixed", B=4, N=5, T=70, mode="params", seed=52, gentle=True,
                                                 extra=dict(is_use_mtv_distance=True, rew_method="ttc_sparse")),
    "mtv_roundabout_2_B4_N10_kw_k3": dict(st="roundabout_2", B=4, N=10, T=40, mode="kwargs", seed=53, gentle=True,
                                         extra=dict(is_use_mtv_distance=True, n_nearing_agents_observed=3)),
    # is_apply_mask (observation_provider_rt.py:638-749): observed neighbours at or beyond 5 agent lengths are masked
    "mask_cpm_entire_B4_N8_k5": dict(st="cpm_entire", B=4, N=8, T=40, mode="params", seed=81,
                                    extra=dict(is_apply_mask=True, n_nearing_agents_observed=5)),
    "mask_roundabout_2_B4_N6_all_k4": dict(st="roundabout_2", B=4, N=6, T=40, mode="kwargs", seed=82, gentle=True,
                                          extra=dict(is_apply_mask=True, n_nearing_agents_observed=4,
                                                     is_observe_vertices=False, is_obs_steering=True,
                                                     is_observe_ref_path_other_agents=True)),
    "mask_cpm_mixed_B4_N5_birdview_k3": dict(st="cpm_mixed", B=4, N=5, T=50, mode="params", seed=83, gentle=True,
                                            extra=dict(is_apply_mask=True, is_ego_view=False,
                                                       n_nearing_agents_observed=3)),
    # more agents than a 4-lane group layout holds: the reference's default n_agents on cpm_entire (15; two lanes per
    # agent on the GPU) and 18 (one lane per agent)
    "big_cpm_entire_B4_N15_ttc_sparse": dict(st="cpm_entire", B=4, N=15, T=30, mode="params", seed=95,
                                            extra=dict(rew_method="ttc_sparse")),
    "big_cpm_entire_B2_N18_distance_k4": dict(st="cpm_entire", B=2, N=18, T=30, mode="kwargs", seed=96, gentle=True,
                                             extra=dict(n_nearing_agents_observed=4)),
    # lanelet-relation mask (map_manager.py:39-119): bird view + is_apply_mask on OSM maps
    "mask_roundabout_2_B4_N8_birdview_lanelets_k4": dict(st="roundabout_2", B=4, N=8, T=40, mode="params", seed=84,
                                                        gentle=True, extra=dict(is_apply_mask=True, is_ego_view=False,
                                                                                n_nearing_agents_observed=4)),
    "mask_intersection_5_B4_N6_birdview_lanelets": dict(st="intersection_5", B=4, N=6, T=40, mode="kwargs", seed=85,
                                                       extra=dict(is_apply_mask=True, is_ego_view=False,
                                                                  is_observe_vertices=False, n_nearing_agents_observed=3)),
    # the remaining maps of constants.py (interchange_1-3, intersection_2-8): open paths, entry / exit respawns
    "map_interchange_1_B4_N4": dict(st="interchange_1", B=4, N=4, T=40, mode="kwargs", seed=61, gentle=True),
    "map_interchange_2_B4_N6": dict(st="interchange_2", B=4, N=6, T=40, mode="params", seed=62, gentle=True,
                                   extra=dict(rew_method="ttc_sparse")),
    "map_interchange_3_B4_N6": dict(st="interchange_3", B=4, N=6, T=30, mode="kwargs", seed=63),
    "map_intersection_2_B4_N3": dict(st="intersection_2", B=4, N=3, T=40, mode="params", seed=64, gentle=True),
    "map_intersection_3_B4_N4": dict(st="intersection_3", B=4, N=4, T=30, mode="kwargs", seed=65),
    "map_intersection_4_B4_N4": dict(st="intersection_4", B=4, N=4, T=40, mode="params", seed=66, gentle=True,
                                    extra=dict(rew_method="distance_sparse")),
    "map_intersection_5_B4_N5": dict(st="intersection_5", B=4, N=5, T=30, mode="kwargs", seed=67),
    "map_intersection_6_B4_N5": dict(st="intersection_6", B=4, N=5, T=40, mode="params", seed=68, gentle=True),
    "map_intersection_7_B4_N4": dict(st="intersection_7", B=4, N=4, T=30, mode="kwargs", seed=69),
    "map_intersection_8_B4_N4": dict(st="intersection_8", B=4, N=4, T=40, mode="params", seed=70, gentle=True,
                                    extra=dict(rew_method="ttc")),
}

# fixtures of features added after the last hardware session: tests/golden/next/ (tests/conftest.py)
NEXT = {"cpm_entire_B4_N3_fixed2s_gentle", "cpm_mixed_B4_N3_fixed1s_testing_gentle", "mtv_cpm_entire_B4_N6_distance",
        "mtv_cpm_mixed_B4_N5_ttc_sparse_gentle", "mtv_roundabout_2_B4_N10_kw_k3"}
NEXT |= {n for n in CONFIGS if n.startswith(("map_", "mask_", "big_", "sets_"))}

OBS_FLAGS = ["is_ego_view", "is_observe_vertices", "is_obs_steering", "is_observe_ref_path_other_agents",
             "is_observe_distance_to_agents", "is_observe_distance_to_center_line",
             "is_observe_distance_to_boundaries", "is_partial_observation", "is_apply_mask", "is_obs_noise"]


def stack_agents(agents, get):
    return torch.stack([get(a) for a in agents], dim=1)


def run(name, st, B, N, T, mode, seed, extra=None, max_steps=128, gentle=False):
    extra = dict(extra or {})
    torch.manual_seed(seed)
    sc = ScenarioRoadTraffic()
    kw = {}
    if mode == "params":
        p = Parameters.from_json("/root/reference/sigmarl/config.json")
        p.scenario_type = st
        p.n_agents = N
        p.num_vmas_envs = B
        p.max_steps = max_steps
        for k, v in extra.items():
            assert hasattr(p, k), k
            setattr(p, k, v)
        sc.parameters = p
    else:
        kw = dict(scenario_type=st, n_agents=N, is_obs_noise=False)
        rew_method = extra.pop("rew_method", None)
        kw.update(extra)
    env = Environment(sc, num_envs=B, device="cpu", max_steps=max_steps, **kw)
    if mode == "kwargs":
        sc.parameters.max_steps = max_steps
        if rew_method is not None:
            sc.parameters.rew_method = rew_method
    ws = sc.world_state
    agents = env.agents
    ur = torch.tensor([float(sc.max_speed), float(sc.max_steering)])
    rec = {}

    def push(k, v):
        rec.setdefault(k, []).append(np.asarray(v.detach().cpu().numpy() if torch.is_tensor(v) else v).copy())

    def snap_state(prefix):
        push(prefix + "pos", stack_agents(agents, lambda a: a.state.pos))
        push(prefix + "rot", stack_agents(agents, lambda a: a.state.rot.squeeze(-1)))
        push(prefix + "speed", stack_agents(agents, lambda a: a.state.speed.squeeze(-1)))
        push(prefix + "steering", stack_agents(agents, lambda a: a.state.steering.squeeze(-1)))
        push(prefix + "path_id", ws.ref_paths_agent_related.path_id.clone())
        push(prefix + "scenario_id", ws.ref_paths_agent_related.scenario_id.clone())

    last_obs = [sc.observation(a).clone() for a in agents]
    for t in range(T):
        snap_state("pre_")
        push("pre_step", sc.timer.step.clone())
        if gentle:
            # pure-pursuit on the 2nd short-term reference point seen in the ego frame (obs[3:5])
            acts = []
            for i in range(N):
                o = last_obs[i]
                if not sc.parameters.is_ego_view or sc.parameters.is_obs_steering:
                    # other layouts (global coordinates / shifted columns): take the same point from the world state
                    d = ws.ref_paths_agent_related.short_term[:, i, 1] - agents[i].state.pos
                    r = agents[i].state.rot[:, 0]
                    o = torch.zeros(B, 5)
                    o[:, 3] = d[:, 0] * torch.cos(r) + d[:, 1] * torch.sin(r)
                    o[:, 4] = d[:, 1] * torch.cos(r) - d[:, 0] * torch.sin(r)
                steer = torch.clamp(1.5 * torch.atan2(o[:, 4], o[:, 3]) + (torch.rand(B) * 2 - 1) * 0.03,
                                    -float(sc.max_steering), float(sc.max_steering))
                acts.append(torch.stack([0.5 + 0.3 * torch.rand(B), steer], dim=1))
        else:
            acts = [(torch.rand(B, 2) * 2 - 1) * ur for _ in range(N)]
        push("action", torch.stack(acts, dim=1))
        pid_before = ws.ref_paths_agent_related.point_id.clone()
        path_before = ws.ref_paths_agent_related.path_id.clone()
        # intercept per-agent respawns inside done()
        respawn = torch.zeros(B, N, dtype=torch.bool)
        orig_reset = sc.reset_world_at

        def spy(env_index=None, agent_index=None, _o=orig_reset, _r=respawn):
            if agent_index is not None:
                _r[int(env_index), int(agent_index)] = True
            return _o(env_index=env_index, agent_index=agent_index)

        sc.reset_world_at = spy
        # the post-step / pre-respawn state must be captured before done() runs: wrap done
        orig_done = sc.done
        holder = {}

        def done_spy(_o=orig_done):
            holder["pos"] = stack_agents(agents, lambda a: a.state.pos).clone()
            holder["rot"] = stack_agents(agents, lambda a: a.state.rot.squeeze(-1)).clone()
            holder["speed"] = stack_agents(agents, lambda a: a.state.speed.squeeze(-1)).clone()
            holder["steering"] = stack_agents(agents, lambda a: a.state.steering.squeeze(-1)).clone()
            holder["vel"] = stack_agents(agents, lambda a: a.state.vel).clone()
            holder["sideslip"] = stack_agents(agents, lambda a: a.state.sideslip_angle.squeeze(-1)).clone()
            holder["col_agents"] = ws.collisions.with_agents.clone()
            holder["col_lane"] = ws.collisions.with_lanelets.clone()
            holder["col_entry"] = ws.collisions.with_entry_segments.clone()
            holder["col_exit"] = ws.collisions.with_exit_segments.clone()
            holder["d_ref"] = ws.distances.ref_paths.clone()
            holder["d_left"] = ws.distances.left_boundaries.clone()
            holder["d_right"] = ws.distances.right_boundaries.clone()
            holder["d_bound"] = ws.distances.boundaries.clone()
            holder["d_agents"] = ws.distances.agents.clone()
            holder["idx_ref"] = ws.distances.closest_point_on_ref_path.clone()
            holder["short_term"] = ws.ref_paths_agent_related.short_term.clone()
            holder["vertices"] = ws.vertices.clone()
            return _o()

        sc.done = done_spy
        obs, rew, done, info = env.step(acts)
        sc.done = orig_done
        sc.reset_world_at = orig_reset
        for k, v in holder.items():
            push(("post_" + k) if k in ("pos", "rot", "speed", "steering", "vel", "sideslip") else k, v)
        push("obs", torch.stack(obs, dim=1))
        push("reward", torch.stack(rew, dim=1))
        push("done", done)
        # info(agent) as vmas returns it (road_traffic.py:1489-1635), stacked over agents: [B, N, ...]
        for k in INFO_KEYS:
            push("info_" + k, torch.stack([torch.as_tensor(d[k]).reshape(B, -1).squeeze(-1) for d in info], dim=1))
        push("num_task_tries", sc.num_task_tries.clone())
        push("task_success_times", sc.task_success_times.clone())
        push("respawn_mask", respawn)
        # post-respawn state of respawned agents (read back from the world)
        push("respawn_pos", stack_agents(agents, lambda a: a.state.pos))
        push("respawn_rot", stack_agents(agents, lambda a: a.state.rot.squeeze(-1)))
        push("respawn_speed", stack_agents(agents, lambda a: a.state.speed.squeeze(-1)))
        push("respawn_path_id", ws.ref_paths_agent_related.path_id.clone())
        push("respawn_point_id", ws.ref_paths_agent_related.point_id.clone())
        # env resets, TorchRL-style: reset_at(i) for each done env, then one observation pass
        reset_obs = torch.zeros(B, N, obs[0].shape[-1])
        for e in torch.where(done)[0]:
            o = env.reset_at(int(e), return_observations=True)[0]
            reset_obs[int(e)] = torch.stack(o, dim=1)[int(e)]
        last_obs = [o.clone() for o in obs]
        if done.any():
            fresh = [sc.observation(a) for a in agents]
            for i in range(N):
                last_obs[i][done] = fresh[i][done]
        push("reset_mask", done)
        push("reset_obs", reset_obs)
        push("reset_pos", stack_agents(agents, lambda a: a.state.pos))
        push("reset_rot", stack_agents(agents, lambda a: a.state.rot.squeeze(-1)))
        push("reset_speed", stack_agents(agents, lambda a: a.state.speed.squeeze(-1)))
        push("reset_vel", stack_agents(agents, lambda a: a.state.vel))
        push("reset_path_id", ws.ref_paths_agent_related.path_id.clone())
        push("reset_point_id", ws.ref_paths_agent_related.point_id.clone())
        push("reset_scenario_id", ws.ref_paths_agent_related.scenario_id.clone())

    out = {k: np.stack(v) for k, v in rec.items()}
    th, pen, nrm = sc.thresholds, sc.penalties, sc.normalizers
    cfg = dict(
        scenario_type=st, B=B, N=N, T=T, mode=mode, seed=seed, dt=float(env.world.dt),
        max_steps=int(sc.parameters.max_steps), rew_method=str(sc.parameters.rew_method),
        n_nearing_agents_observed=int(sc.parameters.n_nearing_agents_observed),
        reward_progress=float(sc.rewards.progress),
        near_boundary_low=float(th.near_boundary_low), near_boundary_high=float(th.near_boundary_high),
        near_other_agents_low=float(th.near_other_agents_low), near_other_agents_high=float(th.near_other_agents_high),
        ttc_low=float(th.ttc_low), ttc_high=float(th.ttc_high),
        penalty_near_boundary=float(pen.near_boundary), penalty_near_other_agents=float(pen.near_other_agents),
        penalty_collide_with_agents=float(pen.collide_with_agents),
        penalty_collide_with_boundaries=float(pen.collide_with_boundaries),
        norm_pos=float(nrm.pos[0]), norm_v=float(nrm.v), norm_rot=float(nrm.rot),
        norm_distance_lanelet=float(nrm.distance_lanelet),
        is_testing_mode=bool(sc.parameters.is_testing_mode),
        reset_agent_fixed_duration=float(sc.parameters.reset_agent_fixed_duration),
        is_use_mtv_distance=bool(sc.parameters.is_use_mtv_distance),
        max_ref_path_points=int(ws.params.max_ref_path_points), gentle=bool(gentle),
        norm_pos_world_x=float(nrm.pos_world[0]), norm_pos_world_y=float(nrm.pos_world[1]),
        norm_distance_agent=float(nrm.distance_agent),
        **{f: bool(getattr(sc.parameters, f)) for f in OBS_FLAGS},
    )
    for k, v in cfg.items():
        out["cfg_" + k] = np.asarray(v)
    out_dir = os.path.join(OUT, "next") if name in NEXT else OUT
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: dones={int(out['done'].sum())} respawns={int(out['respawn_mask'].sum())} "
          f"col_agents={int(out['col_agents'].any(-1).sum())} col_lane={int(out['col_lane'].sum())} "
          f"size={os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    only = sys.argv[1:]
    for name, c in CONFIGS.items():
        if only and name not in only:
            continue
        run(name, **c)
