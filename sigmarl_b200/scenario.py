"""VMAS ``BaseScenario``-shaped facade over the fused CUDA step — the drop-in boundary.

Mirrors ``ScenarioRoadTraffic`` (``sigmarl/scenarios/road_traffic.py``): ``make_world`` (:104),
``reset_world_at`` (:816), ``process_action`` (no-op, ``dynamics.py:194``), ``reward`` (:925),
``observation`` (:1334), ``done`` (:1368), ``info`` (:1489) and ``world.step()``
(``WorldCustom.step``, ``helper_training.py:797``), with the same argument meaning and the same
attribute names on ``world.agents[i]`` (``state.pos/rot/vel/speed/steering/sideslip_angle``,
``action.u``, ``u_range``, ``max_speed``, ``dynamics.needed_action_size``).

How the per-agent VMAS protocol maps onto ONE kernel launch per step: ``world.step()`` launches the fused
kernel, which already produces every agent's reward / observation / flags; ``reward(agent_i)`` /
``observation(agent_i)`` return the ``[:, i]`` slice of the output buffers (VMAS clones them), ``done()``
returns the done mask and — like the reference (:1462-1472) — respawns agents that crossed an entry/exit
segment.  The stale-value semantics of the interleaved call order are reproduced inside
the kernel, so the calls may come in any order after ``world.step()``.

If the real ``vmas`` package is importable the scenario subclasses its ``BaseScenario``; otherwise a
minimal stand-in with the same surface is used (vmas is not installable in the build image).
"""
import math
from typing import Dict, Optional

import numpy as np
import torch

from .config import AGENT_LENGTH, AGENT_WIDTH, EnvConfig, MAX_SPEED, MAX_STEERING, check_fixed_parameters
from .env import RoadTrafficEnv
from . import lib as _lib

try:  # pragma: no cover - vmas is absent in the build image
    from vmas.simulator.scenario import BaseScenario as _VmasBaseScenario
    _HAVE_VMAS = True
except Exception:  # noqa: BLE001
    _VmasBaseScenario = object
    _HAVE_VMAS = False


class _Dynamics:
    """Stands in for ``KinematicBicycleModel`` (dynamics.py): the integration itself runs in the kernel."""
    needed_action_size = 2
    max_speed, max_steering = MAX_SPEED, MAX_STEERING

    def process_action(self):
        pass

    def check_and_process_action(self):
        pass


class _Action:
    def __init__(self, env: RoadTrafficEnv, i: int):
        self._env, self._i = env, i
        self.u_range = [MAX_SPEED, MAX_STEERING]
        self.u_multiplier = [1, 1]

    @property
    def u(self) -> torch.Tensor:
        return self._env.action[:, self._i]

    @u.setter
    def u(self, value: torch.Tensor):
        self._env.action[:, self._i].copy_(value)


class _VehicleState:
    """``VehicleState`` (helper_common.py:290-430) as views into the packed device state."""

    def __init__(self, env: RoadTrafficEnv, i: int):
        self._env, self._i = env, i

    pos = property(lambda s: s._env.pose[:, s._i, 0:2])
    rot = property(lambda s: s._env.pose[:, s._i, 2:3])
    speed = property(lambda s: s._env.pose[:, s._i, 3:4])
    steering = property(lambda s: s._env.aux[:, s._i, 0:1])
    vel = property(lambda s: s._env.aux[:, s._i, 1:3])
    sideslip_angle = property(lambda s: s._env.aux[:, s._i, 3:4])


class Vehicle:
    """``Vehicle`` (helper_common.py) / ``vmas.Agent`` surface used by the trainer and evaluation code."""

    def __init__(self, env: RoadTrafficEnv, i: int):
        self.name = f"agent_{i}"
        self.index = i
        self.state = _VehicleState(env, i)
        self.action = _Action(env, i)
        self.u_range = [MAX_SPEED, MAX_STEERING]
        self.max_speed = MAX_SPEED
        self.dynamics = _Dynamics()
        self.shape = type("Box", (), dict(length=AGENT_LENGTH, width=AGENT_WIDTH))()
        self.action_script = None


class WorldB200:
    """``WorldCustom`` surface (helper_training.py:791-861): ``step()`` is the fused kernel launch."""

    def __init__(self, env: RoadTrafficEnv, scenario):
        self._env, self._scenario = env, scenario
        self.batch_dim, self.device, self.dt = env.B, env.device, env.dt
        self.x_semidim = torch.tensor(env.map.world_x_dim, device=env.device, dtype=torch.float32)
        self.y_semidim = torch.tensor(env.map.world_y_dim, device=env.device, dtype=torch.float32)
        self.agents = [Vehicle(env, i) for i in range(env.N)]
        self.parameters = None

    @property
    def policy_agents(self):
        return self.agents

    @property
    def entities(self):
        return self.agents

    def step(self):
        self._env.step(None)
        self._scenario._stepped = True

    def reset(self, env_index=None):
        pass  # state is overwritten by reset_world_at (vmas zeroes it first; nothing reads the zeros)


class _Observations:
    """``observation_provider.observations`` as far as callers outside the scenario read it: ``nearing_agents_indices``
    [B, N, k] (observation_provider_rt.py:627-636; read by helper_training.py:240, 254 through ``scenario.observations``).
    The kernel selects the k nearest agents on the fly and does not store their indices; this view recomputes them on
    demand from the current positions with the reference's own formula (helper_scenario.py:1012-1029, 1140-1143:
    centre distances, own entry := world diagonal, ``torch.topk(largest=False)``)."""

    def __init__(self, scenario):
        self._sc = scenario

    @property
    def n_nearing_agents(self):
        return int(self._sc.env.cfg.k_near)

    @property
    def nearing_agents_indices(self) -> torch.Tensor:
        e = self._sc.env
        if e.config.is_use_mtv_distance:
            raise NotImplementedError("nearing_agents_indices is recomputed from centre distances; with "
                                      "is_use_mtv_distance the selection uses the MTV distance inside the kernel")
        pos = e.pose[..., 0:2]
        d = torch.sqrt(((pos.unsqueeze(2) - pos.unsqueeze(1)) ** 2).sum(-1))
        diag = math.sqrt(float(self._sc._norm["pos_world"][0]) ** 2 + float(self._sc._norm["pos_world"][1]) ** 2)
        d.diagonal(dim1=-2, dim2=-1).fill_(diag)
        return torch.topk(d, k=self.n_nearing_agents, dim=-1, largest=False).indices


class ScenarioRoadTrafficB200(_VmasBaseScenario):
    def __init__(self):
        if _HAVE_VMAS:  # pragma: no cover
            super().__init__()
        self._world = None
        self.env: Optional[RoadTrafficEnv] = None
        self.i_iter = 0
        self._stepped = False

    # -- vmas BaseScenario plumbing when vmas itself is absent
    if not _HAVE_VMAS:
        @property
        def world(self):
            return self._world

        def env_make_world(self, batch_dim, device, **kwargs):
            self._world = self.make_world(batch_dim, device, **kwargs)
            return self._world

        def env_reset_world_at(self, env_index):
            self.world.reset(env_index)
            self.reset_world_at(env_index)

        def env_process_action(self, agent):
            self.process_action(agent)
            agent.dynamics.check_and_process_action()

        def pre_step(self):
            pass

        def post_step(self):
            pass

    # -- road_traffic.py:104
    def make_world(self, batch_dim: int, device, **kwargs):
        if getattr(self, "_parameters_from_kwargs", False):      # a second make_world on a kwargs-mode scenario
            self.parameters = None
        seed = kwargs.pop("seed", 0)
        env_offset = kwargs.pop("env_offset", 0)
        debug = kwargs.pop("debug", False)
        if hasattr(self, "parameters") and self.parameters is not None:
            cfg = EnvConfig.from_parameters(self.parameters, **{k: v for k, v in kwargs.items()
                                                                if k in EnvConfig.__dataclass_fields__})
        else:
            known = {k: v for k, v in kwargs.items() if k in EnvConfig.__dataclass_fields__}
            known.setdefault("scenario_type", "cpm_entire")
            # the reference's kwargs-mode defaults (road_traffic.py:304-361) where they differ from EnvConfig's own
            # (the raw engine config defaults to noise-free observations): is_obs_noise=True, level 0.2 * agent width
            known.setdefault("is_obs_noise", True)
            cfg = EnvConfig(mode="kwargs", **known)
        self.config = cfg
        # hot-path parameters that exist at one value only: refuse anything else, whichever way it was passed
        _par = getattr(self, "parameters", None)
        check_fixed_parameters(lambda n: kwargs[n] if n in kwargs else getattr(_par, n, None), cfg.scenario_type)
        # evaluation set-up (helper_common.py:110-118; road_traffic.py:842-853): fixed reference paths and start poses
        par = getattr(self, "parameters", None)
        predef = kwargs.get("predefined_ref_path_idx", getattr(par, "predefined_ref_path_idx", None))
        init_state = kwargs.get("init_state", getattr(par, "init_state", None))
        self.env = RoadTrafficEnv(cfg, num_envs=batch_dim, device=device, seed=seed, env_offset=env_offset, debug=debug,
                                  info=True)
        self.n_agents = self.env.N
        self.max_speed, self.max_steering = MAX_SPEED, torch.tensor(MAX_STEERING, device=self.env.device)
        world = WorldB200(self.env, self)
        if getattr(self, "parameters", None) is None:
            # kwargs mode: the reference builds its own Parameters from the kwargs (road_traffic.py:317-361); callers read
            # these attributes through ``scenario.parameters`` (helper_training.py:207-253, 596-641, 709-741)
            import types
            self._parameters_from_kwargs = True
            self.parameters = types.SimpleNamespace(
                **{k: getattr(cfg, k) for k in cfg.__dataclass_fields__ if k not in ("extras", "mode", "exhaustive")})
            self.parameters.n_agents, self.parameters.num_vmas_envs = self.env.N, batch_dim
            self.parameters.dt, self.parameters.device = self.env.dt, self.env.device
            for flag in ("is_using_cbf", "is_using_cbf_training", "is_using_cbf_testing", "is_using_centralized_cbf",
                         "is_apply_cbf_action", "is_using_prioritized_marl", "is_using_opponent_modeling",
                         "is_communication_noise", "is_real_time_rendering"):
                setattr(self.parameters, flag, False)        # policy-side branches outside this library
            self.parameters.communication_noise_level = 0.0
        world.parameters = self.parameters
        self._world = world
        # evaluation counters (road_traffic.py:763-768): incremented by the kernel, same tensors
        self.num_task_tries = self.env.task_tries
        self.task_success_times = self.env.task_success
        dev = self.env.device
        m = self.env.map
        self._lanelet_ids = torch.as_tensor(m.lanelet_ids, device=dev)
        # the reference's path_id counts within the env's path set (world_state_rt_sim.py:313-358: cpm_mixed keeps one
        # list of paths per scenario_id); the library's is global over the map
        gp = np.arange(m.n_paths)
        first = np.asarray([m.set_range[s][0] for s in m.set_names], np.int32)
        self._path_in_set = torch.as_tensor((gp - first[m.set_of_path(gp)]).astype(np.int32), device=dev)
        self._norm = dict(                                                   # road_traffic.py:587-608
            pos_world=torch.tensor([m.world_x_dim, m.world_y_dim], device=dev, dtype=torch.float32),
            v=torch.tensor(MAX_SPEED, device=dev, dtype=torch.float32),
            rot=torch.tensor(2 * math.pi, device=dev, dtype=torch.float32),
            steering=torch.tensor(MAX_STEERING, device=dev, dtype=torch.float32),
            dist=torch.tensor(cfg.lane_width(m) * 3, device=dev, dtype=torch.float32))
        self._zeros_bn = torch.zeros(batch_dim, device=dev, dtype=torch.float32)
        self.observations = _Observations(self)                      # the (stale) name helper_training.py:240 uses
        self.observation_provider = type("ObservationProvider", (), dict(observations=self.observations))()
        self._predef_paths = self._init_state = None
        if predef is not None:
            if init_state is None or len(predef) != self.env.N or len(init_state) != self.env.N:
                raise ValueError("predefined_ref_path_idx needs init_state, both with one entry per agent")
            if getattr(self.env, "per_env_path_sets", False):
                raise NotImplementedError("predefined_ref_path_idx with several weighted cpm_scenario_probabilities: the "
                                          "indices would count within a path set that is drawn per env")
            lo, hi = self.env.path_lo, self.env.path_hi          # indices count within the scenario's path set
            ids = torch.as_tensor([int(i) for i in predef], dtype=torch.int32) + lo
            if int(ids.min()) < lo or int(ids.max()) >= hi:
                raise ValueError(f"predefined_ref_path_idx {list(predef)} outside the {hi - lo} paths of {cfg.scenario_type}")
            self._predef_paths = ids.to(dev)
            self._init_state = torch.as_tensor(init_state, dtype=torch.float32).reshape(self.env.N, 3).to(dev)
        return world

    def _reset_to_init_state(self, env_index: Optional[int]):
        """world_state_rt_sim.py:99-125: every agent at its (x, y, rot) of ``init_state`` on its predefined path, speed,
        velocity, steering and side-slip zero; then what any reset re-derives (sgb_refresh, all-fresh observation)."""
        e = self.env
        sl = slice(None) if env_index is None else int(env_index)
        e.pose[sl, :, 0:3] = self._init_state
        e.pose[sl, :, 3] = 0.0
        e.aux[sl] = 0.0
        e.path_id[sl] = self._predef_paths
        e.step_count[sl] = 0
        e.done[sl] = 0
        mask = None
        if env_index is not None:
            mask = torch.zeros(e.B, dtype=torch.uint8, device=e.device)
            mask[int(env_index)] = 1
        e.refresh(env_mask=mask, write_obs=True)

    # -- road_traffic.py:816
    def reset_world_at(self, env_index: Optional[int] = None, agent_index: Optional[int] = None):
        e = self.env
        if agent_index is None and self._predef_paths is not None:
            self._reset_to_init_state(env_index)     # road_traffic.py:842-853
            return
        if env_index is None:
            e.reset()
            return
        if agent_index is not None:
            # single-agent respawn (any map, any mode); the step-time observation stays, like in the reference
            m = torch.zeros(e.B, e.N, dtype=torch.uint8, device=e.device)
            m[int(env_index), int(agent_index)] = 1
            e.reset_masked(agent_mask=m, write_obs=False, path_range=self._respawn_range(int(agent_index)))
            return
        m = torch.zeros(e.B, dtype=torch.uint8, device=e.device)
        m[int(env_index)] = 1
        e.reset_masked(env_mask=m, write_obs=True)
        e.done[int(env_index)] = 0

    def _respawn_range(self, agent_index: int):
        """Path range a respawn of this agent draws from: its predefined path if there is one, else the env's set."""
        if self._predef_paths is None:
            return None
        p = int(self._predef_paths[agent_index])
        return (p, p + 1)

    def process_action(self, agent):
        pass

    # -- road_traffic.py:925 / :1334 / :1368
    def reward(self, agent) -> torch.Tensor:
        return self.env.reward[:, agent.index]

    def observation(self, agent) -> torch.Tensor:
        return self.env.obs[:, agent.index]

    def done(self) -> torch.Tensor:
        e = self.env
        is_done = e.done.bool().clone()
        if (e.cfg.respawn_on_exit or e.cfg.testing_mode) and self._stepped:
            # :1462-1472 — respawn entry/exit crossers of envs that are NOT done (done envs are reset by the caller);
            # :1435-1447 — in testing mode colliding agents are respawned one by one as well
            which = _lib.SGB_FLAG_ENTRY | _lib.SGB_FLAG_EXIT
            if e.cfg.testing_mode:
                which |= _lib.SGB_FLAG_COLLIDE_AGENT | _lib.SGB_FLAG_COLLIDE_LANE
            crossing = ((e.agent_flags & which) != 0) & ~is_done.unsqueeze(1)
            if self._predef_paths is None:
                e.reset_masked(agent_mask=crossing, write_obs=False)
            else:   # every agent comes back on ITS predefined path (world_state_rt_sim.py:241-242): one launch per agent
                for a in torch.nonzero(crossing.any(dim=0)).flatten().tolist():
                    col = torch.zeros_like(crossing)
                    col[:, a] = crossing[:, a]
                    e.reset_masked(agent_mask=col, write_obs=False, path_range=self._respawn_range(a))
        self._stepped = False
        return is_done

    # -- road_traffic.py:1489-1635: every key of the reference's info dict
    def info(self, agent) -> Dict[str, torch.Tensor]:
        """Same keys, shapes and dtypes as ``ScenarioRoadTraffic.info``.  State entries are views of the packed
        device state; what the step adds (fresh short-term path, distances, reward terms) comes from the
        ``info`` block the fused kernel writes (``include/sigmarl_b200.h``); ``*_nom`` are the reference's
        divisions by its normalizers (:587-608).  CBF / prioritized-MARL entries are out of scope."""
        e, i, nm = self.env, agent.index, self._norm
        fl = e.agent_flags[:, i]
        blk = e.info[:, i]
        two_pi = 2 * math.pi
        rot = torch.remainder(e.pose[:, i, 2:3], two_pi)          # angle_eliminate_two_pi, helper_scenario.py:1276-1289
        rot = torch.where(rot > math.pi, rot - two_pi, rot)
        pos, vel = e.pose[:, i, 0:2], e.aux[:, i, 1:3]
        act_vel, act_steer = e.action[:, i, 0], e.action[:, i, 1]
        ref = blk[:, 0:6]
        d_ref, d_left, d_right = blk[:, 6], blk[:, 7], blk[:, 8]
        z = self._zeros_bn
        return {
            "pos": pos, "pos_nom": pos / nm["pos_world"],
            "rot": rot, "rot_nom": rot / nm["rot"],
            "vel": vel, "vel_nom": vel / nm["v"],
            "act_vel": act_vel, "act_vel_nom": act_vel / nm["v"],
            "act_steer": act_steer, "act_steer_nom": act_steer / nm["steering"],
            "ref": ref, "ref_nom": (ref.reshape(-1, 3, 2) / nm["pos_world"]).reshape(-1, 6),
            "distance_ref": d_ref, "distance_ref_nom": d_ref / nm["dist"],
            "distance_left_b": d_left, "distance_left_b_nom": d_left / nm["dist"],
            "distance_right_b": d_right, "distance_right_b_nom": d_right / nm["dist"],
            "is_collision_with_agents": (fl & _lib.SGB_FLAG_COLLIDE_AGENT) != 0,
            "is_collision_with_lanelets": (fl & _lib.SGB_FLAG_COLLIDE_LANE) != 0,
            "is_reach_goal": (fl & _lib.SGB_FLAG_EXIT) != 0,
            "ref_lanelet_ids": self._lanelet_ids[e.path_id[:, i].long()],
            "path_id": self._path_in_set[e.path_id[:, i].long()],
            # no CBF: nominal == applied == the policy's (clamped) action (:1527-1545)
            "applied_action_vel": act_vel, "applied_action_steer": act_steer,
            "nominal_action_vel": act_vel, "nominal_action_steer": act_steer,
            # RewardInfo (helper_scenario.py:101-114): the reference writes five of the twelve fields
            "rew_progress": z, "rew_reach_goal": blk[:, 12], "rew_speed": z, "rew_centerline": z,
            "rew_near_other_agents": blk[:, 9], "rew_near_left_lane": z, "rew_near_right_lane": z,
            "rew_collide_other_agents": blk[:, 10], "rew_collide_lane": blk[:, 11],
            "rew_energy_acceleration": z, "rew_energy_steering": z, "rew_total": blk[:, 13],
        }

    def extra_render(self, env_index: int = 0):
        return []


class VmasLikeEnvironment:
    """The slice of ``vmas.Environment`` the reference is driven through (step / reset / reset_at /
    get_from_scenario order) for hosts where vmas itself is not installed."""

    def __init__(self, scenario, num_envs, device="cuda:0", max_steps=None, seed=None, **kwargs):
        self.scenario, self.num_envs, self.device, self.max_steps = scenario, num_envs, device, max_steps
        if seed is not None:
            kwargs["seed"] = seed
        if max_steps is not None and not hasattr(scenario, "parameters"):
            kwargs.setdefault("max_steps", max_steps)  # kwargs mode: the scenario's own time limit (road_traffic.py:1413)
        self.world = scenario.env_make_world(num_envs, device, **kwargs)
        self.agents = self.world.policy_agents
        self.n_agents = len(self.agents)
        self.reset()

    def reset(self, return_observations=True):
        self.scenario.env_reset_world_at(None)
        self.steps = torch.zeros(self.num_envs, device=self.scenario.env.device)
        return self.get_from_scenario(return_observations, False, False, False)[0]

    def reset_at(self, index, return_observations=True):
        self.scenario.env_reset_world_at(index)
        self.steps[index] = 0
        return self.get_from_scenario(return_observations, False, False, False)[0]

    def get_from_scenario(self, get_observations, get_rewards, get_infos, get_dones):
        obs, rews, infos = [], [], []
        for agent in self.agents:
            if get_rewards:
                rews.append(self.scenario.reward(agent).clone())
            if get_observations:
                obs.append(self.scenario.observation(agent).clone())
            if get_infos:
                infos.append({k: v.clone() for k, v in self.scenario.info(agent).items()})
        dones = None
        if get_dones:
            dones = self.scenario.done().clone()
            if self.max_steps is not None:
                dones = dones | (self.steps >= self.max_steps)
        return obs, rews, dones, infos

    def step(self, actions):
        for a, agent in zip(actions, self.agents):
            agent.action.u = a.to(torch.float32)
        for agent in self.world.agents:
            self.scenario.env_process_action(agent)
        self.scenario.pre_step()
        self.world.step()
        self.scenario.post_step()
        self.steps += 1
        return self.get_from_scenario(True, True, True, True)


def make_env(scenario_type="cpm_entire", num_envs=32, device="cuda:0", max_steps=128, parameters=None, **kwargs):
    """Convenience: scenario + (vmas or vmas-like) Environment, as ``mappo_cavs.py:166-184`` sets it up."""
    sc = ScenarioRoadTrafficB200()
    if parameters is not None:
        sc.parameters = parameters
    else:
        kwargs.setdefault("scenario_type", scenario_type)
    return VmasLikeEnvironment(sc, num_envs=num_envs, device=device, max_steps=max_steps, **kwargs)
