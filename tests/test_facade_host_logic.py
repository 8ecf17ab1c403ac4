"""CPU: host-side logic of the VMAS-shaped facade (sigmarl_b200/scenario.py) with the CUDA environment replaced by a
recording stand-in.  Nothing is computed here — the stand-in only owns CPU tensors of the right shapes and logs which
library entry points the facade would call with which masks / path ranges — so this checks indexing, broadcasting,
argument plumbing and error behaviour of the facade; the arithmetic behind it is covered by the -m gpu tests
(the product path itself refuses to run without the CUDA library and a device, see test_abi_and_host.py)."""
import numpy as np
import pytest
import torch

from sigmarl_b200 import EnvConfig, lib as L
from sigmarl_b200 import scenario as S
from sigmarl_b200.maps import MapLibrary


class _RecordingEnv:
    """Shape-compatible stand-in for RoadTrafficEnv: state tensors on the CPU, every launch recorded in ``calls``."""

    def __init__(self, cfg, num_envs=4, device="cpu", seed=0, env_offset=0, debug=False, info=False, **kw):
        self.config = cfg
        self.map = MapLibrary(cfg.scenario_type)
        self.cfg = cfg.lower(self.map)            # the real lowering / validation runs
        r = cfg.resolved(cfg.lane_width(self.map), self.map.default_n_agents)
        self.B, self.N, self.dt, self.device = int(num_envs), int(r["n_agents"]), r["dt"], torch.device("cpu")
        self.D = cfg.obs_dim(self.N)
        rng = self.map.default_path_range(cfg.cpm_scenario_probabilities)
        self.per_env_path_sets = rng is None
        self.path_lo, self.path_hi = (-1, 0) if rng is None else rng
        z = lambda *s, dtype=torch.float32: torch.zeros(*s, dtype=dtype)  # noqa: E731
        B, N = self.B, self.N
        self.pose, self.aux, self.carry, self.action = z(B, N, 4), z(B, N, 4), z(B, N, 4), z(B, N, 2)
        self.path_id, self.step_count = z(B, N, dtype=torch.int32), z(B, dtype=torch.int32)
        self.obs, self.reward, self.done = z(B, N, self.D), z(B, N), z(B, dtype=torch.uint8)
        self.agent_flags, self.collide_with = z(B, N, dtype=torch.uint8), z(B, N, dtype=torch.int32)
        self.info = z(B, N, L.SGB_INFO_DIM)
        self.task_tries, self.task_success = z(B, dtype=torch.int32), z(B, dtype=torch.int32)
        self.calls = []

    def step(self, action=None):
        self.calls.append(("step",))
        self.step_count += 1

    def reset(self):
        self.calls.append(("reset",))

    def refresh(self, env_mask=None, write_obs=False):
        self.calls.append(("refresh", None if env_mask is None else env_mask.clone(), bool(write_obs)))

    def reset_masked(self, env_mask=None, agent_mask=None, write_obs=True, path_range=None):
        self.calls.append(("reset_masked", None if env_mask is None else env_mask.clone(),
                           None if agent_mask is None else agent_mask.clone(), bool(write_obs), path_range))


@pytest.fixture
def facade(monkeypatch):
    monkeypatch.setattr(S, "RoadTrafficEnv", _RecordingEnv)

    def make(num_envs=4, **kw):
        sc = S.ScenarioRoadTrafficB200()
        world = sc.env_make_world(num_envs, "cpu", **kw)
        return sc, world, sc.env
    return make


def test_world_surface_and_views(facade):
    """Attribute surface the trainer reads (road_traffic.py:104-110, 770-814; helper_training.py:791-861)."""
    sc, world, e = facade(scenario_type="cpm_entire", n_agents=3)
    assert world.batch_dim == 4 and len(world.agents) == 3 and world.dt == 0.05          # kwargs mode
    a = world.agents[1]
    assert a.dynamics.needed_action_size == 2 and a.u_range == [1.0, pytest.approx(31 * np.pi / 180)]
    e.pose[:, 1] = torch.tensor([1.0, 2.0, 0.5, 0.7])
    e.aux[:, 1] = torch.tensor([0.1, 0.3, 0.4, 0.05])
    assert a.state.pos.shape == (4, 2) and a.state.rot.shape == (4, 1) and a.state.speed.shape == (4, 1)
    assert torch.equal(a.state.pos[0], torch.tensor([1.0, 2.0])) and float(a.state.steering[0]) == pytest.approx(0.1)
    assert torch.equal(a.state.vel[0], torch.tensor([0.3, 0.4])) and float(a.state.sideslip_angle[0]) == pytest.approx(0.05)
    a.action.u = torch.full((4, 2), 0.25)
    assert torch.equal(e.action[:, 1], torch.full((4, 2), 0.25)) and float(e.action[:, 0].abs().max()) == 0.0
    assert sc.reward(a).shape == (4,) and sc.observation(a).shape == (4, e.D)
    info = sc.info(a)
    assert len(info) == 39 and info["pos"].shape == (4, 2) and info["ref"].shape == (4, 6)


def test_reset_world_at_plumbing(facade):
    sc, world, e = facade(scenario_type="cpm_mixed", n_agents=4)
    sc.reset_world_at(None)
    assert e.calls[-1] == ("reset",)
    sc.reset_world_at(2)
    kind, em, am, wo, pr = e.calls[-1]
    assert kind == "reset_masked" and em.tolist() == [0, 0, 1, 0] and am is None and wo and pr is None
    sc.reset_world_at(env_index=1, agent_index=3)
    kind, em, am, wo, pr = e.calls[-1]
    assert em is None and int(am.sum()) == 1 and int(am[1, 3]) == 1 and not wo and pr is None


def test_done_respawns_only_crossers_of_running_envs(facade):
    """road_traffic.py:1462-1472: entry / exit crossers are respawned inside done() unless their env is done anyway;
    cpm_entire has no entries / exits; testing mode respawns colliding agents too (:1435-1447)."""
    sc, world, e = facade(scenario_type="cpm_mixed", n_agents=3)
    e.agent_flags[0, 1] = L.SGB_FLAG_EXIT
    e.agent_flags[1, 2] = L.SGB_FLAG_EXIT | L.SGB_FLAG_COLLIDE_LANE
    e.agent_flags[2, 0] = L.SGB_FLAG_COLLIDE_AGENT
    e.done[1] = 1
    world.step()
    d = sc.done()
    assert d.tolist() == [False, True, False, False]
    kind, em, am, wo, pr = e.calls[-1]
    assert kind == "reset_masked" and am.nonzero().tolist() == [[0, 1]] and not wo
    n = len(e.calls)
    sc.done()                                   # without a step in between nothing is respawned twice
    assert len(e.calls) == n
    sc2, world2, e2 = facade(scenario_type="cpm_entire", n_agents=3)
    e2.agent_flags[0, 1] = L.SGB_FLAG_COLLIDE_LANE
    world2.step(); sc2.done()
    assert [c[0] for c in e2.calls] == ["step"]
    sc3, world3, e3 = facade(scenario_type="cpm_entire", n_agents=3, is_testing_mode=True)
    e3.agent_flags[0, 1] = L.SGB_FLAG_COLLIDE_LANE
    world3.step(); sc3.done()
    assert e3.calls[-1][0] == "reset_masked" and e3.calls[-1][2].nonzero().tolist() == [[0, 1]]


def test_predefined_paths_and_init_state(facade):
    """road_traffic.py:842-853, world_state_rt_sim.py:99-125, 241-242."""
    init = [[1.0, 2.0, 0.1], [3.0, 1.0, -0.2], [2.0, 2.5, 3.0]]
    sc, world, e = facade(scenario_type="cpm_mixed", n_agents=3, predefined_ref_path_idx=[0, 5, 7], init_state=init)
    lo = e.path_lo                                # cpm_mixed: the intersection set does not start at global path 0
    e.pose.fill_(9.0); e.aux.fill_(9.0); e.step_count.fill_(7); e.done.fill_(1)
    sc.reset_world_at(None)
    assert torch.equal(e.pose[..., :3], torch.tensor(init).expand(4, 3, 3)) and float(e.pose[..., 3].abs().max()) == 0
    assert float(e.aux.abs().max()) == 0 and int(e.step_count.max()) == 0 and int(e.done.max()) == 0
    assert e.path_id.tolist() == [[lo, lo + 5, lo + 7]] * 4
    assert e.calls[-1][0] == "refresh" and e.calls[-1][1] is None and e.calls[-1][2]
    e.pose.fill_(9.0); e.step_count.fill_(7)
    sc.reset_world_at(2)                          # one env only
    assert torch.equal(e.pose[2, :, :3], torch.tensor(init)) and float(e.pose[0].min()) == 9.0
    assert e.step_count.tolist() == [7, 7, 0, 7] and e.calls[-1][1].tolist() == [0, 0, 1, 0]
    sc.reset_world_at(env_index=3, agent_index=1)   # respawn on the agent's own path
    assert e.calls[-1][0] == "reset_masked" and e.calls[-1][4] == (lo + 5, lo + 6)
    e.agent_flags[0, 0] = L.SGB_FLAG_EXIT
    e.agent_flags[1, 2] = L.SGB_FLAG_EXIT
    e.agent_flags[3, 2] = L.SGB_FLAG_ENTRY
    e.done.zero_()
    n = len(e.calls)
    world.step(); sc.done()
    new = e.calls[n + 1:]
    assert [(c[0], c[4]) for c in new] == [("reset_masked", (lo, lo + 1)), ("reset_masked", (lo + 7, lo + 8))]
    assert new[0][2].nonzero().tolist() == [[0, 0]] and new[1][2].nonzero().tolist() == [[1, 2], [3, 2]]
    with pytest.raises(ValueError):
        facade(scenario_type="cpm_mixed", n_agents=3, predefined_ref_path_idx=[0, 5], init_state=init)
    with pytest.raises(ValueError):
        facade(scenario_type="cpm_mixed", n_agents=3, predefined_ref_path_idx=[0, 5, 999], init_state=init)


def test_parameters_object_reaches_the_config(facade):
    """mappo_cavs.py:168-169: ``scenario.parameters = parameters`` before make_world; flags of the reference map onto
    EnvConfig, including the ones whose class defaults are ON in helper_common.py (MTV distance, masks, noise)."""
    class P:
        scenario_type, n_agents, dt, max_steps, rew_method = "roundabout_2", 5, 0.1, 64, "ttc_sparse"
        is_use_mtv_distance, is_apply_mask, is_obs_noise, obs_noise_level = True, True, True, 0.05
        is_ego_view, n_nearing_agents_observed, reset_agent_fixed_duration = True, 3, 2
        is_using_cbf = False                      # unknown to EnvConfig: ignored
    sc = S.ScenarioRoadTrafficB200()
    sc.parameters = P()
    sc.env_make_world(8, "cpu")
    c = sc.env.cfg
    assert sc.config.mode == "params" and sc.env.N == 5 and c.use_mtv_distance == 1 and c.reset_fixed_period == 20
    assert c.obs_flags == L.SGB_OBS_APPLY_MASK and abs(c.obs_noise_level - 0.05) < 1e-7 and c.k_near == 3
    assert c.near_agents_low == 0.0 and c.near_agents_high == float(np.float32(0.22))
    assert c.rew_flags == L.SGB_REW_TTC | L.SGB_REW_SPARSE


def test_out_td_recording_layout(facade):
    """helper_common.py:581-611 / helper_training.py:1638-1698: leaves ("agents", "info", key) shaped [B, T, A, F]."""
    from sigmarl_b200.rollout import OUT_TD_KEYS, record_out_td, reduce_out_td, trim_out_td
    monkey_sc = S.ScenarioRoadTrafficB200()
    env = S.VmasLikeEnvironment(monkey_sc, num_envs=1, device="cpu", max_steps=16, scenario_type="cpm_entire", n_agents=3)
    e = monkey_sc.env
    e.pose[0, :, 0] = torch.tensor([1.0, 2.0, 3.0])
    e.agent_flags[0, 2] = L.SGB_FLAG_COLLIDE_LANE
    policy = lambda obs: [torch.full((1, 2), 0.1 * (i + 1)) for i in range(len(obs))]  # noqa: E731
    out = record_out_td(env, policy, T=5)
    info = out["agents"]["info"]
    assert set(info) == set(OUT_TD_KEYS)
    assert info["pos"].shape == (1, 5, 3, 2) and info["rot"].shape == (1, 5, 3, 1) and info["ref"].shape == (1, 5, 3, 6)
    assert info["is_collision_with_lanelets"].shape == (1, 5, 3, 1) and info["is_collision_with_lanelets"].dtype == torch.bool
    assert info["ref_lanelet_ids"].shape == (1, 5, 3, e.map.n_lanelets_all)
    assert out["agents"]["observation"].shape == (1, 5, 3, e.D) and out["agents"]["action"].shape == (1, 5, 3, 2)
    assert out["agents"]["reward"].shape == (1, 5, 3, 1) and out["done"].shape == (1, 5, 1)
    assert torch.equal(out["agents"]["action"][0, :, 1], torch.full((5, 2), 0.2))
    assert info["pos"][0, :, :, 0].tolist() == [[1.0, 2.0, 3.0]] * 5
    red = reduce_out_td(out)
    assert red["pos"].shape == (5, 3, 2) and red["rot"].shape == (5, 3) and red["is_collision_with_lanelets"].shape == (5, 3, 1)
    assert red["is_collision_with_lanelets"][:, 2].all() and not red["is_collision_with_agents"].any()
    assert set(trim_out_td(out)["agents"]["info"]) == set(OUT_TD_KEYS)
    two = {"agents": {"info": {k: torch.cat([v, v]) for k, v in info.items()}}}
    with pytest.raises(ValueError):
        reduce_out_td(two)


def test_nearing_agents_indices_view(facade):
    """scenario.observations.nearing_agents_indices (helper_training.py:240, 254; observation_provider_rt.py:627-636)."""
    sc, world, e = facade(num_envs=2, scenario_type="cpm_entire", n_agents=4, n_nearing_agents_observed=2)
    e.pose[0, :, 0] = torch.tensor([0.0, 1.0, 1.5, 4.0])      # env 0: agents on a line
    e.pose[1, :, 1] = torch.tensor([0.0, 3.0, 0.4, 0.5])
    idx = sc.observations.nearing_agents_indices
    assert idx.shape == (2, 4, 2) and sc.observation_provider.observations is sc.observations
    assert idx[0].tolist() == [[1, 2], [2, 0], [1, 0], [2, 1]]
    assert idx[1].tolist() == [[2, 3], [3, 2], [3, 0], [2, 0]]
    sc2, _, _ = facade(scenario_type="cpm_entire", n_agents=3, is_use_mtv_distance=True)
    with pytest.raises(NotImplementedError):
        sc2.observations.nearing_agents_indices


def test_kwargs_mode_exposes_parameters(facade):
    """helper_training.py:207-253, 709-741 read scenario.parameters.* whatever the construction mode."""
    sc, world, e = facade(num_envs=6, scenario_type="roundabout_2", n_agents=5, n_nearing_agents_observed=3)
    p = sc.parameters
    assert (p.num_vmas_envs, p.n_agents, p.n_nearing_agents_observed, p.scenario_type) == (6, 5, 3, "roundabout_2")
    assert p.dt == 0.05 and not p.is_using_cbf_testing and not p.is_using_prioritized_marl and world.parameters is p
    assert sc.config.mode == "kwargs"
    sc.env_make_world(3, "cpu", scenario_type="cpm_entire", n_agents=2)      # re-made: still kwargs mode
    assert sc.config.mode == "kwargs" and sc.parameters.num_vmas_envs == 3 and sc.env.dt == 0.05


def test_kwargs_mode_defaults_are_the_references(facade):
    """make_world(**kwargs) without explicit values builds the Parameters of road_traffic.py:304-361: dt 0.05, two observed
    neighbours, ego view with vertices / all three distances, no mask, no MTV distance, no steering / neighbour paths —
    and observation noise ON at 0.2 x agent width (:336-339).  (The raw EnvConfig defaults to noise-free; the facade is
    the drop-in and applies the reference's default.)"""
    sc, _, e = facade(scenario_type="cpm_entire")
    p, c = sc.parameters, e.cfg
    assert (p.n_agents, p.dt, p.n_nearing_agents_observed) == (15, 0.05, 2)
    assert p.is_ego_view and p.is_observe_vertices and p.is_observe_distance_to_agents
    assert p.is_observe_distance_to_boundaries and p.is_observe_distance_to_center_line and p.is_partial_observation
    assert not (p.is_apply_mask or p.is_use_mtv_distance or p.is_obs_steering or p.is_observe_ref_path_other_agents)
    assert not p.is_testing_mode and p.reset_agent_fixed_duration == 0 and tuple(p.cpm_scenario_probabilities) == (1.0, 0.0, 0.0)
    assert p.is_obs_noise is True and abs(c.obs_noise_level - np.float32(0.2 * 0.107)) < 1e-9
    assert c.obs_flags == 0 and c.use_mtv_distance == 0 and c.k_near == 2
    # an explicit value wins, in either direction
    sc, _, e = facade(scenario_type="cpm_entire", is_obs_noise=False)
    assert sc.parameters.is_obs_noise is False and e.cfg.obs_noise_level == 0.0
    sc, _, e = facade(scenario_type="cpm_entire", obs_noise_level=0.01)
    assert abs(e.cfg.obs_noise_level - np.float32(0.01)) < 1e-9


def test_single_valued_parameters_are_refused_not_ignored(facade):
    """Parameters that change the step but exist at one value only (config.FIXED_PARAMETERS) fail loudly, whether they
    arrive as kwargs or on a Parameters object; their supported values pass."""
    for kw in (dict(n_points_short_term=5), dict(sample_interval_ref_path=1), dict(is_challenging_initial_state_buffer=True),
               dict(max_speed=2.0)):
        with pytest.raises(NotImplementedError):
            facade(scenario_type="cpm_entire", n_agents=2, **kw)
    # n_observed_steps: accepted wherever the reference accepts it (1 <= observed <= stored, observation_provider_rt.py:
    # 90-98) and, as in the reference — whose get_observation only reads get_latest() — without effect on the observation
    a = facade(scenario_type="cpm_entire", n_agents=2, n_observed_steps=3)[2]
    b = facade(scenario_type="cpm_entire", n_agents=2)[2]
    assert a.D == b.D and bytes(a.cfg) == bytes(b.cfg)
    facade(scenario_type="cpm_entire", n_agents=2, n_observed_steps=5, n_stored_steps=5)
    for kw in (dict(n_observed_steps=0), dict(n_observed_steps=6), dict(n_observed_steps=3, n_stored_steps=2)):
        with pytest.raises(ValueError):
            facade(scenario_type="cpm_entire", n_agents=2, **kw)
    with pytest.raises(NotImplementedError):
        facade(scenario_type="roundabout_2", n_agents=2, lane_width=0.3)       # OSM boundaries depend on it
    facade(scenario_type="cpm_entire", n_agents=2, lane_width=0.3)             # ... the CPM maps do not
    facade(scenario_type="roundabout_2", n_agents=2, lane_width=0.25, n_points_short_term=3,
           is_challenging_initial_state_buffer=False, is_real_time_rendering=True)

    class P:
        scenario_type, n_agents, is_challenging_initial_state_buffer = "cpm_entire", 2, True
    sc = S.ScenarioRoadTrafficB200()
    sc.parameters = P()
    with pytest.raises(NotImplementedError):
        sc.env_make_world(2, "cpu")


class _ToyEnv:
    """Deterministic stand-in with RoadTrafficEnv's rollout surface (obs / reward / done / action buffers, bind, step,
    reset_done): obs' = obs + mean(action), reward = sum(action), done every 3rd step for env b == t % B; a reset writes
    a recognisable fresh observation.  Lets the two collectors of rollout.collect be compared on the CPU."""

    def __init__(self, B=4, N=2, D=3):
        self.B, self.N, self.D, self.device = B, N, D, torch.device("cpu")
        self.obs, self.reward = torch.zeros(B, N, D), torch.zeros(B, N)
        self.done, self.action = torch.zeros(B, dtype=torch.uint8), torch.zeros(B, N, 2)
        self.t = 0

    def bind(self, **tensors):
        for k, v in tensors.items():
            assert k in ("obs", "reward", "done", "action") and v.shape == getattr(self, k).shape
            setattr(self, k, v)

    def step(self, action=None):
        if action is not None:
            self.action.copy_(action)
        prev = self._prev_obs if hasattr(self, "_prev_obs") else torch.zeros(self.B, self.N, self.D)
        self.obs.copy_(prev + self.action.mean(-1, keepdim=True))
        self.reward.copy_(self.action.sum(-1))
        self.done.zero_()
        if self.t % 3 == 2:
            self.done[self.t % self.B] = 1
        self.t += 1
        self._prev_obs = self.obs.clone()
        return self.obs, self.reward, self.done

    def reset_done(self, write_obs=True):
        m = self.done.bool()
        if m.any():
            self.obs[m] = -100.0 - self.t
            self._prev_obs = self.obs.clone()


def test_in_place_and_copying_collectors_agree_on_the_cpu():
    """rollout.collect (helper_training.py:686-788's replacement): with in_place=True the env's outputs are bound to the
    rollout buffer — step t reads buf.action[t], writes buf.reward[t], buf.done[t] and buf.obs[t + 1] — and must leave
    exactly what the copying collector leaves, including the post-reset observations of done envs and the env's own
    buffers being restored afterwards."""
    from sigmarl_b200.rollout import RolloutBuffer, collect
    T = 8
    res = []
    for in_place in (False, True):
        env = _ToyEnv()
        env._prev_obs = torch.ones(env.B, env.N, env.D)
        env.obs.fill_(1.0)
        own = dict(obs=env.obs, reward=env.reward, done=env.done, action=env.action)
        g = torch.Generator().manual_seed(0)
        policy = lambda obs: torch.rand(4, 2, 2, generator=g) + 0.01 * obs[..., :2]  # noqa: E731
        value = lambda obs: obs[..., 0] * 2.0  # noqa: E731
        buf = collect(env, policy, RolloutBuffer(T, env.B, env.N, env.D, "cpu"), value_fn=value, in_place=in_place)
        assert all(getattr(env, k) is v for k, v in own.items())          # bindings restored
        res.append((buf, env.obs.clone(), env.reward.clone(), env.done.clone()))
    a, b = res
    for name in ("obs", "action", "reward", "done", "value", "next_value"):
        assert torch.equal(getattr(a[0], name), getattr(b[0], name)), name
    assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])
    assert int(a[0].done.sum()) == 2 and (a[0].obs[3, 2] <= -100).all()      # the reset env's next observation is the fresh one
    # TorchRL semantics: next_value of the done step is the value of the step-time observation, not of the fresh one
    assert float(a[0].next_value[2, 2].min()) > -100.0
