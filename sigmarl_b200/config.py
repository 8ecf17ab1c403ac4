"""EnvConfig: every constant ``ScenarioRoadTraffic._init_params`` (``road_traffic.py:112-768``) bakes into the
scenario, for its two construction modes, lowered to the POD ``sgb_config`` the kernels read.

mode "params": a ``Parameters`` object was attached before ``make_world`` (``mappo_cavs.py:168-169``); the
thresholds come from ``helper_common.py:129-137`` and ``dt`` from ``config.json:4``.
mode "kwargs": the scenario is built from ``make_world(**kwargs)`` with the defaults of
``road_traffic.py:176-212, 317``.
"""
import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import lib as _lib

# constants.py:628-647 (AGENTS)
AGENT_WIDTH, AGENT_LENGTH = 0.107, 0.22
L_F, L_R, L_WB = 0.075, 0.075, 0.15
MAX_SPEED, MAX_STEERING = 1.0, 31 * math.pi / 180
MAX_ACC, MAX_STEERING_RATE = 5.0, math.pi / 2
R_P_NORMALIZER = 100  # road_traffic.py:126-128
CPM_LANE_WIDTH = 0.15  # constants.py:21

_REW_METHODS = ("distance", "ttc", "sparse", "distance_sparse", "ttc_sparse")


def _f32(x):
    return np.float32(x)


# Parameters / kwargs of the reference that change the environment step but are supported at ONE value only (the
# reference's default): given anything else the host layer refuses instead of silently computing the default.
FIXED_PARAMETERS = dict(
    n_points_short_term=3,                       # road_traffic.py:273-275  (observation width, reward weights)
    sample_interval_ref_path=2,                  # road_traffic.py:316
    n_points_nearing_boundary=5,                 # road_traffic.py:296-298
    is_challenging_initial_state_buffer=False,   # road_traffic.py:857-870; crashes in the reference once it replays
    max_steering=MAX_STEERING, max_speed=MAX_SPEED,   # road_traffic.py:288-294 (AGENTS table)
    lane_width=0.25,                             # helper_common.py:119: the OSM maps were parsed with it (parse_osm.py:283-306)
)


def check_fixed_parameters(get, scenario_type: str):
    """``get(name)`` -> the caller's value or None.  Raises NotImplementedError for an unsupported value.

    ``n_observed_steps`` / ``n_stored_steps`` (road_traffic.py:280-286) are accepted at any value the reference accepts
    (1 <= observed <= stored): the reference keeps ring buffers of ``n_stored_steps`` past states
    (observation_provider_rt.py:49-339) but ``get_observation`` only ever reads ``get_latest()`` (:415-424, :681-919), so
    the number of "observed" steps never reaches the observation — here as there, the latest step is what is observed."""
    n_obs, n_sto = get("n_observed_steps"), get("n_stored_steps")
    n_obs, n_sto = (1 if n_obs is None else int(n_obs)), (5 if n_sto is None else int(n_sto))
    if not (1 <= n_obs <= n_sto):                 # the reference's own asserts, observation_provider_rt.py:90-98
        raise ValueError(f"need 1 <= n_observed_steps ({n_obs}) <= n_stored_steps ({n_sto})")
    for name, want in FIXED_PARAMETERS.items():
        v = get(name)
        if v is None:
            continue
        if name == "lane_width" and scenario_type.startswith("cpm"):
            continue                              # the CPM maps come from the XML file, lane_width is not used for them
        same = (bool(v) == want) if isinstance(want, bool) else abs(float(v) - float(want)) <= 1e-9 * max(1.0, abs(want))
        if not same:
            raise NotImplementedError(f"{name}={v!r}: this library supports {name}={want!r} only (sigmarl_b200/config.py "
                                      f"FIXED_PARAMETERS)")


@dataclass
class EnvConfig:
    scenario_type: str = "cpm_entire"
    n_agents: Optional[int] = None          # None -> SCENARIOS[scenario_type]["n_agents"]
    mode: str = "params"                    # "params" | "kwargs" (see module docstring)
    dt: Optional[float] = None              # None -> 0.1 (params) / 0.05 (kwargs)
    max_steps: int = 128                    # helper_common.py:49
    rew_method: str = "distance"            # helper_common.py:128
    n_nearing_agents_observed: int = 2      # config.json:29
    reward_progress: Optional[float] = None
    threshold_near_boundary_high: Optional[float] = None
    threshold_near_boundary_low: Optional[float] = None
    threshold_near_other_agents_c2c_high: Optional[float] = None
    threshold_near_other_agents_c2c_low: Optional[float] = None
    ttc_low: Optional[float] = None
    ttc_high: Optional[float] = None
    penalty_near_boundary: Optional[float] = None
    penalty_near_other_agents: Optional[float] = None
    penalty_collide_with_agents: float = -100 / R_P_NORMALIZER
    penalty_collide_with_boundaries: float = -100 / R_P_NORMALIZER
    cpm_scenario_probabilities: tuple = (1.0, 0.0, 0.0)
    exhaustive: bool = False                # debug: disable the pruned search (results must not change)
    is_testing_mode: bool = False           # road_traffic.py:1050-1055, 1429-1447; world_state_rt_sim.py:254-261
    reward_reach_goal: float = 100 / R_P_NORMALIZER   # road_traffic.py:217-219
    reset_agent_fixed_duration: float = 0   # seconds, 0 = off (helper_common.py:120; road_traffic.py:1388-1393)
    # observation layout (observation_provider_rt.py:594-925): any combination of these seven is supported
    is_ego_view: bool = True                       # False: bird view (global coordinates / pos_world)
    is_observe_vertices: bool = True               # False: pos, rot, length, width of a neighbour
    is_obs_steering: bool = False
    is_observe_ref_path_other_agents: bool = False
    is_observe_distance_to_agents: bool = True
    is_observe_distance_to_center_line: bool = True
    is_observe_distance_to_boundaries: bool = True # False: 5 + 5 boundary points around the closest ones
    # obs += obs_noise_level * U[0,1) (device generator).  NB the reference's own defaults are ON — Parameters
    # (helper_common.py:76) and make_world(**kwargs) (road_traffic.py:336) — and the drop-in facade applies them
    # (ScenarioRoadTrafficB200.make_world / EnvConfig.from_parameters); this raw engine config defaults to noise-free.
    is_obs_noise: bool = False
    obs_noise_level: Optional[float] = None        # None -> 0.05 (params, helper_common.py:77) / 0.2 * width (kwargs, :337-339)
    obs_noise_seed: int = 0
    # agent distance (road_traffic.py:611-614): False = centre to centre, True = MTV-based (SAT) distance between the
    # rectangles (helper_scenario.py:1030-1138) with its own thresholds (road_traffic.py:264-270; same in both modes)
    is_use_mtv_distance: bool = False
    threshold_near_other_agents_MTV_low: float = 0.0
    threshold_near_other_agents_MTV_high: float = AGENT_LENGTH
    is_apply_mask: bool = False                    # observed neighbours beyond 5 agent lengths show constants (:638-749)
    # flags of the reference that select code OUTSIDE the supported hot path; must keep these values
    is_partial_observation: bool = True            # False crashes in the reference itself (:808)
    extras: dict = field(default_factory=dict)

    def validate(self):
        if self.rew_method not in _REW_METHODS:
            raise NotImplementedError(f"rew_method {self.rew_method!r}: supported {_REW_METHODS} (cbf variants are out of scope)")
        want = dict(is_partial_observation=True)
        for k, v in want.items():
            if getattr(self, k) != v:
                raise NotImplementedError(f"{k}={getattr(self, k)} selects a non-default observation/reset variant "
                                          f"that is outside the accelerated hot path (SURVEY.md §8f-4)")
        if self.mode not in ("params", "kwargs"):
            raise ValueError("mode must be 'params' or 'kwargs'")

    @classmethod
    def from_parameters(cls, p, **over):
        """From a SigmaRL ``Parameters``-like object (``helper_common.py:26-252``); unknown attributes are ignored."""
        kw = dict(mode="params")
        for name in cls.__dataclass_fields__:
            if name in ("mode", "extras", "exhaustive"):
                continue
            if hasattr(p, name) and getattr(p, name) is not None:
                v = getattr(p, name)
                kw[name] = tuple(v) if name == "cpm_scenario_probabilities" else v
        kw.update(over)
        return cls(**kw)

    def resolved(self, lane_width: float, default_n_agents: int):
        """Fill the mode-dependent defaults (returns a plain dict of python floats)."""
        m = self.mode
        pick = lambda v, d: d if v is None else v  # noqa: E731
        if m == "params":
            d = dict(dt=0.1, reward_progress=0.1, nb_high=0.02, nb_low=0.0, na_high=0.3, na_low=0.0,
                     ttc_low=0.0, ttc_high=3.75, pen_nb=-0.2, pen_na=-0.2)
        else:
            d = dict(dt=0.05, reward_progress=10 / R_P_NORMALIZER,
                     nb_high=(lane_width - AGENT_WIDTH) / 2 * 0.9, nb_low=0.0,
                     na_high=AGENT_LENGTH + AGENT_WIDTH, na_low=(AGENT_LENGTH + AGENT_WIDTH) / 2,
                     ttc_low=0.0, ttc_high=3.75, pen_nb=-20 / R_P_NORMALIZER, pen_na=-20 / R_P_NORMALIZER)
        n = self.n_agents or default_n_agents
        if self.is_use_mtv_distance:                       # road_traffic.py:632-648
            na_high, na_low = self.threshold_near_other_agents_MTV_high, self.threshold_near_other_agents_MTV_low
        else:
            na_high = pick(self.threshold_near_other_agents_c2c_high, d["na_high"])
            na_low = pick(self.threshold_near_other_agents_c2c_low, d["na_low"])
        return dict(
            n_agents=n,
            dt=pick(self.dt, d["dt"]),
            reward_progress=pick(self.reward_progress, d["reward_progress"]),
            nb_high=pick(self.threshold_near_boundary_high, d["nb_high"]),
            nb_low=pick(self.threshold_near_boundary_low, d["nb_low"]),
            na_high=na_high, na_low=na_low,
            ttc_low=pick(self.ttc_low, d["ttc_low"]), ttc_high=pick(self.ttc_high, d["ttc_high"]),
            pen_nb=pick(self.penalty_near_boundary, d["pen_nb"]),
            pen_na=pick(self.penalty_near_other_agents, d["pen_na"]),
            k_near=min(self.n_nearing_agents_observed, n - 1),      # road_traffic.py:441-443
        )

    def fixed_period(self, dt: float) -> int:
        """``reset_agent_fixed_duration`` as a period in steps (0 = off).

        The reference ends an env whenever ``t % duration == 0`` with ``t = timer.step * dt`` evaluated in float32
        (int tensor times python float, ``road_traffic.py:1388-1393``).  For the usual pairs (dt 0.1 / 0.05, whole
        seconds) the rounded product is the exact multiple and the test fires at every ``duration / dt``-th step; the
        kernel takes that period.  The float test is replayed here for every step an episode can reach and any pair
        for which it is NOT ``step % period == 0`` is refused instead of being approximated."""
        dur = self.reset_agent_fixed_duration
        if not dur:
            return 0
        if dur < 0:
            raise ValueError("reset_agent_fixed_duration must be >= 0")
        steps = np.arange(0, max(int(self.max_steps), 2), dtype=np.int64)
        t = steps.astype(np.float32) * _f32(dt)                      # fp32 product, rounded once
        hit = (np.fmod(t, _f32(dur)) == 0) & (t != 0)                # torch.remainder == fmod for positive operands
        fired = steps[hit]
        if fired.size == 0:
            return 0                                                 # never fires before the time limit ends the episode
        period = int(fired[0])
        if not np.array_equal(hit, (steps % period == 0) & (steps != 0)):
            raise NotImplementedError(f"reset_agent_fixed_duration={dur} with dt={dt}: the reference's float32 test fires "
                                      f"at steps {fired.tolist()[:8]}..., which is not periodic")
        return period

    def lane_width(self, maplib) -> float:
        """Lane width the reference derives normalisers / kwargs-thresholds from.  Quirk reproduced on purpose:
        ``_init_params`` reads ``kwargs.pop("scenario_type", "cpm_entire")`` (road_traffic.py:116-123), so when the
        scenario is configured through a ``Parameters`` object (no kwargs) it uses the CPM lane width (0.15 m)
        whatever map is loaded."""
        return CPM_LANE_WIDTH if self.mode == "params" else maplib.lane_width

    def lower(self, maplib) -> "_lib.Config":
        """-> ctypes ``sgb_config``; every float is rounded to fp32 exactly where the reference does."""
        self.validate()
        r = self.resolved(self.lane_width(maplib), maplib.default_n_agents)
        c = _lib.Config()
        # torch.linspace(1, 0.2, 3, float32) / sum   (road_traffic.py:536-543)
        w = np.asarray([_f32(1.0), _f32(1.0) + _f32(1.0) * ((_f32(0.2) - _f32(1.0)) / _f32(2.0)), _f32(0.2)], np.float32)
        w = w / (w[0] + w[1] + w[2])
        x, y = _f32(maplib.world_x_dim), _f32(maplib.world_y_dim)
        na_low32 = float(_f32(r["na_low"]))
        vals = dict(
            dt=r["dt"], max_speed=MAX_SPEED, max_steering=MAX_STEERING, max_acc=MAX_ACC,
            max_steering_rate=MAX_STEERING_RATE, l_wb=L_WB, lr_over_lwb=L_R / L_WB,
            half_length=AGENT_LENGTH / 2, half_width=AGENT_WIDTH / 2,
            diag=np.sqrt(x * x + y * y, dtype=np.float32),                 # helper_scenario.py:1140-1143
            w_ref0=w[0], w_ref1=w[1], w_ref2=w[2],
            speed_dt=MAX_SPEED * r["dt"],                                   # road_traffic.py:986
            reward_progress=r["reward_progress"],
            near_boundary_low=r["nb_low"], near_boundary_high=r["nb_high"],
            near_agents_low=r["na_low"], near_agents_high=r["na_high"],
            ttc_low=r["ttc_low"], ttc_high=r["ttc_high"],
            penalty_near_boundary=r["pen_nb"], penalty_near_agents=r["pen_na"],
            penalty_collide_agents=self.penalty_collide_with_agents,
            penalty_collide_lane=self.penalty_collide_with_boundaries,
            norm_pos=AGENT_LENGTH * 10, norm_v=MAX_SPEED, norm_rot=2 * math.pi,
            norm_dist=self.lane_width(maplib) * 3,                               # road_traffic.py:587-608
            dsafe_sq=na_low32 * na_low32,                                   # road_traffic.py:1279,1291
            reset_min_dist_sq=(np.sqrt(_f32(AGENT_LENGTH ** 2 + AGENT_WIDTH ** 2)) * _f32(1.5)) ** 2,
        )
        for k, v in vals.items():
            setattr(c, k, float(_f32(v)))
        rm = self.rew_method
        c.rew_flags = ((_lib.SGB_REW_EXACT_SPARSE if rm == "sparse" else 0) | (_lib.SGB_REW_TTC if "ttc" in rm else 0) |
                       (_lib.SGB_REW_DISTANCE if "distance" in rm else 0) | (_lib.SGB_REW_SPARSE if "sparse" in rm else 0))
        c.k_near = int(r["k_near"])
        c.max_steps = int(self.max_steps)
        c.respawn_on_exit = int(self.scenario_type != "cpm_entire")       # road_traffic.py:1449
        c.exhaustive = int(self.exhaustive)
        c.reward_reach_goal = float(_f32(self.reward_reach_goal))
        c.testing_mode = int(bool(self.is_testing_mode))
        c.reset_fixed_period = self.fixed_period(r["dt"])
        c.use_mtv_distance = int(bool(self.is_use_mtv_distance))
        c.mask_distance = float(_f32(AGENT_LENGTH * 5))                        # road_traffic.py:663
        c.obs_flags = self.obs_flags()
        if self.is_apply_mask and not self.is_ego_view and maplib.lanelet_xy is not None:
            # the reference's lanelet-relation mask is live exactly here: bird view (the lanelet assignment is only
            # computed there, observation_provider_rt.py:585-588) on a map that knows neighbouring lanelets (OSM maps,
            # parse_osm.py:257-262).  Ego view and the CPM maps mask by distance alone, in the reference too.
            c.obs_flags |= _lib.SGB_OBS_MASK_LANELETS
        c.norm_pos_world_x, c.norm_pos_world_y = float(x), float(y)            # road_traffic.py:593-595
        c.norm_dist_agent = float(_f32(AGENT_LENGTH * 10))                     # road_traffic.py:605-607
        level = self.obs_noise_level if self.obs_noise_level is not None else (0.05 if self.mode == "params" else 0.2 * AGENT_WIDTH)
        c.obs_noise_level = float(_f32(level)) if self.is_obs_noise else 0.0
        c.obs_noise_seed = int(self.obs_noise_seed) & 0xffffffff
        return c

    def obs_flags(self) -> int:
        """SGB_OBS_* bits of the configured observation layout (0 = the reference's defaults)."""
        return ((0 if self.is_ego_view else _lib.SGB_OBS_BIRD_VIEW) |
                (0 if self.is_observe_vertices else _lib.SGB_OBS_CENTRES) |
                (_lib.SGB_OBS_STEERING if self.is_obs_steering else 0) |
                (_lib.SGB_OBS_REF_OTHERS if self.is_observe_ref_path_other_agents else 0) |
                (0 if self.is_observe_distance_to_agents else _lib.SGB_OBS_NO_DIST_AGENTS) |
                (0 if self.is_observe_distance_to_center_line else _lib.SGB_OBS_NO_DIST_CENTER) |
                (0 if self.is_observe_distance_to_boundaries else _lib.SGB_OBS_BOUNDARY_POINTS) |
                (_lib.SGB_OBS_APPLY_MASK if self.is_apply_mask else 0))

    def obs_dim(self, n_agents: int) -> int:
        """Observation width of this layout (== sgb_obs_dim of a context built from it)."""
        fl = self.obs_flags()
        k = min(self.n_nearing_agents_observed, n_agents - 1)
        own = (5 if fl & _lib.SGB_OBS_BIRD_VIEW else 1) + (1 if fl & _lib.SGB_OBS_STEERING else 0) + 6 + \
              (0 if fl & _lib.SGB_OBS_NO_DIST_CENTER else 1) + (20 if fl & _lib.SGB_OBS_BOUNDARY_POINTS else 2)
        per = (5 if fl & _lib.SGB_OBS_CENTRES else 8) + 2 + (1 if fl & _lib.SGB_OBS_STEERING else 0) + \
              (0 if fl & _lib.SGB_OBS_NO_DIST_AGENTS else 1) + (6 if fl & _lib.SGB_OBS_REF_OTHERS else 0)
        return own + per * k
