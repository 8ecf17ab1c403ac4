"""TEST INFRASTRUCTURE ONLY — ctypes wrapper around ``oracle/sigmarl_oracle.c``.

The C file is a scalar CPU restatement of the reference's environment step (see its header for
the pinning statement).  This wrapper (a) builds it with gcc, (b) lays the compiled map out in
the padded ``[n_paths, P, 2]`` form the reference keeps per agent (``world_state_rt.py:313-392``),
(c) restates the constants ``road_traffic.py:_init_params`` derives (``:112-768``), and (d) exposes
``OracleWorld`` with numpy views on the C world state.

Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py`` may import
this module.  It must never import ``sigmarl_b200`` (the product) and the product must never
import it.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libsigmarl_oracle.so")
MAPS = os.path.join(REPO, "sigmarl_b200", "maps")

AGENT_WIDTH, AGENT_LENGTH = 0.107, 0.22            # constants.py:630-631
L_WB, L_R = 0.15, 0.075                            # constants.py:635-637
MAX_SPEED, MAX_STEERING = 1.0, 31 * math.pi / 180  # constants.py:638-640
MAX_ACC, MAX_STEERING_RATE = 5.0, math.pi / 2      # constants.py:642-644
N_ST, SAMPLE_INTERVAL = 3, 2                       # road_traffic.py:273-275, 316


def build(force=False):
    src = os.path.join(HERE, "sigmarl_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "libsigmarl_oracle.so"])
    return LIB


class _Map(C.Structure):
    _fields_ = [("n_paths", C.c_int), ("P", C.c_int),
                ("center", C.c_void_p), ("left", C.c_void_p), ("right", C.c_void_p),
                ("n_center", C.c_void_p), ("n_left", C.c_void_p), ("n_right", C.c_void_p),
                ("is_loop", C.c_void_p), ("yaw", C.c_void_p),
                ("n_lanelets", C.c_int), ("lanelet_xy", C.c_void_p), ("lanelet_off", C.c_void_p),
                ("lanelet_adj", C.c_void_p)]


_CFG_FLOATS = ["dt", "max_speed", "max_steering", "max_acc", "max_steering_rate", "l_wb", "lr_over_lwb",
               "half_length", "half_width", "diag", "w_ref0", "w_ref1", "w_ref2", "speed_dt", "reward_progress",
               "nb_low", "nb_high", "na_low", "na_high", "ttc_low", "ttc_high", "pen_near_boundary",
               "pen_near_agents", "pen_collide_agents", "pen_collide_lane", "norm_pos", "norm_v", "norm_rot",
               "norm_dist", "dsafe_sq", "reset_min_dist_sq"]
_CFG_INTS = ["rew_exact_sparse", "rew_has_ttc", "rew_has_distance", "rew_has_sparse", "k_near", "max_steps",
             "is_cpm_entire", "sample_interval", "testing_mode"]


class _Cfg(C.Structure):
    _fields_ = ([(n, C.c_float) for n in _CFG_FLOATS] + [(n, C.c_int) for n in _CFG_INTS] +
                [("reward_reach_goal", C.c_float), ("obs_flags", C.c_int), ("norm_pos_world", C.c_float * 2),
                 ("norm_dist_agent", C.c_float), ("fixed_duration", C.c_float), ("use_mtv", C.c_int),
                 ("mask_distance", C.c_float)])


# observation layout flags (ORC_OBS_* in sigmarl_oracle.c) keyed by the reference's parameter names and the value
# that sets the bit (observation_provider_rt.py:594-925)
OBS_FLAG_BITS = dict(is_ego_view=(1, False), is_observe_vertices=(2, False), is_obs_steering=(4, True),
                     is_observe_ref_path_other_agents=(8, True), is_observe_distance_to_agents=(16, False),
                     is_observe_distance_to_center_line=(32, False), is_observe_distance_to_boundaries=(64, False),
                     is_apply_mask=(128, True))


def obs_flags_from(get):
    """``get(name)`` -> value of the reference parameter (or None when unknown = default)."""
    fl = 0
    for name, (bit, when) in OBS_FLAG_BITS.items():
        v = get(name)
        if v is not None and bool(v) == when:
            fl |= bit
    if (fl & 128) and (fl & 1):
        fl |= 256      # ORC_OBS_MASK_LANELETS: bird view + masks -> the lanelet-relation mask is live (on maps with a table)
    return fl


def f32(x):
    return np.float32(x)


class PaddedMap:
    """Padded per-path polylines, exactly as the reference stores them per (env, agent) slot.

    ``world_state_rt.py:279-311`` (6 extension points ``last + m*(last-prev)``),
    ``:313-392`` (tail padding = last extended point / last boundary point).
    """

    def __init__(self, scenario_type):
        z = np.load(os.path.join(MAPS, f"{scenario_type}.npz"))
        self.scenario_type = scenario_type
        self.sets = ["intersection", "merge_in", "merge_out"] if "cpm_mixed" in scenario_type else ["all"]
        self.world_x_dim = float(z["world_x_dim"])
        self.world_y_dim = float(z["world_y_dim"])
        self.lane_width = float(z["lane_width"])
        paths = []
        self.set_offset, self.set_count = {}, {}
        for s in self.sets:
            off = z[f"{s}_center_off"]
            self.set_offset[s] = len(paths)
            self.set_count[s] = len(off) - 1
            for i in range(len(off) - 1):
                paths.append(dict(
                    center=z[f"{s}_center_xy"][off[i]:off[i + 1]],
                    left=z[f"{s}_left_xy"][z[f"{s}_left_off"][i]:z[f"{s}_left_off"][i + 1]],
                    right=z[f"{s}_right_xy"][z[f"{s}_right_off"][i]:z[f"{s}_right_off"][i + 1]],
                    yaw=z[f"{s}_yaw"][z[f"{s}_yaw_off"][i]:z[f"{s}_yaw_off"][i + 1]],
                    is_loop=bool(z[f"{s}_is_loop"][i])))
        n_ext = N_ST * SAMPLE_INTERVAL
        self.P = P = max(p["center"].shape[0] for p in paths) + n_ext + 2   # road_traffic.py:505-530
        n = len(paths)
        self.n_paths = n
        self.center = np.zeros((n, P, 2), np.float32)
        self.left = np.zeros((n, P, 2), np.float32)
        self.right = np.zeros((n, P, 2), np.float32)
        self.yaw = np.zeros((n, P), np.float32)
        self.n_center = np.zeros(n, np.int32)
        self.n_left = np.zeros(n, np.int32)
        self.n_right = np.zeros(n, np.int32)
        self.is_loop = np.zeros(n, np.uint8)
        m = np.arange(1, n_ext + 1, dtype=np.int32).reshape(-1, 1)
        for i, p in enumerate(paths):
            c = p["center"].astype(np.float32)
            nc = c.shape[0]
            direction = c[-1] - c[-2]
            ext = c[-1] + m.astype(np.float32) * direction          # world_state_rt.py:291-294
            self.center[i, :nc] = c
            self.center[i, nc:nc + n_ext] = ext
            self.center[i, nc + n_ext:] = ext[-1]
            self.n_center[i] = nc
            for key, arr, cnt in (("left", self.left, self.n_left), ("right", self.right, self.n_right)):
                b = p[key].astype(np.float32)
                arr[i, :b.shape[0]] = b
                arr[i, b.shape[0]:] = b[-1]
                cnt[i] = b.shape[0]
            self.yaw[i, :p["yaw"].shape[0]] = p["yaw"]
            self.is_loop[i] = p["is_loop"]
        # lanelet table (OSM maps only): lanelet-relation observation mask, map_manager.py:39-119
        lp = os.path.join(MAPS, f"{scenario_type}.lanelets.npz")
        self.lanelet_xy = self.lanelet_off = self.lanelet_adj = None
        if os.path.exists(lp):
            zl = np.load(lp)
            self.lanelet_xy = np.ascontiguousarray(zl["center_xy"], np.float32)
            self.lanelet_off = np.ascontiguousarray(zl["center_off"], np.int32)
            self.lanelet_adj = np.ascontiguousarray(zl["adjacency"], np.uint8)

    def global_path(self, scenario_id, path_id):
        """(scenario_id, path_id) of the reference (world_state_rt_sim.py:313-358) -> global index."""
        scenario_id = np.asarray(scenario_id)
        path_id = np.asarray(path_id)
        if self.sets == ["all"]:
            return path_id.astype(np.int32)
        offs = np.asarray([0, self.set_offset["intersection"], self.set_offset["merge_in"],
                           self.set_offset["merge_out"]], np.int32)
        return (offs[scenario_id] + path_id).astype(np.int32)


def default_config(scenario_type, pmap, n_agents, mode="params", rew_method="distance", dt=None,
                   max_steps=128, n_nearing_agents_observed=2, **over):
    """Constants of ``road_traffic.py:_init_params`` for the two construction modes.

    mode "params": a ``Parameters`` object is attached (``mappo_cavs.py:168-169``; thresholds from
    ``helper_common.py:129-137``, dt from ``config.json:4``).  mode "kwargs": built from
    ``make_world(**kwargs)`` (``road_traffic.py:176-212, 317``).
    """
    # road_traffic.py:116-123: with a Parameters object and no kwargs, `scenario_type` defaults to
    # "cpm_entire" inside _init_params, so normalisers use the CPM lane width (0.15) on every map.
    lane_width = 0.15 if mode == "params" else pmap.lane_width
    if mode == "params":
        c = dict(reward_progress=0.1, nb_high=0.02, nb_low=0.0, na_high=0.3, na_low=0.0, ttc_low=0.0, ttc_high=3.75,
                 pen_near_boundary=-0.2, pen_near_agents=-0.2, dt=0.1)
    else:
        c = dict(reward_progress=10 / 100, nb_high=(lane_width - AGENT_WIDTH) / 2 * 0.9, nb_low=0.0,
                 na_high=AGENT_LENGTH + AGENT_WIDTH, na_low=(AGENT_LENGTH + AGENT_WIDTH) / 2,
                 ttc_low=0.0, ttc_high=3.75, pen_near_boundary=-20 / 100, pen_near_agents=-20 / 100, dt=0.05)
    if dt is not None:
        c["dt"] = dt
    c.update(pen_collide_agents=-100 / 100, pen_collide_lane=-100 / 100,
             norm_pos=AGENT_LENGTH * 10, norm_v=MAX_SPEED, norm_rot=2 * math.pi, norm_dist=lane_width * 3,
             rew_method=rew_method, max_steps=max_steps,
             k_near=min(n_nearing_agents_observed, n_agents - 1))
    c.update(over)
    return c


def make_cfg(scenario_type, pmap, c):
    cfg = _Cfg()
    dt32 = f32(c["dt"])
    w = np.linspace(1, 0.2, N_ST, dtype=np.float32)
    # torch.linspace(1, 0.2, 3, float32) = [1.0, 0.6, 0.2]; then /= sum   (road_traffic.py:536-543)
    w = np.asarray([f32(1.0), f32(1.0) + f32(1.0) * ((f32(0.2) - f32(1.0)) / f32(2.0)), f32(0.2)], np.float32)
    w = w / (w[0] + w[1] + w[2])
    x = f32(pmap.world_x_dim)
    y = f32(pmap.world_y_dim)
    vals = dict(
        dt=dt32, max_speed=MAX_SPEED, max_steering=MAX_STEERING, max_acc=MAX_ACC,
        max_steering_rate=MAX_STEERING_RATE, l_wb=L_WB, lr_over_lwb=L_R / L_WB,
        half_length=AGENT_LENGTH / 2, half_width=AGENT_WIDTH / 2,
        diag=np.sqrt(x * x + y * y, dtype=np.float32),
        w_ref0=w[0], w_ref1=w[1], w_ref2=w[2], speed_dt=f32(MAX_SPEED * c["dt"]),
        reward_progress=c["reward_progress"], nb_low=c["nb_low"], nb_high=c["nb_high"],
        na_low=c["na_low"], na_high=c["na_high"], ttc_low=c["ttc_low"], ttc_high=c["ttc_high"],
        pen_near_boundary=c["pen_near_boundary"], pen_near_agents=c["pen_near_agents"],
        pen_collide_agents=c["pen_collide_agents"], pen_collide_lane=c["pen_collide_lane"],
        norm_pos=c["norm_pos"], norm_v=c["norm_v"], norm_rot=c["norm_rot"], norm_dist=c["norm_dist"],
        dsafe_sq=float(f32(c["na_low"])) * float(f32(c["na_low"])),
        reset_min_dist_sq=(np.sqrt(f32(AGENT_LENGTH ** 2 + AGENT_WIDTH ** 2)) * f32(1.5)) ** 2,
    )
    for k, v in vals.items():
        setattr(cfg, k, float(f32(v)))
    rm = c["rew_method"]
    cfg.rew_exact_sparse = int(rm == "sparse")
    cfg.rew_has_ttc = int("ttc" in rm)
    cfg.rew_has_distance = int("distance" in rm)
    cfg.rew_has_sparse = int("sparse" in rm)
    cfg.k_near = int(c["k_near"])
    cfg.max_steps = int(c["max_steps"])
    cfg.is_cpm_entire = int(scenario_type == "cpm_entire")
    cfg.sample_interval = SAMPLE_INTERVAL
    cfg.testing_mode = int(bool(c.get("testing_mode", False)))
    cfg.reward_reach_goal = float(f32(c.get("reward_reach_goal", 100 / 100)))     # road_traffic.py:217-219
    cfg.obs_flags = int(c.get("obs_flags", 0))
    cfg.norm_pos_world[0] = float(x)                                               # road_traffic.py:593-595
    cfg.norm_pos_world[1] = float(y)
    cfg.norm_dist_agent = float(f32(AGENT_LENGTH * 10))                            # road_traffic.py:605-607
    cfg.fixed_duration = float(f32(c.get("fixed_duration", 0.0)))                  # road_traffic.py:1388-1393
    cfg.use_mtv = int(bool(c.get("use_mtv", False)))                               # road_traffic.py:611-614
    cfg.mask_distance = float(f32(AGENT_LENGTH * 5))                               # road_traffic.py:663
    return cfg


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_create.restype = C.c_void_p
        _lib.orc_create.argtypes = [C.c_int, C.c_int, C.POINTER(_Map), C.POINTER(_Cfg)]
        _lib.orc_destroy.argtypes = [C.c_void_p]
        _lib.orc_field.restype = C.c_void_p
        _lib.orc_field.argtypes = [C.c_void_p, C.c_char_p]
        _lib.orc_step.argtypes = [C.c_void_p] + [C.c_void_p] * 5 + [C.c_int]
        _lib.orc_refresh_env.argtypes = [C.c_void_p, C.c_int, C.c_int]
        _lib.orc_place.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float]
        _lib.orc_reset_env.restype = C.c_int
        _lib.orc_reset_env.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.c_int]
        _lib.orc_obs_dim.argtypes = [C.c_void_p]
        _lib.orc_fresh_obs.argtypes = [C.c_void_p, C.c_void_p]
        _lib.orc_test_perp.restype = C.c_float
        _lib.orc_test_perp.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
        _lib.orc_test_interx.restype = C.c_int
        _lib.orc_test_interx.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        _lib.orc_test_rect.argtypes = [C.c_float, C.c_float, C.c_void_p, C.c_float, C.c_void_p]
        _lib.orc_test_wrap.restype = C.c_float
        _lib.orc_test_wrap.argtypes = [C.c_float]
        _lib.orc_test_local.argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
        _lib.orc_test_dec.restype = C.c_float
        _lib.orc_test_dec.argtypes = [C.c_float] * 3
        _lib.orc_test_mtv.restype = C.c_float
        _lib.orc_test_mtv.argtypes = [C.c_void_p, C.c_void_p]
        _lib.orc_test_path_points.argtypes = [C.c_void_p] + [C.c_int] * 7 + [C.c_void_p]
    return _lib


_FIELDS = dict(pos=(np.float32, 2), rot=(np.float32, 1), speed=(np.float32, 1), steering=(np.float32, 1),
               vel=(np.float32, 2), sideslip=(np.float32, 1), path_id=(np.int32, 1), vertices=(np.float32, 10),
               d_ref=(np.float32, 1), d_left=(np.float32, 5), d_right=(np.float32, 5), d_bound=(np.float32, 1),
               idx_ref=(np.int32, 1), short_term=(np.float32, 6), prev_pos=(np.float32, 2),
               col_lane=(np.uint8, 1), col_entry=(np.uint8, 1), col_exit=(np.uint8, 1),
               idx_left=(np.int32, 1), idx_right=(np.int32, 1), near_fresh=(np.uint8, 1))


class OracleWorld:
    def __init__(self, scenario_type, B, N, config=None, pmap=None, **cfg_kwargs):
        self.L = lib()
        self.pmap = pmap or PaddedMap(scenario_type)
        self.config = config or default_config(scenario_type, self.pmap, N, **cfg_kwargs)
        self.cfg = make_cfg(scenario_type, self.pmap, self.config)
        p = self.pmap
        self._keep = [np.ascontiguousarray(a) for a in
                      (p.center, p.left, p.right, p.n_center, p.n_left, p.n_right, p.is_loop, p.yaw)]
        m = _Map(p.n_paths, p.P, *[a.ctypes.data for a in self._keep])
        if p.lanelet_xy is not None:
            m.n_lanelets = len(p.lanelet_off) - 1
            m.lanelet_xy, m.lanelet_off, m.lanelet_adj = (p.lanelet_xy.ctypes.data, p.lanelet_off.ctypes.data,
                                                          p.lanelet_adj.ctypes.data)
        self.B, self.N = B, N
        self.h = self.L.orc_create(B, N, C.byref(m), C.byref(self.cfg))
        assert self.h, "orc_create failed"
        self.D = self.L.orc_obs_dim(self.h)
        for name, (dt, w) in _FIELDS.items():
            setattr(self, name, self._view(name, dt, (B, N) if w == 1 else (B, N, w)))
        self.vertices = self.vertices.reshape(B, N, 5, 2)
        self.short_term = self.short_term.reshape(B, N, 3, 2)
        self.d_agents = self._view("d_agents", np.float32, (B, N, N))
        self.col_agents = self._view("col_agents", np.uint8, (B, N, N))
        self.step_count = self._view("step", np.int32, (B,))
        self.rng = C.c_uint64(0x1234567)

    def _view(self, name, dtype, shape):
        ptr = self.L.orc_field(self.h, name.encode())
        n = int(np.prod(shape))
        buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    def set_state(self, pos, rot, speed, steering, path_id, envs=None):
        """Teacher-forcing: inject a state and recompute everything derived from it (refresh)."""
        sl = slice(None) if envs is None else envs
        self.pos[sl] = pos
        self.rot[sl] = rot
        self.speed[sl] = speed
        self.steering[sl] = steering
        self.path_id[sl] = path_id
        for b in (range(self.B) if envs is None else np.atleast_1d(envs)):
            self.L.orc_refresh_env(self.h, int(b), -1)

    def refresh(self, b, agent=-1):
        self.L.orc_refresh_env(self.h, int(b), int(agent))

    def place(self, b, a, path, point, speed):
        self.L.orc_place(self.h, int(b), int(a), int(path), int(point), float(speed))

    def reset_env(self, b, agent=-1, path_lo=0, path_hi=None, max_tries=1000):
        hi = self.pmap.n_paths if path_hi is None else path_hi
        return self.L.orc_reset_env(self.h, int(b), int(agent), int(path_lo), int(hi), C.byref(self.rng), max_tries)

    def step(self, actions, n_threads=1):
        a = np.ascontiguousarray(actions, np.float32).copy()
        obs = np.zeros((self.B, self.N, self.D), np.float32)
        rew = np.zeros((self.B, self.N), np.float32)
        done = np.zeros(self.B, np.uint8)
        resp = np.zeros((self.B, self.N), np.uint8)
        self.L.orc_step(self.h, a.ctypes.data, obs.ctypes.data, rew.ctypes.data, done.ctypes.data, resp.ctypes.data,
                        int(n_threads))
        return obs, rew, done.astype(bool), resp.astype(bool)


def fresh_obs(w):
    """Observation of every env as the reference returns it right after a reset (all values fresh)."""
    obs = np.zeros((w.B, w.N, w.D), np.float32)
    w.L.orc_fresh_obs(w.h, obs.ctypes.data)
    return obs


def config_from_golden(g):
    """Oracle config straight from the constants the reference run itself reported (gen_golden.py)."""
    return dict(dt=float(g["cfg_dt"]), reward_progress=float(g["cfg_reward_progress"]),
                nb_low=float(g["cfg_near_boundary_low"]), nb_high=float(g["cfg_near_boundary_high"]),
                na_low=float(g["cfg_near_other_agents_low"]), na_high=float(g["cfg_near_other_agents_high"]),
                ttc_low=float(g["cfg_ttc_low"]), ttc_high=float(g["cfg_ttc_high"]),
                pen_near_boundary=float(g["cfg_penalty_near_boundary"]),
                pen_near_agents=float(g["cfg_penalty_near_other_agents"]),
                pen_collide_agents=float(g["cfg_penalty_collide_with_agents"]),
                pen_collide_lane=float(g["cfg_penalty_collide_with_boundaries"]),
                norm_pos=float(g["cfg_norm_pos"]), norm_v=float(g["cfg_norm_v"]), norm_rot=float(g["cfg_norm_rot"]),
                norm_dist=float(g["cfg_norm_distance_lanelet"]), rew_method=str(g["cfg_rew_method"]),
                max_steps=int(g["cfg_max_steps"]), k_near=int(g["cfg_n_nearing_agents_observed"]),
                testing_mode=bool(g["cfg_is_testing_mode"]),
                fixed_duration=float(g["cfg_reset_agent_fixed_duration"]) if "cfg_reset_agent_fixed_duration" in g.files else 0.0,
                use_mtv=bool(g["cfg_is_use_mtv_distance"]) if "cfg_is_use_mtv_distance" in g.files else False,
                obs_flags=obs_flags_from(lambda n: g["cfg_" + n] if ("cfg_" + n) in g.files else None))
