"""CPU (-m "not gpu"): the C-ABI library loads and exports every symbol include/sigmarl_b200.h declares,
refuses to run without a GPU (no CPU fallback), and the host-side mirrors (map library, config lowering)
agree with the oracle's independent restatement of the same reference constants."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    import __graft_entry__ as ge
    ge.build()
    from sigmarl_b200 import lib
    return lib.load_library()


def test_header_symbols_are_exported(L):
    from sigmarl_b200 import lib
    hdr = open(os.path.join(REPO, "include", "sigmarl_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(sgb_[a-z_]+)\s*\(", hdr)))
    assert declared == sorted(lib.EXPORTS), (declared, sorted(lib.EXPORTS))
    for name in declared:
        assert hasattr(L, name), name
    assert L.sgb_version() == int(re.search(r"#define SGB_VERSION (\d+)", hdr).group(1))
    # the host-side self-test hooks live in their own header and their own library (test infrastructure): the product
    # library must not export a single one of them, the test library all of them
    thdr = open(os.path.join(REPO, "include", "sigmarl_b200_test.h")).read()
    hooks = sorted(set(re.findall(r"\b(sgb_debug_[a-z_]+)\s*\(", thdr)))
    assert hooks == sorted(lib.TEST_EXPORTS) and "sgb_debug" not in hdr.replace("sgb_debug_*", "")
    T = lib.load_test_library()
    for name in hooks:
        assert hasattr(T, name), name
        with pytest.raises(AttributeError):
            getattr(L, name)
    assert T.sgb_version() == L.sgb_version()


def test_integration_stub_lists_the_structs_as_the_library_binds_them():
    """INTEGRATION.md shows the ctypes stub a maintainer of the reference would add: its sgb_buffers field list and the
    tail of its sgb_config must be the ones sigmarl_b200/lib.py binds (and the header declares, see the layout test)."""
    from sigmarl_b200 import lib
    doc = open(os.path.join(REPO, "INTEGRATION.md")).read()
    blk = doc[doc.index("class Buffers(C.Structure)"):doc.index("ctx = C.c_void_p()")]
    names = re.findall(r'"([a-z_]+)"', blk)
    assert names == lib.BUFFER_FIELDS, (names, lib.BUFFER_FIELDS)
    cfg = doc[doc.index("class Config(C.Structure)"):doc.index("class Buffers(C.Structure)")]
    tail = [n for n, _ in lib.Config._fields_ if n not in lib.CONFIG_FLOATS and not n.startswith("reserved")]
    assert [n for n in re.findall(r'"([a-z_]+)"', cfg) if n in tail] == tail


def test_fused_gae_allgather_refuses_bad_arguments(L):
    """sgb_gae_allgather validates before it touches a device (no compute here): world / rank ranges, NULL buffers,
    NULL peer entries, one multicast mapping without the other."""
    one = (C.c_void_p * 1)(0x1000)
    null1 = (C.c_void_p * 1)(None)
    vp = C.c_void_p
    good = lambda: [4, 8, 2, vp(0x1000), vp(0x2000), vp(0x3000), vp(0x4000), 0.99, 0.9]  # noqa: E731
    assert L.sgb_gae_allgather(*good(), 0, 0, one, one, None, None, None) == -1           # world < 1
    assert L.sgb_gae_allgather(*good(), 17, 0, one, one, None, None, None) == -1          # world > 16
    assert L.sgb_gae_allgather(*good(), 1, 1, one, one, None, None, None) == -1           # rank outside the world
    assert L.sgb_gae_allgather(*good(), 1, 0, None, one, None, None, None) == -1          # no peer table
    assert L.sgb_gae_allgather(*good(), 1, 0, null1, one, None, None, None) == -1         # NULL peer entry
    assert L.sgb_gae_allgather(*good(), 1, 0, one, one, vp(0x5000), None, None) == -1     # half a multicast pair
    bad = good(); bad[3] = None
    assert L.sgb_gae_allgather(*bad, 1, 0, one, one, None, None, None) == -1              # NULL reward
    bad = good(); bad[0] = 0
    assert L.sgb_gae_allgather(*bad, 1, 0, one, one, None, None, None) == -1              # T = 0


def test_struct_layouts_match_header():
    from sigmarl_b200 import lib
    hdr = open(os.path.join(REPO, "include", "sigmarl_b200.h")).read()
    body = hdr[hdr.index("typedef struct {\n    float dt;"):hdr.index("} sgb_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S).replace("typedef struct {", "")
    fields = []   # (name, ctype) in declaration order
    ctypes_of = {"float": C.c_float, "int32_t": C.c_int32, "uint32_t": C.c_uint32}
    for decl in body.split(";"):
        m = re.match(r"\s*(float|int32_t|uint32_t)\s+(.*)", decl.strip().replace("\n", " "))
        if not m:
            continue
        for name in m.group(2).split(","):
            name = name.strip()
            if name == "w_ref[SGB_N_SHORT_TERM]":
                fields += [(f"w_ref{k}", C.c_float) for k in range(3)]
            else:
                fields.append((name, ctypes_of[m.group(1)]))
    assert fields == list(lib.Config._fields_), (fields, lib.Config._fields_)
    assert C.sizeof(lib.Config) == 4 * len(fields)
    bufs = hdr[hdr.index("typedef struct {\n    float*   pose;"):hdr.index("} sgb_buffers;")]
    names = re.findall(r"\*\s*([a-z_]+);", bufs)
    assert names == lib.BUFFER_FIELDS


def test_no_cpu_fallback(L):
    from sigmarl_b200 import EnvConfig, MapLibrary, RoadTrafficEnv, SgbError, lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(SgbError):
        RoadTrafficEnv(EnvConfig(), num_envs=4)
    m = MapLibrary("cpm_entire")
    cfg = EnvConfig(scenario_type="cpm_entire", n_agents=4).lower(m)
    ctx = C.c_void_p()
    d = m.desc()
    rc = L.sgb_create(C.byref(ctx), 0, C.byref(d), C.byref(cfg))
    assert rc == -3 and not ctx.value, "sgb_create must fail with SGB_ERR_NO_DEVICE when there is no GPU"
    assert b"CPU fallback" in L.sgb_status_string(rc)


def test_product_never_imports_oracle():
    for root, _, files in os.walk(os.path.join(REPO, "sigmarl_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "oracle" not in src.replace("oracle/gen_maps.py", "").replace("oracle/gen_golden.py", ""), f


@pytest.mark.parametrize("st", ["cpm_entire", "cpm_mixed", "intersection_1", "on_ramp_2_multilane", "roundabout_2"])
def test_map_library_matches_oracle_padding(st):
    from oracle import oracle as O
    from sigmarl_b200 import MapLibrary
    m, pm = MapLibrary(st), O.PaddedMap(st)
    assert m.n_paths == pm.n_paths and m.max_ref_path_points == pm.P
    assert np.array_equal(m.n_center, pm.n_center)
    for i in range(m.n_paths):
        a, b = m.center_off[i], m.center_off[i + 1]
        assert np.array_equal(m.center_xy[a:b], pm.center[i, :b - a])
        assert np.array_equal(m.left_xy[m.left_off[i]:m.left_off[i + 1]], pm.left[i, :pm.n_left[i]])
    assert np.array_equal(m.is_loop, pm.is_loop)


@pytest.mark.parametrize("mode,rew", [("params", "distance"), ("kwargs", "ttc_sparse"), ("params", "sparse")])
@pytest.mark.parametrize("st", ["cpm_entire", "intersection_1"])
def test_config_lowering_matches_oracle_constants(st, mode, rew):
    """Two independent restatements of road_traffic.py:_init_params must produce identical fp32 constants."""
    from oracle import oracle as O
    from sigmarl_b200 import EnvConfig, MapLibrary, lib
    m, pm = MapLibrary(st), O.PaddedMap(st)
    N = 4
    c = EnvConfig(scenario_type=st, n_agents=N, mode=mode, rew_method=rew).lower(m)
    o = O.make_cfg(st, pm, O.default_config(st, pm, N, mode=mode, rew_method=rew))
    pairs = dict(dt="dt", max_speed="max_speed", max_steering="max_steering", max_acc="max_acc",
                 max_steering_rate="max_steering_rate", l_wb="l_wb", lr_over_lwb="lr_over_lwb",
                 half_length="half_length", half_width="half_width", diag="diag", w_ref0="w_ref0", w_ref1="w_ref1",
                 w_ref2="w_ref2", speed_dt="speed_dt", reward_progress="reward_progress",
                 near_boundary_low="nb_low", near_boundary_high="nb_high", near_agents_low="na_low",
                 near_agents_high="na_high", ttc_low="ttc_low", ttc_high="ttc_high",
                 penalty_near_boundary="pen_near_boundary", penalty_near_agents="pen_near_agents",
                 penalty_collide_agents="pen_collide_agents", penalty_collide_lane="pen_collide_lane",
                 norm_pos="norm_pos", norm_v="norm_v", norm_rot="norm_rot", norm_dist="norm_dist",
                 dsafe_sq="dsafe_sq", reset_min_dist_sq="reset_min_dist_sq")
    for a, b in pairs.items():
        assert getattr(c, a) == getattr(o, b), (a, getattr(c, a), getattr(o, b))
    assert c.k_near == o.k_near and c.max_steps == o.max_steps
    assert bool(c.rew_flags & lib.SGB_REW_TTC) == bool(o.rew_has_ttc)
    assert bool(c.rew_flags & lib.SGB_REW_DISTANCE) == bool(o.rew_has_distance)
    assert bool(c.rew_flags & lib.SGB_REW_SPARSE) == bool(o.rew_has_sparse)
    assert bool(c.rew_flags & lib.SGB_REW_EXACT_SPARSE) == bool(o.rew_exact_sparse)
    assert bool(c.respawn_on_exit) == (not o.is_cpm_entire)


def test_config_matches_reference_constants_in_goldens():
    """...and both agree with what the reference run itself reported (cfg_* entries of the goldens)."""
    import glob
    from sigmarl_b200 import EnvConfig, MapLibrary
    for p in sorted(glob.glob(os.path.join(REPO, "tests", "golden", "*.npz")) +
                    glob.glob(os.path.join(REPO, "tests", "golden", "next", "*.npz"))):
        g = np.load(p)
        st, mode = str(g["cfg_scenario_type"]), str(g["cfg_mode"])
        extra = {}
        if "cfg_is_use_mtv_distance" in g.files:          # MTV thresholds (road_traffic.py:264-270, 632-648)
            extra["is_use_mtv_distance"] = bool(g["cfg_is_use_mtv_distance"])
        if "ttc_sparse" in p and "cpm_mixed" in p:
            extra["threshold_near_other_agents_c2c_low"] = 0.1635
        c = EnvConfig(scenario_type=st, n_agents=int(g["cfg_N"]), mode=mode, rew_method=str(g["cfg_rew_method"]),
                      n_nearing_agents_observed=int(g["cfg_n_nearing_agents_observed"]), **extra).lower(MapLibrary(st))
        f32 = lambda k: float(np.float32(g[k]))  # noqa: E731
        assert c.dt == f32("cfg_dt") and c.reward_progress == f32("cfg_reward_progress"), p
        assert c.near_boundary_high == f32("cfg_near_boundary_high") and c.near_agents_low == f32("cfg_near_other_agents_low"), p
        assert c.near_agents_high == f32("cfg_near_other_agents_high") and c.ttc_high == f32("cfg_ttc_high"), p
        assert c.penalty_near_boundary == f32("cfg_penalty_near_boundary"), p
        assert c.norm_pos == f32("cfg_norm_pos") and c.norm_dist == f32("cfg_norm_distance_lanelet"), p
        assert c.k_near == int(g["cfg_n_nearing_agents_observed"]), p


def test_unsupported_flags_fail_loudly():
    from sigmarl_b200 import EnvConfig, MapLibrary
    m = MapLibrary("cpm_entire")
    for kw in (dict(rew_method="cbf"), dict(is_partial_observation=False)):
        with pytest.raises(NotImplementedError):
            EnvConfig(**{"scenario_type": "cpm_entire", **kw}).lower(m)
    # observation layouts and noise ARE supported (ABI 121 / 122): flags, width and noise level reach sgb_config
    from sigmarl_b200 import lib
    c = EnvConfig(scenario_type="cpm_entire", n_agents=4, is_ego_view=False, is_obs_steering=True,
                  is_observe_distance_to_boundaries=False, is_obs_noise=True)
    low = c.lower(m)
    assert low.obs_flags == lib.SGB_OBS_BIRD_VIEW | lib.SGB_OBS_STEERING | lib.SGB_OBS_BOUNDARY_POINTS
    assert abs(low.obs_noise_level - 0.05) < 1e-7 and c.obs_dim(4) == (5 + 1 + 6 + 1 + 20) + 2 * (8 + 2 + 1 + 1)
    assert EnvConfig(scenario_type="cpm_entire").lower(m).obs_flags == 0 and EnvConfig().obs_dim(8) == 32
    assert EnvConfig(scenario_type="cpm_entire", is_testing_mode=True).lower(m).testing_mode == 1   # supported since ABI 110
    # MTV agent distance (ABI 124): its own thresholds (road_traffic.py:264-270, 632-648) in both construction modes
    for mode in ("params", "kwargs"):
        low = EnvConfig(scenario_type="cpm_entire", mode=mode, is_use_mtv_distance=True).lower(m)
        assert low.use_mtv_distance == 1 and low.near_agents_low == 0.0 and low.dsafe_sq == 0.0
        assert low.near_agents_high == float(np.float32(0.22))
    assert EnvConfig(scenario_type="cpm_entire").lower(m).use_mtv_distance == 0
    # observation masks (ABI 125): by distance; ego view on any map, bird view on the CPM maps
    low = EnvConfig(scenario_type="cpm_entire", is_apply_mask=True, is_ego_view=False).lower(m)
    assert low.obs_flags == lib.SGB_OBS_APPLY_MASK | lib.SGB_OBS_BIRD_VIEW and low.mask_distance == float(np.float32(1.1))
    assert EnvConfig(scenario_type="roundabout_2", is_apply_mask=True).lower(MapLibrary("roundabout_2")).obs_flags == lib.SGB_OBS_APPLY_MASK
    # ... and the lanelet-relation criterion exactly where the reference applies it: bird view on an OSM map (ABI 126)
    low = EnvConfig(scenario_type="roundabout_2", is_apply_mask=True, is_ego_view=False).lower(MapLibrary("roundabout_2"))
    assert low.obs_flags == lib.SGB_OBS_APPLY_MASK | lib.SGB_OBS_BIRD_VIEW | lib.SGB_OBS_MASK_LANELETS
    with pytest.raises(ValueError):
        MapLibrary("no_such_map")


def test_cpm_mixed_path_sets():
    """cpm_scenario_probabilities (road_traffic.py:333-334; world_state_rt_sim.py:313-358): one weighted set -> a plain
    path range; several -> a set is drawn per env (ABI 128: sgb_set_path_sets, path_lo = -1)."""
    from sigmarl_b200 import MapLibrary
    m = MapLibrary("cpm_mixed")
    assert m.set_range == dict(intersection=(0, 24), merge_in=(24, 28), merge_out=(28, 32))
    assert m.default_path_range((1.0, 0.0, 0.0)) == (0, 24) and m.default_path_range([0, 1, 0]) == (24, 28)
    assert m.default_path_range([0, 0, 2.0]) == (28, 32) and m.default_path_range([0.3, 0.3, 0.4]) is None
    lo, hi, pr = m.path_sets([0.3, 0.3, 0.4])
    assert lo.tolist() == [0, 24, 28] and hi.tolist() == [24, 28, 32] and np.allclose(pr, [0.3, 0.3, 0.4])
    assert m.set_of_path([0, 23, 24, 27, 28, 31]).tolist() == [0, 0, 1, 1, 2, 2]
    with pytest.raises(ValueError):
        m.default_path_range([0.5, 0.5])
    with pytest.raises(ValueError):
        m.default_path_range([0, 0, 0])
    e = MapLibrary("cpm_entire")
    assert e.default_path_range([1, 0, 0]) == (0, e.n_paths) and e.path_sets()[2].tolist() == [1.0]


def test_fixed_duration_period_equals_the_references_float_test():
    """EnvConfig.fixed_period: the step period handed to the kernel fires exactly where the reference's
    ``(timer.step * dt) % reset_agent_fixed_duration == 0`` does (road_traffic.py:1388-1393, evaluated with torch the
    way the reference evaluates it: int32 tensor times python float), for every step an episode can reach."""
    import torch
    from sigmarl_b200 import EnvConfig, MapLibrary
    steps = torch.arange(700, dtype=torch.int32)
    for dt in (0.1, 0.05, 0.04, 0.2, 0.07):
        for dur in (1, 2, 3, 15, 0.5, 1.5):
            t = steps * dt
            want = ((t % dur == 0) & (t != 0)).numpy()
            period = EnvConfig(max_steps=700, reset_agent_fixed_duration=dur).fixed_period(dt)
            got = (np.arange(700) % period == 0) & (np.arange(700) != 0) if period else np.zeros(700, bool)
            assert np.array_equal(got, want), (dt, dur, period)
    m = MapLibrary("cpm_entire")
    assert EnvConfig(scenario_type="cpm_entire").lower(m).reset_fixed_period == 0
    assert EnvConfig(scenario_type="cpm_entire", reset_agent_fixed_duration=2).lower(m).reset_fixed_period == 20
    assert EnvConfig(scenario_type="cpm_entire", mode="kwargs", reset_agent_fixed_duration=15,
                     max_steps=600).lower(m).reset_fixed_period == 300      # evaluation_itsc25.py:63-73
    with pytest.raises(ValueError):
        EnvConfig(reset_agent_fixed_duration=-1).fixed_period(0.1)


def test_kernel_mtv_source_matches_reference_vectors_on_the_host():
    """sgb_debug_mtv_distance = the HOST compilation of the source function the MTV kernels use (sgb_kernels.cuh,
    mtv_from_vertices; every product / sum rounded on its own on both sides): bit-exact against the reference's
    get_distances_between_agents("mtv") on the 2 400 known-answer pairs — sign, exact zeros and symmetry included.
    Arithmetic self-test only; the kernel path itself is covered by the -m gpu tests."""
    from sigmarl_b200.lib import load_test_library as load_library
    L = load_library()
    g = np.load(os.path.join(REPO, "tests", "golden", "kat", "mtv.npz"))
    v, want = np.ascontiguousarray(g["mtv_vertices"][:, :, :4], np.float32), g["mtv_dist"]
    B, N = want.shape[:2]
    got = want.copy()
    for b in range(B):
        for i in range(N):
            for j in range(N):
                if i != j:
                    got[b, i, j] = L.sgb_debug_mtv_distance(v[b, i].ctypes.data, v[b, j].ctypes.data)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_every_reference_scenario_type_ships_and_packs():
    """All 18 scenario types of the reference's SCENARIOS table (constants.py:8-626) are shipped as parsed maps
    (oracle/gen_maps.py) and accepted by the library's map packer (host-only part of sgb_create): no degenerate
    polyline, at most 256 segments, and a blob that fits the shared memory of one SM next to the slot arrays."""
    import ctypes as C
    from sigmarl_b200.lib import load_test_library as load_library
    from sigmarl_b200.maps import MapLibrary, available_scenarios
    want = ["cpm_entire", "cpm_mixed", "interchange_1", "interchange_2", "interchange_3"] + \
           [f"intersection_{i}" for i in range(1, 9)] + \
           ["on_ramp_1", "on_ramp_2_multilane", "pseudo_distance_example", "roundabout_1", "roundabout_2"]
    assert available_scenarios() == sorted(want)
    L = load_library()
    blob_bytes = dict(cpm_entire=180368, cpm_mixed=34352, interchange_1=7472, interchange_2=31904, interchange_3=30624,
                      intersection_1=3296, intersection_2=5424, intersection_3=8096, intersection_4=10576,
                      intersection_5=13920, intersection_6=13776, intersection_7=10592, intersection_8=8688,
                      on_ramp_1=4896, on_ramp_2_multilane=29328, pseudo_distance_example=2112, roundabout_1=6736,
                      roundabout_2=36656)
    for st in want:
        m = MapLibrary(st)
        d, n = m.desc(), C.c_int64()
        assert L.sgb_debug_pack_map(C.byref(d), C.byref(n)) == 0, st
        assert n.value == blob_bytes[st], (st, n.value)        # layout pinned: a change here invalidates the hardware runs
        assert m.max_ref_path_points == int(m.n_center.max()) + 8
    assert L.sgb_debug_pack_map(None, None) != 0


def test_kernel_current_lanelet_matches_the_references_formula_on_the_host():
    """sgb_debug_current_lanelet = host build of the kernels' current_lanelet().  Checked against the reference's
    determine_current_lanelet (map_manager.py:39-89) restated with torch exactly as written there — centre lines padded
    with zeros to the longest one, torch.sum((a - c) ** 2), min over points, argmin over lanelets — on random positions
    and on the lanelets' own points (exact ties between lanelets that share an end point) of every OSM map."""
    import torch
    from sigmarl_b200.lib import load_test_library as load_library
    from sigmarl_b200.maps import MapLibrary, available_scenarios
    L = load_library()
    rng = np.random.default_rng(0)
    n_maps = n_ties = 0
    for st in available_scenarios():
        m = MapLibrary(st)
        if m.lanelet_xy is None:
            continue
        n_maps += 1
        off, xy = m.lanelet_off, m.lanelet_xy
        n = len(off) - 1
        max_len = int(np.diff(off).max())
        padded = torch.zeros(n, max_len, 2)
        for l in range(n):
            padded[l, :off[l + 1] - off[l]] = torch.from_numpy(xy[off[l]:off[l + 1]])
        pts = np.concatenate([rng.uniform([-0.5, -0.5], [m.world_x_dim + 0.5, m.world_y_dim + 0.5], (300, 2)),
                              xy[rng.integers(0, len(xy), 100)], np.zeros((1, 2))]).astype(np.float32)
        d = torch.sum((torch.from_numpy(pts)[:, None, None, :] - padded[None]) ** 2, dim=3)
        mind, _ = torch.min(d, dim=2)
        want = torch.argmin(mind, dim=1).numpy()
        n_ties += int((mind == mind.min(dim=1, keepdim=True).values).sum(dim=1).gt(1).sum())
        got = np.asarray([L.sgb_debug_current_lanelet(n, xy.ctypes.data, off.ctypes.data, float(x), float(y)) for x, y in pts])
        assert np.array_equal(got, want), (st, np.where(got != want)[0][:5])
        assert m.lanelet_adj.shape == (n, n) and m.lanelet_adj.diagonal()[:n - 4].all()   # (a lanelet is its own neighbour)
    assert n_maps == 16 and n_ties > 50


def test_packed_blob_certificates_hold_on_every_map():
    """The pruned search is exact only if what the packer stores is conservative: every chunk's box must contain the
    chunk's points, every boundary chunk's direction cone (fp16 mid angle / half width, mod pi) must contain the
    directions of its segments, and the points must be the map's polylines (+ the 6 extension points of a centre line,
    world_state_rt.py:279-311).  Checked on the host for all 18 maps from a copy of the blob sgb_create would upload."""
    import ctypes as C
    from sigmarl_b200.lib import load_test_library as load_library
    from sigmarl_b200.maps import MapLibrary, available_scenarios
    L = load_library()
    K, EXT = 8, 6                                   # kChunk, kExt (sgb_kernels.cuh)
    n_cones = n_wide = 0
    for st in available_scenarios():
        m = MapLibrary(st)
        d, n = m.desc(), C.c_int64()
        assert L.sgb_debug_pack_map(C.byref(d), C.byref(n)) == 0
        raw = np.zeros(n.value, np.uint8)
        assert L.sgb_debug_pack_map_blob(C.byref(d), raw.ctypes.data, n.value) == 0
        hdr = raw[:32].view(np.int32)
        n_paths, path_off, pts_off, box_off, total, cone_off = (int(x) for x in hdr[:6])
        assert n_paths == m.n_paths and total == n.value
        recs = raw[path_off:path_off + 48 * n_paths].view(np.int32).reshape(n_paths, 12)
        pts = raw[pts_off:box_off].view(np.float32).reshape(-1, 2)
        boxes = raw[box_off:cone_off].view(np.float32).reshape(-1, 4)
        cones = raw[cone_off:total].view(np.float16).reshape(-1, 2).astype(np.float64)
        for i in range(n_paths):
            c_off, n_c, l_off, n_l, r_off, n_r, cbox, lbox, rbox, is_loop, lcone, rcone = (int(x) for x in recs[i])
            cen = m.center_xy[m.center_off[i]:m.center_off[i + 1]]
            assert n_c == len(cen) and np.array_equal(pts[c_off:c_off + n_c], cen) and is_loop == int(m.is_loop[i])
            step = cen[-1] - cen[-2]
            ext = cen[-1] + np.arange(1, EXT + 1, dtype=np.float32)[:, None] * step
            assert np.array_equal(pts[c_off + n_c:c_off + n_c + EXT], ext.astype(np.float32))
            for off, cnt, box0, cone0, want in ((c_off, n_c, cbox, None, cen),
                                                (l_off, n_l, lbox, lcone, m.left_xy[m.left_off[i]:m.left_off[i + 1]]),
                                                (r_off, n_r, rbox, rcone, m.right_xy[m.right_off[i]:m.right_off[i + 1]])):
                assert cnt == len(want) and np.array_equal(pts[off:off + cnt], want)
                for c, s0 in enumerate(range(0, cnt - 1, K)):
                    s1 = min(s0 + K, cnt - 1)
                    seg = want[s0:s1 + 1].astype(np.float64)
                    bcx, bcy, bhx, bhy = (float(v) for v in boxes[box0 + c])      # centre + (padded) half extents
                    x0, y0, x1, y1 = bcx - bhx, bcy - bhy, bcx + bhx, bcy + bhy
                    tol = 4e-7 * max(1.0, float(np.abs(seg).max()))               # a few fp32 ulps of a coordinate
                    assert bhx - 0.5 * np.ptp(seg[:, 0]) < tol and bhy - 0.5 * np.ptp(seg[:, 1]) < tol   # and tight
                    assert (seg[:, 0] >= x0).all() and (seg[:, 0] <= x1).all() and (seg[:, 1] >= y0).all() and (seg[:, 1] <= y1).all()
                    if cone0 is None:
                        continue
                    mid, half = cones[cone0 + c]
                    n_cones += 1
                    if half >= np.pi / 2:
                        n_wide += 1
                        continue
                    dv = np.diff(seg, axis=0)
                    ang = np.arctan2(dv[:, 1], dv[:, 0])
                    rel = np.mod(ang - mid + np.pi / 2, np.pi) - np.pi / 2          # difference of undirected lines
                    assert np.abs(rel).max() <= half - np.arcsin(0.01) + 1e-3, (st, i, c, np.abs(rel).max(), half)
    assert n_cones > 1500 and n_wide < 0.05 * n_cones


def _scan_batch(L, m, path, pos, psi, hint, exhaustive):
    import ctypes as C
    d = m.desc()
    xs, ys = np.ascontiguousarray(pos[:, 0]), np.ascontiguousarray(pos[:, 1])
    out = np.zeros((len(path), 16), np.float32)
    rc = L.sgb_debug_scan_batch(C.byref(d), len(path), path.ctypes.data, xs.ctypes.data, ys.ctypes.data, psi.ctypes.data,
                                hint.ctypes.data, C.c_float(0.11), C.c_float(0.0535), int(exhaustive), out.ctypes.data)
    assert rc == 0
    return out


def _scan_poses(m, n, rng):
    """Poses around the map's centre lines: lateral / longitudinal noise of 6 cm (half of them touch a boundary),
    headings along the path +- 0.5 rad (a tenth exactly along it: collinear edges), hints from exact to random."""
    path = rng.integers(0, m.n_paths, n).astype(np.int32)
    nc = m.n_center[path]
    k = (rng.random(n) * (nc - 1)).astype(np.int64)
    yaw_off = np.concatenate([[0], np.cumsum(m.n_center - 1)])
    yaw = m.center_yaw[np.minimum(yaw_off[path] + k, yaw_off[path + 1] - 1)]
    pos = (m.center_xy[m.center_off[path] + k] + rng.normal(0, 0.06, (n, 2))).astype(np.float32)
    psi = (yaw + rng.normal(0, 0.5, n)).astype(np.float32)
    psi[:n // 10] = yaw[:n // 10]
    hint = np.where(rng.random(n) < 0.7, k + 1 + rng.integers(-3, 4, n), rng.integers(0, nc)).astype(np.int32)
    return path, pos, psi, hint


def test_pruned_scans_equal_exhaustive_scans_on_the_host_for_every_map(oracle_mod):
    """sgb_debug_scan_batch runs the kernels' OWN scan_center / scan_boundary source (host build, one lane per agent)
    through the phase-B glue on a freshly packed blob.  (1) On all 18 maps the pruned search — hint chunk, box votes,
    direction cones / bands, crossing gate — returns bit for bit what the exhaustive one returns: closest index,
    centre / vertex distances, lane-crossing flags (about half of the poses cross a boundary).  (2) The result is the
    reference's: closest index and crossing flags equal the oracle's get_perpendicular_distances / interX restatement
    exactly, the centre-line distance bit for bit, the boundary distances (reciprocal form) to 3e-6."""
    import ctypes as C
    from sigmarl_b200.lib import load_test_library as load_library
    from sigmarl_b200.maps import MapLibrary, available_scenarios
    L, O = load_library(), oracle_mod
    ol = O.lib()
    rng = np.random.default_rng(0)
    n_hits = n_checked = 0
    for st in available_scenarios():
        m = MapLibrary(st)
        path, pos, psi, hint = _scan_poses(m, 1500, rng)
        pruned, full = (_scan_batch(L, m, path, pos, psi, hint, ex) for ex in (0, 1))
        assert np.array_equal(pruned.view(np.uint32), full.view(np.uint32)), st
        # the product mode (no debug buffer: only the minimum over the four vertices is kept exact, which lets the
        # chunk vote use the best vertex instead of the worst): everything downstream consumes equals the exhaustive scan
        lean = _scan_batch(L, m, path, pos, psi, hint, 2)
        want = full.copy()
        for o in (3, 10):
            want[:, o:o + 4] = full[:, o:o + 4].min(1, keepdims=True)
        assert np.array_equal(lean.view(np.uint32), want.view(np.uint32)), st
        n_hits += int(full[:, 7].sum() + full[:, 14].sum())
        pm = O.PaddedMap(st)
        if pm.n_paths != m.n_paths:
            continue                                    # cpm_mixed: the oracle numbers the three path sets separately
        for i in range(0, 1500, 3):
            p = int(path[i])
            idx = C.c_int()
            pt = np.ascontiguousarray(pos[i])
            d = ol.orc_test_perp(pt.ctypes.data, pm.center[p].ctypes.data, pm.P, int(pm.n_center[p]), C.byref(idx))
            assert idx.value == int(full[i, 1]) and d == full[i, 0], (st, i)      # the pinned centre scan: bit-identical
            rect = np.zeros((5, 2), np.float32)
            ol.orc_test_rect(C.c_float(0.11), C.c_float(0.0535), pt.ctypes.data, C.c_float(float(psi[i])), rect.ctypes.data)
            for side, poly, cnt in ((0, pm.left, pm.n_left), (1, pm.right, pm.n_right)):
                hit = ol.orc_test_interx(rect.ctypes.data, 5, poly[p].ctypes.data, pm.P)
                assert bool(hit) == bool(full[i, 7 + 7 * side]), (st, i, side)
                dcg = ol.orc_test_perp(pt.ctypes.data, poly[p].ctypes.data, pm.P, int(cnt[p]), C.byref(idx))
                assert abs(dcg - full[i, 2 + 7 * side]) <= 3e-6, (st, i, side)   # reciprocal form: a few ulps of a 4 m coordinate
            n_checked += 1
    assert n_hits > 10000 and n_checked > 8000


def test_kernel_helpers_match_the_references_known_answers_on_the_host():
    """Host builds of the kernels' small helper functions (sgb_debug_helper / sgb_debug_short_term) against the vectors the
    UNMODIFIED reference helpers produced (oracle/gen_kat.py -> tests/golden/kat/helpers.npz): angle_eliminate_two_pi
    (helper_scenario.py:1276-1289) and decreasing_fcn (:960-996) to 1e-6 (fmodf / division), the short-term reference
    path gather (:892-957, loop and open paths, closest point = last point) bit for bit; and kth_nearest against
    torch.topk(largest=False) incl. ties."""
    import ctypes as C
    import torch
    from sigmarl_b200.lib import load_test_library as load_library
    L = load_library()
    k = np.load(os.path.join(REPO, "tests", "golden", "kat", "helpers.npz"))
    out = np.zeros(8, np.float32)
    for x, want in zip(k["wrap_in"], k["wrap_out"]):
        assert L.sgb_debug_helper(0, np.float32([x]).ctypes.data, 1, out.ctypes.data) == 0
        assert abs(out[0] - want) <= 1e-6 or abs(abs(out[0] - want) - 2 * np.pi) <= 1e-6, (x, out[0], want)
    for x, want in zip(k["dec_x"], k["dec_lin_0_03"]):
        assert L.sgb_debug_helper(1, np.float32([x, 0.0, 0.3]).ctypes.data, 3, out.ctypes.data) == 0
        assert abs(out[0] - want) <= 1e-6, (x, out[0], want)
    poly, n_c, loop, idx0 = k["st_poly"], k["st_n"], k["st_loop"], k["st_idx0"]
    for i in range(poly.shape[0]):
        p = np.ascontiguousarray(poly[i], np.float32)
        assert L.sgb_debug_short_term(p.ctypes.data, int(n_c[i]), int(loop[i]), int(idx0[i]), out.ctypes.data) == 0
        assert np.array_equal(out[:6].reshape(3, 2), k["st_pts"][i]), i
    rng = np.random.default_rng(0)
    for _ in range(300):
        n = int(rng.integers(2, 33))
        d = rng.random(n).astype(np.float32)
        if n > 3:
            d[rng.integers(0, n)] = d[rng.integers(0, n)]          # exact ties
        kk = int(rng.integers(0, n))
        vals, idx = torch.topk(torch.from_numpy(d), k=kk + 1, largest=False)
        assert L.sgb_debug_helper(2, np.concatenate([[np.float32(kk)], d]).astype(np.float32).ctypes.data, n + 1, out.ctypes.data) == 0
        assert out[1] == float(vals[kk]), (d, kk)
        # torch.topk does not promise an order among equal values; the kernel (like a stable sort) takes the lower index
        assert int(out[0]) == int(np.argsort(d, kind="stable")[kk]), (d, kk)


def test_rectangle_pair_crossing_on_the_host(oracle_mod):
    """Host build of the kernels' rect_cross_rect == the oracle's interX restatement on rectangle pairs from touching to
    far apart (incl. identical / parallel / collinear ones), and the pair gate's certificate: a pair the gate skips
    (far and not collinear) never crosses."""
    import ctypes as C
    from sigmarl_b200.lib import load_test_library as load_library
    L, ol = load_library(), oracle_mod.lib()
    rng = np.random.default_rng(3)
    n = 60000
    lo = np.zeros((n, 3), np.float32)
    lo[:, :2] = rng.uniform(0, 4, (n, 2)); lo[:, 2] = rng.uniform(-7, 7, n)
    hi = lo.copy()
    hi[:, :2] += rng.normal(0, 1, (n, 2)).astype(np.float32) * rng.choice([0.05, 0.15, 0.4, 2.0], (n, 1)).astype(np.float32)
    hi[:, 2] = np.where(rng.random(n) < 0.3, lo[:, 2] + rng.integers(0, 4, n) * np.float32(np.pi / 2), rng.uniform(-7, 7, n))
    hi[:50] = lo[:50]                                             # identical rectangles
    out = np.zeros(n, np.uint8)
    assert L.sgb_debug_pair_batch(n, lo.ctypes.data, hi.ctypes.data, C.c_float(0.11), C.c_float(0.0535), out.ctypes.data) == 0
    cross, skip = (out & 1) != 0, (out & 2) != 0
    assert not (cross & skip).any()
    assert 0.1 < cross.mean() < 0.9 and 0.1 < skip.mean() < 0.9
    ra, rb = np.zeros((5, 2), np.float32), np.zeros((5, 2), np.float32)
    for i in range(0, n, 20):
        pa, pb = np.ascontiguousarray(lo[i, :2]), np.ascontiguousarray(hi[i, :2])     # (named: the pointers must stay valid)
        ol.orc_test_rect(C.c_float(0.11), C.c_float(0.0535), pa.ctypes.data, C.c_float(float(lo[i, 2])), ra.ctypes.data)
        ol.orc_test_rect(C.c_float(0.11), C.c_float(0.0535), pb.ctypes.data, C.c_float(float(hi[i, 2])), rb.ctypes.data)
        assert bool(ol.orc_test_interx(ra.ctypes.data, 5, rb.ctypes.data, 5)) == bool(cross[i]), i
