"""ctypes binding of ``libsigmarl_b200.so`` (C-ABI in ``include/sigmarl_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C sigmarl_b200/csrc``.  If it is
missing, or a call fails, this module raises — nothing here falls back to another implementation.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libsigmarl_b200.so"

SGB_MAX_AGENTS = 32
SGB_FLAG_COLLIDE_AGENT, SGB_FLAG_COLLIDE_LANE, SGB_FLAG_ENTRY, SGB_FLAG_EXIT = 1, 2, 4, 8
SGB_REW_EXACT_SPARSE, SGB_REW_TTC, SGB_REW_DISTANCE, SGB_REW_SPARSE = 1, 2, 4, 8
(SGB_OBS_BIRD_VIEW, SGB_OBS_CENTRES, SGB_OBS_STEERING, SGB_OBS_REF_OTHERS, SGB_OBS_NO_DIST_AGENTS,
 SGB_OBS_NO_DIST_CENTER, SGB_OBS_BOUNDARY_POINTS, SGB_OBS_APPLY_MASK, SGB_OBS_MASK_LANELETS) = 1, 2, 4, 8, 16, 32, 64, 128, 256
CARRY_IDX_MASK, CARRY_FRESH_BIT = 0x3fffffff, 0x40000000   # carry.w with SGB_OBS_BOUNDARY_POINTS (see the header)

# every symbol include/sigmarl_b200.h declares (tests/test_abi_and_host.py checks the two lists agree)
EXPORTS = ["sgb_create", "sgb_destroy", "sgb_obs_dim", "sgb_max_ref_path_points", "sgb_step", "sgb_refresh",
           "sgb_place", "sgb_reset", "sgb_reset_all", "sgb_reset_masked", "sgb_step_host", "sgb_gae", "sgb_launch_count", "sgb_map_bytes",
           "sgb_status_string", "sgb_last_error", "sgb_version", "sgb_set_lanelets", "sgb_set_env_offset", "sgb_step_reset_host",
           "sgb_set_path_sets", "sgb_gae_allgather"]
# ... and include/sigmarl_b200_test.h: host-side self-test hooks, present only in libsigmarl_b200_test.so (test suite)
TEST_EXPORTS = ["sgb_debug_mtv_distance", "sgb_debug_pack_map", "sgb_debug_current_lanelet", "sgb_debug_pack_map_blob",
                "sgb_debug_scan_batch", "sgb_debug_scan_counters", "sgb_debug_helper", "sgb_debug_short_term",
                "sgb_debug_pair_batch"]

class SgbError(RuntimeError):
    pass


class MapDesc(C.Structure):
    _fields_ = [("n_paths", C.c_int32),
                ("center_xy", C.c_void_p), ("center_off", C.c_void_p),
                ("left_xy", C.c_void_p), ("left_off", C.c_void_p),
                ("right_xy", C.c_void_p), ("right_off", C.c_void_p),
                ("center_yaw", C.c_void_p), ("is_loop", C.c_void_p)]


CONFIG_FLOATS = ["dt", "max_speed", "max_steering", "max_acc", "max_steering_rate", "l_wb", "lr_over_lwb",
                 "half_length", "half_width", "diag", "w_ref0", "w_ref1", "w_ref2", "speed_dt", "reward_progress",
                 "near_boundary_low", "near_boundary_high", "near_agents_low", "near_agents_high", "ttc_low",
                 "ttc_high", "penalty_near_boundary", "penalty_near_agents", "penalty_collide_agents",
                 "penalty_collide_lane", "norm_pos", "norm_v", "norm_rot", "norm_dist", "dsafe_sq",
                 "reset_min_dist_sq"]


class Config(C.Structure):
    _fields_ = ([(n, C.c_float) for n in CONFIG_FLOATS] +
                [("rew_flags", C.c_uint32), ("k_near", C.c_int32), ("max_steps", C.c_int32),
                 ("respawn_on_exit", C.c_int32), ("exhaustive", C.c_int32), ("reward_reach_goal", C.c_float),
                 ("testing_mode", C.c_int32), ("obs_flags", C.c_uint32), ("norm_pos_world_x", C.c_float),
                 ("norm_pos_world_y", C.c_float), ("norm_dist_agent", C.c_float), ("obs_noise_level", C.c_float),
                 ("obs_noise_seed", C.c_uint32), ("reset_fixed_period", C.c_int32), ("use_mtv_distance", C.c_uint32), ("mask_distance", C.c_float), ("reserved0", C.c_uint32), ("reserved1", C.c_uint32), ("reserved2", C.c_uint32)])


BUFFER_FIELDS = ["pose", "aux", "path_id", "carry", "action", "step_count", "obs", "reward", "done",
                 "agent_flags", "collide_with", "info", "task_tries", "task_success", "dbg", "scenario_id", "nan_flags"]
SGB_INFO_DIM = 16


class Buffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in BUFFER_FIELDS]


_lib = None


def library_path():
    return os.environ.get("SGB_LIBRARY", os.path.join(HERE, _LIB_NAME))


def load_library():
    """Load libsigmarl_b200.so and declare prototypes.  Raises SgbError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise SgbError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       f"or `make -C sigmarl_b200/csrc` (there is no CPU fallback)")
    L = C.CDLL(path)
    vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
    L.sgb_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(MapDesc), C.POINTER(Config)]
    L.sgb_destroy.argtypes = [vp]
    L.sgb_set_env_offset.argtypes = [vp, i64]
    L.sgb_set_path_sets.argtypes = [vp, i32, vp, vp, vp]
    L.sgb_obs_dim.argtypes = [vp]
    L.sgb_max_ref_path_points.argtypes = [vp]
    L.sgb_step.argtypes = [vp, i32, i32, C.POINTER(Buffers), vp]
    L.sgb_refresh.argtypes = [vp, i32, i32, C.POINTER(Buffers), vp, i32, vp]
    L.sgb_place.argtypes = [vp, i32, i32, C.POINTER(Buffers), vp, vp, vp, vp, vp]
    L.sgb_reset.argtypes = [vp, i32, i32, C.POINTER(Buffers), i32, i32, u64, u64, i64, i32, i32, vp, vp]
    L.sgb_reset_all.argtypes = [vp, i32, i32, C.POINTER(Buffers), i32, i32, u64, u64, i64, i32, vp, vp]
    L.sgb_reset_masked.argtypes = [vp, i32, i32, C.POINTER(Buffers), vp, vp, i32, i32, u64, u64, i64, i32, i32, vp, vp]
    L.sgb_step_host.argtypes = [vp, i32, i32, C.POINTER(Buffers), vp, vp, vp, vp, vp]
    L.sgb_step_reset_host.argtypes = [vp, i32, i32, C.POINTER(Buffers), vp, vp, vp, vp, i32, i32, u64, u64, i64, i32, vp, vp]
    L.sgb_gae.argtypes = [i32, i32, i32, vp, vp, vp, vp, C.c_float, C.c_float, vp, vp, vp]
    L.sgb_gae_allgather.argtypes = [i32, i32, i32, vp, vp, vp, vp, C.c_float, C.c_float, i32, i32, vp, vp, vp, vp, vp]
    L.sgb_launch_count.argtypes = [vp]
    L.sgb_launch_count.restype = i64
    L.sgb_map_bytes.argtypes = [vp]
    L.sgb_map_bytes.restype = i64
    L.sgb_status_string.argtypes = [C.c_int]
    L.sgb_status_string.restype = C.c_char_p
    L.sgb_last_error.restype = C.c_char_p
    L.sgb_version.restype = C.c_int
    L.sgb_set_lanelets.argtypes = [vp, i32, vp, vp, vp]
    _lib = L
    return L


_test_lib = None


def load_test_library():
    """TEST INFRASTRUCTURE: libsigmarl_b200_test.so = the same sources compiled with the host-side self-test hooks
    (include/sigmarl_b200_test.h; `make -C sigmarl_b200/csrc test`).  Nothing in this package calls it."""
    global _test_lib
    if _test_lib is not None:
        return _test_lib
    path = os.environ.get("SGB_TEST_LIBRARY", os.path.join(HERE, "libsigmarl_b200_test.so"))
    if not os.path.exists(path):
        raise SgbError(f"{path} not found: build it with `make -C sigmarl_b200/csrc test`")
    L = C.CDLL(path)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.sgb_version.restype = C.c_int
    L.sgb_debug_current_lanelet.argtypes = [i32, vp, vp, C.c_float, C.c_float]
    L.sgb_debug_scan_batch.argtypes = [C.POINTER(MapDesc), i32, vp, vp, vp, vp, vp, C.c_float, C.c_float, i32, vp]
    L.sgb_debug_pair_batch.argtypes = [i32, vp, vp, C.c_float, C.c_float, vp]
    L.sgb_debug_helper.argtypes = [i32, vp, i32, vp]
    L.sgb_debug_short_term.argtypes = [vp, i32, i32, i32, vp]
    L.sgb_debug_scan_counters.argtypes = [vp, i32]
    L.sgb_debug_scan_counters.restype = None
    L.sgb_debug_pack_map_blob.argtypes = [C.POINTER(MapDesc), vp, i64]
    L.sgb_debug_pack_map.argtypes = [C.POINTER(MapDesc), C.POINTER(i64)]
    L.sgb_debug_mtv_distance.argtypes = [vp, vp]
    L.sgb_debug_mtv_distance.restype = C.c_float
    _test_lib = L
    return L


def check(rc, what):
    if rc != 0:
        L = load_library()
        raise SgbError(f"{what} failed: {L.sgb_status_string(rc).decode()} ({rc}) {L.sgb_last_error().decode()}")
