"""Map library: the reference's parsed maps as flat polylines (one shared, read-only copy).

The ``.npz`` files in ``sigmarl_b200/maps/`` are produced once, in the build container, by running the
reference's own parsers (``oracle/gen_maps.py``; ``map_manager.py:13-40``).  Parsing is a one-off host
job and is not re-implemented.  The reference copies a path's polylines into a
``[B, N, P, 2]`` slot per agent (``world_state_rt.py:155-175, 313-392``); here every agent carries an int
``path_id`` into this library instead.
"""
import ctypes as C
import os

import numpy as np

from .lib import MapDesc

MAP_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "maps")
N_POINTS_SHORT_TERM = 3       # road_traffic.py:273-275
SAMPLE_INTERVAL_REF_PATH = 2  # road_traffic.py:316


def available_scenarios():
    return sorted(f[:-4] for f in os.listdir(MAP_DIR) if f.endswith(".npz") and not f.endswith(".lanelets.npz"))


class MapLibrary:
    """All reference paths of one ``scenario_type`` with a global path numbering.

    ``cpm_mixed`` keeps three path sets (``world_state_rt_sim.py:313-358``: scenario_id 1 = intersection,
    2 = merge-in, 3 = merge-out); every other type has the single set ``all`` (scenario_id 0).
    ``set_range[name] = (lo, hi)`` is the global index range of a set.
    """

    def __init__(self, scenario_type: str):
        path = os.path.join(MAP_DIR, f"{scenario_type}.npz")
        if not os.path.exists(path):
            raise ValueError(f"unknown scenario_type {scenario_type!r}; have {available_scenarios()}")
        z = np.load(path)
        self.scenario_type = scenario_type
        self.world_x_dim = float(z["world_x_dim"])
        self.world_y_dim = float(z["world_y_dim"])
        self.lane_width = float(z["lane_width"])
        self.default_n_agents = int(z["default_n_agents"])
        self.set_names = ["intersection", "merge_in", "merge_out"] if "cpm_mixed" in scenario_type else ["all"]
        cen, lef, rig, yaw, loop, lids = [], [], [], [], [], []
        c_off, l_off, r_off = [0], [0], [0]
        self.set_range = {}
        for s in self.set_names:
            n = len(z[f"{s}_center_off"]) - 1
            lo = len(c_off) - 1
            for i in range(n):
                a, b = z[f"{s}_center_off"][i:i + 2]
                cen.append(z[f"{s}_center_xy"][a:b])
                c_off.append(c_off[-1] + (b - a))
                a, b = z[f"{s}_left_off"][i:i + 2]
                lef.append(z[f"{s}_left_xy"][a:b])
                l_off.append(l_off[-1] + (b - a))
                a, b = z[f"{s}_right_off"][i:i + 2]
                rig.append(z[f"{s}_right_xy"][a:b])
                r_off.append(r_off[-1] + (b - a))
                a, b = z[f"{s}_yaw_off"][i:i + 2]
                yaw.append(z[f"{s}_yaw"][a:b])
                a, b = z[f"{s}_lanelet_off"][i:i + 2]
                lids.append(z[f"{s}_lanelet_ids"][a:b])
            loop.append(z[f"{s}_is_loop"])
            self.set_range[s] = (lo, lo + n)
        self.center_xy = np.ascontiguousarray(np.concatenate(cen), np.float32)
        self.left_xy = np.ascontiguousarray(np.concatenate(lef), np.float32)
        self.right_xy = np.ascontiguousarray(np.concatenate(rig), np.float32)
        self.center_yaw = np.ascontiguousarray(np.concatenate(yaw), np.float32)
        self.center_off = np.asarray(c_off, np.int32)
        self.left_off = np.asarray(l_off, np.int32)
        self.right_off = np.asarray(r_off, np.int32)
        self.is_loop = np.ascontiguousarray(np.concatenate(loop), np.uint8)
        self.n_paths = len(c_off) - 1
        # ref_lanelet_ids as the reference holds them per agent: zero-padded to len(lanelets_all)
        # (world_state_rt.py:152, 242-246, 411-417)
        self.n_lanelets_all = int(z["n_lanelets_all"])
        self.lanelet_ids = np.zeros((self.n_paths, self.n_lanelets_all), np.int32)
        for i, ids in enumerate(lids):
            self.lanelet_ids[i, :len(ids)] = ids
        n_c = np.diff(self.center_off)
        # road_traffic.py:505-530
        self.max_ref_path_points = int(n_c.max()) + N_POINTS_SHORT_TERM * SAMPLE_INTERVAL_REF_PATH + 2
        self.has_loops = bool(self.is_loop.any())
        # lanelet table (lanelet-relation observation mask, map_manager.py:39-119): OSM maps only
        lp = os.path.join(MAP_DIR, f"{scenario_type}.lanelets.npz")
        self.lanelet_xy = self.lanelet_off = self.lanelet_adj = None
        if os.path.exists(lp):
            zl = np.load(lp)
            self.lanelet_xy = np.ascontiguousarray(zl["center_xy"], np.float32)
            self.lanelet_off = np.ascontiguousarray(zl["center_off"], np.int32)
            self.lanelet_adj = np.ascontiguousarray(zl["adjacency"], np.uint8)

    @property
    def n_center(self):
        return np.diff(self.center_off)

    def global_path(self, scenario_id, path_id):
        """Reference (scenario_id, path_id) -> global path index."""
        scenario_id = np.asarray(scenario_id)
        path_id = np.asarray(path_id)
        if self.set_names == ["all"]:
            return path_id.astype(np.int32)
        offs = np.asarray([0] + [self.set_range[s][0] for s in self.set_names], np.int32)
        return (offs[scenario_id] + path_id).astype(np.int32)

    def default_path_range(self, cpm_scenario_probabilities=(1.0, 0.0, 0.0)):
        """Path range a reset samples from, or ``None`` when every env draws its own path set (``path_sets``).
        ``cpm_mixed`` with probabilities (1,0,0) (``config.json:35``) draws from the intersection set."""
        if self.set_names == ["all"]:
            return self.set_range["all"]
        p = [float(x) for x in cpm_scenario_probabilities]
        if len(p) != len(self.set_names) or min(p) < 0 or sum(p) <= 0:
            raise ValueError(f"cpm_scenario_probabilities must be {len(self.set_names)} non-negative weights, got {p}")
        nz = [i for i, x in enumerate(p) if x > 0]
        if len(nz) == 1:
            return self.set_range[self.set_names[nz[0]]]
        return None

    def path_sets(self, cpm_scenario_probabilities=(1.0, 0.0, 0.0)):
        """(lo[], hi[], probability[]) of the map's path sets for ``sgb_set_path_sets``: the reference draws one set
        per env at every full reset (``world_state_rt_sim.py:313-358``)."""
        lo = np.asarray([self.set_range[s][0] for s in self.set_names], np.int32)
        hi = np.asarray([self.set_range[s][1] for s in self.set_names], np.int32)
        p = np.asarray(list(cpm_scenario_probabilities) if len(self.set_names) > 1 else [1.0], np.float32)
        return lo, hi, p

    def set_of_path(self, path_id):
        """Index of the path set a global path index belongs to (reference: scenario_id - 1 on cpm_mixed)."""
        path_id = np.asarray(path_id)
        out = np.zeros(path_id.shape, np.int32)
        for k, s in enumerate(self.set_names):
            lo, hi = self.set_range[s]
            out[(path_id >= lo) & (path_id < hi)] = k
        return out

    def desc(self):
        """ctypes ``sgb_map_desc`` over this object's numpy arrays (keep ``self`` alive while it is used)."""
        d = MapDesc()
        d.n_paths = self.n_paths
        for name in ("center_xy", "center_off", "left_xy", "left_off", "right_xy", "right_off", "center_yaw", "is_loop"):
            setattr(d, name, getattr(self, name).ctypes.data_as(C.c_void_p))
        return d
