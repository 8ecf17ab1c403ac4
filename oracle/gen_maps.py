#!/usr/bin/env python
"""TEST/BUILD INFRASTRUCTURE — compile the reference's parsed maps into flat arrays.

Runs the reference's own map parsers (``sigmarl/map_manager.py:13-40`` ->
``parse_xml.py`` / ``parse_osm.py``) ONCE, in this container, behind the import shim
(``oracle/refshim/install_shims.py``) and dumps every reference path of a scenario type
as flat float32 polylines into ``sigmarl_b200/maps/<scenario_type>.npz``.

Parsing is a one-off host job and explicitly not re-implemented;
the committed ``.npz`` files are what travels to the GPU box (``/root/reference`` does
not exist there).  Re-run:  ``python oracle/gen_maps.py``

Layout of one ``.npz`` (path sets: ``all`` = ``parser.reference_paths``,
``intersection`` / ``merge_in`` / ``merge_out`` = the ``cpm_mixed`` sub-scenarios,
``world_state_rt_sim.py:313-358``):
  <set>_center_xy [sum n_c, 2]  <set>_center_off [n_paths+1]   centre lines      (parse_xml.py:785-797)
  <set>_left_xy / _left_off,  <set>_right_xy / _right_off        *_boundary_shared
  <set>_yaw [sum (n_c-1)]   <set>_yaw_off [n_paths+1]            center_line_yaw
  <set>_is_loop [n_paths] uint8
  meta: world_x_dim, world_y_dim (float64), lane_width (SCENARIOS[...]["lane_width"]),
        default_n_agents, osm_lane_width (the Parameters.lane_width the OSM polylines were built with)
"""
import os
import sys

os.environ["CICD_TESTING"] = "true"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "refshim"))
import install_shims  # noqa: E402,F401

import numpy as np  # noqa: E402
import torch  # noqa: E402
from sigmarl.constants import SCENARIOS  # noqa: E402
from sigmarl.map_manager import MapManager  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "sigmarl_b200", "maps")
SCENARIO_TYPES = list(SCENARIOS)   # every scenario_type of constants.py:8-626 (17 maps + pseudo_distance_example)
OSM_LANE_WIDTH = 0.25  # Parameters.lane_width default (helper_common.py:119)


def flatten(paths, key):
    off = [0]
    chunks = []
    for p in paths:
        a = p[key].detach().cpu().numpy().astype(np.float32)
        chunks.append(a)
        off.append(off[-1] + a.shape[0])
    if chunks:
        flat = np.concatenate(chunks, axis=0)
    else:
        flat = np.zeros((0, 2), np.float32)
    return flat, np.asarray(off, np.int32)


def main():
    os.makedirs(OUT, exist_ok=True)
    for st in SCENARIO_TYPES:
        m = MapManager(scenario_type=st, device="cpu", lane_width=OSM_LANE_WIDTH)
        p = m.parser
        sets = {
            "all": p.reference_paths,
            "intersection": p.reference_paths_intersection,
            "merge_in": p.reference_paths_merge_in,
            "merge_out": p.reference_paths_merge_out,
        }
        out = {}
        for name, paths in sets.items():
            out[f"{name}_center_xy"], out[f"{name}_center_off"] = flatten(paths, "center_line")
            out[f"{name}_left_xy"], out[f"{name}_left_off"] = flatten(paths, "left_boundary_shared")
            out[f"{name}_right_xy"], out[f"{name}_right_off"] = flatten(paths, "right_boundary_shared")
            yaw, yoff = flatten([{"y": q["center_line_yaw"].reshape(-1, 1)} for q in paths], "y")
            out[f"{name}_yaw"], out[f"{name}_yaw_off"] = yaw.reshape(-1), yoff
            out[f"{name}_is_loop"] = np.asarray([bool(q["is_loop"]) for q in paths], np.uint8)
            # lanelet IDs of every path (info()["ref_lanelet_ids"], world_state_rt.py:411-417)
            ids = [np.asarray(q["lanelet_IDs"], np.int32).reshape(-1) for q in paths]
            out[f"{name}_lanelet_ids"] = np.concatenate(ids) if ids else np.zeros(0, np.int32)
            out[f"{name}_lanelet_off"] = np.asarray(np.concatenate([[0], np.cumsum([len(a) for a in ids])]), np.int32)
        out["world_x_dim"] = np.float64(p.bounds["world_x_dim"])
        out["world_y_dim"] = np.float64(p.bounds["world_y_dim"])
        out["lane_width"] = np.float64(SCENARIOS[st]["lane_width"])
        out["default_n_agents"] = np.int32(SCENARIOS[st]["n_agents"])
        out["osm_lane_width"] = np.float64(OSM_LANE_WIDTH)
        out["n_lanelets_all"] = np.int32(len(p.lanelets_all))   # width of ref_lanelet_ids (world_state_rt.py:152)
        np.savez_compressed(os.path.join(OUT, f"{st}.npz"), **out)
        # lanelet table for the lanelet-relation observation mask (map_manager.py:39-119; only the OSM parser knows
        # neighbouring lanelets, parse_osm.py:257-262): every lanelet's centre line + adjacency matrix, kept in a file of
        # its own so that the polyline files above stay byte-stable
        nb = p.neighboring_lanelets_idx
        if len(nb):
            cl = [np.asarray(l["center_line"].detach().cpu().numpy(), np.float32) for l in p.lanelets_all]
            adj = np.zeros((len(cl), len(cl)), np.uint8)
            for i, lst in enumerate(nb):
                for j in lst:
                    if 0 <= j < len(cl):
                        adj[i, j] = 1
            np.savez_compressed(os.path.join(OUT, f"{st}.lanelets.npz"), center_xy=np.concatenate(cl),
                                center_off=np.asarray(np.concatenate([[0], np.cumsum([len(a) for a in cl])]), np.int32),
                                adjacency=adj)
        nmax = int(np.diff(out["all_center_off"]).max())
        print(f"{st}: {len(p.reference_paths)} paths, max centre pts {nmax}, loops {int(out['all_is_loop'].sum())}")


if __name__ == "__main__":
    if sys.argv[1:]:
        SCENARIO_TYPES = sys.argv[1:]
    main()
