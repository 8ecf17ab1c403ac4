#!/usr/bin/env python
"""TEST INFRASTRUCTURE — function-level known-answer vectors from the UNMODIFIED reference helpers.

Calls the reference's own geometry primitives (``sigmarl/helper_scenario.py``, imported from
``/root/reference`` behind ``oracle/refshim/install_shims.py``) on seeded random inputs plus the edge cases
the environment step meets (a point exactly on a polyline vertex, touching / identical / collinear
rectangles, angles at +-pi, short-term indices at the end of loop and open paths) and stores inputs and
outputs in ``tests/golden/kat/helpers.npz``.  ``tests/test_oracle_kat.py`` replays them through the C
oracle's primitives ("function-level known-answer tests against the reference functions").

Re-run (here only):  ``python oracle/gen_kat.py``
"""
import os
import sys

os.environ["CICD_TESTING"] = "true"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "refshim"))
import install_shims  # noqa: E402,F401

import numpy as np  # noqa: E402
import torch  # noqa: E402
from sigmarl import helper_scenario as H  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "kat", "helpers.npz")
L, W = 0.22, 0.107  # constants.py:630-631


def main():
    torch.manual_seed(1234)
    out = {}

    # ---- get_perpendicular_distances (helper_scenario.py:829-889): smooth random polylines padded to P points
    K, P = 96, 40
    t = torch.linspace(0, 1, P).unsqueeze(0)
    ph = torch.rand(K, 1) * 6.28
    poly = torch.stack([2.0 * t + 0.3 * torch.sin(4 * t + ph), 0.8 * torch.cos(3 * t + ph) + 0.2 * t], dim=-1)  # [K,P,2]
    n_pts = torch.randint(8, P - 3, (K,))
    for k in range(K):      # tail padding as world_state_rt.py:313-392 (repeat the last real point)
        poly[k, n_pts[k]:] = poly[k, n_pts[k] - 1]
    pts = poly[torch.arange(K), torch.randint(0, 8, (K,))] + 0.15 * torch.randn(K, 2)
    pts[:16] = poly[torch.arange(16), torch.randint(1, 7, (16,))]        # exactly on a vertex: two segments tie at 0
    pts[16:24] = poly[torch.arange(16, 24), n_pts[16:24] - 1] + 0.05     # beyond the last real point (fix-up :877-879)
    d, idx = H.get_perpendicular_distances(point=pts.clone(), polyline=poly.clone(), n_points_long_term=n_pts.clone())
    out.update(perp_poly=poly, perp_n=n_pts, perp_point=pts, perp_dist=d, perp_idx=idx)

    # ---- get_rectangle_vertices (:695-826)
    c = torch.randn(64, 2) * 2
    yaw = (torch.rand(64, 1) * 2 - 1) * 7.0
    yaw[:4, 0] = torch.tensor([0.0, torch.pi / 2, -torch.pi, 3 * torch.pi])
    v = H.get_rectangle_vertices(center=c, yaw=yaw, width=W, length=L, is_close_shape=True)
    out.update(rect_center=c, rect_yaw=yaw.squeeze(-1), rect_vertices=v)

    # ---- interX (:1148-1229): rectangle vs rectangle and rectangle vs polyline
    n = 256
    ca = torch.randn(n, 2) * 0.2
    cb = ca + torch.randn(n, 2) * 0.18
    ya, yb = torch.rand(n, 1) * 6.28, torch.rand(n, 1) * 6.28
    cb[:8], yb[:8] = ca[:8], ya[:8]                                       # identical rectangles
    cb[8:16] = ca[8:16] + torch.tensor([L, 0.0]); ya[8:16] = 0.0; yb[8:16] = 0.0      # edge-touching, collinear sides
    cb[16:24] = ca[16:24] + torch.tensor([L, W]); ya[16:24] = 0.0; yb[16:24] = 0.0    # corner-touching
    cb[24:32] = ca[24:32] + torch.tensor([0.0, W / 2]); ya[24:32] = 0.0; yb[24:32] = 0.0  # overlapping, parallel
    ra = H.get_rectangle_vertices(center=ca, yaw=ya, width=W, length=L, is_close_shape=True)
    rb = H.get_rectangle_vertices(center=cb, yaw=yb, width=W, length=L, is_close_shape=True)
    out.update(ix_a=ra, ix_b=rb, ix_rr=H.interX(ra.clone(), rb.clone(), False))
    cp = poly
    cr = cp[torch.arange(K), torch.randint(0, 8, (K,))] + 0.04 * torch.randn(K, 2)
    rr = H.get_rectangle_vertices(center=cr, yaw=torch.rand(K, 1) * 6.28, width=W, length=L, is_close_shape=True)
    out.update(ix_rect=rr, ix_poly=cp, ix_rp=H.interX(rr.clone(), cp.clone(), False))

    # ---- angle_eliminate_two_pi (:1276-1289)
    a = torch.cat([torch.randn(200) * 8, torch.tensor([0.0, torch.pi, -torch.pi, 2 * torch.pi, -2 * torch.pi,
                                                        3 * torch.pi, 1e-7, -1e-7, 6.2831855, 3.1415927, 3.1415925])])
    out.update(wrap_in=a.clone(), wrap_out=H.angle_eliminate_two_pi(a.clone()))

    # ---- transform_from_global_to_local_coordinate (:1241-1273), batched form
    pi_, pj = torch.randn(64, 2), torch.randn(64, 5, 2)
    ri = (torch.rand(64, 1) * 2 - 1) * 4
    out.update(loc_pi=pi_, loc_pj=pj, loc_rot=ri.squeeze(-1),
               loc_out=H.transform_from_global_to_local_coordinate(pos_i=pi_, pos_j=pj, rot_i=ri))

    # ---- get_short_term_reference_path (:892-957): loop / open, both parameterisations used by the scenario
    n_c = torch.randint(12, P - 8, (K,))
    is_loop = torch.rand(K) < 0.5
    idx0 = torch.stack([torch.randint(1, int(m), ()) for m in n_c])
    idx0[:12] = n_c[:12] - 1                                              # closest point = last point
    for name, kw in (("st", dict(n_points_to_return=3, sample_interval=2, n_points_shift=1)),
                     ("nb", dict(n_points_to_return=5, sample_interval=1, n_points_shift=-2)),
                     ("nbr", dict(n_points_to_return=5, sample_interval=1, n_points_shift=1))):
        sp, fi = H.get_short_term_reference_path(polyline=poly.clone(), index_closest_point=idx0.clone(),
                                                 is_polyline_a_loop=is_loop.clone(), n_points_long_term=n_c.clone(), **kw)
        out[name + "_pts"], out[name + "_idx"] = sp, fi
    out.update(st_poly=poly, st_n=n_c, st_loop=is_loop, st_idx0=idx0)

    # ---- structural near-ties: an agent that sits exactly on a centre point (every reset pose, world_state_rt_sim.py:
    #      215-311) against its lane boundaries — the foot of the perpendicular is next to a boundary vertex, two
    #      segments are within an ulp of each other and torch.norm's rounding (fma) decides the argmin
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle.oracle import PaddedMap  # noqa: E402  (map layout only: the reference's padded [P,2] polylines)
    for st in ("cpm_mixed", "on_ramp_2_multilane"):
        pm = PaddedMap(st)
        rows = []
        for p in range(pm.n_paths):
            for k in range(3, int(pm.n_center[p]) // 2):
                pt = torch.tensor(pm.center[p, k])
                for side, arr, cnt in ((0, pm.left, pm.n_left), (1, pm.right, pm.n_right)):
                    _, i1 = H.get_perpendicular_distances(point=pt.clone(), polyline=torch.tensor(arr[p]),
                                                          n_points_long_term=torch.tensor(int(cnt[p])))
                    rows.append((p, k, side, int(i1[0])))
        out["spawn_idx_" + st] = np.asarray(rows, np.int32)

    # ---- decreasing_fcn (:960-996), linear
    x = torch.cat([torch.rand(100) * 0.6 - 0.1, torch.tensor([0.0, 0.3, 0.02])])
    out.update(dec_x=x, dec_lin_0_03=H.decreasing_fcn(x.clone(), torch.tensor(0.0), torch.tensor(0.3), "linear"))

    np.savez_compressed(OUT, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()})
    print("wrote", OUT, {k: tuple(np.asarray(v).shape) for k, v in out.items() if k.endswith(("dist", "rr", "rp", "out"))},
          "rect-rect crossings", int(out["ix_rr"].sum()), "rect-poly crossings", int(out["ix_rp"].sum()))


def main_mtv():
    """get_distances_between_agents(distance_type="mtv") (helper_scenario.py:1030-1138) -> tests/golden/kat/mtv.npz:
    rectangles of vehicle size at random poses from far apart to deeply overlapping, plus identical, edge-touching,
    corner-touching and parallel-overlapping ones; 6 agents per env so that every call holds 15 pairs."""
    torch.manual_seed(4321)
    B, N = 160, 6
    c = torch.randn(B, 1, 2) * 0.5 + torch.randn(B, N, 2) * torch.linspace(0.02, 0.45, B).reshape(B, 1, 1)
    yaw = torch.rand(B, N, 1) * 6.28
    yaw[:40] = (yaw[:40] * 0.05)                                          # near-parallel traffic
    c[0, 1], yaw[0, 1] = c[0, 0], yaw[0, 0]                               # identical rectangles
    yaw[1, :2] = 0.0; c[1, 1] = c[1, 0] + torch.tensor([L, 0.0])          # edge-touching, collinear sides
    yaw[2, :2] = 0.0; c[2, 1] = c[2, 0] + torch.tensor([L, W])            # corner-touching
    yaw[3, :2] = 0.0; c[3, 1] = c[3, 0] + torch.tensor([0.0, W / 2])      # overlapping, parallel
    yaw[4, 0] = 0.0; yaw[4, 1] = torch.pi / 2; c[4, 1] = c[4, 0] + torch.tensor([0.05, 0.0])   # crossed
    v = torch.stack([H.get_rectangle_vertices(center=c[:, a], yaw=yaw[:, a], width=W, length=L, is_close_shape=True)
                     for a in range(N)], dim=1)                           # [B,N,5,2] like WorldStateRT.vertices
    d = H.get_distances_between_agents(data=v.clone(), distance_type="mtv", is_set_diagonal=True,
                                       x_semidim=torch.tensor(4.5), y_semidim=torch.tensor(4.0))
    out = dict(mtv_vertices=v.numpy(), mtv_dist=d.numpy())
    path = os.path.join(os.path.dirname(OUT), "mtv.npz")
    np.savez_compressed(path, **out)
    off = ~np.eye(N, dtype=bool)
    print("wrote", path, "pairs", B * N * (N - 1) // 2, "negative", int((d.numpy()[:, off] < 0).sum() // 2),
          "zero", int((d.numpy()[:, off] == 0).sum() // 2), "min", float(d.min()), "max off-diag", float(d.numpy()[:, off].max()))


if __name__ == "__main__":
    if sys.argv[1:] == ["mtv"]:
        main_mtv()
    else:
        main()
