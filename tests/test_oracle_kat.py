"""CPU: function-level known-answer tests of the C oracle's primitives against the UNMODIFIED reference helpers
(``sigmarl/helper_scenario.py``; vectors from ``oracle/gen_kat.py`` -> ``tests/golden/kat/helpers.npz``).

Bit-exact for indices, masks and gathered points; 1e-6 abs for fp32 values that go through libm (glibc here,
ATen / Sleef in the reference)."""
import ctypes as C
import os

import numpy as np
import pytest

KAT = os.path.join(os.path.dirname(__file__), "golden", "kat", "helpers.npz")
L, W = 0.22, 0.107


@pytest.fixture(scope="module")
def kat():
    return np.load(KAT)


def _p(a):
    return np.ascontiguousarray(a, np.float32)


def test_perpendicular_distance_and_argmin(oracle_mod, kat):
    """helper_scenario.py:829-889 incl. the tail fix-up; on-vertex points tie at distance 0 -> first index."""
    lib = oracle_mod.lib()
    poly, n, pts = _p(kat["perp_poly"]), kat["perp_n"], _p(kat["perp_point"])
    P = poly.shape[1]
    for k in range(poly.shape[0]):
        idx = C.c_int()
        d = lib.orc_test_perp(pts[k].ctypes.data, poly[k].ctypes.data, P, int(n[k]), C.byref(idx))
        assert idx.value == int(kat["perp_idx"][k]), (k, idx.value, int(kat["perp_idx"][k]))
        assert abs(d - float(kat["perp_dist"][k])) <= 1e-6, (k, d, float(kat["perp_dist"][k]))
    assert (kat["perp_dist"][:16] == 0).all()          # the fixture really holds the exact ties


@pytest.mark.parametrize("st", ["cpm_mixed", "on_ramp_2_multilane"])
def test_boundary_argmin_at_spawn_poses(oracle_mod, kat, st):
    """Every reset pose (agent exactly on a centre point) against both lane boundaries: structural near-ties that only
    torch.norm's exact rounding — sqrt(fma(ey, ey, ex*ex)) — resolves like the reference (plain ex*ex + ey*ey gets
    4 of the 780 cpm_mixed cases wrong)."""
    lib = oracle_mod.lib()
    pm = oracle_mod.PaddedMap(st)
    rows = kat["spawn_idx_" + st]
    assert len(rows) > 500
    for p, k, side, want in rows:
        arr, cnt = (pm.right, pm.n_right) if side else (pm.left, pm.n_left)
        idx = C.c_int()
        lib.orc_test_perp(np.ascontiguousarray(pm.center[p, k]).ctypes.data, np.ascontiguousarray(arr[p]).ctypes.data,
                          pm.P, int(cnt[p]), C.byref(idx))
        assert idx.value == want, (st, p, k, side, idx.value, want)


def test_rectangle_vertices(oracle_mod, kat):
    """helper_scenario.py:695-826 (closed, 5 vertices)."""
    lib = oracle_mod.lib()
    c, yaw, want = _p(kat["rect_center"]), kat["rect_yaw"], kat["rect_vertices"]
    out = np.zeros((5, 2), np.float32)
    worst = 0.0
    for k in range(c.shape[0]):
        lib.orc_test_rect(np.float32(L / 2), np.float32(W / 2), c[k].ctypes.data, float(yaw[k]), out.ctypes.data)
        worst = max(worst, float(np.abs(out - want[k]).max()))
    assert worst <= 1e-6, worst


def test_interx_masks_are_bit_exact(oracle_mod, kat):
    """helper_scenario.py:1148-1229: strict crossing predicate, incl. identical / touching / collinear rectangles."""
    lib = oracle_mod.lib()
    a, b = _p(kat["ix_a"]), _p(kat["ix_b"])
    got = np.array([lib.orc_test_interx(a[k].ctypes.data, 5, b[k].ctypes.data, 5) for k in range(a.shape[0])], bool)
    assert np.array_equal(got, kat["ix_rr"]), np.where(got != kat["ix_rr"])[0]
    r, p = _p(kat["ix_rect"]), _p(kat["ix_poly"])
    got = np.array([lib.orc_test_interx(r[k].ctypes.data, 5, p[k].ctypes.data, p.shape[1]) for k in range(r.shape[0])], bool)
    assert np.array_equal(got, kat["ix_rp"]), np.where(got != kat["ix_rp"])[0]
    assert kat["ix_rr"].sum() > 20 and (~kat["ix_rr"]).sum() > 20 and kat["ix_rp"].sum() > 10


def test_angle_wrap(oracle_mod, kat):
    """helper_scenario.py:1276-1289, incl. +-pi, multiples of 2 pi and the fp32 neighbours of pi."""
    lib = oracle_mod.lib()
    got = np.array([lib.orc_test_wrap(float(x)) for x in kat["wrap_in"]], np.float32)
    assert np.abs(got - kat["wrap_out"]).max() <= 1e-6
    assert np.array_equal(got[-11:], kat["wrap_out"][-11:])      # the special values bit for bit


def test_global_to_local_transform(oracle_mod, kat):
    """helper_scenario.py:1241-1273."""
    lib = oracle_mod.lib()
    pi, pj, rot, want = _p(kat["loc_pi"]), _p(kat["loc_pj"]), kat["loc_rot"], kat["loc_out"]
    out = np.zeros(2, np.float32)
    worst = 0.0
    for k in range(pi.shape[0]):
        for q in range(pj.shape[1]):
            lib.orc_test_local(pi[k].ctypes.data, float(rot[k]), pj[k, q].ctypes.data, out.ctypes.data)
            worst = max(worst, float(np.abs(out - want[k, q]).max()))
    assert worst <= 2e-6, worst


@pytest.mark.parametrize("name,count,interval,shift", [("st", 3, 2, 1), ("nb", 5, 1, -2), ("nbr", 5, 1, 1)])
def test_short_term_and_nearing_points(oracle_mod, kat, name, count, interval, shift):
    """helper_scenario.py:892-957 with the three parameterisations the scenario uses: short-term reference path
    (world_state_rt.py:668-684), nearing boundary points in a step (:686-725) and at a reset (:531-576)."""
    lib = oracle_mod.lib()
    poly, n, loop, idx0 = _p(kat["st_poly"]), kat["st_n"], kat["st_loop"], kat["st_idx0"]
    P = poly.shape[1]
    out = np.zeros((count, 2), np.float32)
    for k in range(poly.shape[0]):
        lib.orc_test_path_points(poly[k].ctypes.data, P, int(n[k]), int(loop[k]), int(idx0[k]), count, interval, shift,
                                 out.ctypes.data)
        assert np.array_equal(out, kat[name + "_pts"][k]), (name, k, kat[name + "_idx"][k])


def test_decreasing_fcn(oracle_mod, kat):
    """helper_scenario.py:960-996, linear."""
    lib = oracle_mod.lib()
    got = np.array([lib.orc_test_dec(float(x), 0.0, np.float32(0.3)) for x in kat["dec_x"]], np.float32)
    assert np.abs(got - kat["dec_lin_0_03"]).max() <= 1e-7


def test_mtv_distance_between_rectangles(oracle_mod):
    """helper_scenario.py:1030-1138 get_distances_between_agents("mtv") — vectors from ``oracle/gen_kat.py mtv``: 2 400
    rectangle pairs from far apart to deeply overlapping, incl. identical / touching / crossed ones.  Pure fp32
    arithmetic (no libm besides sqrt), restated in the reference's operation order: bit-exact, sign and exact zeros
    ("collision" in MTV mode, world_state_rt_sim.py:394-396) included."""
    lib = oracle_mod.lib()
    g = np.load(os.path.join(os.path.dirname(KAT), "mtv.npz"))
    v, want = _p(g["mtv_vertices"]), g["mtv_dist"]
    B, N = want.shape[:2]
    got = np.zeros_like(want)
    for b in range(B):
        for i in range(N):
            for j in range(N):
                got[b, i, j] = want[b, i, i] if i == j else lib.orc_test_mtv(v[b, i].ctypes.data, v[b, j].ctypes.data)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    off = ~np.eye(N, dtype=bool)
    assert (want[:, off] < 0).sum() > 500 and (want[:, off] == 0).sum() > 0 and np.array_equal(want, want.transpose(0, 2, 1))
