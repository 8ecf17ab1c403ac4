#!/usr/bin/env python
"""CPU tool (test infrastructure; the oracle generates realistic states): how many chunk evaluations the boundary scans
need per agent-step under different chunk layouts — one aligned 8-segment grid (the round-1 kernel), two grids offset by
half a chunk (pick the one in which the hint is central), 4-segment chunks evaluated in pairs.  Numpy model of the vote
(hint chunk -> bound -> chunks whose box is within the bound), including the per-warp maximum over the 8 agents of an
env that SIMT execution pays for.      python tests/tools/chunk_sim.py [scenario] [n_agents]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import oracle as O
from sigmarl_b200.maps import MapLibrary

st = sys.argv[1] if len(sys.argv) > 1 else "cpm_entire"
B, N = 256, int(sys.argv[2]) if len(sys.argv) > 2 else 8
HL, HW = 0.11, 0.0535
R = float(np.hypot(HL, HW)) * 1.0001
m = MapLibrary(st)
w = O.OracleWorld(st, B, N, mode="params", rew_method="distance")
for b in range(B): w.reset_env(b)
for b in range(B): w.refresh(b)
rng = np.random.default_rng(0)
ur = np.float32([1.0, 31 * np.pi / 180])

def seg_dists(P, poly):           # P [5,2], poly [n,2] -> [5, n-1]
    a, e = poly[:-1], poly[1:]
    l = e - a
    v = P[:, None, :] - a[None]
    t = np.clip((v * l[None]).sum(-1) / (l * l).sum(-1)[None], 0, 1)
    c = a[None] + l[None] * t[..., None]
    return np.linalg.norm(c - P[:, None, :], axis=-1)

def boxes_of(poly, starts, K):
    out = []
    nseg = len(poly) - 1
    for s0 in starts:
        s1 = min(s0 + K, nseg)
        seg = poly[max(s0, 0):s1 + 1]
        out.append((seg[:, 0].min(), seg[:, 1].min(), seg[:, 0].max(), seg[:, 1].max(), max(s0, 0), s1))
    return out

def lb(box, p):
    dx = max(box[0] - p[0], p[0] - box[2], 0.0); dy = max(box[1] - p[1], p[1] - box[3], 0.0)
    return (dx * dx + dy * dy) ** 0.5

def vote(D, boxes, first, p, P=None, minbound=False):
    """first: list of box indices evaluated first; returns the extra boxes voted.  P given: per-point refinement (a box
    is needed only if it is within some point's own running best, or within reach of the rectangle for crossings)"""
    segs = np.zeros(D.shape[1], bool)
    for i in first: segs[boxes[i][4]:boxes[i][5]] = True
    best = D[:, segs].min(1)
    thr = max(best.max() + R + 1e-4, R + 0.01 + 1e-4)
    if minbound:   # only the centre's minimum and the minimum over the four vertices are consumed
        thr = max(best[1:].min() + R, best[0], R + 0.01) + 1e-4
    out = [i for i, b in enumerate(boxes) if i not in first and lb(b, p) <= thr]
    if P is not None:
        # (a chunk within reach of the rectangle stays in: it is a crossing candidate)
        out = [i for i in out if lb(boxes[i], p) <= R + 0.01 + 1e-4 or any(lb(boxes[i], P[v]) <= best[v] + 1e-4 for v in range(5))]
    return out

cache = {}
stats = {k: [] for k in ("A", "AB", "K4", "App", "ABpp", "K4pp", "Amin", "ABmin", "K4min")}
wrong = {k: 0 for k in stats}
n_scans = 0
for t in range(8):
    act = ((rng.random((B, N, 2), np.float32) * 2 - 1) * ur).astype(np.float32)
    if t % 2: act[..., 1] *= 0.2
    obs, rew, done, _ = w.step(act, n_threads=8)
    per_env = {k: np.zeros((B, N, 2), int) for k in stats}
    for b in range(B):
        for a in range(N):
            p = int(w.path_id[b, a]); c = w.pos[b, a].astype(np.float64); psi = float(w.rot[b, a])
            h2 = int(w.idx_ref[b, a]) - 1
            cs, sn = np.cos(psi), np.sin(psi)
            P = np.array([c] + [c + np.array([cs * x - sn * y, sn * x + cs * y]) for x, y in ((HL, HW), (HL, -HW), (-HL, -HW), (-HL, HW))])
            for side in (0, 1):
                key = (p, side)
                if key not in cache:
                    poly = (m.left_xy[m.left_off[p]:m.left_off[p + 1]] if side == 0 else m.right_xy[m.right_off[p]:m.right_off[p + 1]]).astype(np.float64)
                    nseg = len(poly) - 1
                    cache[key] = (poly, boxes_of(poly, range(0, nseg, 8), 8), boxes_of(poly, range(-4, nseg, 8), 8), boxes_of(poly, range(0, nseg, 4), 4))
                poly, bA, bB, b4 = cache[key]
                nseg = len(poly) - 1
                D = seg_dists(P, poly)
                h = min(max(h2, 0), nseg - 1)
                true_seg = int(D[0].argmin())
                # A: aligned grid
                c0 = h // 8
                ex = vote(D, bA, [c0], c); per_env["A"][b, a, side] = 1 + len(ex); wrong["A"] += len(ex) > 0
                # AB: pick the grid in which the hint is central
                if 2 <= h % 8 < 6: ex = vote(D, bA, [h // 8], c)
                else:
                    k = (h + 4) // 8              # bB[k] covers [8k-4, 8k+4)
                    ex = vote(D, bB, [min(k, len(bB) - 1)], c)
                per_env["AB"][b, a, side] = 1 + len(ex); wrong["AB"] += len(ex) > 0
                # K4: pairs of 4-chunks; first pass = hint's chunk + the neighbour on the nearer side
                c4 = h // 4
                nb = c4 + 1 if h % 4 >= 2 else c4 - 1
                nb = min(max(nb, 0), len(b4) - 1)
                first = sorted({c4, nb})
                ex = vote(D, b4, first, c); per_env["K4"][b, a, side] = 1 + (len(ex) + 1) // 2; wrong["K4"] += len(ex) > 0
                ex = vote(D, bA, [c0], c, P); per_env["App"][b, a, side] = 1 + len(ex); wrong["App"] += len(ex) > 0
                if 2 <= h % 8 < 6: ex = vote(D, bA, [h // 8], c, P)
                else: ex = vote(D, bB, [min((h + 4) // 8, len(bB) - 1)], c, P)
                per_env["ABpp"][b, a, side] = 1 + len(ex); wrong["ABpp"] += len(ex) > 0
                ex = vote(D, b4, first, c, P); per_env["K4pp"][b, a, side] = 1 + (len(ex) + 1) // 2; wrong["K4pp"] += len(ex) > 0
                ex = vote(D, bA, [c0], c, minbound=True); per_env["Amin"][b, a, side] = 1 + len(ex); wrong["Amin"] += len(ex) > 0
                if 2 <= h % 8 < 6: ex = vote(D, bA, [h // 8], c, minbound=True)
                else: ex = vote(D, bB, [min((h + 4) // 8, len(bB) - 1)], c, minbound=True)
                per_env["ABmin"][b, a, side] = 1 + len(ex); wrong["ABmin"] += len(ex) > 0
                ex = vote(D, b4, first, c, minbound=True); per_env["K4min"][b, a, side] = 1 + (len(ex) + 1) // 2; wrong["K4min"] += len(ex) > 0
                n_scans += 1
    for k in stats: stats[k].append(per_env[k])
    for b in np.where(done)[0]:
        w.reset_env(int(b)); w.refresh(int(b))
print(f"{st} N={N}: {n_scans} boundary scans")
for k in stats:
    x = np.concatenate(stats[k])           # [T*B, N, 2] passes (8 segment evaluations per lane-group each)
    print(f"{k:3s}: passes per scan mean {x.mean():.3f} | warp cost (sum over sides of max over agents) {x.max(1).sum(-1).mean():.3f} "
          f"vs mean-sum {x.mean(1).sum(-1).mean():.3f} | merged L/R walk (max over agents of L+R) {x.sum(-1).max(1).mean():.3f} | "
          f"flat work list (hint passes + extras dealt over the warp) {2 + np.ceil((x - 1).sum((1, 2)) / 8).mean():.3f} | "
          f"scans needing extra chunks {wrong[k] / n_scans:.3f}")
