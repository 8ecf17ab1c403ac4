/*
 * sigmarl_oracle.c — TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT PATH.
 *
 * A plain-C, scalar-fp32 CPU restatement of SigmaRL's vectorised road-traffic environment
 * step, written in the reference's own *sequential* order (world.step -> for each agent
 * {reward, observation} -> done) with the reference's own persistent world state, so that the
 * one-step-stale values the VMAS call order produces fall out of the ordering
 * instead of being coded as formulas.  Every function cites the reference file:line it follows
 * (paths are relative to /root/reference/sigmarl/).
 *
 * PARITY PINNING: the reference's own tests hold no golden vector for this path,
 * and two pieces of arithmetic live in un-vendored dependencies (torchdiffeq==0.2.5 fixed-grid
 * Euler, vmas==1.4.3 call order).  This oracle is therefore pinned against outputs of the
 * UNMODIFIED reference code run in the build container behind an import shim
 * (oracle/refshim/install_shims.py, oracle/gen_golden.py -> tests/golden/ *.npz); see
 * tests/test_oracle_golden.py.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline legs may load this file.
 *
 * Build: gcc -O2 -pthread -ffp-contract=off -fno-fast-math -shared -fPIC (oracle/Makefile).
 * -ffp-contract=off matters: the reference evaluates every ATen op with its own rounding, and
 * argmin / strict-sign decisions below depend on it.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAX_AGENTS 32
#define ORC_NST 3 /* n_points_short_term */

typedef struct {
    int n_paths, P;                        /* P = max_ref_path_points (road_traffic.py:505-530) */
    const float *center, *left, *right;    /* [n_paths][P][2], padded as world_state_rt.py:313-392 */
    const int *n_center, *n_left, *n_right;
    const uint8_t *is_loop;
    const float *yaw;                      /* [n_paths][P] center_line_yaw (n_center-1 valid) */
    /* lanelet table for the lanelet-relation observation mask (map_manager.py:39-119); n_lanelets = 0: none */
    int n_lanelets;
    const float *lanelet_xy;               /* centre lines of all lanelets, concatenated [lanelet_off[n]][2] */
    const int *lanelet_off;                /* [n_lanelets + 1] */
    const uint8_t *lanelet_adj;            /* [n_lanelets][n_lanelets]: j in neighboring_lanelets_idx[i] */
} orc_map;

typedef struct {
    float dt, max_speed, max_steering, max_acc, max_steering_rate;
    float l_wb, lr_over_lwb, half_length, half_width, diag;
    float w_ref[ORC_NST];
    float speed_dt;          /* float32(max_speed * dt)            road_traffic.py:986 */
    float reward_progress;
    float nb_low, nb_high, na_low, na_high, ttc_low, ttc_high;
    float pen_near_boundary, pen_near_agents, pen_collide_agents, pen_collide_lane;
    float norm_pos, norm_v, norm_rot, norm_dist;
    float dsafe_sq;          /* float32(d_safe*d_safe in double)    road_traffic.py:1291 */
    float reset_min_dist_sq; /* reset_agent_min_distance**2         world_state_rt_sim.py:305 */
    int rew_exact_sparse, rew_has_ttc, rew_has_distance, rew_has_sparse;
    int k_near, max_steps, is_cpm_entire, sample_interval;
    int testing_mode;        /* parameters.is_testing_mode          road_traffic.py:1050-1055, 1429-1447 */
    float reward_reach_goal; /* rewards.reach_goal                  road_traffic.py:217-219 */
    /* observation layout (observation_provider_rt.py:594-925); 0 = the default flags */
    int obs_flags;           /* ORC_OBS_* */
    float norm_pos_world[2]; /* normalizers.pos_world (bird view)   road_traffic.py:593-595 */
    float norm_dist_agent;   /* normalizers.distance_agent          road_traffic.py:605-607 */
    float fixed_duration;    /* reset_agent_fixed_duration [s], 0 = off   road_traffic.py:1388-1393 */
    int use_mtv;             /* is_use_mtv_distance: distances.type == "mtv"  road_traffic.py:611-614 */
    float mask_distance;     /* thresholds.distance_mask_agents (5 agent lengths)   road_traffic.py:663 */
} orc_cfg;

#define ORC_OBS_BIRD_VIEW 1       /* is_ego_view = False */
#define ORC_OBS_CENTRES 2         /* is_observe_vertices = False: pos, rot, length, width instead of 4 vertices */
#define ORC_OBS_STEERING 4        /* is_obs_steering */
#define ORC_OBS_REF_OTHERS 8      /* is_observe_ref_path_other_agents */
#define ORC_OBS_NO_DIST_AGENTS 16 /* is_observe_distance_to_agents = False */
#define ORC_OBS_NO_DIST_CENTER 32 /* is_observe_distance_to_center_line = False */
#define ORC_OBS_BOUNDARY_POINTS 64 /* is_observe_distance_to_boundaries = False: 5 points of each boundary */
#define ORC_OBS_MASK 128          /* is_apply_mask: observed neighbours farther than mask_distance show constants */
#define ORC_OBS_MASK_LANELETS 256 /* + mask neighbours on non-adjacent lanelets (bird view, OSM maps; :646-664) */
#define ORC_NNB 5                  /* n_points_nearing_boundary (road_traffic.py:296-298) */

typedef struct {
    int B, N;
    orc_map map;
    orc_cfg cfg;
    /* agent state (helper_common.py:290-430) */
    float *pos, *rot, *speed, *steering, *vel, *sideslip; /* [B][N][2|1] */
    int *path_id;                                         /* global path index */
    /* world state (world_state_rt.py:119-277, world_state_rt_sim.py:36-55) */
    float *vertices;                                      /* [B][N][5][2] */
    float *d_agents;                                      /* [B][N][N] */
    float *d_ref, *d_left, *d_right, *d_bound;            /* [B][N], [B][N][5] */
    int *idx_ref;
    int *idx_left, *idx_right;                            /* distances.closest_point_on_left_b / right_b [B][N] */
    uint8_t *near_fresh;                                  /* [B][N] 1: nearing boundary points last written by a
                                                             reset (n_points_shift = +1, world_state_rt.py:531-576),
                                                             0: by a step (shift = -2, :686-725) */
    float *short_term;                                    /* [B][N][3][2] */
    float *prev_pos;                                      /* state_buffer latest */
    uint8_t *col_agents, *col_lane, *col_entry, *col_exit;
    int *step;                                            /* timer.step [B] */
} orc_world;

/* ---------------------------------------------------------------- geometry primitives */

/* torch.norm(x, dim=-1) over a dimension of size 2, as ATen's CPU reduction evaluates it on an FMA-capable x86
 * (AVX2 / AVX512 dispatch): acc = fma(x0, x0, 0); acc = fma(x1, x1, acc); sqrt(acc) — i.e. the second square is
 * NOT rounded on its own.  Measured against torch 2.11 on 1.5e6 random pairs (contiguous and strided): 0
 * mismatches; plain sqrtf(x0*x0 + x1*x1) differs by 1 ulp in 8 % of them, which decides the argmin below when the
 * foot of the perpendicular sits next to a polyline vertex (e.g. every spawn pose against its lane boundaries).
 * torch.sum(a*b, dim) and sqrt(sum(d**2)) (helper_scenario.py:861-862, :1022-1028) round every product first. */
static float orc_norm2(float x0, float x1) { return sqrtf(fmaf(x1, x1, x0 * x0)); }

/* helper_scenario.py:829-889 get_perpendicular_distances.  Returns min distance, *idx = argmin+1. */
static float orc_perp(const float p[2], const float *poly, int P, int n, int *idx) {
    float d[1024];
    int S = P - 1;
    for (int s = 0; s < S; s++) {
        float ax = poly[2 * s], ay = poly[2 * s + 1];
        float bx = poly[2 * s + 2], by = poly[2 * s + 3];
        float lx = bx - ax, ly = by - ay;                 /* line_vecs  :857 */
        float px = p[0] - ax, py = p[1] - ay;             /* point_vecs :858 */
        float len2 = lx * lx + ly * ly;                   /* :861 */
        float t = (px * lx + py * ly) / len2;             /* :862 */
        /* torch.clamp(x, 0, 1) propagates NaN :865 */
        if (t < 0.0f) t = 0.0f;
        else if (t > 1.0f) t = 1.0f;
        float cx = ax + lx * t, cy = ay + ly * t;         /* :868 */
        float ex = cx - p[0], ey = cy - p[1];
        d[s] = orc_norm2(ex, ey);                         /* torch.norm :871 */
    }
    for (int s = n - 1; s < S; s++) d[s] = d[n - 2];      /* :873-879 */
    int best = 0;
    float bd = d[0];
    for (int s = 1; s < S; s++) {                         /* torch.min: first minimal index :883 */
        if (d[s] < bd || (d[s] != d[s] && bd == bd)) { bd = d[s]; best = s; }
    }
    *idx = best + 1;                                      /* :885-887 */
    return bd;
}

/* helper_scenario.py:1148-1229 interX (is_return_points=False): any strict segment crossing. */
static int orc_interx(const float *L1, int n1, const float *L2, int n2) {
    int hit = 0;
    for (int i = 0; i + 1 < n1; i++) {
        float x1a = L1[2 * i], y1a = L1[2 * i + 1], x1b = L1[2 * i + 2], y1b = L1[2 * i + 3];
        float dx1 = x1b - x1a, dy1 = y1b - y1a;
        float S1 = dx1 * y1a - dy1 * x1a;                 /* :1176 */
        for (int j = 0; j + 1 < n2; j++) {
            float x2a = L2[2 * j], y2a = L2[2 * j + 1], x2b = L2[2 * j + 2], y2b = L2[2 * j + 3];
            float dx2 = x2b - x2a, dy2 = y2b - y2a;
            float S2 = dx2 * y2a - dy2 * x2a;             /* :1177 */
            float f_a = (dx1 * y2a - dy1 * x2a) - S1;     /* :1183-1189 */
            float f_b = (dx1 * y2b - dy1 * x2b) - S1;
            int C1 = (f_a * f_b) < 0.0f;
            float g_a = (y1a * dx2 - x1a * dy2) - S2;     /* :1190-1198 */
            float g_b = (y1b * dx2 - x1b * dy2) - S2;
            int C2 = (g_a * g_b) < 0.0f;
            hit |= (C1 & C2);
        }
    }
    return hit;
}

/* helper_scenario.py:695-826 get_rectangle_vertices (closed, 5 vertices); the 2x2 . 2x5 matmul
 * goes through ATen's small-matrix kernel: r = c*vx, r += (-s)*vy with float accumulation. */
static void orc_rect(const orc_cfg *c, const float pos[2], float yaw, float *out) {
    const float hl = c->half_length, hw = c->half_width;
    const float bx[5] = {hl, hl, -hl, -hl, hl};
    const float by[5] = {hw, -hw, -hw, hw, hw};
    float cy = cosf(yaw), sy = sinf(yaw);
    float nsy = -sy;
    for (int v = 0; v < 5; v++) {
        float rx = cy * bx[v] + nsy * by[v];
        float ry = sy * bx[v] + cy * by[v];
        out[2 * v] = rx + pos[0];
        out[2 * v + 1] = ry + pos[1];
    }
}

/* helper_scenario.py:1276-1289 angle_eliminate_two_pi */
static float orc_wrap(float a) {
    const float two_pi = (float)(2.0 * M_PI);
    float m = fmodf(a, two_pi);                           /* torch `%`: sign of divisor */
    if (m != 0.0f && m < 0.0f) m += two_pi;
    if (m > (float)M_PI) m -= two_pi;
    return m;
}

/* helper_scenario.py:1241-1273 transform_from_global_to_local_coordinate (one point) */
static void orc_local(const float pi[2], float rot_i, const float pj[2], float out[2]) {
    float vx = pj[0] - pi[0], vy = pj[1] - pi[1];
    float a = orc_norm2(vx, vy);                          /* pos_vec.norm(dim=2) :1263 */
    float r = atan2f(vy, vx) - rot_i;
    out[0] = cosf(r) * a;
    out[1] = sinf(r) * a;
}

/* helper_scenario.py:960-996 decreasing_fcn(type="linear") */
static float orc_dec(float x, float x0, float x1) {
    if (x < x0) x = x0;
    else if (x > x1) x = x1;
    float denom = x1 - x0;
    return 1.0f - (x - x0) / denom;
}

/* helper_scenario.py:892-957 get_short_term_reference_path on one padded polyline [P][2]:
 * idx_k = k * interval + idx + shift (:928-932); loops wrap with (idx + 1) % n for idx >= n - 1 (:941-946);
 * a negative index is python's "from the end" of the P-point array. */
static void orc_path_points(const float *poly, int P, int n, int is_loop, int idx, int count, int interval, int shift,
                            float *out) {
    for (int k = 0; k < count; k++) {
        int fi = k * interval + idx + shift;
        if (is_loop && fi >= n - 1) fi = (fi + 1) % n;
        if (fi < 0) fi += P;
        out[2 * k] = poly[2 * fi];
        out[2 * k + 1] = poly[2 * fi + 1];
    }
}

/* the agent's short-term reference path: 3 points, sample interval 2, n_points_shift = 1 (world_state_rt.py:668-684) */
static void orc_short_term(const orc_world *w, int path, int idx, float *out) {
    const orc_map *m = &w->map;
    orc_path_points(m->center + (size_t)path * m->P * 2, m->P, m->n_center[path], m->is_loop[path], idx, ORC_NST,
                    w->cfg.sample_interval, 1, out);
}

/* world_state_rt.py:689-725: nearing boundary points = get_short_term_reference_path on the PADDED boundary
 * array [P][2] with sample_interval 1, n_points_shift -2 and — as the reference passes it — the CENTRE line's
 * point count for the loop wrap.  Index -1 is python's "last element" (tail padding = last boundary point). */
static void orc_nearing_points(const orc_world *w, int path, const float *poly, int idx, int shift, float *out) {
    const orc_map *m = &w->map;
    orc_path_points(poly, m->P, m->n_center[path], m->is_loop[path], idx, ORC_NNB, 1, shift, out);
}

/* ---------------------------------------------------------------- world-state updates */

#define AG(b, a) ((size_t)(b) * N + (a))

/* helper_scenario.py:1030-1138: MTV-based ("minimum translation vector") distance between two rectangles given by
 * their first four vertices [4][2], in the reference's operation order.  Per rectangle the two edge directions
 * v1-v0, v2-v1 are normalised (:1042-1045); all vertices are projected on the axes of the OTHER rectangle; a vertex
 * outside the other's projection interval on an axis contributes the (signed) gap on that axis, the per-vertex
 * distance is the Euclidean norm of its two gaps (:1072-1077); the pair's distance is the smallest of the eight
 * (:1113-1118).  If any vertex lies strictly inside the other rectangle the distance is minus the smaller of the two
 * rectangles' smallest projection overlaps (:1079-1084, :1119-1124). */
static void orc_mtv_axes(const float *v, float ax[2][2]) {
    for (int k = 0; k < 2; k++) {
        float dx = v[2 * (k + 1)] - v[2 * k], dy = v[2 * (k + 1) + 1] - v[2 * k + 1];   /* torch.diff :1042 */
        float n = orc_norm2(dx, dy);                                                     /* torch.norm :1043 */
        ax[k][0] = dx / n;
        ax[k][1] = dy / n;
    }
}

/* vertices of `va` against the axes `axb` of rectangle `vb`: pos[4] = per-vertex Euclidean gap, *min_overlap =
 * min over the two axes of the projection overlap, returns 1 if some vertex of a is strictly inside b */
static int orc_mtv_half(const float *va, const float *vb, float axb[2][2], float pos[4], float *min_overlap) {
    float pbb[4][2], pab[4][2], mx_b[2], mn_b[2], mx_a[2], mn_a[2];
    for (int v = 0; v < 4; v++)
        for (int k = 0; k < 2; k++) {
            pbb[v][k] = vb[2 * v] * axb[k][0] + vb[2 * v + 1] * axb[k][1];              /* (.*.).sum(dim=3) :1060-1062 */
            pab[v][k] = va[2 * v] * axb[k][0] + va[2 * v + 1] * axb[k][1];              /* :1066-1068 */
        }
    for (int k = 0; k < 2; k++) {
        mx_b[k] = mn_b[k] = pbb[0][k];
        mx_a[k] = mn_a[k] = pab[0][k];
        for (int v = 1; v < 4; v++) {
            if (pbb[v][k] > mx_b[k]) mx_b[k] = pbb[v][k];
            if (pbb[v][k] < mn_b[k]) mn_b[k] = pbb[v][k];
            if (pab[v][k] > mx_a[k]) mx_a[k] = pab[v][k];
            if (pab[v][k] < mn_a[k]) mn_a[k] = pab[v][k];
        }
    }
    float ov[2];
    for (int k = 0; k < 2; k++)                                                          /* :1079 */
        ov[k] = (mx_b[k] < mx_a[k] ? mx_b[k] : mx_a[k]) - (mn_b[k] > mn_a[k] ? mn_b[k] : mn_a[k]);
    *min_overlap = ov[0] < ov[1] ? ov[0] : ov[1];
    int inside_any = 0;
    for (int v = 0; v < 4; v++) {
        float gap[2];
        int inside = 1;
        for (int k = 0; k < 2; k++) {                                                    /* :1072-1076 */
            float lo = (pab[v][k] - mn_b[k]) * (float)(pab[v][k] <= mn_b[k]);
            float hi = (mx_b[k] - pab[v][k]) * (float)(pab[v][k] >= mx_b[k]);
            gap[k] = lo + hi;
            inside &= (pab[v][k] > mn_b[k]) && (pab[v][k] < mx_b[k]);                    /* :1081-1083 */
        }
        pos[v] = orc_norm2(gap[0], gap[1]);                                              /* torch.norm(dim=2) :1077 */
        /* MTVs_Euclidean_negative = -min_overlap * inside; "negative" iff its absolute value is > 0 (:1119-1121) */
        if (inside && fabsf(-*min_overlap) > 0.0f) inside_any = 1;
    }
    return inside_any;
}

static float orc_mtv(const float *vi, const float *vj) {
    float axi[2][2], axj[2][2], pos_ij[4], pos_ji[4], ov_j, ov_i;
    orc_mtv_axes(vi, axi);
    orc_mtv_axes(vj, axj);
    int neg = orc_mtv_half(vi, vj, axj, pos_ij, &ov_j);     /* i's vertices on j's axes */
    neg |= orc_mtv_half(vj, vi, axi, pos_ji, &ov_i);        /* j's vertices on i's axes */
    float d = pos_ij[0];
    for (int v = 1; v < 4; v++) if (pos_ij[v] < d) d = pos_ij[v];
    for (int v = 0; v < 4; v++) if (pos_ji[v] < d) d = pos_ji[v];
    if (neg) d = -(ov_j < ov_i ? ov_j : ov_i);                                           /* :1122-1124 */
    return d;
}

/* world_state_rt_sim.py:360-373 + helper_scenario.py:1012-1029,1140-1143 (c2c) / :1030-1138 (mtv).  MTV mode reads
 * the stored rectangles: inside a step they are still LAST step's (update_vertices runs after update_distances,
 * world_state_rt_sim.py:432-448), after a reset they are fresh (world_state_rt.py:469-475). */
static void orc_mutual(orc_world *w, int b) {
    int N = w->N;
    if (w->cfg.use_mtv) {
        for (int i = 0; i < N; i++) {
            w->d_agents[(AG(b, i)) * N + i] = w->cfg.diag;
            for (int j = i + 1; j < N; j++) {
                float d = orc_mtv(&w->vertices[AG(b, i) * 10], &w->vertices[AG(b, j) * 10]);
                w->d_agents[(AG(b, i)) * N + j] = d;
                w->d_agents[(AG(b, j)) * N + i] = d;
            }
        }
        return;
    }
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) {
            float dx = w->pos[AG(b, i) * 2] - w->pos[AG(b, j) * 2];
            float dy = w->pos[AG(b, i) * 2 + 1] - w->pos[AG(b, j) * 2 + 1];
            float d = sqrtf(dx * dx + dy * dy);
            if (i == j) d = w->cfg.diag;
            w->d_agents[(AG(b, i)) * N + j] = d;
        }
}

/* world_state_rt.py:582-656 update_distances (the per-agent part) */
static void orc_update_distances(orc_world *w, int b, int i) {
    int N = w->N;
    const orc_map *m = &w->map;
    size_t g = AG(b, i);
    int path = w->path_id[g];
    const float *cen = m->center + (size_t)path * m->P * 2;
    const float *lef = m->left + (size_t)path * m->P * 2;
    const float *rig = m->right + (size_t)path * m->P * 2;
    int dummy;
    w->d_ref[g] = orc_perp(&w->pos[g * 2], cen, m->P, m->n_center[path], &w->idx_ref[g]);
    w->d_left[g * 5] = orc_perp(&w->pos[g * 2], lef, m->P, m->n_left[path], &w->idx_left[g]) - w->cfg.half_width;
    w->d_right[g * 5] = orc_perp(&w->pos[g * 2], rig, m->P, m->n_right[path], &w->idx_right[g]) - w->cfg.half_width;
    for (int c = 0; c < 4; c++) {
        const float *v = &w->vertices[(g * 5 + c) * 2];
        w->d_left[g * 5 + c + 1] = orc_perp(v, lef, m->P, m->n_left[path], &dummy);
        w->d_right[g * 5 + c + 1] = orc_perp(v, rig, m->P, m->n_right[path], &dummy);
    }
    float mn = w->d_left[g * 5];
    for (int c = 0; c < 5; c++) {
        if (w->d_left[g * 5 + c] < mn) mn = w->d_left[g * 5 + c];
        if (w->d_right[g * 5 + c] < mn) mn = w->d_right[g * 5 + c];
    }
    w->d_bound[g] = mn;
}

/* world_state_rt_sim.py:379-424 update_collisions */
static void orc_update_collisions(orc_world *w, int b) {
    int N = w->N;
    const orc_map *m = &w->map;
    for (int i = 0; i < N; i++) {
        size_t g = AG(b, i);
        const float *vi = &w->vertices[g * 10];
        if (w->cfg.use_mtv) {          /* :394-396: agents collide iff their mtv-based distance is exactly zero */
            for (int j = 0; j < N; j++) w->col_agents[g * N + j] = (uint8_t)(w->d_agents[g * N + j] == 0.0f);
        } else
        for (int j = i + 1; j < N; j++) {
            if (orc_interx(vi, 5, &w->vertices[AG(b, j) * 10], 5)) {
                w->col_agents[g * N + j] = 1;
                w->col_agents[AG(b, j) * N + i] = 1;
            }
        }
        int path = w->path_id[g];
        const float *lef = m->left + (size_t)path * m->P * 2;
        const float *rig = m->right + (size_t)path * m->P * 2;
        if (orc_interx(vi, 5, lef, m->P) | orc_interx(vi, 5, rig, m->P)) w->col_lane[g] = 1;
        /* entry/exit only when no env has this agent on a loop (:414); a map set is all-loop or
         * no-loop, so this is a per-path flag here. */
        if (!m->is_loop[path]) {
            float entry[4] = {lef[0], lef[1], rig[0], rig[1]};   /* world_state_rt.py:394-406 */
            int nl = m->n_left[path], nr = m->n_right[path];
            float exit_[4] = {lef[2 * (nl - 1)], lef[2 * (nl - 1) + 1], rig[2 * (nr - 1)], rig[2 * (nr - 1) + 1]};
            w->col_entry[g] = (uint8_t)orc_interx(vi, 5, entry, 2);
            w->col_exit[g] = (uint8_t)orc_interx(vi, 5, exit_, 2);
        }
    }
}

static void orc_reset_collisions(orc_world *w, int b) {
    int N = w->N;
    memset(&w->col_agents[AG(b, 0) * N], 0, (size_t)N * N);
    memset(&w->col_lane[AG(b, 0)], 0, N);
    memset(&w->col_entry[AG(b, 0)], 0, N);
    memset(&w->col_exit[AG(b, 0)], 0, N);
}

/* world_state_rt.py:422-529 reset_init_distances_and_short_term_ref_path: everything fresh */
static void orc_refresh_agent(orc_world *w, int b, int i) {
    int N = w->N;
    size_t g = AG(b, i);
    orc_rect(&w->cfg, &w->pos[g * 2], w->rot[g], &w->vertices[g * 10]);
    orc_update_distances(w, b, i); /* same 11 scans, but vertices are fresh here (:470-476) */
    orc_short_term(w, w->path_id[g], w->idx_ref[g], &w->short_term[g * 6]);
    w->near_fresh[g] = 1;          /* :531-576 */
}

/* road_traffic.py:897-923: after (re)placing agents of env b (all agents, or one respawned agent). */
void orc_refresh_env(orc_world *w, int b, int agent /* -1 = all */) {
    int N = w->N;
    for (int i = 0; i < N; i++)
        if (agent < 0 || agent == i) orc_refresh_agent(w, b, i);
    orc_mutual(w, b);
    orc_reset_collisions(w, b);
    for (int i = 0; i < N; i++) { /* state_buffer.reset(); add(current) :910-923 */
        w->prev_pos[AG(b, i) * 2] = w->pos[AG(b, i) * 2];
        w->prev_pos[AG(b, i) * 2 + 1] = w->pos[AG(b, i) * 2 + 1];
    }
}

/* ---------------------------------------------------------------- dynamics */

/* helper_training.py:797-861 WorldCustom.step + dynamics.py:62-192 (Euler, one tick) */
static void orc_dynamics(orc_world *w, int b, int a, float *action /* in/out: clamped */) {
    int N = w->N;
    const orc_cfg *c = &w->cfg;
    size_t g = AG(b, a);
    float u0 = action[0], u1 = action[1];
    if (u0 < -c->max_speed) u0 = -c->max_speed; else if (u0 > c->max_speed) u0 = c->max_speed;
    if (u1 < -c->max_steering) u1 = -c->max_steering; else if (u1 > c->max_steering) u1 = c->max_steering;
    action[0] = u0; action[1] = u1;                       /* written back :808-818 */
    float v = w->speed[g], delta = w->steering[g], psi = w->rot[g];
    float acc = (u0 - v) / c->dt;                         /* :821 */
    float rate = (u1 - delta) / c->dt;
    if (acc < -c->max_acc) acc = -c->max_acc; else if (acc > c->max_acc) acc = c->max_acc;
    if (rate < -c->max_steering_rate) rate = -c->max_steering_rate;
    else if (rate > c->max_steering_rate) rate = c->max_steering_rate;
    float td = tanf(delta);
    float beta = atanf(c->lr_over_lwb * td);              /* dynamics.py:102 */
    float f0 = v * cosf(psi + beta);
    float f1 = v * sinf(psi + beta);
    float f2 = ((v / c->l_wb) * td) * cosf(beta);         /* :108-110 */
    /* torchdiffeq fixed-grid Euler: y1 = y0 + (t1 - t0) * f(t0, y0) */
    float x1 = w->pos[g * 2] + c->dt * f0;
    float y1 = w->pos[g * 2 + 1] + c->dt * f1;
    float psi1 = psi + c->dt * f2;
    float v1 = v + c->dt * acc;
    float d1 = delta + c->dt * rate;
    const float pi_f = (float)M_PI, two_pi = (float)(2.0 * M_PI);
    float t = d1 + pi_f;                                  /* :158 */
    float mth = fmodf(t, two_pi);
    if (mth != 0.0f && mth < 0.0f) mth += two_pi;
    d1 = mth - pi_f;
    float beta1 = atanf(c->lr_over_lwb * tanf(d1));       /* :161-163 */
    float course = psi1 + beta1;
    w->pos[g * 2] = x1; w->pos[g * 2 + 1] = y1;
    w->rot[g] = psi1; w->speed[g] = v1; w->steering[g] = d1;
    w->vel[g * 2] = v1 * cosf(course);
    w->vel[g * 2 + 1] = v1 * sinf(course);
    w->sideslip[g] = beta1;
}

/* ---------------------------------------------------------------- reward / observation */

/* road_traffic.py:1255-1332 */
static float orc_ttc_penalty(const orc_world *w, int b, int i) {
    int N = w->N;
    const orc_cfg *c = &w->cfg;
    const float eps = 1e-6f;
    float sum = 0.0f;
    for (int j = 0; j < N; j++) {
        float px = w->pos[AG(b, j) * 2] - w->pos[AG(b, i) * 2];
        float py = w->pos[AG(b, j) * 2 + 1] - w->pos[AG(b, i) * 2 + 1];
        float vx = w->vel[AG(b, j) * 2] - w->vel[AG(b, i) * 2];
        float vy = w->vel[AG(b, j) * 2 + 1] - w->vel[AG(b, i) * 2 + 1];
        float qa = vx * vx + vy * vy;
        float qb = 2.0f * (px * vx + py * vy);
        float pp = px * px + py * py;
        float qc = pp - c->dsafe_sq;
        float disc = qb * qb - (4.0f * qa) * qc;
        float sq = sqrtf(disc < 0.0f ? 0.0f : disc);
        float dist = sqrtf(pp < 0.0f ? 0.0f : pp);
        int valid = (qa > eps) && (disc > 0.0f) && (qb < 0.0f);
        float cand = (-qb - sq) / (2.0f * qa + eps);
        float ttc = INFINITY;
        if (valid && cand > 0.0f) ttc = cand;
        if (dist <= c->na_low) ttc = 0.0f;
        if (j == i) ttc = INFINITY;
        if (!(dist <= c->na_high)) ttc = INFINITY;
        if (ttc > c->ttc_high) ttc = c->ttc_high;
        sum += orc_dec(ttc, c->ttc_low, c->ttc_high);
    }
    float risk = sum / (float)(N - 1 > 1 ? N - 1 : 1);
    return risk * c->pen_near_agents;
}

/* road_traffic.py:925-1253 reward(agent i) for env b, incl. the state updates it triggers */
static float orc_reward(orc_world *w, int b, int i) {
    int N = w->N;
    const orc_cfg *c = &w->cfg;
    size_t g = AG(b, i);
    if (i == 0) w->step[b] += 1;                          /* :954-962 */
    /* update_state_before_rewarding world_state_rt_sim.py:432-448 */
    if (i == 0) orc_mutual(w, b);
    orc_update_distances(w, b, i);
    w->near_fresh[g] = 0;   /* update_ref_paths_agent_related (world_state_rt_sim.py:454) rewrites the nearing points */
    if (i == 0) {
        orc_reset_collisions(w, b);
        for (int a = 0; a < N; a++) orc_rect(c, &w->pos[AG(b, a) * 2], w->rot[AG(b, a)], &w->vertices[AG(b, a) * 10]);
        orc_update_collisions(w, b);
    }
    /* forward movement :971-989 */
    float mvx = w->pos[g * 2] - w->prev_pos[g * 2], mvy = w->pos[g * 2 + 1] - w->prev_pos[g * 2 + 1];
    float acc = 0.0f;
    for (int k = 0; k < ORC_NST; k++) {
        float rx = w->short_term[g * 6 + 2 * k] - w->prev_pos[g * 2];
        float ry = w->short_term[g * 6 + 2 * k + 1] - w->prev_pos[g * 2 + 1];
        float mp = mvx * rx + mvy * ry;
        acc += mp * c->w_ref[k];                          /* torch.matmul [B,3].[3] */
    }
    float rew = 0.0f;
    rew += (acc / c->speed_dt) * c->reward_progress;
    int any_a2a = 0;
    for (int j = 0; j < N; j++) any_a2a |= w->col_agents[g * N + j];
    float pen_a2a = (float)any_a2a * c->pen_collide_agents;
    float pen_lane = (float)w->col_lane[g] * c->pen_collide_lane;
    float pen_nb = orc_dec(w->d_bound[g], c->nb_low, c->nb_high) * c->pen_near_boundary; /* :1040-1048 */
    if (c->testing_mode) {                                                              /* :1050-1055 */
        rew += (float)w->col_exit[g] * c->reward_reach_goal;
        rew += pen_a2a; rew += pen_lane;
    }
    if (!c->testing_mode && c->rew_exact_sparse) { rew += pen_a2a; rew += pen_lane; }  /* :1058-1062 */
    if (!c->testing_mode && c->rew_has_ttc) {                                           /* :1064-1085 */
        rew += orc_ttc_penalty(w, b, i);
        rew += pen_nb;
        rew += pen_a2a; rew += pen_lane;
        if (c->rew_has_sparse) { rew += pen_a2a; rew += pen_lane; }
    }
    if (!c->testing_mode && c->rew_has_distance) {                                      /* :1087-1112 */
        float s = 0.0f;
        for (int j = 0; j < N; j++) s += orc_dec(w->d_agents[g * N + j], c->na_low, c->na_high);
        rew += s * c->pen_near_agents;
        rew += pen_nb;
        if (c->rew_has_sparse) { rew += pen_a2a; rew += pen_lane; }
    }
    if (i == N - 1)                                       /* state_buffer.add :1226-1240 */
        for (int a = 0; a < N; a++) {
            w->prev_pos[AG(b, a) * 2] = w->pos[AG(b, a) * 2];
            w->prev_pos[AG(b, a) * 2 + 1] = w->pos[AG(b, a) * 2 + 1];
        }
    orc_short_term(w, w->path_id[g], w->idx_ref[g], &w->short_term[g * 6]); /* :1243, world_state_rt.py:668 */
    if (rew < -1.0f) rew = -1.0f; else if (rew > 1.0f) rew = 1.0f;          /* :1249 */
    return rew;
}

/* what observation_provider_rt.py:345-588 update_state snapshots at observation(agent 0) time and
 * that later reward(i>=1) calls would otherwise overwrite */
typedef struct {
    float short_term[ORC_MAX_AGENTS][6];
    float d_ref[ORC_MAX_AGENTS], min_l[ORC_MAX_AGENTS], min_r[ORC_MAX_AGENTS];
    float near_l[ORC_MAX_AGENTS][2 * ORC_NNB], near_r[ORC_MAX_AGENTS][2 * ORC_NNB];
    int lanelet[ORC_MAX_AGENTS];          /* map.current_lanelet_idx (determine_current_lanelet at update_state, :585-588) */
} orc_snap;

/* map_manager.py:39-89 determine_current_lanelet for one position: the lanelet whose centre line holds the closest
 * point (squared distance, torch.sum((a - c)**2)), first minimal lanelet index.  Centre lines are padded with ZEROS to
 * the longest one (:58-66), so every shorter lanelet also "has" the point (0, 0). */
static int orc_current_lanelet(const orc_map *m, const float pos[2]) {
    int max_len = 0, best = 0;
    float best_d = INFINITY;
    for (int l = 0; l < m->n_lanelets; l++)
        if (m->lanelet_off[l + 1] - m->lanelet_off[l] > max_len) max_len = m->lanelet_off[l + 1] - m->lanelet_off[l];
    for (int l = 0; l < m->n_lanelets; l++) {
        float dmin = INFINITY;
        int n = m->lanelet_off[l + 1] - m->lanelet_off[l];
        for (int k = 0; k < n; k++) {
            float dx = pos[0] - m->lanelet_xy[2 * (m->lanelet_off[l] + k)];
            float dy = pos[1] - m->lanelet_xy[2 * (m->lanelet_off[l] + k) + 1];
            float d = dx * dx + dy * dy;
            if (d < dmin) dmin = d;
        }
        if (n < max_len) {
            float d = pos[0] * pos[0] + pos[1] * pos[1];
            if (d < dmin) dmin = d;
        }
        if (dmin < best_d) { best_d = dmin; best = l; }
    }
    return best;
}

static void orc_take_snapshot(const orc_world *w, int b, orc_snap *s) {
    int N = w->N;
    for (int j = 0; j < N; j++) {
        size_t g = AG(b, j);
        memcpy(s->short_term[j], &w->short_term[g * 6], 6 * sizeof(float));
        s->d_ref[j] = w->d_ref[g];
        float ml = w->d_left[g * 5], mr = w->d_right[g * 5];
        for (int c = 1; c < 5; c++) {
            if (w->d_left[g * 5 + c] < ml) ml = w->d_left[g * 5 + c];
            if (w->d_right[g * 5 + c] < mr) mr = w->d_right[g * 5 + c];
        }
        s->min_l[j] = ml; s->min_r[j] = mr;
        if ((w->cfg.obs_flags & ORC_OBS_MASK_LANELETS) && w->map.n_lanelets > 0)
            s->lanelet[j] = orc_current_lanelet(&w->map, &w->pos[g * 2]);
        if (w->cfg.obs_flags & ORC_OBS_BOUNDARY_POINTS) {
            /* ref_paths_agent_related.nearing_points_*[:, j]: refreshed together with short_term[:, j]
             * (world_state_rt.py:668-725), i.e. from the closest boundary index as it stands now */
            int path = w->path_id[g];
            int shift = w->near_fresh[g] ? 1 : -2;
            orc_nearing_points(w, path, w->map.left + (size_t)path * w->map.P * 2, w->idx_left[g], shift, s->near_l[j]);
            orc_nearing_points(w, path, w->map.right + (size_t)path * w->map.P * 2, w->idx_right[g], shift, s->near_r[j]);
        }
    }
}

/* observation_provider_rt.py:594-925 get_observation for the flag combinations of ORC_OBS_* (partial
 * observation, distances to the boundaries, no mask / noise).  Default flags: D = 10 + 11 k.
 * update_state (:345-588) runs at observation(agent 0) time: poses, velocities, steering and vertices of
 * ALL agents are post-step values; short-term paths / centre / boundary distances are the snapshot `s`
 * (fresh for agent 0, one step old for agents >= 1). */
static void orc_observe(const orc_world *w, int b, int i, const orc_snap *s, float *obs) {
    int N = w->N;
    const orc_cfg *c = &w->cfg;
    const int fl = c->obs_flags;
    const int bird = fl & ORC_OBS_BIRD_VIEW;
    size_t g = AG(b, i);
    const float *pi = &w->pos[g * 2];
    float rot_i = w->rot[g];
    int o = 0;
    if (bird) {
        obs[o++] = pi[0] / c->norm_pos_world[0];          /* past_pos[b, i]  :545-552, :866-872 */
        obs[o++] = pi[1] / c->norm_pos_world[1];
        obs[o++] = orc_wrap(rot_i) / c->norm_rot;         /* past_rot[b, i]  :556-558, :873-879 */
        obs[o++] = w->vel[g * 2] / c->norm_v;             /* past_vel[b, i]  :553-555 */
        obs[o++] = w->vel[g * 2 + 1] / c->norm_v;
    } else {
        /* own speed: past_vel[i,i,0] = |vel_i| * cos(wrap(0)) / v_norm  (:434-441, :880-882) */
        float vabs = orc_norm2(w->vel[g * 2], w->vel[g * 2 + 1]);    /* torch.norm(vel, dim=1) :444 */
        float rr = orc_wrap(rot_i - rot_i);
        obs[o++] = (vabs * cosf(rr)) / c->norm_v;
    }
    if (fl & ORC_OBS_STEERING) obs[o++] = orc_wrap(w->steering[g]) / c->norm_rot;    /* :354-358, :394, :883-887 */
    for (int k = 0; k < ORC_NST; k++) {                   /* own short-term path :444-452 / :567-574 */
        if (bird) {
            obs[o++] = s->short_term[i][2 * k] / c->norm_pos_world[0];
            obs[o++] = s->short_term[i][2 * k + 1] / c->norm_pos_world[1];
        } else {
            float loc[2];
            orc_local(pi, rot_i, &s->short_term[i][2 * k], loc);
            obs[o++] = loc[0] / c->norm_pos;
            obs[o++] = loc[1] / c->norm_pos;
        }
    }
    if (!(fl & ORC_OBS_NO_DIST_CENTER)) obs[o++] = s->d_ref[i] / c->norm_dist;       /* :373-375 */
    if (fl & ORC_OBS_BOUNDARY_POINTS) {                   /* :453-470 / :575-588, :903-925 */
        for (int side = 0; side < 2; side++)
            for (int k = 0; k < ORC_NNB; k++) {
                const float *q = side ? &s->near_r[i][2 * k] : &s->near_l[i][2 * k];
                if (bird) {
                    obs[o++] = q[0] / c->norm_pos_world[0];
                    obs[o++] = q[1] / c->norm_pos_world[1];
                } else {
                    float loc[2];
                    orc_local(pi, rot_i, q, loc);
                    obs[o++] = loc[0] / c->norm_pos;
                    obs[o++] = loc[1] / c->norm_pos;
                }
            }
    } else {
        obs[o++] = s->min_l[i] / c->norm_dist;            /* :376-383 */
        obs[o++] = s->min_r[i] / c->norm_dist;
    }
    /* torch.topk(distances.agents[:, i], k, largest=False) :627-636 */
    int used[ORC_MAX_AGENTS] = {0};
    for (int kk = 0; kk < c->k_near; kk++) {
        int bj = -1;
        float bd = INFINITY;
        for (int j = 0; j < N; j++)
            if (!used[j] && (bj < 0 || w->d_agents[g * N + j] < bd)) { bd = w->d_agents[g * N + j]; bj = j; }
        used[bj] = 1;
        size_t gj = AG(b, bj);
        const float *pj = &w->pos[gj * 2];
        float rr = orc_wrap(w->rot[gj] - rot_i);          /* :427 */
        /* is_apply_mask :638-668: a neighbour at or beyond distance_mask_agents is masked (position-like entries := 1,
         * angles / velocities := 0, :682-749).  The lanelet-relation mask is live only where the lanelet assignment is
         * computed — bird view (:585-588) — and only on maps that know neighbouring lanelets (OSM, parse_osm.py:
         * 257-262): ORC_OBS_MASK_LANELETS, set by the host layer for exactly that case. */
        int masked = (fl & ORC_OBS_MASK) && (bd >= c->mask_distance);
        /* :646-664 + map_manager.py:91-119: also masked if its lanelet is not the ego's or a neighbour of it */
        if ((fl & ORC_OBS_MASK_LANELETS) && w->map.n_lanelets > 0)
            masked |= !w->map.lanelet_adj[(size_t)s->lanelet[i] * w->map.n_lanelets + s->lanelet[bj]];
#define MSK(v, m) (masked ? (m) : (v))
        if (fl & ORC_OBS_CENTRES) {                       /* :826-836: pos, rot, length, width */
            if (bird) {
                obs[o++] = MSK(pj[0] / c->norm_pos_world[0], 1.0f);
                obs[o++] = MSK(pj[1] / c->norm_pos_world[1], 1.0f);
                obs[o++] = MSK(orc_wrap(w->rot[gj]) / c->norm_rot, 0.0f);
            } else {
                float loc[2];
                orc_local(pi, rot_i, pj, loc);            /* :418-424 */
                obs[o++] = MSK(loc[0] / c->norm_pos, 1.0f);
                obs[o++] = MSK(loc[1] / c->norm_pos, 1.0f);
                obs[o++] = MSK(rr / c->norm_rot, 0.0f);
            }
            obs[o++] = (2.0f * c->half_length) / c->norm_dist_agent;  /* :359-371, :388-393; 2*fl(L/2) == fl(L) */
            obs[o++] = (2.0f * c->half_width) / c->norm_dist_agent;
        } else {
            for (int v = 0; v < 4; v++) {                 /* vertices :484-492 / :559-566 */
                const float *pv = &w->vertices[(gj * 5 + v) * 2];
                if (bird) {
                    obs[o++] = MSK(pv[0] / c->norm_pos_world[0], 1.0f);
                    obs[o++] = MSK(pv[1] / c->norm_pos_world[1], 1.0f);
                } else {
                    float loc[2];
                    orc_local(pi, rot_i, pv, loc);
                    obs[o++] = MSK(loc[0] / c->norm_pos, 1.0f);
                    obs[o++] = MSK(loc[1] / c->norm_pos, 1.0f);
                }
            }
        }
        if (bird) {
            obs[o++] = MSK(w->vel[gj * 2] / c->norm_v, 0.0f);        /* :553-555 */
            obs[o++] = MSK(w->vel[gj * 2 + 1] / c->norm_v, 0.0f);
        } else {
            float vabs = orc_norm2(w->vel[gj * 2], w->vel[gj * 2 + 1]);
            obs[o++] = MSK((vabs * cosf(rr)) / c->norm_v, 0.0f);     /* :432-441 */
            obs[o++] = MSK((vabs * sinf(rr)) / c->norm_v, 0.0f);
        }
        if (fl & ORC_OBS_STEERING) obs[o++] = MSK(orc_wrap(w->steering[gj]) / c->norm_rot, 0.0f);   /* :692-700 */
        if (!(fl & ORC_OBS_NO_DIST_AGENTS)) obs[o++] = MSK(bd / c->norm_dist, 1.0f);                /* :369-371 */
        if (fl & ORC_OBS_REF_OTHERS)                                                     /* :443-452, :716-724 */
            for (int k = 0; k < ORC_NST; k++) {
                if (bird) {
                    obs[o++] = MSK(s->short_term[bj][2 * k] / c->norm_pos_world[0], 1.0f);
                    obs[o++] = MSK(s->short_term[bj][2 * k + 1] / c->norm_pos_world[1], 1.0f);
                } else {
                    float loc[2];
                    orc_local(pi, rot_i, &s->short_term[bj][2 * k], loc);
                    obs[o++] = MSK(loc[0] / c->norm_pos, 1.0f);
                    obs[o++] = MSK(loc[1] / c->norm_pos, 1.0f);
                }
            }
    }
}

#undef MSK
/* width of the observation for the configured flags */
int orc_obs_dim(const orc_world *w) {
    const int fl = w->cfg.obs_flags;
    int own = ((fl & ORC_OBS_BIRD_VIEW) ? 5 : 1) + ((fl & ORC_OBS_STEERING) ? 1 : 0) + 2 * ORC_NST +
              ((fl & ORC_OBS_NO_DIST_CENTER) ? 0 : 1) + ((fl & ORC_OBS_BOUNDARY_POINTS) ? 4 * ORC_NNB : 2);
    int per = ((fl & ORC_OBS_CENTRES) ? 5 : 8) + 2 + ((fl & ORC_OBS_STEERING) ? 1 : 0) +
              ((fl & ORC_OBS_NO_DIST_AGENTS) ? 0 : 1) + ((fl & ORC_OBS_REF_OTHERS) ? 2 * ORC_NST : 0);
    return own + per * w->cfg.k_near;
}

/* ---------------------------------------------------------------- the step */

/* vmas Environment.step order: world.step(); for agent i: reward(i), observation(i);
 * done().  Outputs: obs [B][N][D], reward [B][N], done [B], respawn_request [B][N]. */
typedef struct {
    orc_world *w;
    float *actions, *obs, *reward;
    uint8_t *done, *respawn_request;
    int b0, b1;
} orc_job;

static void *orc_step_range(void *arg) {
    orc_job *jb = (orc_job *)arg;
    orc_world *w = jb->w;
    int N = w->N, D = orc_obs_dim(w);
    for (int b = jb->b0; b < jb->b1; b++) {
        for (int a = 0; a < N; a++) orc_dynamics(w, b, a, &jb->actions[AG(b, a) * 2]);
        orc_snap snap;
        for (int i = 0; i < N; i++) {
            jb->reward[AG(b, i)] = orc_reward(w, b, i);
            if (i == 0) orc_take_snapshot(w, b, &snap);
            orc_observe(w, b, i, &snap, &jb->obs[AG(b, i) * D]);
        }
        /* done() road_traffic.py:1368-1487: training mode :1449-1457, testing mode :1429-1447 */
        int any = (w->step[b] == w->cfg.max_steps - 1);
        if (w->cfg.fixed_duration > 0.0f) {   /* :1388-1393: t = timer.step * dt (int tensor * python float -> fp32) */
            volatile float t = (float)w->step[b] * w->cfg.dt;
            any |= (fmodf(t, w->cfg.fixed_duration) == 0.0f) && (t != 0.0f);
        }
        if (!w->cfg.testing_mode)
            for (int a = 0; a < N; a++) {
                any |= w->col_lane[AG(b, a)];
                for (int j = 0; j < N; j++) any |= w->col_agents[AG(b, a) * N + j];
            }
        jb->done[b] = (uint8_t)any;
        for (int a = 0; a < N; a++) {
            int leave = w->col_entry[AG(b, a)] | w->col_exit[AG(b, a)];
            if (w->cfg.testing_mode) {
                int hit = w->col_lane[AG(b, a)];
                for (int j = 0; j < N; j++) hit |= w->col_agents[AG(b, a) * N + j];
                jb->respawn_request[AG(b, a)] = (uint8_t)(!any && (hit | leave));
            } else {
                jb->respawn_request[AG(b, a)] = (uint8_t)(!w->cfg.is_cpm_entire && !any && leave);
            }
        }
    }
    return NULL;
}

/* n_threads <= 1: run inline; otherwise envs are split in contiguous ranges over pthreads (envs are
 * independent; the image's gcc has no libgomp). */
void orc_step(orc_world *w, float *actions, float *obs, float *reward, uint8_t *done, uint8_t *respawn_request,
              int n_threads) {
    int B = w->B;
    if (n_threads > B) n_threads = B;
    if (n_threads <= 1) {
        orc_job jb = {w, actions, obs, reward, done, respawn_request, 0, B};
        orc_step_range(&jb);
        return;
    }
    pthread_t th[256];
    orc_job jobs[256];
    if (n_threads > 256) n_threads = 256;
    for (int t = 0; t < n_threads; t++) {
        jobs[t] = (orc_job){w, actions, obs, reward, done, respawn_request,
                            (int)((long long)B * t / n_threads), (int)((long long)B * (t + 1) / n_threads)};
        pthread_create(&th[t], NULL, orc_step_range, &jobs[t]);
    }
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
}

/* Observation pass right after a reset (vmas Environment.reset_at -> get_from_scenario(obs only)):
 * observation(0) snapshots a world state in which everything is fresh. */
void orc_fresh_obs(orc_world *w, float *obs) {
    int N = w->N, D = orc_obs_dim(w);
    for (int b = 0; b < w->B; b++) {
        orc_snap snap;
        orc_take_snapshot(w, b, &snap);
        for (int i = 0; i < N; i++) orc_observe(w, b, i, &snap, &obs[AG(b, i) * D]);
    }
}

/* ---------------------------------------------------------------- reset (own RNG; distribution of
 * world_state_rt_sim.py:215-311: uniform path, uniform point in [3, n/2), >= min distance apart) */
static uint64_t orc_rng_next(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* place agent `a` of env b at (path, point) with speed; world_state_rt_sim.py:143-213 */
void orc_place(orc_world *w, int b, int a, int path, int point, float speed) {
    int N = w->N;
    const orc_map *m = &w->map;
    size_t g = AG(b, a);
    w->path_id[g] = path;
    w->pos[g * 2] = m->center[((size_t)path * m->P + point) * 2];
    w->pos[g * 2 + 1] = m->center[((size_t)path * m->P + point) * 2 + 1];
    float yaw = m->yaw[(size_t)path * m->P + point];
    w->rot[g] = yaw;
    w->steering[g] = 0.0f;
    w->sideslip[g] = 0.0f;
    w->speed[g] = speed;
    w->vel[g * 2] = speed * cosf(0.0f + yaw);
    w->vel[g * 2 + 1] = speed * sinf(0.0f + yaw);
}

/* full reset of env b (agent < 0) or respawn of one agent; returns 0, or -1 if no feasible placement
 * was found in `max_tries` (the reference would spin forever). */
int orc_reset_env(orc_world *w, int b, int agent, int path_lo, int path_hi, uint64_t *rng, int max_tries) {
    int N = w->N;
    const orc_map *m = &w->map;
    for (int a = 0; a < N; a++) {
        if (agent >= 0 && a != agent) continue;
        int ok = 0;
        for (int tries = 0; tries < max_tries && !ok; tries++) {
            int path = path_lo + (int)(orc_rng_next(rng) % (uint64_t)(path_hi - path_lo));
            int end = m->n_center[path] / 2;
            int point = 3 + (int)(orc_rng_next(rng) % (uint64_t)(end - 3));
            const float *p = &m->center[((size_t)path * m->P + point) * 2];
            ok = 1;
            int lim = (agent >= 0) ? N : a;
            for (int o = 0; o < lim && ok; o++) {
                if (o == a) continue;
                float dx = p[0] - w->pos[AG(b, o) * 2], dy = p[1] - w->pos[AG(b, o) * 2 + 1];
                float dsq = dx * dx + dy * dy;
                if (!(dsq >= w->cfg.reset_min_dist_sq)) ok = 0;
            }
            if (ok) {
                float speed = (float)((orc_rng_next(rng) >> 40) * (1.0 / 16777216.0)) * w->cfg.max_speed;
                orc_place(w, b, a, path, point, speed);
            }
        }
        if (!ok) return -1;
    }
    if (agent < 0) w->step[b] = 0;
    orc_refresh_env(w, b, agent);
    return 0;
}

/* ---------------------------------------------------------------- lifetime */

orc_world *orc_create(int B, int N, const orc_map *map, const orc_cfg *cfg) {
    if (N > ORC_MAX_AGENTS || map->P > 1024) return NULL;
    orc_world *w = (orc_world *)calloc(1, sizeof(orc_world));
    w->B = B; w->N = N; w->map = *map; w->cfg = *cfg;
    size_t BN = (size_t)B * N;
    w->pos = calloc(BN * 2, 4); w->rot = calloc(BN, 4); w->speed = calloc(BN, 4);
    w->steering = calloc(BN, 4); w->vel = calloc(BN * 2, 4); w->sideslip = calloc(BN, 4);
    w->path_id = calloc(BN, 4); w->vertices = calloc(BN * 10, 4); w->d_agents = calloc(BN * N, 4);
    w->d_ref = calloc(BN, 4); w->d_left = calloc(BN * 5, 4); w->d_right = calloc(BN * 5, 4);
    w->d_bound = calloc(BN, 4); w->idx_ref = calloc(BN, 4); w->short_term = calloc(BN * 6, 4);
    w->idx_left = calloc(BN, 4); w->idx_right = calloc(BN, 4); w->near_fresh = calloc(BN, 1);
    w->prev_pos = calloc(BN * 2, 4); w->col_agents = calloc(BN * N, 1); w->col_lane = calloc(BN, 1);
    w->col_entry = calloc(BN, 1); w->col_exit = calloc(BN, 1); w->step = calloc(B, 4);
    return w;
}

void orc_destroy(orc_world *w) {
    if (!w) return;
    free(w->pos); free(w->rot); free(w->speed); free(w->steering); free(w->vel); free(w->sideslip);
    free(w->path_id); free(w->vertices); free(w->d_agents); free(w->d_ref); free(w->d_left);
    free(w->d_right); free(w->d_bound); free(w->idx_ref); free(w->short_term); free(w->prev_pos);
    free(w->col_agents); free(w->col_lane); free(w->col_entry); free(w->col_exit); free(w->step);
    free(w->idx_left); free(w->idx_right); free(w->near_fresh);
    free(w);
}

/* raw field access for the ctypes wrapper (name -> pointer) */
void *orc_field(orc_world *w, const char *name) {
#define F(n) if (!strcmp(name, #n)) return (void *)w->n;
    F(pos) F(rot) F(speed) F(steering) F(vel) F(sideslip) F(path_id) F(vertices) F(d_agents) F(d_ref)
    F(d_left) F(d_right) F(d_bound) F(idx_ref) F(short_term) F(prev_pos) F(col_agents) F(col_lane)
    F(col_entry) F(col_exit) F(step) F(idx_left) F(idx_right) F(near_fresh)
#undef F
    return NULL;
}

/* function-level entry points for known-answer tests */
float orc_test_perp(const float *p, const float *poly, int P, int n, int *idx) { return orc_perp(p, poly, P, n, idx); }
int orc_test_interx(const float *L1, int n1, const float *L2, int n2) { return orc_interx(L1, n1, L2, n2); }
float orc_test_wrap(float a) { return orc_wrap(a); }
void orc_test_local(const float *pi, float rot_i, const float *pj, float *out) { orc_local(pi, rot_i, pj, out); }
float orc_test_dec(float x, float x0, float x1) { return orc_dec(x, x0, x1); }
void orc_test_path_points(const float *poly, int P, int n, int is_loop, int idx, int count, int interval, int shift,
                          float *out) {
    orc_path_points(poly, P, n, is_loop, idx, count, interval, shift, out);
}
void orc_test_rect(float hl, float hw, const float *pos, float yaw, float *out) {
    orc_cfg c; memset(&c, 0, sizeof c); c.half_length = hl; c.half_width = hw; orc_rect(&c, pos, yaw, out);
}
/* MTV distance of two rectangles given as [>=4][2] vertex arrays (helper_scenario.py:1030-1138) */
float orc_test_mtv(const float *vi, const float *vj) { return orc_mtv(vi, vj); }
