// sgb_kernels.cuh — device code of libsigmarl_b200: the fused road-traffic environment step for sm_100a.
//
// One persistent CTA per SM (256 agent slots, 256 * G threads).  Each CTA bulk-copies (TMA, cp.async.bulk +
// mbarrier) the packed map blob into shared memory once, then loops over tiles of whole envs; groups of warps walk
// the phases together (named barriers) so that one phase's code stays in the instruction cache:
//   phase A  one thread per agent : kinematic-bicycle Euler tick, rectangle vertices  -> smem
//   phase B  G lanes per agent    : point->polyline distances (centre line, left/right boundary) and
//                                   rectangle-vs-boundary crossing tests out of the smem map: hint chunk, vote
//                                   over per-chunk boxes / direction cones, walk of the voted chunks;
//                                   warp-shuffle reductions; then the rectangle pairs of the env
//   phase C  G lanes per agent    : centre distances / TTC against the other agents of the env, k-nearest
//                                   selection, reward, observation (written in place), next carry, info block
//   phase D  per env              : done flag, step counter, evaluation counters (warp ballots)
// In refresh mode (MODE 1) the same kernel rebuilds carry / observation from the current pose; after a device
// reset phase B is skipped (spawn table, see place_agent).
//
// Arithmetic contract (see DESIGN.md "Exactness"): IEEE sqrt/div, no fast-math.  FMA contraction is
// left ON for the compiler, but every value that feeds an argmin, a strict-sign collision predicate
// or the integrated state is written with never-contracted __fmul_rn/__fadd_rn/__fsub_rn in the
// reference's operation order (helper_scenario.py:829-889, :1148-1229), so those decisions are bit-identical
// to the reference given the same inputs.  Pruning never changes a result: distance chunks are skipped only
// when a lower bound exceeds the running best by a margin 30x larger than the fp32 evaluation error, and the
// exact crossing predicate is skipped only for segments that are certified unable to fire it (far from the
// rectangle and not collinear with an edge, or whose line misses the rectangle) — cfg.exhaustive = 1 bypasses
// all of it and must give bit-identical outputs.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sigmarl_b200.h"

namespace sgb {

#ifndef SGB_THREADS
#define SGB_THREADS 1024
#endif
#ifndef SGB_SYNC_WARPS          // warps per phase-aligned group in the step kernel (0/1 = free-running warps)
#define SGB_SYNC_WARPS 8
#endif
#ifndef SGB_SYNC_WARPS_REFRESH  // same for the refresh kernel.  Free-running: what it mostly runs is the spawn-table refresh of
#define SGB_SYNC_WARPS_REFRESH 0 // reset envs (no scans, little code), where the group barriers only cost (reset + refresh at the
#endif                           // headline shape 0.0668 -> 0.0647 ms); with scans, groups of 8 were 3 % faster (0.137 -> 0.133)
// Threads per CTA (one CTA per SM: the map blob fills most of the shared memory) for G = 4 lanes per agent.  The CTA
// always holds kSlots agent slots — that fixes the shared-memory footprint next to the map — so with fewer lanes per
// agent it has fewer threads, and more registers each: G = 4 -> 1024 threads, G = 2 -> 512, G = 1 -> 256.
constexpr int kThreads = SGB_THREADS;
constexpr int kSlots = kThreads / 4;
__host__ __device__ constexpr int cta_threads(int G) { return kSlots * G; }
constexpr int kChunk = 8;            // polyline segments per bounding-box chunk
constexpr int kExt = 6;              // extension points behind a centre line (3 short-term pts x interval 2)
constexpr float kDistMargin = 1e-4f; // [m]  >> fp32 error of a point-segment distance (~3e-6)
constexpr float kSignMargin = 1e-4f; // [m^2] >> fp32 error of an edge-line sign function (~6e-6)
// Crossing tests against FAR segments (DESIGN.md "Exactness"): a segment whose distance to the rectangle exceeds
// kFarMargin cannot truly cross it, and interX can only fire on it through fp32 sign noise, which needs the
// segment's line and an edge's line to be collinear within (eps_i + eps_j) / kFarMargin ~ 2e-3 rad (eps ~ 1e-5 m =
// evaluation error of a line function / its gradient).  kCollinear is 5x that bound.
constexpr float kFarMargin = 0.01f;  // [m]
constexpr float kCollinear = 0.01f;  // |sin(angle between segment and edge direction)|

// ---- packed map blob (global memory -> shared memory, byte-identical) ---------------------------------
struct BlobHeader {      // 32 bytes
    int32_t n_paths;
    int32_t path_off;    // byte offsets from blob start
    int32_t pts_off;
    int32_t box_off;
    int32_t total_bytes; // multiple of 16
    int32_t cone_off;    // half2 (mid angle mod pi, half width incl. slack) per BOUNDARY chunk
    int32_t pad[2];
};
struct PathRec {         // 48 bytes; point offsets in float2 units, box offsets in float4 units
    int32_t c_off, n_c;  // centre line: n_c real points followed by kExt extension points
    int32_t l_off, n_l;
    int32_t r_off, n_r;
    int32_t cbox, lbox, rbox;
    int32_t is_loop;
    int32_t lcone, rcone; // cone offsets (half2 units) of the left / right boundary chunks
};

struct Params {
    sgb_config cfg;
    sgb_buffers buf;
    const unsigned char* blob; // device copy of the packed map
    const int32_t* env_list;   // refresh: compacted list of env indices (NULL = all B envs, in order)
    const int32_t* env_count;  // device pointer: number of entries of env_list
    int32_t B, N, D;
    int32_t blob_bytes;
    int32_t mode;              // 0 = step, 1 = refresh
    int32_t write_obs;
    // refresh after a device reset: every agent of a listed env sits on a spawn point, whose centre / boundary
    // distances come from the spawn table (reset_kernel stored them): phase B is skipped
    int32_t skip_scan;
    const float* fresh;        // [B,N,4] dLc, dRc, m4L, m4R of the spawn pose (written by reset_kernel)
    // derived on the host once per launch: kernel parameters live in the constant bank and cost no registers
    float rect_radius;         // circumradius of the rectangle * 1.0001 (conservative reach for the pruning bounds)
    float near2;               // (rect_radius + kFarMargin)^2
    float r_pos, r_v, r_dist;  // reciprocals of the observation normalisers
    // lanelet table (SGB_OBS_MASK_LANELETS; sgb_set_lanelets), global memory — appended, so nothing above moves
    const float2* lanelet_xy;      // centre lines of all lanelets, concatenated
    const int32_t* lanelet_off;    // [n_lanelets + 1]
    const uint8_t* lanelet_adj;    // [n_lanelets][n_lanelets]
    int32_t n_lanelets, lanelet_max_len;
    // observation noise key (appended): API-call counter and the global index of this launch's env 0
    uint64_t noise_epoch;
    int64_t env_base;
    float band_l, band_w;      // half_length / half_width + 1 mm: the edge-line bands of the far-candidate vote
};

// ---- small helpers ---------------------------------------------------------------------------------------
// @region small helpers (msub2 etc.)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// Programmatic dependent launch (sm_90+): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// start while its predecessor in the stream is still running — as soon as every CTA of the predecessor has executed
// launch_dependents (or exited) and an SM has room.  Everything before pdl_wait() must touch nothing the predecessor
// writes (here: mbarrier set-up and the bulk copy of the read-only map); pdl_wait() returns once the predecessor has
// completed and its writes are visible.  EVERY CTA waits before it exits, so that "this grid has completed" implies "its
// predecessors have" for the grid that comes next.  Launched without the attribute both are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(phase)
        : "memory");
}

// Separately-rounded fp32 operations (never contracted into FMA).  The file is compiled with FMA contraction ON;
// every value that feeds an argmin, a strict-sign predicate or the integrated state is written with these so
// that it rounds exactly like the reference's one-ATen-op-at-a-time evaluation.
#ifdef __CUDA_ARCH__
#define SGB_UNROLL _Pragma("unroll")
#define SGB_INF __int_as_float(0x7f800000)
#define SGB_FFS(x) __ffs(x)
#define SGB_SHFL_XOR(v, m) __shfl_xor_sync(0xffffffffu, v, m)
#else
// host build of the scan functions (one lane per agent: the group reductions are the identity) — backs the
// sgb_debug_scan_* hooks, which check pruned == exhaustive without a device
#define SGB_UNROLL
#define SGB_INF __builtin_huge_valf()
#define SGB_FFS(x) __builtin_ffs((int)(x))
#define SGB_SHFL_XOR(v, m) (v)
#endif
#define SGB_HD __host__ __device__
// work counters of the host build (sgb_debug_scan_counters): 0 segment evaluations of the centre scan, 1 of the
// boundary scans (each covers 5 points), 2 chunk boxes tested in the votes, 3 exact crossing predicates, 4 scans
#ifdef SGB_TEST_HOOKS
static thread_local long long g_scan_counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // host build of the test library only
#endif
#if defined(SGB_TEST_HOOKS) && !defined(__CUDA_ARCH__)
#define SGB_COUNT(i, n) (g_scan_counters[i] += (n))
#else
#define SGB_COUNT(i, n)
#endif
// (__host__ too: the host build of the same source backs the arithmetic self-test hook sgb_debug_mtv_distance; the host
// compiler runs with -ffp-contract=off, so the plain expressions round separately there as well.)
#ifdef __CUDA_ARCH__
__host__ __device__ __forceinline__ float mulr(float a, float b) { return __fmul_rn(a, b); }
__host__ __device__ __forceinline__ float addr(float a, float b) { return __fadd_rn(a, b); }
__host__ __device__ __forceinline__ float subr(float a, float b) { return __fsub_rn(a, b); }
__host__ __device__ __forceinline__ float fmar(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__host__ __device__ __forceinline__ float divr(float a, float b) { return __fdiv_rn(a, b); }
#else
__host__ __device__ __forceinline__ float mulr(float a, float b) { return a * b; }
__host__ __device__ __forceinline__ float addr(float a, float b) { return a + b; }
__host__ __device__ __forceinline__ float subr(float a, float b) { return a - b; }
__host__ __device__ __forceinline__ float fmar(float a, float b, float c) { return fmaf(a, b, c); }
__host__ __device__ __forceinline__ float divr(float a, float b) { return a / b; }
#endif
// a*b - c*d and a*b + c*d with three roundings
__host__ __device__ __forceinline__ float msub2(float a, float b, float c, float d) { return subr(mulr(a, b), mulr(c, d)); }
__host__ __device__ __forceinline__ float madd2(float a, float b, float c, float d) { return addr(mulr(a, b), mulr(c, d)); }

__device__ __noinline__ void sincos_ool(float x, float* sn, float* cs) { sincosf(x, sn, cs); }
__device__ __noinline__ float tan_ool(float x) { return tanf(x); }
__device__ __noinline__ float atan_ool(float x) { return atanf(x); }

SGB_HD __forceinline__ float clampf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

// torch `%` (sign of the divisor) for a positive divisor
SGB_HD __forceinline__ float pymod(float a, float m) {
    float r = fmodf(a, m);
    if (r != 0.0f && r < 0.0f) r += m;
    return r;
}

// helper_scenario.py:960-996 decreasing_fcn(type="linear")
SGB_HD __forceinline__ float dec_lin(float x, float x0, float x1) {
    x = clampf(x, x0, x1);
    return 1.0f - (x - x0) / (x1 - x0);
}

// Running minimum of d = sqrt(q) with torch.min's "first minimal index" rule, for segments visited in ANY
// order (the hint chunk is scanned first): a candidate wins if d is smaller, or equal with a smaller index.
// sqrt is monotone, so min d = sqrt(min q); distinct q within ~1.2e-7 relative can round to the same d, hence
// the exact (d, idx) comparison runs for every q <= qmin * (1 + 5e-7) and is skipped (no sqrt) otherwise.
// @region Best (centre argmin)
struct Best {
    float qmin, qlim, d;
    int idx;
    SGB_HD __forceinline__ void init() { qmin = SGB_INF; qlim = qmin; d = qmin; idx = 0x7fffffff; }
    SGB_HD __forceinline__ void upd(float qq, int s) {
        if (qq <= qlim) {
            float dd = sqrtf(qq);
            if (dd < d || (dd == d && s < idx)) { d = dd; idx = s; }
            if (qq < qmin) { qmin = qq; qlim = qq * 1.0000005f; }
        }
    }
};
// boundary distances feed only continuous outputs (observation / reward, 1e-5 tolerance), never an argmin
// or a predicate, so they are tracked as min q = min d^2 and square-rooted once: min sqrt(q) == sqrt(min q).
// @region BestQ
struct BestQ {
    float q;
    SGB_HD __forceinline__ void init() { q = SGB_INF; }
    SGB_HD __forceinline__ void upd(float qq) { q = fminf(q, qq); }
    // conservative "a box at squared distance lb2 cannot improve on q": lb > sqrt(q) + 3e-6 is implied
    SGB_HD __forceinline__ bool box_useless(float lb2) const { return lb2 > q * 1.01f + 1e-7f; }
};

// squared point-segment distance, operation order of helper_scenario.py:856-871 (IEEE division): used for the  @region seg_q exact (centre)
// centre line, whose argmin must be bit-identical to the reference
SGB_HD __forceinline__ float seg_q(float ax, float ay, float lx, float ly, float len2, float px, float py) {
    const float vx = subr(px, ax), vy = subr(py, ay);
    float t = madd2(vx, lx, vy, ly) / len2;
    t = clampf(t, 0.0f, 1.0f);
    const float cx = addr(ax, mulr(lx, t)), cy = addr(ay, mulr(ly, t));
    const float ex = subr(cx, px), ey = subr(cy, py);
    // torch.norm over the two components (:871) is sqrt(fma(ey, ey, ex * ex)) in ATen's CPU reduction (FMA-capable
    // x86): the second square is not rounded on its own.  1 ulp from ex*ex + ey*ey in 8 % of the cases — which decides
    // the argmin when the foot of the perpendicular sits next to a vertex (measured on 1.5e6 pairs: 0 mismatches, DESIGN.md "Exactness")
    return fmar(ey, ey, mulr(ex, ex));
}
// same with the projection parameter computed as dot * (1/len2): t differs from the reference's quotient by  @region seg_q_r (boundary)
// <= 1.5 ulp, i.e. the distance by ~1e-8 m.  Used for the boundaries only (one reciprocal shared by 5 points).
SGB_HD __forceinline__ float sat01(float x) {
#ifdef __CUDA_ARCH__
    return __saturatef(x);            // folds into the producing FMUL as .SAT
#else
    return fminf(fmaxf(x, 0.0f), 1.0f);
#endif
}
// (sx, sy) = l / |l|^2: the projection parameter is v . s, scaled once per segment instead of once per point
SGB_HD __forceinline__ float seg_q_r(float ax, float ay, float lx, float ly, float sx, float sy, float px, float py) {
    const float vx = px - ax, vy = py - ay;
    const float t = sat01(vx * sx + vy * sy);
    const float ex = fmaf(lx, t, -vx), ey = fmaf(ly, t, -vy);   // (a + l t) - p, one FMA per component
    return ex * ex + ey * ey;
}

// 1/x with MUFU.RCP (<= 1 ulp): only used where the result feeds continuous outputs  @region rcp/box_lb
SGB_HD __forceinline__ float rcp_fast(float x) {
#ifdef __CUDA_ARCH__
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}

// sqrt with MUFU.SQRT (max relative error 2^-23): pruning bounds (which carry a 1e-4 m margin) and the boundary
// distances (continuous outputs, 1e-5 tolerance) — never a value that feeds an argmin or a predicate
SGB_HD __forceinline__ float sqrt_fast(float x) {
#ifdef __CUDA_ARCH__
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return sqrtf(x);
#endif
}

// squared distance from a point to an axis-aligned box (lower bound for every polyline point inside it).  Boxes are
// stored as (centre x, centre y, half extent x, half extent y) — half extents rounded up so that the box still holds
// every point after the centre's rounding (pack_map) — so the test is |p - c| - h per axis, and the crossing vote
// reuses p - c.
SGB_HD __forceinline__ float box_lb2_rel(float4 bx, float rx, float ry) {   // (rx, ry) = box centre - point
    const float dx = fmaxf(fabsf(rx) - bx.z, 0.0f), dy = fmaxf(fabsf(ry) - bx.w, 0.0f);
    return dx * dx + dy * dy;
}
SGB_HD __forceinline__ float box_lb2(float4 bx, float px, float py) { return box_lb2_rel(bx, bx.x - px, bx.y - py); }
SGB_HD __forceinline__ float box_lb(float4 bx, float px, float py) { return sqrtf(box_lb2(bx, px, py)); }

// The agent's rectangle as interX sees it (helper_scenario.py:1148-1229): closed 5-vertex polyline.  @region Rect.finish
struct Rect {
    float vx[4], vy[4];    // vertices 0..3 (vertex 4 == vertex 0)
    float dx[4], dy[4], S[4]; // per edge i: v[i] -> v[i+1]
    SGB_HD __forceinline__ void finish() {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int j = (i + 1) & 3;
            dx[i] = vx[j] - vx[i];
            dy[i] = vy[j] - vy[i];
            S[i] = msub2(dx[i], vy[i], dy[i], vx[i]);
        }
    }
};

// interX of the rectangle (as L1) against ONE segment a->b of L2; exact predicate.  C2 first: when all four  @region rect_cross_seg_L1
// vertices lie strictly on one side of (or on) the segment's line no edge can cross it, and the four C1 terms
// are skipped (same values as the reference would compute, just not evaluated).
SGB_HD __forceinline__ bool rect_cross_seg_L1(const float* rvx, const float* rvy, float ax, float ay, float bx, float by,
                                                  bool no_filter = false) {
    // Takes the bare vertices: the edge vectors / S_i and the bounding box are rebuilt here, in the same fp32
    // operations as Rect::finish(), because this function runs for few segments (see the gate in scan_boundary)
    // and the scan must not carry them in registers.
    const float dx2 = bx - ax, dy2 = by - ay;
    if (!no_filter) {
        // Cheapest filter first (fused arithmetic, certified by a margin): g at the bounding-box centre, and the
        // largest change of g over the box.  |g(c)| - (|dx2| hy + |dy2| hx) > 1e-5 >> fp32 error of g (~3e-7)
        // => all four vertices are strictly on one side of the segment's line => no C2 term can be true.
        const float x0 = fminf(fminf(rvx[0], rvx[1]), fminf(rvx[2], rvx[3])), x1 = fmaxf(fmaxf(rvx[0], rvx[1]), fmaxf(rvx[2], rvx[3]));
        const float y0 = fminf(fminf(rvy[0], rvy[1]), fminf(rvy[2], rvy[3])), y1 = fmaxf(fmaxf(rvy[0], rvy[1]), fmaxf(rvy[2], rvy[3]));
        const float bcx = 0.5f * (x0 + x1), bcy = 0.5f * (y0 + y1);
        const float bhx = 0.5f * (x1 - x0) + 1e-6f, bhy = 0.5f * (y1 - y0) + 1e-6f;   // + rounding slack of the centre
        const float gc = dx2 * (bcy - ay) - dy2 * (bcx - ax);
        if (fabsf(gc) > fabsf(dx2) * bhy + fabsf(dy2) * bhx + 1e-5f) return false;
    }
    const float S2 = msub2(dx2, ay, dy2, ax);
    float g[4];
#pragma unroll
    for (int i = 0; i < 4; i++) g[i] = subr(msub2(rvy[i], dx2, rvx[i], dy2), S2);
    bool c2[4];
    bool any2 = false;
#pragma unroll
    for (int i = 0; i < 4; i++) { c2[i] = (g[i] * g[(i + 1) & 3]) < 0.0f; any2 |= c2[i]; }
    if (!any2) return false;
    bool hit = false;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (!c2[i]) continue;            // C1[i] only matters where C2[i] holds (typically two of the four edges)
        const int j = (i + 1) & 3;
        const float dxi = rvx[j] - rvx[i], dyi = rvy[j] - rvy[i];
        const float Si = msub2(dxi, rvy[i], dyi, rvx[i]);
        const float fa = subr(msub2(dxi, ay, dyi, ax), Si);
        const float fb = subr(msub2(dxi, by, dyi, bx), Si);
        hit |= (fa * fb) < 0.0f;
    }
    return hit;
}

// One out-of-line copy for the boundary scans: the predicate runs for ~0.2 segments per agent-step but was inlined twice
// into scan_boundary (two segments per lane-iteration), in the middle of the hottest loop of a kernel that lives on its
// instruction cache (0.2949 -> 0.2908 ms).  Scalars by value: an array argument would force the caller's vertex
// registers into local memory.  (Moving the rectangle-pair predicate and the rank >= 2 neighbour search out of line as
// well made the kernel slower again, 0.2962 ms: the pair predicate runs too often for a call.)
#ifdef __CUDA_ARCH__
__device__ __noinline__ bool rect_cross_seg_ool(float x0, float x1, float x2, float x3, float y0, float y1, float y2, float y3,
                                                float ax, float ay, float bx, float by) {
    const float rvx[4] = {x0, x1, x2, x3}, rvy[4] = {y0, y1, y2, y3};
    return rect_cross_seg_L1(rvx, rvy, ax, ay, bx, by, true);
}
#define SGB_RECT_CROSS_SEG(rvx, rvy, ax, ay, bx, by) \
    rect_cross_seg_ool(rvx[0], rvx[1], rvx[2], rvx[3], rvy[0], rvy[1], rvy[2], rvy[3], ax, ay, bx, by)
#else
#define SGB_RECT_CROSS_SEG(rvx, rvy, ax, ay, bx, by) rect_cross_seg_L1(rvx, rvy, ax, ay, bx, by, true)
#endif

// interX(L1 = rectangle lo, L2 = rectangle hi), 4 x 4 edge pairs.  C1 first: f_i at hi's four vertices; if no  @region rect_cross_rect
// edge line of lo separates two consecutive vertices of hi there is no crossing and C2 is not evaluated.
SGB_HD __forceinline__ bool rect_cross_rect(const Rect& lo, const float* hx, const float* hy) {
    uint32_t c1 = 0; // bit 4*i + j
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float f[4];
#pragma unroll
        for (int j = 0; j < 4; j++) f[j] = subr(msub2(lo.dx[i], hy[j], lo.dy[i], hx[j]), lo.S[i]);
#pragma unroll
        for (int j = 0; j < 4; j++) c1 |= ((f[j] * f[(j + 1) & 3]) < 0.0f) ? (1u << (4 * i + j)) : 0u;
    }
    if (!c1) return false;
    bool hit = false;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int jn = (j + 1) & 3;
        const float dx2 = hx[jn] - hx[j], dy2 = hy[jn] - hy[j];
        const float S2 = msub2(dx2, hy[j], dy2, hx[j]);
        float g[4];
#pragma unroll
        for (int i = 0; i < 4; i++) g[i] = subr(msub2(lo.vy[i], dx2, lo.vx[i], dy2), S2);
#pragma unroll
        for (int i = 0; i < 4; i++) hit |= (((g[i] * g[(i + 1) & 3]) < 0.0f) & (((c1 >> (4 * i + j)) & 1u) != 0u));
    }
    return hit;
}

// map_manager.py:39-89 determine_current_lanelet for one position: the lanelet whose centre line holds the closest  @region lanelet
// point — squared distance, every product rounded (torch.sum((a - c)**2)), first minimal lanelet.  The reference pads
// the centre lines with ZEROS to the longest one (:58-66): every shorter lanelet also "has" the point (0, 0).
__host__ __device__ inline int current_lanelet(const float2* xy, const int32_t* off, int n_lanelets, int max_len, float px, float py) {
    int best = 0;
    float best_d = 3.402823466e38f;
    for (int l = 0; l < n_lanelets; l++) {
        const int o0 = off[l], n = off[l + 1] - o0;
        float dmin = 3.402823466e38f;
        for (int k = 0; k < n; k++) {
            const float2 c = xy[o0 + k];
            const float dx = subr(px, c.x), dy = subr(py, c.y);
            dmin = fminf(dmin, madd2(dx, dx, dy, dy));
        }
        if (n < max_len) dmin = fminf(dmin, madd2(px, px, py, py));
        if (dmin < best_d) { best_d = dmin; best = l; }
    }
    return best;
}
__device__ __noinline__ int current_lanelet_ool(const float2* xy, const int32_t* off, int n_lanelets, int max_len, float px, float py) {
    return current_lanelet(xy, off, n_lanelets, max_len, px, py);
}

// MTV-based distance between two rectangles (is_use_mtv_distance; helper_scenario.py:1030-1138), in the reference's  @region mtv
// operation order with every product / sum rounded on its own (the file compiles with FMA contraction on).  Per
// rectangle the two edge directions v1-v0, v2-v1 are normalised (:1042-1045); the vertices of a are projected on the
// axes of b; outside b's projection interval a vertex contributes its signed gap on that axis, its distance is the
// norm of the two gaps (torch.norm: sqrt(fma(g1, g1, g0*g0)), see seg_q) (:1072-1077).  A vertex strictly inside b
// makes the pair's distance negative: minus the smallest projection overlap (:1079-1084, :1119-1124).
SGB_HD __forceinline__ void mtv_half(const float* ax, const float* ay, const float* bx, const float* by,
                                         float& pos_min, float& ov_min, bool& neg) {
    float ux[2], uy[2], mnb[2], mxb[2];
SGB_UNROLL
    for (int k = 0; k < 2; k++) {
        const float dx = subr(bx[k + 1], bx[k]), dy = subr(by[k + 1], by[k]);             // torch.diff :1042
        const float n = sqrtf(fmar(dy, dy, mulr(dx, dx)));                                // torch.norm :1043
        ux[k] = divr(dx, n);
        uy[k] = divr(dy, n);
        float mna = 0.0f, mxa = 0.0f;
SGB_UNROLL
        for (int v = 0; v < 4; v++) {
            const float pb = madd2(bx[v], ux[k], by[v], uy[k]);                           // (.*.).sum(dim=3) :1060-1068
            const float pa = madd2(ax[v], ux[k], ay[v], uy[k]);
            if (v == 0) { mnb[k] = mxb[k] = pb; mna = mxa = pa; }
            else { mnb[k] = fminf(mnb[k], pb); mxb[k] = fmaxf(mxb[k], pb); mna = fminf(mna, pa); mxa = fmaxf(mxa, pa); }
        }
        const float ov = subr(fminf(mxb[k], mxa), fmaxf(mnb[k], mna));                    // :1079
        ov_min = (k == 0) ? ov : fminf(ov_min, ov);
    }
SGB_UNROLL
    for (int v = 0; v < 4; v++) {
        float gap[2];
        bool inside = true;
SGB_UNROLL
        for (int k = 0; k < 2; k++) {
            const float pa = madd2(ax[v], ux[k], ay[v], uy[k]);                           // same ops as above: same value
            const float lo = (pa <= mnb[k]) ? subr(pa, mnb[k]) : 0.0f;                    // :1072-1076
            const float hi = (pa >= mxb[k]) ? subr(mxb[k], pa) : 0.0f;
            gap[k] = addr(lo, hi);
            inside = inside && (pa > mnb[k]) && (pa < mxb[k]);                            // :1081-1083
        }
        const float d = sqrtf(fmar(gap[1], gap[1], mulr(gap[0], gap[0])));                // torch.norm(dim=2) :1077
        pos_min = fminf(pos_min, d);
        neg = neg || (inside && fabsf(ov_min) > 0.0f);                                    // (-overlap * inside).abs() > 0
    }
}

// rectangle of a pose exactly as phase A builds it (get_rectangle_vertices, helper_scenario.py:742-826)
SGB_HD __forceinline__ void rect_of_pose(float x, float y, float cy, float sy, float hl, float hw, float* vx, float* vy) {
    const float nsy = -sy;
    const float bxs[4] = {hl, hl, -hl, -hl};
    const float bys[4] = {hw, -hw, -hw, hw};
SGB_UNROLL
    for (int k = 0; k < 4; k++) {
        vx[k] = addr(madd2(cy, bxs[k], nsy, bys[k]), x);
        vy[k] = addr(madd2(sy, bxs[k], cy, bys[k]), y);
    }
}

SGB_HD __forceinline__ float mtv_from_vertices(const float* ax, const float* ay, const float* bx, const float* by) {
    float pos_min = 3.402823466e38f, ov_j, ov_i;    // every per-vertex distance is finite
    bool neg = false;
    mtv_half(ax, ay, bx, by, pos_min, ov_j, neg);   // i's vertices on j's axes
    mtv_half(bx, by, ax, ay, pos_min, ov_i, neg);   // j's vertices on i's axes
    return neg ? -fminf(ov_j, ov_i) : pos_min;      // :1113-1124
}

// out of line: called from the MTV instantiations only, keeps their register allocation at the level of the others
__device__ __noinline__ float mtv_distance(float xi, float yi, float ci, float si, float xj, float yj, float cj, float sj,
                                           float hl, float hw) {
    float ax[4], ay[4], bx[4], by[4];
    rect_of_pose(xi, yi, ci, si, hl, hw, ax, ay);
    rect_of_pose(xj, yj, cj, sj, hl, hw, bx, by);
    return mtv_from_vertices(ax, ay, bx, by);
}

// @region group shuffles
template <int G>
SGB_HD __forceinline__ float group_min(float v) {
#pragma unroll
    for (int m = 1; m < G; m <<= 1) v = fminf(v, SGB_SHFL_XOR(v, m));
    return v;
}
template <int G>
SGB_HD __forceinline__ uint32_t group_or(uint32_t v) {
#pragma unroll
    for (int m = 1; m < G; m <<= 1) v |= SGB_SHFL_XOR(v, m);
    return v;
}
template <int G>
SGB_HD __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int m = 1; m < G; m <<= 1) v += SGB_SHFL_XOR(v, m);
    return v;
}

// ---- per-tile shared arrays (SoA, one slot per agent of the tile) --------------------------------------  @region tile smem
struct TileSmem {
    float* px;  float* py;      // post-step (or current, in refresh mode) centre position
    float* ox;  float* oy;      // pre-step position (reward progress)
    float* cs;  float* sn;      // cos / sin of the heading
    float* vx;  float* vy;  float* vabs;
    float* psim;                // heading mod pi (direction-cone tests)
    float* vtx;                 // [8][A]: x0..x3, y0..y3
    float* car;                 // [4][A]: carry of the pre-step pose
    float* sc;                  // [6][A]: phase-B results: d_ref, idx(int), dLcg, dRcg, min4L, min4R
    float* dij;                 // [A][N] centre distances
    int* path;
    int* flags;
    int* env;                   // global env index of the slot (-1: inactive)
    int* coll;                  // bit j: rectangle crossing with agent j of the same env
};

// Dynamic shared memory layout: [ slot arrays (size fixed at compile time) | map blob | dij [A][N] ].  With the slot
// arrays first, every t.x[...] is an LDS/STS at a compile-time constant address + index: no pointer registers.
constexpr int kSlotFloats = 10 + 8 + 4 + 6 + 4;   // per slot: 10 scalars, vtx[8], car[4], sc[6], path/flags/env/coll
__host__ __device__ constexpr size_t tile_fixed_bytes(int A) { return ((size_t)A * kSlotFloats * sizeof(float) + 127) & ~(size_t)127; }
template <int A>
SGB_HD __forceinline__ void carve_tile(unsigned char* base, TileSmem& t) {
    float* f = reinterpret_cast<float*>(base);
    t.px = f; f += A; t.py = f; f += A; t.ox = f; f += A; t.oy = f; f += A;
    t.cs = f; f += A; t.sn = f; f += A; t.vx = f; f += A; t.vy = f; f += A; t.vabs = f; f += A; t.psim = f; f += A;
    t.vtx = f; f += 8 * A; t.car = f; f += 4 * A; t.sc = f; f += 6 * A;
    t.path = reinterpret_cast<int*>(f); f += A; t.flags = reinterpret_cast<int*>(f); f += A;
    t.env = reinterpret_cast<int*>(f); f += A;
    t.coll = reinterpret_cast<int*>(f); f += A;
}
__host__ __device__ inline size_t tile_smem_bytes(int A, int N) {
    return tile_fixed_bytes(A) + sizeof(float) * (size_t)A * N;   // + dij, placed behind the blob
}

// ---- phase B building blocks ------------------------------------------------------------------------------

// Phase-B scans.  Per polyline: (0) the G lanes of the agent's group scan the hint chunk together, which gives
// a tight upper bound; (1) the lanes split the chunk boxes and vote two bitmasks — chunks whose box could hold
// a closer segment, chunks whose box is not sign-certified against the rectangle's edge lines; (2) the group
// walks the set bits together, every lane taking segments lane, lane+G, ... of the chunk.  All lanes of a
// group stay busy; only the number of candidate chunks differs between the groups of a warp.

// centre line: min distance + closest index from (px,py)  @region scan_center
template <int G>
SGB_HD __forceinline__ void scan_center(const float2* __restrict__ pts, const float4* __restrict__ boxes, int n_c,
                                            int hint_seg, bool exhaustive, float px, float py, int lane, float& d_out,
                                            int& idx_out) {
    const int nseg = n_c - 1;
    const int nch = (nseg + kChunk - 1) / kChunk;
    static_assert(kChunk == 8, "hint chunk: shift by log2(kChunk)");
    const int c0 = max(0, min(hint_seg >> 3, nch - 1));
    Best b;
    b.init();
    // Every lane takes two segments per iteration (s and s+G): two independent dependency chains in flight.
    // When s+G runs past the chunk the index is clamped, i.e. a segment is evaluated twice — harmless for a min.
    // (fixed trip count kChunk / 2G — ONE pass with four lanes per agent; indices past a short last chunk are clamped)
    constexpr int NIT = kChunk / (2 * G);
    auto chunk = [&](int c) {
        const int s_last = min(c * kChunk + kChunk, nseg) - 1;
#pragma unroll 1
        for (int it = 0; it < NIT; it++) {
            const int s = min(c * kChunk + lane + it * 2 * G, s_last), s2 = min(s + G, s_last);
            const float2 a = pts[s], e = pts[s + 1], a2 = pts[s2], e2 = pts[s2 + 1];
            const float lx = e.x - a.x, ly = e.y - a.y, lx2 = e2.x - a2.x, ly2 = e2.y - a2.y;
            const float q1 = seg_q(a.x, a.y, lx, ly, madd2(lx, lx, ly, ly), px, py);
            const float q2 = seg_q(a2.x, a2.y, lx2, ly2, madd2(lx2, lx2, ly2, ly2), px, py);
            b.upd(q1, s);
            b.upd(q2, s2);
            SGB_COUNT(0, 2);
        }
    };
    chunk(c0);
    float thr = group_min<G>(b.d) + kDistMargin;
    thr = thr * thr;
    uint32_t m = 0;
    for (int c = lane; c < nch; c += 2 * G) {   // two boxes per lane and iteration
        const int c2 = c + G;
        m |= (box_lb2(boxes[c], px, py) > thr) ? 0u : (1u << c);
        m |= (c2 >= nch || box_lb2(boxes[c2 < nch ? c2 : c], px, py) > thr) ? 0u : (1u << (c2 & 31));
    }
    SGB_COUNT(2, (nch + G - 1) / G);
    SGB_COUNT(4, 1);
    if (exhaustive) m = 0xffffffffu;
    m = group_or<G>(m) & (nch >= 32 ? 0xffffffffu : ((1u << nch) - 1u)) & ~(1u << c0);
#pragma unroll 1
    while (m) {
        const int c = SGB_FFS(m) - 1;
        m &= m - 1;
        chunk(c);
    }
    // (d, idx) lexicographic min across the group == torch.min's first minimal index
    float d = b.d;
    int idx = b.idx;
#pragma unroll
    for (int k = 1; k < G; k <<= 1) {
        float od = SGB_SHFL_XOR(d, k);
        int oi = SGB_SHFL_XOR(idx, k);
        if (od < d || (od == d && oi < idx)) { d = od; idx = oi; }
    }
    d_out = d;
    idx_out = idx + 1; // helper_scenario.py:885-887
}

// boundary: distances of the centre + 4 vertices, and the rectangle-crossing flag.  Pass 0 scans the hint  @region scan_boundary
// chunk (distances + crossing gate), pass 1 the voted chunks; the segment body exists once (code size matters:
// warps run different phases at the same time and share the instruction caches).  A voted distance chunk
// evaluates all five points: a per-point refinement of the vote (which points can still improve in this box)
// was measured to cost more — instructions, registers, shuffles — than the evaluations it saved.
template <int G>
SGB_HD __forceinline__ void scan_boundary(const float2* __restrict__ pts, const float4* __restrict__ boxes,
                                              const __half2* __restrict__ cones, int n_b, int hint_seg, bool exhaustive,
                                              float px, float py, const float* cs_s, const float* sn_s,
                                              const float* psi_m_s, const float* rvx, const float* rvy, float rect_radius,
                                              float near2, float half_l, float half_w, float band_l, float band_w, bool want_dv, int lane,
                                              float& d_cg,
                                              float dv[4], float& m4_out, bool& hit_out) {
    const int nseg = n_b - 1;
    const int nch = (nseg + kChunk - 1) / kChunk;
    static_assert(kChunk == 8, "hint chunk: shift by log2(kChunk)");
    const int c0 = max(0, min(hint_seg >> 3, nch - 1));
    const float near_r = rect_radius + kFarMargin;   // segments farther than this from the centre cannot touch the rectangle
    BestQ bq[5]; // 0 = centre, 1..4 = vertices
#pragma unroll
    for (int v = 0; v < 5; v++) bq[v].init();
    bool hit = false;
    uint32_t md = 1u << c0, mx = 1u << c0;   // chunks to scan for distances / for crossings
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
        uint32_t m = md | mx;
#pragma unroll 1
        while (m) {
            const int c = SGB_FFS(m) - 1;
            m &= m - 1;
            const bool do_x = (mx >> c) & 1u, do_d = (md >> c) & 1u;
            const int s_last = min(c * kChunk + kChunk, nseg) - 1;
#pragma unroll 1
            for (int it = 0; it < kChunk / (2 * G); it++) {
                // two segments per lane-iteration (s, s+G): independent chains; a clamped duplicate is harmless (and
                // the trip count is fixed: a single pass with four lanes per agent)
                const int s = min(c * kChunk + lane + it * 2 * G, s_last), s2 = min(s + G, s_last);
                const float2 a = pts[s], e = pts[s + 1], a2 = pts[s2], e2 = pts[s2 + 1];
                const float lx = e.x - a.x, ly = e.y - a.y, len2 = lx * lx + ly * ly;
                const float lx2 = e2.x - a2.x, ly2 = e2.y - a2.y, len2b = lx2 * lx2 + ly2 * ly2;
                float q0a = SGB_INF, q0b = q0a;   // centre -> segment, +inf when not evaluated
                if (do_d) {
                    SGB_COUNT(1, 2);
                    const float rl = rcp_fast(len2), rl2 = rcp_fast(len2b);
                    const float sx = lx * rl, sy = ly * rl, sx2 = lx2 * rl2, sy2 = ly2 * rl2;
                    q0a = seg_q_r(a.x, a.y, lx, ly, sx, sy, px, py);
                    q0b = seg_q_r(a2.x, a2.y, lx2, ly2, sx2, sy2, px, py);
                    bq[0].upd(fminf(q0a, q0b));
#pragma unroll
                    for (int v = 0; v < 4; v++)
                        bq[v + 1].upd(fminf(seg_q_r(a.x, a.y, lx, ly, sx, sy, rvx[v], rvy[v]),
                                            seg_q_r(a2.x, a2.y, lx2, ly2, sx2, sy2, rvx[v], rvy[v])));
                }
                if (do_x) {
                    // Gate of the exact predicate: the segment is near the rectangle (then its chunk is a near chunk,
                    // hence a distance chunk, and q0 was evaluated), or it is collinear with an edge direction within
                    // kCollinear — the only way fp32 sign noise can fire interX on a far segment ...
                    const float cs = *cs_s, sn = *sn_s;   // heading; read here so that it is not live across the scan
                    const float cr = lx * sn - ly * cs, dt = lx * cs + ly * sn;
                    const float cr2 = lx2 * sn - ly2 * cs, dt2 = lx2 * cs + ly2 * sn;
                    const bool cola = fminf(cr * cr, dt * dt) <= (kCollinear * kCollinear) * len2;
                    const bool colb = fminf(cr2 * cr2, dt2 * dt2) <= (kCollinear * kCollinear) * len2b;
                    bool ga = (q0a <= near2) | cola, gb = (q0b <= near2) | colb;
                    // ... and its LINE must pass through the rectangle, or no C2 term can be true: the segment's line
                    // function at the centre, against its largest change over the oriented rectangle
                    // (half_l |d x u| + half_w |d . u|), with 1e-5 of slack >> the 4e-7 evaluation error of g.
                    const float gca = lx * (py - a.y) - ly * (px - a.x), gcb = lx2 * (py - a2.y) - ly2 * (px - a2.x);
                    ga &= fabsf(gca) <= half_l * fabsf(cr) + half_w * fabsf(dt) + 1e-5f;
                    gb &= fabsf(gcb) <= half_l * fabsf(cr2) + half_w * fabsf(dt2) + 1e-5f;
                    // Last, for the few segments still in: a segment whose two endpoints lie beyond the same side of the
                    // rectangle by kFarMargin (in the vehicle frame) is "far" in the sense of the certificate above
                    // even though it is close to the centre — e.g. the lane boundary next to a centred vehicle — and,
                    // not being collinear, cannot fire.
                    auto separated = [&](float ax, float ay, float ddx, float ddy) {
                        const float rx = ax - px, ry = ay - py;
                        const float lat = cs * ry - sn * rx, lon = cs * rx + sn * ry;     // endpoint a in the vehicle frame
                        const float late = lat + (cs * ddy - sn * ddx), lone = lon + (cs * ddx + sn * ddy);   // endpoint b
                        const float wl = half_w + kFarMargin, ll = half_l + kFarMargin;
                        return (fminf(lat, late) > wl) | (fmaxf(lat, late) < -wl) | (fminf(lon, lone) > ll) | (fmaxf(lon, lone) < -ll);
                    };
                    if (exhaustive || (ga && (cola || !separated(a.x, a.y, lx, ly)))) {
                        SGB_COUNT(3, 1);
                        hit |= SGB_RECT_CROSS_SEG(rvx, rvy, a.x, a.y, e.x, e.y);
                    }
                    if (exhaustive || (gb && (colb || !separated(a2.x, a2.y, lx2, ly2)))) {
                        SGB_COUNT(3, 1);
                        hit |= SGB_RECT_CROSS_SEG(rvx, rvy, a2.x, a2.y, e2.x, e2.y);
                    }
                }
            }
        }
        if (pass) break;
        // Vote.  Group-wide bound after the hint chunk (lanes that got no segment hold +inf): every point is within
        // rect_radius of the centre, so a chunk can matter for some point only if lb(centre) <= max best + radius.
        float thr;
        if (want_dv) {   // debug buffer: every vertex' own minimum must be exact
            float gq = group_min<G>(bq[0].q);
#pragma unroll
            for (int v = 1; v < 5; v++) gq = fmaxf(gq, group_min<G>(bq[v].q));
            thr = sqrt_fast(gq) + rect_radius;
        } else {
            // Consumers read the centre's minimum and the minimum OVER the four vertices (carry, observation, reward,
            // info), never a single vertex' distance.  A chunk can lower the vertex minimum m4 only if some vertex is
            // within m4 of it, i.e. lb(centre) <= m4 + radius; it can lower the centre's minimum only if lb(centre) is
            // within that.  Bounding with the BEST vertex instead of the worst one cuts the chunk evaluations per scan
            // from 2.08 to 1.57, and what a warp pays (maximum over its 8 agents) by a quarter (tests/tools/chunk_sim.py).
            const float q0 = group_min<G>(bq[0].q);
            const float q4 = group_min<G>(fminf(fminf(bq[1].q, bq[2].q), fminf(bq[3].q, bq[4].q)));
            thr = fmaxf(sqrt_fast(q4) + rect_radius, sqrt_fast(q0));
        }
        thr = fmaxf(thr + kDistMargin, near_r + kDistMargin);   // near chunks are distance chunks
        thr = thr * thr;
        md = 0; mx = 0;
        const float pi_f = 3.14159274f, half_pi = 1.57079637f;
        const float psi_m = *psi_m_s, cs = *cs_s, sn = *sn_s;
        const float acs = fabsf(cs), asn = fabsf(sn);
        SGB_COUNT(2, (nch + G - 1) / G);
        SGB_COUNT(5, 1);
        // Crossing candidates: near chunks, and far chunks on which interX could fire through sign noise.  That
        // needs a segment collinear (within kCollinear) with an edge whose LINE also passes through the chunk
        // (DESIGN.md "Exactness").  The side edges lie on the two lines parallel to the heading at lateral offset
        // +-half_width from the centre, the front/back edges on the two lines along the normal at +-half_length:
        // the chunk qualifies if its direction cone contains that direction and its box reaches into the band
        // between / around the two lines (band widened by 1 mm >> the 1e-7 m error of vertices and cos/sin).
        auto test_box = [&](int c, uint32_t bit) {   // branch-free; c0 is masked out below
            const float4 bx = boxes[c];
            const float2 cone = __half22float2(cones[c]);
            const float bcx = bx.x - px, bcy = bx.y - py;   // box centre, relative
            const float hx = bx.z, hy = bx.w;
            const float lb2 = box_lb2_rel(bx, bcx, bcy);
            md |= (lb2 > thr) ? 0u : bit;
            float da = fabsf(psi_m - cone.x);
            da = fminf(da, pi_f - da);                       // angle between heading and cone axis, mod pi
            const bool side_band = fabsf(cs * bcy - sn * bcx) <= hx * asn + (hy * acs + band_w);
            const bool face_band = fabsf(cs * bcx + sn * bcy) <= hx * acs + (hy * asn + band_l);
            const bool far_cand = ((da <= cone.y) & side_band) | (((half_pi - da) <= cone.y) & face_band);
            mx |= (!(lb2 > near2) | far_cand) ? bit : 0u;
        };
        // two boxes per lane and iteration: independent dependency chains, half the loop overhead
        for (int c = lane; c < nch; c += 2 * G) {
            const int c2 = c + G;
            test_box(c, 1u << c);
            test_box(c2 < nch ? c2 : c, c2 < nch ? (1u << c2) : 0u);
        }
        if (exhaustive) { md = 0xffffffffu; mx = 0xffffffffu; }
        {
            const uint32_t live = (nch >= 32 ? 0xffffffffu : ((1u << nch) - 1u)) & ~(1u << c0);
            md &= live; mx &= live;
        }
        md = group_or<G>(md);
        mx = group_or<G>(mx);
    }
    d_cg = sqrt_fast(group_min<G>(bq[0].q));
    // consumers read the minimum over the vertices: ONE root of the minimal square in either mode, so that the value
    // does not depend on whether a debug buffer is bound (the spawn table is built with one)
    m4_out = sqrt_fast(group_min<G>(fminf(fminf(bq[1].q, bq[2].q), fminf(bq[3].q, bq[4].q))));
    if (want_dv) {   // per-vertex distances are only reported in the debug buffer
#pragma unroll
        for (int v = 0; v < 4; v++) dv[v] = sqrt_fast(group_min<G>(bq[v + 1].q));
    } else {
        dv[0] = dv[1] = dv[2] = dv[3] = m4_out;
    }
    hit_out = group_or<G>(hit ? 1u : 0u) != 0u;
}

// index (and distance) of the rank-kk nearest agent: torch.topk(k, largest=False) order, ties -> lower index  @region kth_nearest/short_term
SGB_HD __forceinline__ int kth_nearest(const float* dij, int N, int kk, float* d_out) {
    uint32_t used = 0;
    int bj = 0;
    float bd = 0.0f;
    for (int q = 0; q <= kk; q++) {
        bj = -1;
        bd = SGB_INF;
        for (int j = 0; j < N; j++) {
            float dj = dij[j];
            if (!((used >> j) & 1u) && (bj < 0 || dj < bd)) { bd = dj; bj = j; }
        }
        used |= 1u << bj;
    }
    if (d_out) *d_out = bd;
    return bj;
}

// helper_scenario.py:892-957 short-term reference path (n_points_shift = 1, sample interval 2).  Loop paths wrap with
// (fi + 1) % n_c for fi >= n_c - 1 (:941-946).  idx is a closest-point index, 1 <= idx <= n_c - 1, so fi + 1 <= n_c + 5 <
// 2 n_c (pack_map refuses paths with fewer than 8 points) and the modulo is ONE conditional subtraction — an integer
// remainder by a run-time divisor costs ~25 instructions, and this one was unrolled into 2.7 KB of a kernel that lives
// on its instruction cache.  The clamp keeps a corrupt carry index inside the path's points + extension slots.
SGB_HD __forceinline__ int short_term_index(int fi, int n_c, bool is_loop) {
    if (is_loop && fi >= n_c - 1) { fi += 1; fi = fi >= n_c ? fi - n_c : fi; }
    return max(0, min(fi, n_c + kExt - 1));
}
SGB_HD __forceinline__ void short_term(const float2* __restrict__ cpts, int n_c, bool is_loop, int idx, float2 out[3]) {
SGB_UNROLL
    for (int k = 0; k < SGB_N_SHORT_TERM; k++) out[k] = cpts[short_term_index(2 * k + idx + 1, n_c, is_loop)];
}

// splitmix64 finaliser: the counter-based generator of the reset kernels and of the observation noise
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// helper_scenario.py:1276-1289 angle_eliminate_two_pi (fp32: the python scalars are cast to the tensor's dtype)
SGB_HD __forceinline__ float wrap_pi(float a) {
    const float pi_f = 3.14159274101257324f, two_pi = 6.28318548202514648f;
    float m = pymod(a, two_pi);
    if (m > pi_f) m -= two_pi;
    return m;
}

// carry.w = closest centre index of the pre-step pose.  With SGB_OBS_BOUNDARY_POINTS bit 30 additionally says that the
// agent's pose was written by a reset / respawn rather than by a step (the reference then holds nearing boundary
// points taken with n_points_shift = +1 instead of -2, world_state_rt.py:531-576 vs :686-725).
constexpr int kCarryIdxMask = 0x3fffffff;
constexpr int kCarryFreshBit = 0x40000000;
constexpr int kNearPts = 5;      // n_points_nearing_boundary (road_traffic.py:296-298)

// Observation width for a layout (observation_provider_rt.py:594-925; see SGB_OBS_* in the header)
__host__ __device__ inline int obs_dim_of(uint32_t fl, int k_near) {
    const int own = ((fl & SGB_OBS_BIRD_VIEW) ? 5 : 1) + ((fl & SGB_OBS_STEERING) ? 1 : 0) + 2 * SGB_N_SHORT_TERM +
                    ((fl & SGB_OBS_NO_DIST_CENTER) ? 0 : 1) + ((fl & SGB_OBS_BOUNDARY_POINTS) ? 4 * kNearPts : 2);
    const int per = ((fl & SGB_OBS_CENTRES) ? 5 : 8) + 2 + ((fl & SGB_OBS_STEERING) ? 1 : 0) +
                    ((fl & SGB_OBS_NO_DIST_AGENTS) ? 0 : 1) + ((fl & SGB_OBS_REF_OTHERS) ? 2 * SGB_N_SHORT_TERM : 0);
    return own + per * k_near;
}

// ---- the fused kernel -------------------------------------------------------------------------------------  @region kernel prologue
// Work decomposition: a WARP owns whole envs.  With G lanes per agent an env takes N*G lanes, a warp holds
// EW = 32 / (N*G) envs (N = 8, G = 4: exactly one env per warp).  After the CTA-wide map staging there is no
// CTA barrier any more: every warp runs phases A-D of its envs on its own (only __syncwarp), so warps drift
// apart and overlap their ALU / MUFU / shared-memory phases.
// MODE 0 = step, MODE 1 = refresh (rebuild carry / observation from the current pose; no dynamics, no reward).
// OV 0 = the reference's default observation layout (hard-wired, the tuned path), OV 1 = layout assembled from
// cfg.obs_flags (write_obs_general); everything else is the same code.  OV 2 = OV 1 + the MTV agent distance
// (cfg.use_mtv_distance): distances.agents from the PRE-step rectangles instead of centre distances, agents collide
// iff that distance is exactly zero, no interX between rectangles (world_state_rt_sim.py:360-396).
// MB = resident CTAs per SM the registers are budgeted for: 1 everywhere except the two-lane kernels (512 threads) on maps
// small enough for two CTAs' shared memory — 64 instead of 127 registers, twice the resident warps (8 192 x 12 on the
// on-ramp / roundabout maps: 0.089 -> 0.081 / 0.094 -> 0.086 ms; on cpm_entire, where only one CTA fits, the 127-register
// build is 5 % faster and stays in use).
template <int G, int MODE, int OV, int MB = 1>
__global__ void __launch_bounds__(cta_threads(G), MB) env_step_kernel(const Params p) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long mbar;

    const int tid = threadIdx.x;
    const int N = p.N, D = p.D;
    const sgb_config& cfg = p.cfg;
    constexpr bool step_mode = (MODE == 0);
    constexpr bool MTV = (OV == 2);
    // closest-index part of carry.w (the flag-driven layouts may keep a history bit above it, see kCarryFreshBit)
    auto idx_of = [](float w) { return __float_as_int(w) & kCarryIdxMask; };   // (used by the OV = 1 instantiations only)
    const bool bpoints = OV != 0 && (p.cfg.obs_flags & SGB_OBS_BOUNDARY_POINTS) != 0;
    constexpr int kWarps = cta_threads(G) / 32;
    constexpr int SPW = 32 / G;                 // agent slots per warp
    const int env_lanes = N * G;                // lanes per env
    const int EW = 32 / env_lanes;              // envs per warp (>= 1, checked by the host)

    // --- stage the map: one elected thread arms the mbarrier and issues the bulk copies -------------
    const uint32_t bar = smem_u32(&mbar);
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_launch_dependents();   // the next kernel of the stream may move in as soon as this grid's CTAs leave their SMs
    // Stage the map BEFORE waiting for the predecessor: the blob is read-only, so its 180 KB per SM travel while the
    // previous kernel (reset / refresh / the previous step) drains.  Only when the number of envs is known up front: over
    // a compacted list (length only known on the device, often short) a CTA stages once it knows that it has work.
    auto stage_map = [&]() {
        mbar_expect_tx(bar, (uint32_t)p.blob_bytes);
        const uint32_t dst = smem_u32(smem + tile_fixed_bytes(kSlots));
        for (int off = 0; off < p.blob_bytes; off += 32768) {
            int n = min(32768, p.blob_bytes - off);
            tma_bulk_g2s(dst + off, p.blob + off, (uint32_t)n, bar);
        }
    };
    const bool early = p.env_list == nullptr && (int)blockIdx.x * kWarps < (p.B + EW - 1) / EW;
    if (early && tid == 0) stage_map();
    pdl_wait();                // state buffers, env list and its length are the predecessor's outputs
    const int n_envs = p.env_list ? *p.env_count : p.B;
    const int n_wt = (n_envs + EW - 1) / EW;    // warp-tiles
    if ((int)blockIdx.x * kWarps >= n_wt) return;  // nothing to do for this CTA (then nothing was staged either)
    // Spawn-table refresh (observation of freshly reset envs) of a SMALL batch: phase B is skipped, all it reads of the map
    // are a path record and three centre points per agent — straight from global memory / L2 instead of staging 180 KB
    // per SM.  The host asks for it (skip_scan == 2) when the list cannot hold more than about a wave of envs: measured
    // reset + refresh at 8 192 x 8: 0.0275 -> 0.0259 ms, at 65 536 x 8 (3.6 waves of dependent L2 loads): 0.074 -> 0.076,
    // so large batches keep staging.  Compile-time false in the step kernel.
    const bool map_global = !step_mode && p.skip_scan == 2 && !bpoints;
    if (!early && !map_global && tid == 0) stage_map();
    bool map_ready = map_global;

    constexpr int AS = kSlots;                  // slot stride of the SoA arrays (all warps)
    unsigned char* const blob_s = smem + tile_fixed_bytes(AS);
    const unsigned char* const blob_b = map_global ? p.blob : blob_s;     // where this launch reads the map from
    const BlobHeader* hdr = reinterpret_cast<const BlobHeader*>(blob_b);
    TileSmem ts;
    carve_tile<AS>(smem, ts);
    ts.dij = reinterpret_cast<float*>(blob_s + ((p.blob_bytes + 127) & ~127));
    // MTV only: cos / sin of the PRE-step heading per slot, behind dij (the default layout of the slot arrays is untouched)
    float* const ocs_a = ts.dij + (size_t)AS * p.N;
    float* const osn_a = ocs_a + AS;

    const int w = tid >> 5, ln = tid & 31;
    const int slot0 = w * SPW;                  // first slot of this warp
    const int n_slots = EW * N;                 // slots this warp uses
    const int lane = ln % G;                    // lane within the agent's group
    const int sl_l = ln / G;                    // slot (within the warp) this lane works for in phases B/C
    const int i_of_lane = sl_l % N;             // its agent index within the env
    const float rect_radius = p.rect_radius;
    const float r_pos = p.r_pos, r_v = p.r_v, r_dist = p.r_dist;

    // Phase alignment (DESIGN.md "Instruction cache"): the kernel's code (~55 KB executed) does not fit the SM's
    // 32 KB instruction cache, and 32 free-running warps keep all of it live at once.  Groups of SYNCW warps
    // therefore walk the phases together (named barrier per group): at any time a group executes one phase's
    // code only.  Every warp of the CTA runs the same number of tile iterations so that the barriers match.
    constexpr int SYNCW = (step_mode ? SGB_SYNC_WARPS : SGB_SYNC_WARPS_REFRESH) * G / 4;   // always 4 groups per CTA
    static_assert(SYNCW <= 1 || (kWarps % (SYNCW > 0 ? SYNCW : 1) == 0 && kWarps / (SYNCW > 0 ? SYNCW : 1) <= 15), "phase-aligned groups must tile the CTA (named barriers 1..15)");
    auto phase_sync = [&]() {
        if (SYNCW >= kWarps) __syncthreads();
        else if (SYNCW > 1) asm volatile("bar.sync %0, %1;" ::"r"(1 + w / (SYNCW > 0 ? SYNCW : 1)), "r"(SYNCW * 32) : "memory");
        else __syncwarp();
    };
    // rectangle pairs of an env: pair index -> (lo, hi), lo < hi, row-major over the upper triangle (pairs before row
    // lo number lo (2N - 1 - lo) / 2; the row comes from a float root, corrected by at most one step); packed lo | hi << 8
    auto decode_pair = [&](int pi) {
        int lo = (int)((float)(2 * N - 1) * 0.5f - sqrt_fast((float)((2 * N - 1) * (2 * N - 1)) * 0.25f - 2.0f * (float)pi));
        lo = max(0, min(lo, N - 2));
        if (lo * (2 * N - 1 - lo) / 2 > pi) lo--;
        else if ((lo + 1) * (2 * N - 2 - lo) / 2 <= pi) lo++;
        return lo | ((lo + 1 + (pi - lo * (2 * N - 1 - lo) / 2)) << 8);
    };
    const int pair_first = (N >= 2) ? decode_pair(min(ln % env_lanes, N * (N - 1) / 2 - 1)) : 0;
    // first lane of every env of the warp: slot of the env's agent 0 (phase D), -1 for all other lanes — computed once,
    // the two integer divisions by a run-time divisor cost 44 instructions per tile iteration otherwise
    const int lead_slot = (ln < EW * env_lanes && ln % env_lanes == 0) ? slot0 + (ln / env_lanes) * N : -1;
    const int stride_wt = gridDim.x * kWarps;
    const int n_iter = SYNCW > 1 ? (n_wt - (int)blockIdx.x * kWarps + stride_wt - 1) / stride_wt : (1 << 30);
    for (int it = 0, wt = blockIdx.x * kWarps + w; it < n_iter && (SYNCW > 1 || wt < n_wt); it++, wt += stride_wt) {
        // ================= phase A: one lane per agent ============================================  @region phase A
        // Which agent this thread integrates.  Free-running warps: lane l of a warp takes slot l of its own warp
        // (n_slots of 32 lanes busy).  Phase-aligned groups: the group's SYNCW * n_slots agents are dealt to the
        // first threads of the group, so whole warps are busy and the others go straight to the barrier (the
        // transcendental-heavy phase then issues 32/n_slots x fewer warp instructions).
        int a_w = w, a_sl = ln;
        bool a_on = ln < n_slots;
        if (SYNCW > 1 && SYNCW < kWarps + 1) {
            const int gw0 = w - w % (SYNCW > 0 ? SYNCW : 1);                       // first warp of this thread's group (consecutive grouping)
            const int tg = (w - gw0) * 32 + ln;                  // thread index within the group
            a_on = tg < SYNCW * n_slots;
            a_w = gw0 + tg / n_slots;
            a_sl = tg % n_slots;
        }
        if (a_on) {
            const int st = a_w * SPW + a_sl;
            const int ei = (wt - w + a_w) * EW + a_sl / N;
            const int i = a_sl % N;
            const bool active = ei < n_envs;
            int e = -1;
            if (active) {
                e = p.env_list ? p.env_list[ei] : ei;
                const uint32_t g = (uint32_t)e * (uint32_t)N + (uint32_t)i;   // B * N * D < 2^32 is checked by the host
                float4 pose = reinterpret_cast<const float4*>(p.buf.pose)[g];
                float delta = p.buf.aux[4 * g];
                float4 car = reinterpret_cast<const float4*>(p.buf.carry)[g];
                float x = pose.x, y = pose.y, psi = pose.z, v = pose.w;
                ts.ox[st] = x;
                ts.oy[st] = y;
                if (MTV) {   // rectangle the reference still holds when it updates the mutual distances (pre-step pose)
                    float osy, ocy;
                    sincos_ool(psi, &osy, &ocy);
                    ocs_a[st] = ocy; osn_a[st] = osy;
                }
                if (step_mode) {
                    // helper_training.py:807-836
                    float2 u = reinterpret_cast<const float2*>(p.buf.action)[g];
                    u.x = clampf(u.x, -cfg.max_speed, cfg.max_speed);
                    u.y = clampf(u.y, -cfg.max_steering, cfg.max_steering);
                    reinterpret_cast<float2*>(p.buf.action)[g] = u;
                    float acc = clampf((u.x - v) / cfg.dt, -cfg.max_acc, cfg.max_acc);
                    float rate = clampf((u.y - delta) / cfg.dt, -cfg.max_steering_rate, cfg.max_steering_rate);
                    // dynamics.py:62-118 + fixed-grid Euler (torchdiffeq)
                    float td = tan_ool(delta);
                    float beta = atan_ool(mulr(cfg.lr_over_lwb, td));
                    float sb, cb;
                    sincos_ool(psi + beta, &sb, &cb);
                    float f0 = v * cb, f1 = v * sb;
                    float sb2, cb2;
                    sincos_ool(beta, &sb2, &cb2);
                    float f2 = ((v / cfg.l_wb) * td) * cb2;
                    x = x + cfg.dt * f0;
                    y = y + cfg.dt * f1;
                    psi = psi + cfg.dt * f2;
                    v = v + cfg.dt * acc;
                    delta = delta + cfg.dt * rate;
                    const float pi_f = 3.14159274101257324f, two_pi = 6.28318548202514648f;
                    delta = pymod(delta + pi_f, two_pi) - pi_f; // dynamics.py:158
                }
                const float beta1 = atan_ool(mulr(cfg.lr_over_lwb, tan_ool(delta)));   // dynamics.py:161-163
                float sc_, cc_;
                sincos_ool(psi + beta1, &sc_, &cc_);
                float vx = mulr(v, cc_), vy = mulr(v, sc_);
                reinterpret_cast<float4*>(p.buf.pose)[g] = make_float4(x, y, psi, v);
                reinterpret_cast<float4*>(p.buf.aux)[g] = make_float4(delta, vx, vy, beta1);
                // helper_scenario.py:742-826 rectangle vertices
                float sy, cy;
                sincos_ool(psi, &sy, &cy);
                const float hl = cfg.half_length, hw = cfg.half_width;
                const float nsy = -sy;
                const float bxs[4] = {hl, hl, -hl, -hl};
                const float bys[4] = {hw, -hw, -hw, hw};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    ts.vtx[k * AS + st] = addr(madd2(cy, bxs[k], nsy, bys[k]), x);
                    ts.vtx[(4 + k) * AS + st] = addr(madd2(sy, bxs[k], cy, bys[k]), y);
                }
                ts.px[st] = x; ts.py[st] = y; ts.cs[st] = cy; ts.sn[st] = sy;
                ts.psim[st] = fmaf(-3.14159274f, floorf(psi * 0.318309873f), psi);
                ts.vx[st] = vx; ts.vy[st] = vy; ts.vabs[st] = sqrtf(__fmaf_rn(vy, vy, mulr(vx, vx)));   // torch.norm(vel): see seg_q
                ts.car[0 * AS + st] = car.x; ts.car[1 * AS + st] = car.y;
                ts.car[2 * AS + st] = car.z; ts.car[3 * AS + st] = car.w;
                ts.path[st] = p.buf.path_id[g];
            }
            ts.flags[st] = active ? 0 : -1; // -1 marks an inactive slot
            ts.env[st] = e;
            ts.coll[st] = 0;
        }
        phase_sync();
        if (!map_ready) { mbar_wait(bar, 0); map_ready = true; }

        const PathRec* paths = reinterpret_cast<const PathRec*>(blob_b + hdr->path_off);
        const float2* pts = reinterpret_cast<const float2*>(blob_b + hdr->pts_off);
        const float4* boxes = reinterpret_cast<const float4*>(blob_b + hdr->box_off);
        const __half2* cones = reinterpret_cast<const __half2*>(blob_b + hdr->cone_off);

        // ================= phase B: G lanes per agent, polyline queries out of the smem map =======  @region phase B glue
        const bool slot_ok = (sl_l < n_slots) && (ts.flags[slot0 + sl_l] >= 0);
        __syncwarp();   // every lane has read the slot's "active" mark before lane 0 of the group overwrites it below
        // a warp-tile past the end of a short batch: skip it (with phase alignment it runs on dummy-safe data
        // instead, so that every warp arrives at every barrier)
        if (SYNCW <= 1 && !__any_sync(0xffffffffu, slot_ok)) continue;
        // all 32 lanes take part in the shuffles, so lanes without a live slot run on dummy-safe data (slot 0)
        const int sl = slot_ok ? slot0 + sl_l : slot0;
        int path = slot_ok ? ts.path[sl] : 0;
        path = (path < 0 || path >= hdr->n_paths) ? 0 : path;
        // Path record fields are read from shared memory where they are used, and every scan result goes to the
        // slot arrays as soon as it exists: nothing but the path index stays live across the three scans (the
        // kernel runs at the 64-register cap of a 1024-thread CTA; what is live across a scan gets spilled).
        const PathRec* prp = paths + path;
        if (!step_mode && p.skip_scan) {
            // spawn-table refresh: the scan results of this pose were computed once, at context creation
            if (slot_ok && lane == 0) {
                const float4 fr = reinterpret_cast<const float4*>(p.fresh)[(size_t)ts.env[sl] * N + i_of_lane];
                ts.sc[0 * AS + sl] = ts.car[0 * AS + sl];
                if (OV) ts.sc[1 * AS + sl] = __int_as_float(idx_of(ts.car[3 * AS + sl]));
                else ts.sc[1 * AS + sl] = ts.car[3 * AS + sl];
                ts.sc[2 * AS + sl] = fr.x; ts.sc[3 * AS + sl] = fr.y;
                ts.sc[4 * AS + sl] = fr.z; ts.sc[5 * AS + sl] = fr.w;
                ts.flags[sl] = 0;
            }
        } else {
            const float px = slot_ok ? ts.px[sl] : 0.0f, py = slot_ok ? ts.py[sl] : 0.0f;
            float rvx[4], rvy[4];   // the agent's rectangle (vertices only: keeps the scans' register footprint small)
#pragma unroll
            for (int k = 0; k < 4; k++) {
                rvx[k] = slot_ok ? ts.vtx[k * AS + sl] : 0.0f;
                rvy[k] = slot_ok ? ts.vtx[(4 + k) * AS + sl] : 0.0f;
            }
            const bool ex = cfg.exhaustive != 0;
            const bool writer = slot_ok && lane == 0;
            float* dbg = (p.buf.dbg && writer) ? p.buf.dbg + ((size_t)ts.env[sl] * N + i_of_lane) * 16 : nullptr;
            // hint = last closest segment (step) / the spawn point written by place_agent (refresh); any value
            // is valid, a good one lets the first chunk scanned set a tight pruning bound
            const int hint = OV ? idx_of(slot_ok ? ts.car[3 * AS + sl] : 0.0f) - 1
                                : __float_as_int(slot_ok ? ts.car[3 * AS + sl] : 0.0f) - 1;
            int h2;
            {
                float d_ref;
                int idx_ref;
                scan_center<G>(pts + prp->c_off, boxes + prp->cbox, prp->n_c, hint, ex, px, py, lane, d_ref, idx_ref);
                if (writer) {
                    ts.sc[0 * AS + sl] = d_ref;
                    ts.sc[1 * AS + sl] = __int_as_float(idx_ref);
                    if (dbg) { dbg[0] = d_ref; dbg[1] = __int_as_float(idx_ref); }
                }
                // the boundaries run alongside the centre line: reuse its closest segment as the hint
                h2 = idx_ref - 1;
            }
            int fl = 0;
            // One rolled loop over {left, right}: a single copy of the scan in the instruction stream.
#pragma unroll 1
            for (int side = 0; side < 2; side++) {
                float dc, dvv[4], m4;
                bool hit;
                scan_boundary<G>(pts + (side ? prp->r_off : prp->l_off), boxes + (side ? prp->rbox : prp->lbox),
                                 cones + (side ? prp->rcone : prp->lcone), side ? prp->n_r : prp->n_l, h2, ex, px, py,
                                 ts.cs + sl, ts.sn + sl, ts.psim + sl, rvx, rvy, rect_radius, p.near2, cfg.half_length,
                                 cfg.half_width, p.band_l, p.band_w, p.buf.dbg != nullptr, lane, dc, dvv, m4, hit);
                if (hit) fl = (int)SGB_FLAG_COLLIDE_LANE;
                if (writer) {
                    dc = dc - cfg.half_width;                                   // world_state_rt.py:608-610
                    ts.sc[(2 + side) * AS + sl] = dc;
                    ts.sc[(4 + side) * AS + sl] = m4;
                    if (dbg) {
                        dbg[side ? 7 : 2] = dc;
#pragma unroll
                        for (int v = 0; v < 4; v++) dbg[(side ? 8 : 3) + v] = dvv[v];
                    }
                }
            }
            if (!prp->is_loop && lane == 0) {
                // entry / exit segments (world_state_rt.py:394-406, world_state_rt_sim.py:412-424)
                const float2* L = pts + prp->l_off;
                const float2* R = pts + prp->r_off;
#pragma unroll 1
                for (int k = 0; k < 2; k++) {
                    const float2 a = L[k ? prp->n_l - 1 : 0], e = R[k ? prp->n_r - 1 : 0];
                    if (rect_cross_seg_L1(rvx, rvy, a.x, a.y, e.x, e.y)) fl |= (int)(k ? SGB_FLAG_EXIT : SGB_FLAG_ENTRY);
                }
            }
            if (writer) ts.flags[sl] = step_mode ? fl : 0;
            if (bpoints) {
                // Boundary points instead of boundary distances: the observation needs distances.closest_point_on_left_b
                // / right_b (world_state_rt.py:597-622), an ARGMIN, so the exact centre-line scan runs on the two
                // boundaries as well.  What agent i's row shows is the fresh index for i == 0 and last step's for
                // i >= 1 — a function of the PRE-step position, which is still in the slot arrays.
                const bool stale = step_mode && i_of_lane != 0;
                const float qx = (stale && slot_ok) ? ts.ox[sl] : px, qy = (stale && slot_ok) ? ts.oy[sl] : py;
                const int hb = stale ? hint : h2;
                int packed = 0;
#pragma unroll 1
                for (int side = 0; side < 2; side++) {
                    float dd;
                    int ii;
                    scan_center<G>(pts + (side ? prp->r_off : prp->l_off), boxes + (side ? prp->rbox : prp->lbox),
                                   side ? prp->n_r : prp->n_l, hb, ex, qx, qy, lane, dd, ii);
                    packed |= ii << (16 * side);
                }
                if (writer) ts.psim[sl] = __int_as_float(packed);   // psim (heading mod pi) is dead after the scans
            }
        }
        // ---- rectangle-rectangle crossings: the N(N-1)/2 unordered pairs of an env are dealt round-robin to  @region pairs
        //      the env's N*G lanes; interX(vertices[lo], vertices[hi]) with lo < hi exactly as
        //      world_state_rt_sim.py:384-393, result OR-ed into both agents' masks
        __syncwarp();   // flags / scan results of all groups of this warp are in shared memory (racecheck-clean)
        if (step_mode && !MTV && ln < EW * env_lanes) {
            const int el = ln / env_lanes;             // env within the warp
            const int q = ln - el * env_lanes;         // lane within the env
            const int sbase = slot0 + el * N;
            if (ts.flags[sbase] >= 0) {
                const int n_pairs = N * (N - 1) / 2;
                for (int pi = q; pi < n_pairs; pi += env_lanes) {
                    // (lo, hi) of a lane's FIRST pair was decoded once, before the tile loop
                    const int packed = (pi == q) ? pair_first : decode_pair(pi);
                    const int lo = packed & 0xff, hi = packed >> 8;
                    {
                        // Gate (same certificate as for boundary segments): rectangles whose centres are farther
                        // apart than two circumradii + kFarMargin cannot touch, and interX can then only fire
                        // through fp32 sign noise on edges that are collinear within kCollinear.
                        const float ddx = ts.px[sbase + hi] - ts.px[sbase + lo], ddy = ts.py[sbase + hi] - ts.py[sbase + lo];
                        const float c1 = ts.cs[sbase + lo], s1 = ts.sn[sbase + lo], c2 = ts.cs[sbase + hi], s2 = ts.sn[sbase + hi];
                        const float cr = c1 * s2 - s1 * c2, dt = c1 * c2 + s1 * s2;   // sin / cos of the heading difference
                        const float reach = 2.0f * rect_radius + kFarMargin;
                        if (!cfg.exhaustive && ddx * ddx + ddy * ddy > reach * reach &&
                            fminf(fabsf(cr), fabsf(dt)) > kCollinear)
                            continue;
                    }
                    Rect rl;
                    float hx[4], hy[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        rl.vx[k] = ts.vtx[k * AS + sbase + lo]; rl.vy[k] = ts.vtx[(4 + k) * AS + sbase + lo];
                        hx[k] = ts.vtx[k * AS + sbase + hi]; hy[k] = ts.vtx[(4 + k) * AS + sbase + hi];
                    }
                    rl.finish();
                    if (rect_cross_rect(rl, hx, hy)) {
                        atomicOr(&ts.coll[sbase + lo], 1 << hi);
                        atomicOr(&ts.coll[sbase + hi], 1 << lo);
                    }
                }
            }
        }
#ifdef SGB_NO_BC_SYNC
        __syncwarp();
#else
        phase_sync();   // alignment only (phase C reads this warp's slots), but dropping it costs 12 %: I-cache
#endif

        // ================= phase C: interactions inside the env, reward, observation ===============  @region phase C1
        {
            const int i = slot_ok ? i_of_lane : 0;   // (sl - slot0) % N, hoisted out of the tile loop
            const int base = sl - i; // slot of agent 0 of this env
            const float pix = ts.px[sl], piy = ts.py[sl];
            uint32_t coll = MTV ? 0u : (uint32_t)ts.coll[sl];
            float ttc_sum = 0.0f, near_sum = 0.0f;
            // ---- C1: the lanes of the group split the other agents j ----
            if (slot_ok) {
                for (int j = lane; j < N; j += G) {
                    const int sj = base + j;
                    const float pjx = ts.px[sj], pjy = ts.py[sj];
                    const float dx = pix - pjx, dy = piy - pjy;
                    const float pp = madd2(dx, dx, dy, dy);
                    float dist = (j == i) ? cfg.diag : sqrtf(pp);  // helper_scenario.py:1012-1029, :1140-1143
                    if (MTV && j != i) {
                        // helper_scenario.py:1030-1138 on the rectangles of the pre-step poses (== current ones in a refresh)
                        dist = mtv_distance(ts.ox[sl], ts.oy[sl], ocs_a[sl], osn_a[sl], ts.ox[sj], ts.oy[sj], ocs_a[sj], osn_a[sj],
                                            cfg.half_length, cfg.half_width);
                        if (step_mode && dist == 0.0f) coll |= 1u << j;   // world_state_rt_sim.py:394-396
                    }
                    ts.dij[sl * N + j] = dist;
                    if (!step_mode) continue;
                    near_sum += dec_lin(dist, cfg.near_agents_low, cfg.near_agents_high);
                    if (cfg.rew_flags & SGB_REW_TTC) {
                        // road_traffic.py:1255-1332 (p_rel = p_j - p_i)
                        const float eps = 1e-6f;
                        const float rx = pjx - pix, ry = pjy - piy;
                        const float wx = ts.vx[sj] - ts.vx[sl], wy = ts.vy[sj] - ts.vy[sl];
                        const float qa = madd2(wx, wx, wy, wy);
                        const float qb = mulr(2.0f, madd2(rx, wx, ry, wy));
                        const float rr = madd2(rx, rx, ry, ry);
                        const float qc = subr(rr, cfg.dsafe_sq);
                        const float disc = subr(mulr(qb, qb), mulr(mulr(4.0f, qa), qc));
                        const float sq = sqrtf(fmaxf(disc, 0.0f));
                        const float dd = sqrtf(fmaxf(rr, 0.0f));
                        const bool valid = (qa > eps) && (disc > 0.0f) && (qb < 0.0f);
                        const float cand = subr(-qb, sq) / addr(mulr(2.0f, qa), eps);
                        float ttc = __int_as_float(0x7f800000);
                        if (valid && cand > 0.0f) ttc = cand;
                        if (dd <= cfg.near_agents_low) ttc = 0.0f;
                        if (j == i) ttc = __int_as_float(0x7f800000);
                        if (!(dd <= cfg.near_agents_high)) ttc = __int_as_float(0x7f800000);
                        ttc = fminf(ttc, cfg.ttc_high);
                        ttc_sum += dec_lin(ttc, cfg.ttc_low, cfg.ttc_high);
                    }
                }
            }
            ttc_sum = group_sum<G>(ttc_sum);
            near_sum = group_sum<G>(near_sum);
            if (MTV) coll = group_or<G>(coll);
            __syncwarp(); // dij of this group is complete

            // ---- C2: every lane of the group picks the k nearest (same result in all lanes);  @region C2 topk+obs
            //          torch.topk(k, largest=False), ties -> lower index.
            const int k_near = cfg.k_near;
            int nb_j[2] = {0, 0};          // the first two neighbours stay in registers, the rest is re-derived
            float nb_d[2] = {0.0f, 0.0f};
            if (G == 4) {
                // Four lanes per agent, N <= 8: lane l owns the candidates j = l and l + 4.  A candidate is the 64-bit
                // key (distance bits, j): distances are >= 0 (MTV: order-preserving bit flip), so the unsigned order of
                // the keys IS the lexicographic (d, j) order of torch.topk's "smallest first, lower index on ties".
                // Two group minima (the second with the first winner struck out) = 4 shuffle rounds in all.
                uint64_t key[2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int j = lane + 4 * h;
                    key[h] = ~0ull;
                    if (slot_ok && j < N) {
                        uint32_t b = __float_as_uint(ts.dij[sl * N + j]);
                        if (MTV) b ^= (b >> 31) ? 0xffffffffu : 0x80000000u;
                        key[h] = ((uint64_t)b << 32) | (uint32_t)j;
                    }
                }
                uint64_t first = ~0ull;
#pragma unroll
                for (int kk = 0; kk < 2; kk++) {
                    uint64_t b = key[0] < key[1] ? key[0] : key[1];
#pragma unroll
                    for (int m = 1; m < 4; m <<= 1) {
                        const uint64_t o = __shfl_xor_sync(0xffffffffu, b, m);
                        b = o < b ? o : b;
                    }
                    int bj = (int)(uint32_t)b;
                    if (b == ~0ull) bj = 0;                                   // inactive slot
                    const float bd = ts.dij[(slot_ok ? sl : slot0) * N + bj];
                    if (kk == 0) { nb_j[0] = bj; nb_d[0] = bd; first = b; } else { nb_j[1] = bj; nb_d[1] = bd; }
                    // strike the winner out (j is unique, the low word identifies it)
                    if ((uint32_t)key[0] == (uint32_t)first && first != ~0ull) key[0] = ~0ull;
                    if ((uint32_t)key[1] == (uint32_t)first && first != ~0ull) key[1] = ~0ull;
                }
            } else {
            uint32_t used = 0;
            // The lanes of the group split j (lane, lane + G, ...), then a lexicographic (d, j) minimum over the group:
            // same result as the ascending scan with strict '<' (first minimal index).  All lanes shuffle.
            for (int kk = 0; kk < k_near && kk < 2; kk++) {
                int bj = 0x7fffffff;
                float bd = __int_as_float(0x7f800000);
                if (slot_ok)
                    for (int j = lane; j < N; j += G) {
                        const float dj = ts.dij[sl * N + j];
                        if (!((used >> j) & 1u) && (dj < bd || bj == 0x7fffffff)) { bd = dj; bj = j; }
                    }
#pragma unroll
                for (int m = 1; m < G; m <<= 1) {
                    const float od = __shfl_xor_sync(0xffffffffu, bd, m);
                    const int oj = __shfl_xor_sync(0xffffffffu, bj, m);
                    if (oj != 0x7fffffff && (bj == 0x7fffffff || od < bd || (od == bd && oj < bj))) { bd = od; bj = oj; }
                }
                if (bj == 0x7fffffff) bj = 0;   // inactive slot
                used |= 1u << bj;
                if (kk == 0) { nb_j[0] = bj; nb_d[0] = bd; } else { nb_j[1] = bj; nb_d[1] = bd; }   // no dynamic index
            }
            }
            if (slot_ok) {
                const uint32_t g = (uint32_t)ts.env[sl] * (uint32_t)N + (uint32_t)i;   // 32-bit: see launch_env
                const float d_ref_n = ts.sc[0 * AS + sl];
                const int idx_n = __float_as_int(ts.sc[1 * AS + sl]);
                const float dLc = ts.sc[2 * AS + sl], dRc = ts.sc[3 * AS + sl];
                const float m4L = ts.sc[4 * AS + sl], m4R = ts.sc[5 * AS + sl];
                const float c_dref = ts.car[0 * AS + sl], c_mL = ts.car[1 * AS + sl], c_mR = ts.car[2 * AS + sl];
                const int c_idx = OV ? idx_of(ts.car[3 * AS + sl]) : __float_as_int(ts.car[3 * AS + sl]);
                const bool write_obs = step_mode || p.write_obs;
                float* o = p.buf.obs + g * (uint32_t)D;   // written in place: 128 B per agent, L2 merges the partial sectors
                const float cs = ts.cs[sl], sn = ts.sn[sl];
                const float2* cpts = pts + prp->c_off;
                const int pr_nc = prp->n_c;
                const bool pr_loop = prp->is_loop != 0;
                // What agent i's observation sees: after a reset everything is fresh; in a
                // step agent 0 sees fresh centre queries + vertex queries of the OLD rectangle, agents >= 1 see
                // last step's values (observation_provider_rt.py:857-925).
                float o_dref, o_mL, o_mR;
                int o_idx;
                if (!step_mode) { o_dref = d_ref_n; o_mL = fminf(dLc, m4L); o_mR = fminf(dRc, m4R); o_idx = idx_n; }
                else if (i == 0) { o_dref = d_ref_n; o_mL = fminf(dLc, c_mL); o_mR = fminf(dRc, c_mR); o_idx = idx_n; }
                else { o_dref = c_dref; o_mL = c_mL; o_mR = c_mR; o_idx = c_idx; }
                if (OV != 0 && write_obs) {
                    // ---- general layout (cfg.obs_flags): the last lane of the group writes the agent's own part, the
                    //      observed neighbours are dealt over the lanes.  What update_state snapshots at
                    //      observation(agent 0) time (observation_provider_rt.py:345-588): poses, velocities, steering
                    //      and vertices of ALL agents are post-step; the short-term path of agent j is fresh for
                    //      j == 0 and one step old for j >= 1 (all fresh after a reset).  Heading and steering are
                    //      read back from pose / aux (stored in phase A by this phase-aligned group; barrier since).
                    const uint32_t ofl = cfg.obs_flags;
                    const bool bird = (ofl & SGB_OBS_BIRD_VIEW) != 0;
                    const uint32_t g0 = g - (uint32_t)i;         // agent 0 of this env
                    const float psi_i = p.buf.pose[4 * g + 2];
                    const float nwx = cfg.norm_pos_world_x, nwy = cfg.norm_pos_world_y;
                    // a global point as the observation holds it: ego frame / norm_pos, or global / pos_world
                    auto put_point = [&](float* dst, float qx, float qy) {
                        if (bird) { dst[0] = qx / nwx; dst[1] = qy / nwy; }
                        else {
                            const float dx = qx - pix, dy = qy - piy;
                            dst[0] = (dx * cs + dy * sn) * r_pos;
                            dst[1] = (dy * cs - dx * sn) * r_pos;
                        }
                    };
                    if (lane == G - 1) {
                        float* q = o;
                        if (bird) {
                            q[0] = pix / nwx; q[1] = piy / nwy;
                            q[2] = wrap_pi(psi_i) / cfg.norm_rot;
                            q[3] = ts.vx[sl] * r_v; q[4] = ts.vy[sl] * r_v;
                            q += 5;
                        } else {
                            *q++ = ts.vabs[sl] * r_v;
                        }
                        if (ofl & SGB_OBS_STEERING) *q++ = wrap_pi(p.buf.aux[4 * g]) / cfg.norm_rot;
                        float2 st[3];
                        short_term(cpts, pr_nc, pr_loop, o_idx, st);
#pragma unroll
                        for (int k = 0; k < 3; k++) { put_point(q, st[k].x, st[k].y); q += 2; }
                        if (!(ofl & SGB_OBS_NO_DIST_CENTER)) *q++ = o_dref * r_dist;
                        if (ofl & SGB_OBS_BOUNDARY_POINTS) {
                            // world_state_rt.py:686-725 (step: shift -2) / :531-576 (reset: shift +1); the loop wrap uses
                            // the CENTRE line's point count, as the reference passes it; an index outside the boundary
                            // lands in the reference's tail padding (= last boundary point; -1 is python's last element)
                            const int packed = __float_as_int(ts.psim[sl]);
                            const bool from_reset = !step_mode || (i != 0 && (__float_as_int(ts.car[3 * AS + sl]) & kCarryFreshBit));
                            const int shift = from_reset ? 1 : -2;
#pragma unroll 1
                            for (int side = 0; side < 2; side++) {
                                const float2* bp = pts + (side ? prp->r_off : prp->l_off);
                                const int n_s = side ? prp->n_r : prp->n_l;
                                const int ib = (packed >> (16 * side)) & 0xffff;
                                for (int k = 0; k < kNearPts; k++) {
                                    int fi = k + ib + shift;
                                    if (pr_loop && fi >= pr_nc - 1) fi = (fi + 1) % pr_nc;
                                    if (fi < 0 || fi >= n_s) fi = n_s - 1;
                                    const float2 b = bp[fi];
                                    put_point(q, b.x, b.y);
                                    q += 2;
                                }
                            }
                        } else {
                            q[0] = o_mL * r_dist;
                            q[1] = o_mR * r_dist;
                        }
                    }
                    const int own = obs_dim_of(ofl, 0), per = obs_dim_of(ofl, 1) - own;
                    for (int kk = lane; kk < k_near; kk += G) {
                        int bj;
                        float bd;
                        if (kk == 0) { bj = nb_j[0]; bd = nb_d[0]; }
                        else if (kk == 1) { bj = nb_j[1]; bd = nb_d[1]; }
                        else bj = kth_nearest(ts.dij + sl * N, N, kk, &bd);
                        const int sj = base + bj;
                        const uint32_t gj = g0 + (uint32_t)bj;
                        float* q = o + own + per * kk;
                        const float psi_j = p.buf.pose[4 * gj + 2];
                        if (ofl & SGB_OBS_CENTRES) {
                            put_point(q, ts.px[sj], ts.py[sj]);
                            q[2] = (bird ? wrap_pi(psi_j) : wrap_pi(psi_j - psi_i)) / cfg.norm_rot;
                            q[3] = (2.0f * cfg.half_length) / cfg.norm_dist_agent;
                            q[4] = (2.0f * cfg.half_width) / cfg.norm_dist_agent;
                            q += 5;
                        } else {
#pragma unroll
                            for (int v = 0; v < 4; v++) put_point(q + 2 * v, ts.vtx[v * AS + sj], ts.vtx[(4 + v) * AS + sj]);
                            q += 8;
                        }
                        if (bird) { q[0] = ts.vx[sj] * r_v; q[1] = ts.vy[sj] * r_v; }
                        else {
                            const float cj = ts.cs[sj], sj_ = ts.sn[sj];
                            q[0] = (ts.vabs[sj] * (cj * cs + sj_ * sn)) * r_v;
                            q[1] = (ts.vabs[sj] * (sj_ * cs - cj * sn)) * r_v;
                        }
                        q += 2;
                        if (ofl & SGB_OBS_STEERING) *q++ = wrap_pi(p.buf.aux[4 * gj]) / cfg.norm_rot;
                        if (!(ofl & SGB_OBS_NO_DIST_AGENTS)) *q++ = bd * r_dist;
                        if (ofl & SGB_OBS_REF_OTHERS) {
                            int pj = ts.path[sj];
                            pj = (pj < 0 || pj >= hdr->n_paths) ? 0 : pj;
                            const PathRec* prj = paths + pj;
                            const int idx_j = idx_of((step_mode && bj != 0) ? ts.car[3 * AS + sj] : ts.sc[1 * AS + sj]);
                            float2 st[3];
                            short_term(pts + prj->c_off, prj->n_c, prj->is_loop != 0, idx_j, st);
#pragma unroll
                            for (int k = 0; k < 3; k++) { put_point(q, st[k].x, st[k].y); q += 2; }
                        }
                        bool masked = (ofl & SGB_OBS_APPLY_MASK) && bd >= cfg.mask_distance;
                        if (!masked && (ofl & SGB_OBS_MASK_LANELETS)) {
                            // + the lanelet relation (observation_provider_rt.py:646-664, map_manager.py:91-119): masked
                            // unless the neighbour's current lanelet is the ego's or adjacent to it (post-step positions)
                            const int li = current_lanelet_ool(p.lanelet_xy, p.lanelet_off, p.n_lanelets, p.lanelet_max_len, pix, piy);
                            const int lj = current_lanelet_ool(p.lanelet_xy, p.lanelet_off, p.n_lanelets, p.lanelet_max_len,
                                                               ts.px[sj], ts.py[sj]);
                            masked = p.lanelet_adj[li * p.n_lanelets + lj] == 0;
                        }
                        if (masked) {
                            // is_apply_mask (observation_provider_rt.py:638-749): a far neighbour shows constants —
                            // positions / vertices / reference path / distance 1, heading / steering / velocity 0
                            float* m = o + own + per * kk;
                            if (ofl & SGB_OBS_CENTRES) { m[0] = 1.0f; m[1] = 1.0f; m[2] = 0.0f; m += 5; }   // length, width stay
                            else {
#pragma unroll
                                for (int v = 0; v < 8; v++) m[v] = 1.0f;
                                m += 8;
                            }
                            m[0] = 0.0f; m[1] = 0.0f; m += 2;
                            if (ofl & SGB_OBS_STEERING) *m++ = 0.0f;
                            if (!(ofl & SGB_OBS_NO_DIST_AGENTS)) *m++ = 1.0f;
                            if (ofl & SGB_OBS_REF_OTHERS) {
#pragma unroll
                                for (int k = 0; k < 6; k++) m[k] = 1.0f;
                            }
                        }
                        if (p.buf.dbg && kk < 2) p.buf.dbg[g * 16 + 13 + kk] = (float)bj;
                    }
                    if (cfg.obs_noise_level > 0.0f) {
                        // observation_provider_rt.py:611-617: obs + level * U[0,1) on every element, fresh draws on every
                        // call (torch.rand_like).  The reference draws from torch's global generator; here the draw is a
                        // counter-based hash of (seed, API-call counter, GLOBAL env index, agent, column): independent
                        // between envs, agents, columns and steps — also for agents whose state does not change, and for
                        // envs started from the same initial state — reproducible, and independent of how envs are
                        // sharded over GPUs (distribution-equivalent, like the reset draws).
                        __syncwarp(((G >= 32) ? 0xffffffffu : ((1u << G) - 1u)) << (ln - lane));   // the row is complete
                        uint64_t key = mix64((uint64_t)cfg.obs_noise_seed ^ mix64(p.noise_epoch));
                        key = mix64(key ^ (uint64_t)(p.env_base + ts.env[sl]));
                        key = mix64(key ^ ((uint64_t)i << 48));
                        for (int d = lane; d < D; d += G) {
                            const float u = (float)(mix64(key + (uint64_t)d) >> 40) * (1.0f / 16777216.0f);
                            o[d] += cfg.obs_noise_level * u;
                        }
                    }
                }
                if (OV == 0 && write_obs) {
                    // ---- point transforms into the ego frame, split evenly over the lanes: 3 short-term points
                    //      of the agent itself + 4 vertices of each observed neighbour (rotation form of
                    //      helper_scenario.py:1241-1273, cos/sin of the heading come from phase A; normalisers
                    //      applied as reciprocals: <= 1 ulp from the reference's division, tolerance is 1e-5)
                    const int n_pts = 3 + 4 * k_near;
                    for (int t = lane; t < n_pts; t += G) {
                        float qx, qy;
                        float* dst;
                        if (t < 3) {
                            const float2 q = cpts[short_term_index(2 * t + o_idx + 1, pr_nc, pr_loop)];   // helper_scenario.py:928-946
                            qx = q.x; qy = q.y;
                            dst = o + 1 + 2 * t;
                        } else {
                            const int kk = (t - 3) >> 2, v = (t - 3) & 3;
                            const int bj = kk == 0 ? nb_j[0] : (kk == 1 ? nb_j[1] : kth_nearest(ts.dij + sl * N, N, kk, nullptr));
                            const int sj = base + bj;
                            qx = ts.vtx[v * AS + sj]; qy = ts.vtx[(4 + v) * AS + sj];
                            dst = o + 10 + 11 * kk + 2 * v;
                        }
                        const float dx = qx - pix, dy = qy - piy;
                        dst[0] = (dx * cs + dy * sn) * r_pos;
                        dst[1] = (dy * cs - dx * sn) * r_pos;
                    }
                    if (lane == 0) {
                        o[0] = ts.vabs[sl] * r_v;
                        o[7] = o_dref * r_dist;
                        o[8] = o_mL * r_dist;
                        o[9] = o_mR * r_dist;
                    }
                    {   // the neighbours' velocity / distance entries, one neighbour per lane
                        for (int kk = lane; kk < k_near; kk += G) {
                            int bj;
                            float bd;
                            if (kk == 0) { bj = nb_j[0]; bd = nb_d[0]; }
                            else if (kk == 1) { bj = nb_j[1]; bd = nb_d[1]; }
                            else bj = kth_nearest(ts.dij + sl * N, N, kk, &bd);
                            const int sj = base + bj;
                            float* ob = o + 10 + 11 * kk;
                            // |v_j| * (cos, sin)(psi_j - psi_i) via the stored cos/sin of both headings
                            const float cj = ts.cs[sj], sj_ = ts.sn[sj];
                            ob[8] = (ts.vabs[sj] * (cj * cs + sj_ * sn)) * r_v;
                            ob[9] = (ts.vabs[sj] * (sj_ * cs - cj * sn)) * r_v;
                            ob[10] = bd * r_dist;
                            if (p.buf.dbg && kk < 2) p.buf.dbg[g * 16 + 13 + kk] = (float)bj;
                        }
                    }
                }
                // @region reward/carry
                if (lane == 2 % G) {
                    // ---- reward (road_traffic.py:947-1253), flags, next step's carry ----
                    int fl = ts.flags[sl];
                    if (coll) fl |= (int)SGB_FLAG_COLLIDE_AGENT;
                    ts.flags[sl] = fl;
                    float d_bound;
                    if (!step_mode || i != 0) d_bound = fminf(fminf(dLc, m4L), fminf(dRc, m4R));
                    else d_bound = fminf(fminf(dLc, c_mL), fminf(dRc, c_mR)); // agent 0: stale vertices
                    float pen_near_agents = 0.0f, pen_a2a = 0.0f, pen_lane = 0.0f, rew_goal = 0.0f, rew_total = 0.0f;
                    if (step_mode) {
                        float2 st_old[3];   // short-term path of the PREVIOUS step
                        short_term(cpts, pr_nc, pr_loop, c_idx, st_old);
                        const float oxp = ts.ox[sl], oyp = ts.oy[sl];
                        float mvx = pix - oxp, mvy = piy - oyp;
                        float acc = 0.0f;
#pragma unroll
                        for (int k = 0; k < 3; k++) {
                            float rx = st_old[k].x - oxp, ry = st_old[k].y - oyp;
                            acc += (mvx * rx + mvy * ry) * cfg.w_ref[k];
                        }
                        float rew = 0.0f;
                        rew += (acc / cfg.speed_dt) * cfg.reward_progress;
                        pen_a2a = (coll ? 1.0f : 0.0f) * cfg.penalty_collide_agents;
                        pen_lane = ((fl & SGB_FLAG_COLLIDE_LANE) ? 1.0f : 0.0f) * cfg.penalty_collide_lane;
                        rew_goal = ((fl & SGB_FLAG_EXIT) ? 1.0f : 0.0f) * cfg.reward_reach_goal;   // :996-997
                        const float pen_nb = dec_lin(d_bound, cfg.near_boundary_low, cfg.near_boundary_high) * cfg.penalty_near_boundary;
                        if (cfg.testing_mode) {                                              // :1050-1055
                            rew += rew_goal; rew += pen_a2a; rew += pen_lane;
                        } else {
                            if (cfg.rew_flags & SGB_REW_EXACT_SPARSE) { rew += pen_a2a; rew += pen_lane; }
                            if (cfg.rew_flags & SGB_REW_TTC) {
                                float risk = ttc_sum / (float)(N - 1 > 1 ? N - 1 : 1);
                                pen_near_agents = risk * cfg.penalty_near_agents;
                                rew += pen_near_agents;
                                rew += pen_nb;
                                rew += pen_a2a; rew += pen_lane;
                                if (cfg.rew_flags & SGB_REW_SPARSE) { rew += pen_a2a; rew += pen_lane; }
                            }
                            if (cfg.rew_flags & SGB_REW_DISTANCE) {
                                pen_near_agents = near_sum * cfg.penalty_near_agents;
                                rew += pen_near_agents;
                                rew += pen_nb;
                                if (cfg.rew_flags & SGB_REW_SPARSE) { rew += pen_a2a; rew += pen_lane; }
                            }
                        }
                        rew_total = clampf(rew, -1.0f, 1.0f);
                        p.buf.reward[g] = rew_total;
                        p.buf.agent_flags[g] = (uint8_t)fl;
                        // health word (optional): the reference asserts that positions / rewards hold no NaN / inf
                        // (road_traffic.py:1245-1246); a non-finite term makes the sum non-finite
                        if (p.buf.nan_flags && !(fabsf(pix + piy + ts.vabs[sl] + rew_total + d_ref_n + d_bound) < SGB_INF))
                            atomicOr(p.buf.nan_flags, 1u);
                        if (p.buf.collide_with) p.buf.collide_with[g] = coll;
                    } else {
                        p.buf.agent_flags[g] = 0;
                        if (p.buf.collide_with) p.buf.collide_with[g] = 0;
                    }
                    float4 nc;
                    nc.x = d_ref_n;
                    nc.y = (i == 0) ? m4L : fminf(dLc, m4L);
                    nc.z = (i == 0) ? m4R : fminf(dRc, m4R);
                    if (OV) nc.w = __int_as_float(idx_n | ((bpoints && !step_mode) ? kCarryFreshBit : 0));
                    else nc.w = __int_as_float(idx_n);
                    reinterpret_cast<float4*>(p.buf.carry)[g] = nc;
                    if (p.buf.dbg) p.buf.dbg[g * 16 + 12] = d_bound;
                    if (p.buf.info) {
                        // what info(agent_i) reads right after reward(i) / observation(i) (road_traffic.py:1574-1633):
                        // agent i's own distances and short-term path are FRESH here (for i == 0 the vertex part of
                        // the boundary distances is the stale one, as in its observation)
                        float4* io = reinterpret_cast<float4*>(p.buf.info + g * SGB_INFO_DIM);
                        float2 st_new[3];
                        short_term(cpts, pr_nc, pr_loop, idx_n, st_new);
                        const bool stale0 = step_mode && i == 0;
                        io[0] = make_float4(st_new[0].x, st_new[0].y, st_new[1].x, st_new[1].y);
                        io[1] = make_float4(st_new[2].x, st_new[2].y, d_ref_n, fminf(dLc, stale0 ? c_mL : m4L));
                        io[2] = make_float4(fminf(dRc, stale0 ? c_mR : m4R), pen_near_agents, pen_a2a, pen_lane);
                        io[3] = make_float4(rew_goal, rew_total, d_bound, 0.0f);
                    }
                }
            }
        }
        __syncwarp();   // phase D reads this warp's own slots only

        // ================= phase D: per-env outputs ===============================================  @region phase D
        if (step_mode) {
            // one ballot per quantity over the warp (lane 0 of every agent group votes its agent's flags), then the
            // first lane of each env reads its env's bits
            const int myfl = (slot_ok && lane == 0) ? ts.flags[sl] : 0;
            const uint32_t b_hit = __ballot_sync(0xffffffffu, (myfl & (int)(SGB_FLAG_COLLIDE_AGENT | SGB_FLAG_COLLIDE_LANE)) != 0);
            const uint32_t b_exit = __ballot_sync(0xffffffffu, (myfl & (int)SGB_FLAG_EXIT) != 0);
            if (lead_slot >= 0 && ts.flags[lead_slot] >= 0) {
                const uint32_t em = (env_lanes >= 32 ? 0xffffffffu : ((1u << env_lanes) - 1u)) << ln;   // this env's lanes
                const int e = ts.env[lead_slot];
                const int tries = __popc((b_hit | b_exit) & em);                                                // :1029-1035
                const int succ = __popc(b_exit & em);                                                           // :998-1002
                if (p.buf.task_tries && tries) p.buf.task_tries[e] += tries;
                if (p.buf.task_success && succ) p.buf.task_success[e] += succ;
                const int step = p.buf.step_count[e] + 1;          // road_traffic.py:954-962
                p.buf.step_count[e] = step;
                // road_traffic.py:1451-1457 (training mode) / :1429-1433 (testing mode: only the time limit ends an env)
                // :1388-1393: reset_agent_fixed_duration (both modes) — periodic in timer.step, see sgb_config
                const bool dn = (step == cfg.max_steps - 1) || (!cfg.testing_mode && (b_hit & em) != 0u) ||
                                (cfg.reset_fixed_period > 0 && step % cfg.reset_fixed_period == 0);
                p.buf.done[e] = dn ? 1 : 0;
            }
        }
        phase_sync();   // the next tile's phase A (dealt over the whole group) overwrites these slots
    }
    if (!map_ready) mbar_wait(bar, 0); // never leave with a bulk copy in flight
}

// ---- placement / reset kernels ------------------------------------------------------------------------------  @region other kernels
// counter-based: one 64-bit draw per (seed, epoch, env, agent, try, which)
__device__ __forceinline__ uint64_t draw(uint64_t seed, uint64_t epoch, uint64_t env, uint32_t agent, uint32_t tr, uint32_t which) {
    uint64_t h = mix64(seed ^ mix64(epoch));
    h = mix64(h ^ env);
    h = mix64(h ^ (((uint64_t)agent << 40) | ((uint64_t)tr << 8) | which));
    return h;
}

struct PlaceParams {
    sgb_config cfg;
    sgb_buffers buf;
    const unsigned char* blob;
    const float* yaw;          // [pts of centre lines] indexed like the blob's centre points (c_off + point)
    const uint8_t* agent_mask;
    const int32_t* path;
    const int32_t* point;
    const float* speed;
    int32_t B, N;
};

// spawn table: per centre point (indexed like `yaw`) d_ref, (int) idx_ref, dLc, dRc, m4L, m4R, 0, 0 of an agent placed
// there — what sgb_refresh computes for that pose, evaluated once at context creation (sgb_api.cu: build_spawn_table)
__device__ __forceinline__ void place_agent(const sgb_config& cfg, const sgb_buffers& buf, const unsigned char* blob,
                                            const float* yaw, size_t g, int path, int point, float speed,
                                            const float* spawn_tab = nullptr, int agent = 0, float* fresh = nullptr) {
    const BlobHeader* hdr = reinterpret_cast<const BlobHeader*>(blob);
    const PathRec* paths = reinterpret_cast<const PathRec*>(blob + hdr->path_off);
    const float2* pts = reinterpret_cast<const float2*>(blob + hdr->pts_off);
    const PathRec pr = paths[path];
    float2 c = pts[pr.c_off + point];
    float psi = yaw[pr.c_off + point];
    // world_state_rt_sim.py:189-211: steering 0, sideslip 0, vel = speed * (cos, sin)(0 + yaw)
    reinterpret_cast<float4*>(buf.pose)[g] = make_float4(c.x, c.y, psi, speed);
    float s, co;
    sincosf(0.0f + psi, &s, &co);
    reinterpret_cast<float4*>(buf.aux)[g] = make_float4(0.0f, speed * co, speed * s, 0.0f);
    if (spawn_tab) {
        const float4 t0 = reinterpret_cast<const float4*>(spawn_tab)[2 * (size_t)(pr.c_off + point)];
        const float4 t1 = reinterpret_cast<const float4*>(spawn_tab)[2 * (size_t)(pr.c_off + point) + 1];
        // the carry of agent 0 holds the vertex part only (its centre part is always fresh)
        reinterpret_cast<float4*>(buf.carry)[g] = make_float4(t0.x, agent == 0 ? t1.x : fminf(t0.z, t1.x),
                                                               agent == 0 ? t1.y : fminf(t0.w, t1.y),
                                                               (cfg.obs_flags & SGB_OBS_BOUNDARY_POINTS)
                                                                   ? __int_as_float(__float_as_int(t0.y) | kCarryFreshBit) : t0.y);
        if (fresh) reinterpret_cast<float4*>(fresh)[g] = make_float4(t0.z, t0.w, t1.x, t1.y);
    } else {
        reinterpret_cast<float4*>(buf.carry)[g] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(point + 1)); // search hint
    }
    buf.path_id[g] = path;
}

__global__ void place_kernel(const PlaceParams p) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (size_t)p.B * p.N) return;
    if (p.agent_mask && !p.agent_mask[g]) return;
    place_agent(p.cfg, p.buf, p.blob, p.yaw, g, p.path[g], p.point[g], p.speed[g]);
}

struct ResetParams {
    sgb_config cfg;
    sgb_buffers buf;
    const unsigned char* blob;
    const float* yaw;
    int32_t* list;             // [B] out: compacted indices of the envs that need a refresh
    int32_t* count;            // out: number of entries (zero on entry: cleared by the previous reset, see count_next)
    int32_t* count_next;       // the counter the NEXT reset on this scratch appends to: cleared here (no memset between the kernels)
    int32_t* n_failed;
    uint64_t seed, epoch;
    int64_t env_offset;
    int32_t B, N, path_lo, path_hi, max_tries, all;
    const uint8_t* env_mask;   // explicit selection (sgb_reset_masked): envs to reset fully (may be NULL) ...
    const uint8_t* agent_mask; // ... and agents to respawn; when either is given, done / flags / config are ignored
    int32_t explicit_sel;
    int32_t epw;               // envs per warp of reset_kernel
    const float* spawn_tab;    // see place_agent
    float* fresh;              // [B,N,4] scratch for the observation refresh of fully reset envs
    int32_t list_full_only;    // 1: only fully reset envs go into `list` (they get a fresh observation)
    // path sets (sgb_set_path_sets; cpm_mixed): n_sets > 0 -> a full reset draws the env's set, a respawn keeps it
    int32_t n_sets;
    int32_t set_lo[4], set_hi[4];
    float set_cum[4];          // cumulative probabilities, set_cum[n_sets - 1] >= 1
};

// One SUB-WARP of W lanes per env (W = 8 or 32, W >= N), 32 / W envs in flight per warp: bounded rejection sampling
// (world_state_rt_sim.py:215-311).  Agents are placed one after the other (each must keep its distance from the ones
// before it), but the TRIES of an agent run in parallel: lane l of the sub-warp evaluates try W*r + l of round r, a ballot
// picks the first feasible one — exactly the try the reference's sequential loop would accept — and the winner is
// broadcast.  Lane a holds agent a's position; at the end every lane places its own agent (parallel loads from the spawn
// table, parallel stores).  Every shuffle / ballot names its sub-warp's lanes only, so the sub-warps of a warp run
// independently (different envs need different numbers of tries).  The draws are keyed by (seed, epoch, GLOBAL env,
// agent, try): which lanes evaluate them does not change a result.  History: one thread per env 0.155 ms per masked reset
// at the headline shape and 0.14 ms on crowded maps; one warp per env 0.049 -> 0.0465; four envs per warp (W = 8; large
// batches of up to 8 agents) 0.037.
// Register budget: the warp-per-env form is latency-bound and runs best at 40 registers / 48 resident warps per SM
// (0.0500 -> 0.0465 ms per masked reset at the headline shape); the sub-warp forms keep more state (their masks and lane
// arithmetic) and run best without spills at 64 registers / 32 warps (0.0415 with 40 registers, 0.0371 with 64).
template <int W>
__global__ void __launch_bounds__(256, W == 32 ? 6 : 4) reset_kernel(const ResetParams p) {
    pdl_launch_dependents();
    pdl_wait();                // done / flags / poses are the step kernel's outputs
    if (blockIdx.x == 0 && threadIdx.x == 0 && p.count_next) *p.count_next = 0;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int ln = threadIdx.x & 31;
    const int N = p.N;
    constexpr int S = 32 / W;                       // envs in flight per warp
    const int sg = ln / W;                          // sub-warp of this lane
    const int sl = ln - sg * W;                     // lane within the sub-warp (= agent index for sl < N)
    const int base = sg * W;                        // first lane of the sub-warp
    const uint32_t wmask = W == 32 ? 0xffffffffu : ((1u << W) - 1u);
    const uint32_t sgmask = wmask << base;          // the sub-warp's lanes: the mask of all its shuffles / ballots
    // a warp looks after `epw` consecutive envs: lane l finds out whether env l needs work, then the touched ones are
    // dealt to the sub-warps, S at a time
    const int e_first = w * p.epw;
    if (e_first >= p.B) return;
    bool l_full = false;
    uint32_t l_respawn = 0;
    if (ln < p.epw && e_first + ln < p.B) {
        const int el = e_first + ln;
        l_full = p.explicit_sel ? (p.env_mask && p.env_mask[el]) : (p.all || p.buf.done[el]);
        if (!l_full) {
            if (p.explicit_sel) {
                if (p.agent_mask)
                    for (int a = 0; a < N; a++) l_respawn |= p.agent_mask[(size_t)el * N + a] ? (1u << a) : 0u;
            } else if (p.cfg.respawn_on_exit || p.cfg.testing_mode) {
                // training mode (:1449-1472): agents that crossed an entry / exit segment, maps with open paths only;
                // testing mode (:1435-1447): every colliding or leaving agent, on every map
                const uint32_t which = p.cfg.testing_mode ? (SGB_FLAG_COLLIDE_AGENT | SGB_FLAG_COLLIDE_LANE | SGB_FLAG_ENTRY | SGB_FLAG_EXIT)
                                                          : (SGB_FLAG_ENTRY | SGB_FLAG_EXIT);
                for (int a = 0; a < N; a++) l_respawn |= (p.buf.agent_flags[(size_t)el * N + a] & which) ? (1u << a) : 0u;
            }
        }
    }
    uint32_t touched = __ballot_sync(0xffffffffu, l_full || l_respawn != 0u);
  while (touched) {
    // sub-warp s takes the s-th lowest touched env of this round (warp-uniform arithmetic, no communication)
    int src_l = -1;
#pragma unroll
    for (int s = 0; s < S; s++) {
        if (touched) {
            const int bit = __ffs(touched) - 1;
            touched &= touched - 1;
            if (s == sg) src_l = bit;
        }
    }
    // the env's selection sits in lane src_l of the WARP: fetch it before the sub-warps go their own ways
    const bool full_w = __shfl_sync(0xffffffffu, (int)l_full, max(src_l, 0)) != 0;
    const uint32_t respawn_w = __shfl_sync(0xffffffffu, l_respawn, max(src_l, 0));
    if (src_l >= 0) {
    const int e = e_first + src_l;
    const bool full = full_w;
    const uint32_t respawn = respawn_w;
    const uint32_t todo = full ? (N >= 32 ? 0xffffffffu : ((1u << N) - 1u)) : respawn;
    const BlobHeader* hdr = reinterpret_cast<const BlobHeader*>(p.blob);
    const PathRec* paths = reinterpret_cast<const PathRec*>(p.blob + hdr->path_off);
    const float2* pts = reinterpret_cast<const float2*>(p.blob + hdr->pts_off);
    float qx = 0.0f, qy = 0.0f;            // lane a of the sub-warp: position of agent a
    if (sl < N) {
        const float4 ps = reinterpret_cast<const float4*>(p.buf.pose)[(size_t)e * N + sl];
        qx = ps.x; qy = ps.y;
    }
    const uint64_t env_g = (uint64_t)(p.env_offset + e);
    // The paths this env draws from.  With path sets (cpm_mixed, world_state_rt_sim.py:313-358) a full reset draws
    // ONE set for the whole env — multinomial(cpm_scenario_probabilities) in the reference, the env's own counter-based
    // draw here (every lane computes the same value) — and a respawn keeps the env's set.
    int path_lo = p.path_lo, path_hi = p.path_hi;
    if (p.n_sets > 0) {
        int sid;
        if (full) {
            const float u = (float)(draw(p.seed, p.epoch, env_g, 0, 0, 3) >> 40) * (1.0f / 16777216.0f);
            sid = 0;
            while (sid < p.n_sets - 1 && !(u < p.set_cum[sid])) sid++;
            if (sl == 0) p.buf.scenario_id[e] = sid;
        } else {
            sid = min(max(p.buf.scenario_id[e], 0), p.n_sets - 1);
        }
        path_lo = p.set_lo[sid];
        path_hi = p.set_hi[sid];
    }
    int my_path = path_lo, my_point = 3;  // lane a: where agent a goes (if it is placed)
    int failed = 0;
    auto candidate = [&](int a, int tr, int& path, int& point) {
        path = path_lo + (int)(draw(p.seed, p.epoch, env_g, a, tr, 0) % (uint64_t)(path_hi - path_lo));
        const int2 on = *reinterpret_cast<const int2*>(&paths[path].c_off);   // c_off, n_c
        int end = on.y / 2;
        // testing mode: the range starts as [3, 4) and grows by the try count (world_state_rt_sim.py:254-261)
        if (p.cfg.testing_mode) end = min(end, 3 + (tr + 1) * (tr + 2) / 2);
        point = 3 + (int)(draw(p.seed, p.epoch, env_g, a, tr, 1) % (uint64_t)(end - 3 > 0 ? end - 3 : 1));
        return pts[on.x + point];
    };
    // stage 1: the FIRST try of every agent, all agents in parallel (lane a = agent a) — on roomy maps it is
    // accepted nine times out of ten, so this is where the draws and the dependent map loads should overlap
    float2 c0 = make_float2(0.0f, 0.0f);
    if (sl < N && ((todo >> sl) & 1u)) c0 = candidate(sl, 0, my_path, my_point);
    // stage 2: accept / retry, agent by agent (each must keep its distance from the ones placed before it)
    for (int a = 0; a < N; a++) {
        if (!((todo >> a) & 1u)) continue;     // (uniform within the sub-warp)
        // full reset: against agents 0..a-1 (agent 0 always feasible); respawn: against all others
        const uint32_t others = (full ? ((1u << a) - 1u) : (N >= 32 ? 0xffffffffu : ((1u << N) - 1u))) & ~(1u << a);
        const float cx = __shfl_sync(sgmask, c0.x, base + a), cy = __shfl_sync(sgmask, c0.y, base + a);
        // try 0: every lane o tests the candidate against ITS agent's position, one ballot
        const float ddx = cx - qx, ddy = cy - qy;
        const uint32_t bad = ((__ballot_sync(sgmask, !(madd2(ddx, ddx, ddy, ddy) >= p.cfg.reset_min_dist_sq)) >> base) & wmask) & others;
        if (!bad || p.max_tries <= 1) {
            if (bad) failed++;
            if (sl == a) { qx = cx; qy = cy; }
            continue;
        }
        // tries 1, 2, ...: W at a time, lane l evaluates try 1 + r0 + l; a ballot picks the first feasible one,
        // exactly the try a sequential loop would accept
        bool placed = false;
        int w_path = path_lo, w_point = 3;
        float w_x = 0.0f, w_y = 0.0f;
        for (int r0 = 1; r0 < p.max_tries && !placed; r0 += W) {
            const int tr = r0 + sl;
            const bool live = tr < p.max_tries;
            int path = path_lo, point = 3;
            float2 c = make_float2(0.0f, 0.0f);
            if (live) c = candidate(a, tr, path, point);
            bool ok = live;
            for (int o = 0; o < N; o++) {            // sub-warp-uniform loop: every lane tests ITS candidate against agent o
                const float ox = __shfl_sync(sgmask, qx, base + o), oy = __shfl_sync(sgmask, qy, base + o);
                if (!((others >> o) & 1u)) continue;
                const float dx = c.x - ox, dy = c.y - oy;
                if (!(madd2(dx, dx, dy, dy) >= p.cfg.reset_min_dist_sq)) ok = false;
            }
            const uint32_t okm = (__ballot_sync(sgmask, ok) >> base) & wmask;
            // if this was the last round and nothing is feasible, the last try is kept (rather than spinning
            // forever) and reported
            const bool last_round = r0 + W >= p.max_tries;
            const int src = okm ? (__ffs(okm) - 1) : (last_round ? (p.max_tries - 1 - r0) : -1);
            if (src >= 0) {
                w_path = __shfl_sync(sgmask, path, base + src);
                w_point = __shfl_sync(sgmask, point, base + src);
                w_x = __shfl_sync(sgmask, c.x, base + src);
                w_y = __shfl_sync(sgmask, c.y, base + src);
                placed = true;
                if (!okm) failed++;
            }
        }
        if (sl == a) { qx = w_x; qy = w_y; my_path = w_path; my_point = w_point; }
    }
    if (sl < N && ((todo >> sl) & 1u)) {
        const float u = (float)(draw(p.seed, p.epoch, env_g, sl, 0, 2) >> 40) * (1.0f / 16777216.0f);
        place_agent(p.cfg, p.buf, p.blob, p.yaw, (size_t)e * N + sl, my_path, my_point, u * p.cfg.max_speed, p.spawn_tab, sl,
                    p.fresh);
    }
    // collision masks of a touched env are cleared (road_traffic.py:907)
    if (sl < N) {
        p.buf.agent_flags[(size_t)e * N + sl] = 0;
        if (p.buf.collide_with) p.buf.collide_with[(size_t)e * N + sl] = 0;
    }
    if (sl == 0) {
        if (full) p.buf.step_count[e] = 0; // road_traffic.py:875-877
        if (full || !p.list_full_only) p.list[atomicAdd(p.count, 1)] = e;
        if (failed && p.n_failed) atomicAdd(p.n_failed, failed);
    }
    }
    __syncwarp();   // the sub-warps meet again before the next round is dealt
  }
}

// Generalised advantage estimation over a rollout, one thread per (env, agent), reverse scan over T.
// TorchRL GAE semantics used by the reference (optimization_module.py:62-67, mappo_cavs.py:342-378):
//   delta_t = r_t + gamma * V(s_{t+1}) * (1 - terminated_t) - V(s_t)
//   A_t     = delta_t + gamma * lambda * (1 - done_t) * A_{t+1},  A_T = 0;   value_target = A + V
// terminated == done (per env, broadcast over agents).  Layout [T, B*N]; consecutive threads touch consecutive
// addresses at every t, so all five streams are coalesced.  Pure HBM streaming: 13 B read + 8 B written per element.
__global__ void gae_kernel(int T, int BN, int N, const float* __restrict__ reward, const float* __restrict__ value,
                           const float* __restrict__ next_value, const uint8_t* __restrict__ done, float gamma,
                           float lmbda, float* __restrict__ adv, float* __restrict__ target) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= BN) return;
    const int b = k / N;
    const int B = BN / N;
    float a = 0.0f;
    for (int t = T - 1; t >= 0; t--) {
        const size_t o = (size_t)t * BN + k;
        const float nd = 1.0f - (float)done[(size_t)t * B + b];
        const float v = value[o];
        const float delta = reward[o] + gamma * next_value[o] * nd - v;
        a = delta + gamma * lmbda * nd * a;
        adv[o] = a;
        target[o] = a + v;
    }
}

// The same scan FUSED with the design's one collective (all-gather of advantage / value target at
// PPO-update time): every rank keeps [world, T, B*N] gather buffers in peer-mapped (symmetric) memory and each value is
// stored, as soon as it exists, into slot `rank` of EVERY rank's buffer — plain stores to the peers' mappings over
// NVLink / NVSwitch, or ONE multimem.st to the buffers' multicast address, which the switch replicates to all ranks
// (NVLS).  No staging copy, no separate collective launch; the scan's loads and the remote stores overlap element by
// element.  Ordering is the caller's job: a cross-rank barrier on the stream before (everybody is done reading the last
// rollout's values) and after (all stores have landed) — sgb_gae_allgather in the header.
constexpr int kMaxPeers = 16;
struct GaePeers {
    float* adv[kMaxPeers];      // base of rank w's [world, T, BN] advantage buffer as mapped on THIS device
    float* tgt[kMaxPeers];
    float* adv_mc;              // multicast mapping of the same buffers (NULL: store to every peer)
    float* tgt_mc;
};
__device__ __forceinline__ void multimem_st(float* p, float v) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
// V = 4: one thread scans four consecutive (env, agent) columns (same env: N % 4 == 0) with 16-byte loads and stores —
// a warp's store to a peer is then 512 contiguous bytes, which the links carry better than 128 (measured on 4 B200s:
// see DESIGN.md §5).  V = 1 is the general layout.
template <bool MC, int V>
__global__ void __launch_bounds__(256) gae_allgather_kernel(int T, int BN, int N, const float* __restrict__ reward,
                                                            const float* __restrict__ value,
                                                            const float* __restrict__ next_value,
                                                            const uint8_t* __restrict__ done, float gamma, float lmbda,
                                                            const GaePeers peers, int world, int rank) {
    pdl_launch_dependents();
    pdl_wait();
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) * V;
    if (k >= BN) return;
    const int b = k / N;
    const int B = BN / N;
    const size_t slot = (size_t)rank * T * BN;
    float a[V];
#pragma unroll
    for (int j = 0; j < V; j++) a[j] = 0.0f;
#pragma unroll 2
    for (int t = T - 1; t >= 0; t--) {
        const size_t o = (size_t)t * BN + k;
        const float nd = 1.0f - (float)done[(size_t)t * B + b];
        float r[V], v[V], nv[V], tg[V];
        if (V == 4) {
            *reinterpret_cast<float4*>(r) = *reinterpret_cast<const float4*>(reward + o);
            *reinterpret_cast<float4*>(v) = *reinterpret_cast<const float4*>(value + o);
            *reinterpret_cast<float4*>(nv) = *reinterpret_cast<const float4*>(next_value + o);
        } else {
            r[0] = reward[o]; v[0] = value[o]; nv[0] = next_value[o];
        }
#pragma unroll
        for (int j = 0; j < V; j++) {
            const float delta = r[j] + gamma * nv[j] * nd - v[j];
            a[j] = delta + gamma * lmbda * nd * a[j];
            tg[j] = a[j] + v[j];
        }
        if (MC) {
#pragma unroll
            for (int j = 0; j < V; j++) {
                multimem_st(peers.adv_mc + slot + o + j, a[j]);
                multimem_st(peers.tgt_mc + slot + o + j, tg[j]);
            }
        } else {
            for (int w = 0; w < world; w++) {
                const int dst = (rank + w) % world;          // own copy first, then round the ring: spreads the links
                if (V == 4) {
                    *reinterpret_cast<float4*>(peers.adv[dst] + slot + o) = *reinterpret_cast<const float4*>(a);
                    *reinterpret_cast<float4*>(peers.tgt[dst] + slot + o) = *reinterpret_cast<const float4*>(tg);
                } else {
                    peers.adv[dst][slot + o] = a[0];
                    peers.tgt[dst][slot + o] = tg[0];
                }
            }
        }
    }
}

// byte mask -> compacted index list (order is irrelevant: envs are independent)
__global__ void mask_to_list_kernel(const uint8_t* mask, int B, int32_t* list, int32_t* count) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < B && mask[e]) list[atomicAdd(count, 1)] = e;
}

} // namespace sgb
