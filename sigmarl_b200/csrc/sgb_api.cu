// sgb_api.cu — the C-ABI of libsigmarl_b200.so (include/sigmarl_b200.h): context, map packing, launches.
// There is deliberately NO CPU implementation here: without a CUDA device sgb_create fails.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only; a no-op unless a profiler is attached

#include "sgb_kernels.cuh"
#ifdef SGB_TEST_HOOKS
#include "../../include/sigmarl_b200_test.h"
#endif

using namespace sgb;

struct sgb_ctx {
    int device = 0;
    sgb_config cfg{};
    unsigned char* d_blob = nullptr;
    float* d_yaw = nullptr;          // yaw per centre point, indexed like the blob's centre points
    float* d_spawn = nullptr;        // spawn table [centre points][8] (sgb_kernels.cuh: place_agent)
    float* d_fresh = nullptr;        // [fresh_cap][4] scratch: spawn-pose boundary distances for the observation refresh
    float* d_lanelet_xy = nullptr;   // sgb_set_lanelets: centre lines of all lanelets / offsets / adjacency matrix
    int32_t* d_lanelet_off = nullptr;
    uint8_t* d_lanelet_adj = nullptr;
    int32_t n_lanelets = 0, lanelet_max_len = 0;
    int64_t fresh_cap = 0;           // capacity of d_fresh in agents
    int32_t n_points = 0;            // centre points incl. extension slots (rows of d_yaw / d_spawn)
    int32_t* d_list = nullptr;       // [cap] compacted env indices for a masked refresh
    int32_t* d_count = nullptr;      // [4]: 0 / 1 the reset's list length (alternating, see ResetParams::count_next), 2 sgb_refresh's
    int count_phase = 0;
    bool pdl = true;                 // programmatic dependent launch between the library's kernels (SGB_NO_PDL=1 turns it off)
    int32_t list_cap = 0;
    int32_t blob_bytes = 0;
    int32_t n_paths = 0;
    int32_t max_center = 0;
    int32_t num_sms = 0;
    int32_t max_smem_optin = 0;
    int64_t launches = 0;
    int32_t n_sets = 0;              // sgb_set_path_sets
    int32_t set_lo[4] = {0, 0, 0, 0}, set_hi[4] = {0, 0, 0, 0};
    float set_cum[4] = {1.0f, 1.0f, 1.0f, 1.0f};
    uint64_t noise_epoch = 0;        // counts API calls that can write an observation: part of the noise key
    int64_t env_offset = 0;          // global index of env 0 (sgb_set_env_offset / the reset entry points)
    size_t smem_configured[6][5] = {};   // dynamic-smem opt-in done for <MODE + 2 * OV, G (and MB = 2)> on this device
    int32_t smem_per_sm = 0;
    cudaStream_t pipe_stream[2] = {nullptr, nullptr};   // sgb_step_host: chunked copy/compute pipeline
    cudaEvent_t pipe_event[2] = {nullptr, nullptr};
    cudaEvent_t pipe_start = nullptr;
    bool pipe_ready = false;
    int32_t* pipe_list[2] = {nullptr, nullptr};         // per-stream reset scratch of sgb_step_reset_host
    int32_t* pipe_count[2] = {nullptr, nullptr};         // [2] each, alternating like d_count
    int pipe_phase[2] = {0, 0};
    int32_t pipe_list_cap = 0;
};

static thread_local char g_err[256] = "";

static int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
    return SGB_ERR_CUDA;
}
#define CK(call)                                             \
    do {                                                     \
        cudaError_t _e = (call);                             \
        if (_e != cudaSuccess) return cuda_fail(_e, #call);  \
    } while (0)

// Every entry point that touches the device runs on the CONTEXT's device and leaves the caller's current device as it
// found it (a host framework such as PyTorch keeps its own notion of the current device).
struct DeviceGuard {
    int prev = -1, dev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device) : dev(device) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        if (prev >= 0 && prev != dev) cudaSetDevice(prev);
    }
};
// NVTX range over an API call (the reference's counterpart: timer.step_duration, road_traffic.py:955-962): shows up as
// "sgb_step" / "sgb_reset" / ... on the CPU timeline of nsys / ncu --nvtx, brackets the launches the call enqueues
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

#define GUARD(ctx)                                                   \
    DeviceGuard _guard((ctx)->device);                               \
    if (_guard.err != cudaSuccess) return cuda_fail(_guard.err, "cudaSetDevice(context device)")

extern "C" const char* sgb_last_error(void) { return g_err; }
extern "C" int sgb_version(void) { return SGB_VERSION; }
#ifdef SGB_TEST_HOOKS   // host-side self-test hooks: libsigmarl_b200_test.so only (include/sigmarl_b200_test.h)
extern "C" float sgb_debug_mtv_distance(const float* vi, const float* vj) {
    // HOST build of the very source the MTV kernels compile (mtv_from_vertices): arithmetic self-test without a GPU
    float ax[4], ay[4], bx[4], by[4];
    for (int k = 0; k < 4; k++) { ax[k] = vi[2 * k]; ay[k] = vi[2 * k + 1]; bx[k] = vj[2 * k]; by[k] = vj[2 * k + 1]; }
    return sgb::mtv_from_vertices(ax, ay, bx, by);
}
extern "C" int sgb_debug_current_lanelet(int32_t n, const float* xy, const int32_t* off, float x, float y) {
    if (n <= 0 || !xy || !off) return SGB_ERR_ARG;
    int max_len = 0;
    for (int l = 0; l < n; l++) max_len = std::max(max_len, off[l + 1] - off[l]);
    return sgb::current_lanelet(reinterpret_cast<const float2*>(xy), off, n, max_len, x, y);   // host build of the kernels' function
}
#endif
extern "C" const char* sgb_status_string(int s) {
    switch (s) {
        case SGB_OK: return "ok";
        case SGB_ERR_ARG: return "invalid argument";
        case SGB_ERR_CUDA: return "CUDA runtime error";
        case SGB_ERR_NO_DEVICE: return "no usable CUDA device (this library has no CPU fallback)";
        case SGB_ERR_MAP: return "map rejected (degenerate polyline or does not fit in shared memory)";
        case SGB_ERR_UNSUPPORTED: return "configuration outside the supported hot path";
        default: return "unknown status";
    }
}

// ---- map packing (host) -------------------------------------------------------------------------------
namespace {

struct Packed {
    std::vector<unsigned char> blob;
    std::vector<float> yaw;
    int max_center = 0;
};

void add_boxes(std::vector<float>& boxes, const float* xy, int n_pts) {
    const int nseg = n_pts - 1;
    for (int s0 = 0; s0 < nseg; s0 += kChunk) {
        const int s1 = std::min(s0 + kChunk, nseg);
        float x0 = xy[2 * s0], x1 = x0, y0 = xy[2 * s0 + 1], y1 = y0;
        for (int k = s0; k <= s1; k++) {
            x0 = std::min(x0, xy[2 * k]); x1 = std::max(x1, xy[2 * k]);
            y0 = std::min(y0, xy[2 * k + 1]); y1 = std::max(y1, xy[2 * k + 1]);
        }
        // stored as centre + half extents; the half extents are rounded UP from the exact distances between the
        // (rounded) centre and the extremes, so the stored box contains every point of the chunk
        const float cx = 0.5f * (x0 + x1), cy = 0.5f * (y0 + y1);
        auto up = [](double v) { float f = (float)v; return (double)f < v ? std::nextafterf(f, INFINITY) : f; };
        boxes.insert(boxes.end(), {cx, cy, up(std::max((double)cx - x0, (double)x1 - cx)),
                                   up(std::max((double)cy - y0, (double)y1 - cy))});
    }
}

// Direction cone of a chunk's segments as undirected lines (angles mod pi): mid angle and half width, widened by
// asin(kCollinear) + the fp16 storage error.  A far segment can only fire interX through fp32 noise if it is
// collinear with an edge of the rectangle within kCollinear (sgb_kernels.cuh), so a chunk whose cone contains
// neither the heading nor its normal holds no such segment.  half >= pi/2 means "always test".
void add_cones(std::vector<__half2>& cones, const float* xy, int n_pts) {
    const double PI = 3.14159265358979323846;
    const int nseg = n_pts - 1;
    for (int s0 = 0; s0 < nseg; s0 += kChunk) {
        const int s1 = std::min(s0 + kChunk, nseg);
        double a0 = 0.0, lo = 0.0, hi = 0.0;
        for (int k = s0; k < s1; k++) {
            const double a = std::atan2((double)xy[2 * k + 3] - xy[2 * k + 1], (double)xy[2 * k + 2] - xy[2 * k]);
            if (k == s0) { a0 = a; continue; }
            double rel = std::fmod(a - a0, PI);          // (-pi, pi)
            if (rel > PI / 2) rel -= PI;
            if (rel <= -PI / 2) rel += PI;                // (-pi/2, pi/2]
            lo = std::min(lo, rel); hi = std::max(hi, rel);
        }
        double half = 0.5 * (hi - lo), mid = a0 + 0.5 * (hi + lo);
        // a chunk that turns by more than ~80 degrees may not unwrap uniquely: always test it
        if (hi - lo > 1.4) half = 4.0;
        else half += std::asin((double)kCollinear) + 4e-3;  // + fp16 rounding of mid (<= 1e-3) and half, fp32 psi mod pi
        mid = std::fmod(mid, PI);
        if (mid < 0) mid += PI;
        cones.push_back(__halves2half2(__float2half_rn((float)mid), __float2half_rn((float)half)));
    }
}

bool degenerate(const float* xy, int n) {
    if (n < 2) return true;
    for (int s = 0; s + 1 < n; s++)
        if (xy[2 * s] == xy[2 * s + 2] && xy[2 * s + 1] == xy[2 * s + 3]) return true; // zero-length segment -> NaN in the reference
    return false;
}

int pack_map(const sgb_map_desc* m, Packed& out) {
    if (!m || m->n_paths <= 0 || !m->center_xy || !m->center_off || !m->left_xy || !m->left_off || !m->right_xy ||
        !m->right_off || !m->center_yaw || !m->is_loop)
        return SGB_ERR_ARG;
    const int n = m->n_paths;
    std::vector<PathRec> recs(n);
    std::vector<float> pts, boxes;
    std::vector<__half2> cones;
    out.yaw.clear();
    int yaw_in = 0;
    for (int i = 0; i < n; i++) {
        PathRec& r = recs[i];
        std::memset(&r, 0, sizeof r);
        const int nc = m->center_off[i + 1] - m->center_off[i];
        const int nl = m->left_off[i + 1] - m->left_off[i];
        const int nr = m->right_off[i + 1] - m->right_off[i];
        const float* c = m->center_xy + 2 * (size_t)m->center_off[i];
        const float* l = m->left_xy + 2 * (size_t)m->left_off[i];
        const float* rr = m->right_xy + 2 * (size_t)m->right_off[i];
        if (nc < 8 || degenerate(c, nc) || degenerate(l, nl) || degenerate(rr, nr)) return SGB_ERR_MAP;
        if (std::max(nc, std::max(nl, nr)) - 1 > 32 * kChunk) return SGB_ERR_MAP; // candidate-chunk masks are 32-bit
        out.max_center = std::max(out.max_center, nc);
        r.is_loop = m->is_loop[i] ? 1 : 0;
        r.c_off = (int)(pts.size() / 2);
        r.n_c = nc;
        r.cbox = (int)(boxes.size() / 4);
        add_boxes(boxes, c, nc);
        pts.insert(pts.end(), c, c + 2 * nc);
        // world_state_rt.py:279-311: extension points last + m * (last - prev), m = 1..6 (fp32, no contraction)
        {
            volatile float dx = c[2 * (nc - 1)] - c[2 * (nc - 2)];
            volatile float dy = c[2 * (nc - 1) + 1] - c[2 * (nc - 2) + 1];
            for (int k = 1; k <= kExt; k++) {
                volatile float mx = (float)k * dx, my = (float)k * dy;
                volatile float ex = c[2 * (nc - 1)] + mx, ey = c[2 * (nc - 1) + 1] + my;
                pts.push_back((float)ex);
                pts.push_back((float)ey);
            }
        }
        // yaw aligned with the centre points (n-1 values + padding for the extension slots)
        for (int k = 0; k < nc - 1; k++) out.yaw.push_back(m->center_yaw[yaw_in + k]);
        yaw_in += nc - 1;
        for (int k = 0; k < kExt + 1; k++) out.yaw.push_back(0.0f);
        r.l_off = (int)(pts.size() / 2);
        r.n_l = nl;
        r.lbox = (int)(boxes.size() / 4);
        add_boxes(boxes, l, nl);
        r.lcone = (int)cones.size();
        add_cones(cones, l, nl);
        pts.insert(pts.end(), l, l + 2 * nl);
        r.r_off = (int)(pts.size() / 2);
        r.n_r = nr;
        r.rbox = (int)(boxes.size() / 4);
        add_boxes(boxes, rr, nr);
        r.rcone = (int)cones.size();
        add_cones(cones, rr, nr);
        pts.insert(pts.end(), rr, rr + 2 * nr);
    }
    // yaw must be indexable by (c_off + point): rebuild it on the blob's point numbering
    std::vector<float> yaw_by_pt(pts.size() / 2, 0.0f);
    {
        size_t src = 0;
        for (int i = 0; i < n; i++) {
            for (int k = 0; k < recs[i].n_c + kExt; k++) yaw_by_pt[recs[i].c_off + k] = out.yaw[src + k];
            src += recs[i].n_c + kExt;
        }
    }
    out.yaw.swap(yaw_by_pt);
    auto align16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    BlobHeader h{};
    h.n_paths = n;
    h.path_off = (int32_t)align16(sizeof(BlobHeader));
    h.pts_off = (int32_t)align16(h.path_off + sizeof(PathRec) * n);
    h.box_off = (int32_t)align16(h.pts_off + sizeof(float) * pts.size());
    h.cone_off = (int32_t)align16(h.box_off + sizeof(float) * boxes.size());
    h.total_bytes = (int32_t)align16(h.cone_off + sizeof(__half2) * cones.size());
    out.blob.assign(h.total_bytes, 0);
    std::memcpy(out.blob.data(), &h, sizeof h);
    std::memcpy(out.blob.data() + h.path_off, recs.data(), sizeof(PathRec) * n);
    std::memcpy(out.blob.data() + h.pts_off, pts.data(), sizeof(float) * pts.size());
    std::memcpy(out.blob.data() + h.box_off, boxes.data(), sizeof(float) * boxes.size());
    std::memcpy(out.blob.data() + h.cone_off, cones.data(), sizeof(__half2) * cones.size());
    return SGB_OK;
}

int check_buffers(const sgb_buffers* b, int step) {
    if (!b || !b->pose || !b->aux || !b->path_id || !b->carry || !b->agent_flags) return SGB_ERR_ARG;
    if (step && (!b->action || !b->step_count || !b->obs || !b->reward || !b->done)) return SGB_ERR_ARG;
    return SGB_OK;
}

// lanes per agent: a warp must hold whole envs (N * G <= 32); as many lanes as still fit, at most 4
int pick_group(int N) {
    if (N <= 8) return 4;
    if (N <= 16) return 2;
    return 1;
}

int ensure_list(sgb_ctx* c, int B) {
    if (c->list_cap >= B) return SGB_OK;
    cudaFree(c->d_list);
    c->d_list = nullptr;
    CK(cudaMalloc(&c->d_list, sizeof(int32_t) * (size_t)B));
    if (!c->d_count) {
        CK(cudaMalloc(&c->d_count, 4 * sizeof(int32_t)));
        CK(cudaMemset(c->d_count, 0, 4 * sizeof(int32_t)));
    }
    c->list_cap = B;
    return SGB_OK;
}

int ensure_fresh(sgb_ctx* c, int64_t agents) {
    if (c->fresh_cap >= agents) return SGB_OK;
    cudaFree(c->d_fresh);
    c->d_fresh = nullptr;
    CK(cudaMalloc(&c->d_fresh, sizeof(float) * 4 * (size_t)agents));
    c->fresh_cap = agents;
    return SGB_OK;
}

// Launch of a kernel that calls pdl_wait() (sgb_kernels.cuh): with programmatic stream serialization it may begin — map
// staging, block scheduling — while the previous kernel of the stream drains.  Harmless after anything that is not such a
// kernel (the dependency is then the usual full one).
template <typename P>
int launch_chained(sgb_ctx* ctx, void (*kernel)(const P), dim3 grid, dim3 block, size_t smem, cudaStream_t st, const P& p) {
    cudaLaunchConfig_t lc{};
    lc.gridDim = grid; lc.blockDim = block; lc.dynamicSmemBytes = smem; lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at;
    lc.numAttrs = ctx->pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&lc, kernel, p));
    return SGB_OK;
}

template <int G, int MODE, int OV, int MB>
int launch_env_kernel_mb(sgb_ctx* ctx, Params& p, cudaStream_t st, size_t smem, int n_wt) {
    // the opt-in is per (kernel, device): remember it in the context, which is bound to one device
    size_t& configured = ctx->smem_configured[MODE + 2 * OV][G == 4 ? 0 : (G == 2 ? 1 : 2) + (MB == 2 ? 2 : 0)];
    if (configured < smem) {
        CK(cudaFuncSetAttribute(env_step_kernel<G, MODE, OV, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int warps = cta_threads(G) / 32;
    // persistent grid: as many CTAs as are resident at once
    const int grid = std::min((n_wt + warps - 1) / warps, ctx->num_sms * MB);
    const int rc = launch_chained(ctx, env_step_kernel<G, MODE, OV, MB>, dim3(grid), dim3(cta_threads(G)), smem, st, p);
    if (rc) return rc;
    ctx->launches++;
    return SGB_OK;
}

template <int G, int MODE, int OV>
int launch_env_kernel(sgb_ctx* ctx, Params& p, cudaStream_t st) {
    const int slots = kSlots;
    const int envs_per_warp = 32 / (p.N * G);
    if (envs_per_warp < 1) return SGB_ERR_ARG;
    const int n_wt = (p.B + envs_per_warp - 1) / envs_per_warp;   // upper bound (a list may hold fewer envs)
    // OV 2 (MTV distance) keeps cos / sin of the pre-step heading per slot behind dij
    const size_t smem = ((size_t)ctx->blob_bytes + 127) / 128 * 128 + tile_smem_bytes(slots, p.N) + 128 +
                        (OV == 2 ? 2 * sizeof(float) * (size_t)slots : 0);
    if ((int64_t)smem > ctx->max_smem_optin) {
        snprintf(g_err, sizeof g_err, "map blob %d B + tile arrays need %zu B of shared memory, device offers %d B",
                 ctx->blob_bytes, smem, ctx->max_smem_optin);
        return SGB_ERR_MAP;
    }
    // two- and one-lane kernels (512 / 256 threads): two CTAs per SM when the map leaves room for two sets of tile arrays
    // (each CTA also costs 1 KB of system-reserved shared memory) — see env_step_kernel's MB
    if (G <= 2 && 2 * (smem + 1024) <= (size_t)ctx->smem_per_sm)
        return launch_env_kernel_mb<G, MODE, OV, (G <= 2 ? 2 : 1)>(ctx, p, st, smem, n_wt);
    return launch_env_kernel_mb<G, MODE, OV, 1>(ctx, p, st, smem, n_wt);
}

template <int MODE, int OV>
int launch_env_group(sgb_ctx* ctx, Params& p, cudaStream_t st, int g) {
    if (g == 4) return launch_env_kernel<4, MODE, OV>(ctx, p, st);
#ifdef SGB_DEV_ONLY_G4   // development builds (profiles/scripts/variants.sh): compile the headline instantiations only
    snprintf(g_err, sizeof g_err, "development build: only the four-lanes-per-agent kernels were compiled");
    return SGB_ERR_UNSUPPORTED;
#else
    if (g == 2) return launch_env_kernel<2, MODE, OV>(ctx, p, st);
    return launch_env_kernel<1, MODE, OV>(ctx, p, st);
#endif
}

int launch_env(sgb_ctx* ctx, int B, int N, const sgb_buffers* buf, int mode, const int32_t* env_list,
               const int32_t* env_count, int write_obs, cudaStream_t st, int skip_scan = 0,
               const sgb_config* cfg_override = nullptr, int64_t env_first = 0, const float* fresh = nullptr) {
    Params p{};
    p.noise_epoch = ctx->noise_epoch;
    p.env_base = ctx->env_offset + env_first;   // global index of buf's env 0 (env_first: chunked launches)
    p.cfg = cfg_override ? *cfg_override : ctx->cfg;
    // spawn-table refresh: the table holds no boundary indices, so a layout with boundary points runs the scans
    p.skip_scan = (p.cfg.obs_flags & SGB_OBS_BOUNDARY_POINTS) ? 0 : skip_scan;
    {   // small batches read the few map entries a spawn-table refresh needs from global memory (see map_global)
        const int g = pick_group(N);
        const int wave = ctx->num_sms * (cta_threads(g) / 32) * std::max(1, 32 / (N * g));
        if (p.skip_scan && B <= 4 * wave) p.skip_scan = 2;
    }
    p.fresh = fresh ? fresh : ctx->d_fresh;
    p.buf = *buf;
    p.blob = ctx->d_blob;
    p.env_list = env_list;
    p.env_count = env_count;
    p.B = B; p.N = N; p.D = obs_dim_of(p.cfg.obs_flags, p.cfg.k_near);
    p.blob_bytes = ctx->blob_bytes;
    p.mode = mode;
    p.write_obs = write_obs;
    p.rect_radius = std::sqrt(ctx->cfg.half_length * ctx->cfg.half_length + ctx->cfg.half_width * ctx->cfg.half_width) * 1.0001f;
    p.near2 = (p.rect_radius + kFarMargin) * (p.rect_radius + kFarMargin);
    p.band_l = ctx->cfg.half_length + 1e-3f;
    p.band_w = ctx->cfg.half_width + 1e-3f;
    p.r_pos = 1.0f / ctx->cfg.norm_pos;
    p.r_v = 1.0f / ctx->cfg.norm_v;
    p.r_dist = 1.0f / ctx->cfg.norm_dist;
    if ((p.cfg.obs_flags & SGB_OBS_MASK_LANELETS) && (p.cfg.k_near == 0 || (mode != 0 && !write_obs))) {
        // nobody is observed in this launch (e.g. the spawn-table build inside sgb_create, which runs before the caller
        // can upload a lanelet table): the lanelet criterion has nothing to act on
        p.cfg.obs_flags &= ~SGB_OBS_MASK_LANELETS;
    }
    if (p.cfg.obs_flags & SGB_OBS_MASK_LANELETS) {
        if (!ctx->d_lanelet_xy || !(p.cfg.obs_flags & SGB_OBS_APPLY_MASK)) {
            snprintf(g_err, sizeof g_err, "SGB_OBS_MASK_LANELETS needs SGB_OBS_APPLY_MASK and a lanelet table (sgb_set_lanelets)");
            return SGB_ERR_ARG;
        }
        p.lanelet_xy = reinterpret_cast<const float2*>(ctx->d_lanelet_xy);
        p.lanelet_off = ctx->d_lanelet_off;
        p.lanelet_adj = ctx->d_lanelet_adj;
        p.n_lanelets = ctx->n_lanelets;
        p.lanelet_max_len = ctx->lanelet_max_len;
    }
    // the kernels index agents and observation elements with 32 bits
    if ((uint64_t)B * (uint64_t)N * (uint64_t)std::max(p.D, (int32_t)SGB_INFO_DIM) >= (1ull << 32)) {
        snprintf(g_err, sizeof g_err, "B * N * max(D, 16) = %llu does not fit 32-bit indexing: shard the batch",
                 (unsigned long long)B * N * std::max(p.D, (int32_t)SGB_INFO_DIM));
        return SGB_ERR_ARG;
    }
    const int g = pick_group(N);
#ifdef SGB_DEV_ONLY_G4
    if (p.cfg.use_mtv_distance || p.cfg.obs_flags != 0 || p.cfg.obs_noise_level > 0.0f) {
        snprintf(g_err, sizeof g_err, "development build: only the default observation layout was compiled");
        return SGB_ERR_UNSUPPORTED;
    }
    return mode == 0 ? launch_env_group<0, 0>(ctx, p, st, g) : launch_env_group<1, 0>(ctx, p, st, g);
#endif
    // MTV agent distance: its own instantiation (flag-driven writer + SAT distance in phase C1)
    if (p.cfg.use_mtv_distance) return mode == 0 ? launch_env_group<0, 2>(ctx, p, st, g) : launch_env_group<1, 2>(ctx, p, st, g);
    // the default observation layout runs the hard-wired (tuned) writer, any other one the flag-driven writer
    if (p.cfg.obs_flags == 0 && !(p.cfg.obs_noise_level > 0.0f)) return mode == 0 ? launch_env_group<0, 0>(ctx, p, st, g) : launch_env_group<1, 0>(ctx, p, st, g);
    return mode == 0 ? launch_env_group<0, 1>(ctx, p, st, g) : launch_env_group<1, 1>(ctx, p, st, g);
}

// Spawn table.  A device reset puts an agent exactly on a centre-line point with the path's yaw, so everything
// sgb_refresh would derive from that pose (centre / boundary distances, closest index) depends on (path, point)
// only.  It is computed ONCE here — by the refresh kernel itself, on one single-agent env per centre point, so the
// values are bit-identical to a refresh of the reset pose — and looked up by reset_kernel afterwards: a masked
// reset then needs no polyline scan at all.
int build_spawn_table(sgb_ctx* c, const Packed& pk) {
    const BlobHeader* h = reinterpret_cast<const BlobHeader*>(pk.blob.data());
    const PathRec* recs = reinterpret_cast<const PathRec*>(pk.blob.data() + h->path_off);
    const int M = c->n_points;
    std::vector<int32_t> path(M, 0), point(M, 0);
    std::vector<uint8_t> mask(M, 0);
    for (int i = 0; i < h->n_paths; i++)
        for (int k = 0; k < recs[i].n_c; k++) {
            const int r = recs[i].c_off + k;     // row of d_yaw / d_spawn; NB c_off counts blob points (incl. boundaries)
            if (r < 0 || r >= M) return SGB_ERR_MAP;
            path[r] = i; point[r] = k; mask[r] = 1;
        }
    float *d_f = nullptr;                       // pose, aux, carry [M,4] each, speed [M], dbg [M,16]
    int32_t* d_i = nullptr;                     // path_id, path, point [M] each
    uint8_t* d_b = nullptr;                     // agent_flags, mask [M] each
    auto cleanup = [&]() { cudaFree(d_f); cudaFree(d_i); cudaFree(d_b); };
    if (cudaMalloc(&d_f, sizeof(float) * (size_t)M * (12 + 1 + 16)) != cudaSuccess ||
        cudaMalloc(&d_i, sizeof(int32_t) * (size_t)M * 3) != cudaSuccess ||
        cudaMalloc(&d_b, (size_t)M * 2) != cudaSuccess) { cleanup(); return cuda_fail(cudaErrorMemoryAllocation, "spawn table"); }
    cudaMemset(d_f, 0, sizeof(float) * (size_t)M * 29);
    cudaMemset(d_i, 0, sizeof(int32_t) * (size_t)M * 3);
    cudaMemcpy(d_i + M, path.data(), sizeof(int32_t) * M, cudaMemcpyHostToDevice);
    cudaMemcpy(d_i + 2 * M, point.data(), sizeof(int32_t) * M, cudaMemcpyHostToDevice);
    cudaMemcpy(d_b + M, mask.data(), M, cudaMemcpyHostToDevice);
    sgb_buffers b{};
    b.pose = d_f; b.aux = d_f + 4 * (size_t)M; b.carry = d_f + 8 * (size_t)M; b.dbg = d_f + 13 * (size_t)M;
    b.path_id = d_i; b.agent_flags = d_b;
    PlaceParams pp{};
    pp.cfg = c->cfg; pp.buf = b; pp.blob = c->d_blob; pp.yaw = c->d_yaw; pp.agent_mask = d_b + M;
    pp.path = d_i + M; pp.point = d_i + 2 * M; pp.speed = d_f + 12 * (size_t)M; pp.B = M; pp.N = 1;
    place_kernel<<<(M + 255) / 256, 256>>>(pp);
    sgb_config one = c->cfg;
    one.k_near = 0;                              // single-agent envs: nobody to observe
    int rc = launch_env(c, M, 1, &b, 1, nullptr, nullptr, 0, nullptr, 0, &one);
    if (rc == SGB_OK && cudaDeviceSynchronize() != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "spawn table kernels");
    if (rc != SGB_OK) { cleanup(); return rc; }
    std::vector<float> dbg((size_t)M * 16), car((size_t)M * 4), tab((size_t)M * 8, 0.0f);
    cudaMemcpy(dbg.data(), b.dbg, sizeof(float) * dbg.size(), cudaMemcpyDeviceToHost);
    cudaMemcpy(car.data(), b.carry, sizeof(float) * car.size(), cudaMemcpyDeviceToHost);
    for (int r = 0; r < M; r++) {
        const float* d = &dbg[(size_t)r * 16];
        float* t = &tab[(size_t)r * 8];
        t[0] = d[0]; t[1] = d[1];                // d_ref, idx_ref (int bits)
        t[2] = d[2]; t[3] = d[7];                // dLc, dRc (already minus half width)
        // minimum over the vertices: what the refresh wrote into the carry of agent 0 (the very value a generic
        // refresh of this pose computes — one root of the minimal squared distance)
        t[4] = car[(size_t)r * 4 + 1];
        t[5] = car[(size_t)r * 4 + 2];
    }
    cleanup();
    CK(cudaMalloc(&c->d_spawn, sizeof(float) * tab.size()));
    CK(cudaMemcpy(c->d_spawn, tab.data(), sizeof(float) * tab.size(), cudaMemcpyHostToDevice));
    c->launches = 0;                             // set-up launches are not part of anyone's step accounting
    return SGB_OK;
}

} // namespace

// device half of sgb_create; on any failure the caller destroys the context (sgb_destroy copes with a partial one)
static int init_device_state(sgb_ctx* c, const Packed& pk) {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, c->device));
    c->num_sms = prop.multiProcessorCount;
    c->max_smem_optin = (int32_t)prop.sharedMemPerBlockOptin;
    c->smem_per_sm = (int32_t)prop.sharedMemPerMultiprocessor;
    if (prop.major < 10) {
        snprintf(g_err, sizeof g_err, "device %d is sm_%d%d; this library is built for sm_100a only", c->device, prop.major, prop.minor);
        return SGB_ERR_NO_DEVICE;
    }
    CK(cudaMalloc(&c->d_blob, pk.blob.size()));
    CK(cudaMemcpy(c->d_blob, pk.blob.data(), pk.blob.size(), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&c->d_yaw, pk.yaw.size() * sizeof(float)));
    CK(cudaMemcpy(c->d_yaw, pk.yaw.data(), pk.yaw.size() * sizeof(float), cudaMemcpyHostToDevice));
    c->n_points = (int32_t)pk.yaw.size();
    return build_spawn_table(c, pk);
}

// ---- C-ABI ---------------------------------------------------------------------------------------------
#ifdef SGB_TEST_HOOKS
extern "C" int sgb_debug_pack_map(const sgb_map_desc* map, int64_t* blob_bytes) {
    Packed pk;
    const int rc = pack_map(map, pk);
    if (blob_bytes) *blob_bytes = rc == SGB_OK ? (int64_t)pk.blob.size() : 0;
    return rc;
}
// Host run of the kernels' own polyline scans (scan_center<1>, scan_boundary<1>: one lane per agent, the group
// reductions are the identity) on a freshly packed blob — the phase-B glue of env_step_kernel for `n` independent poses.
// out[16 * i]: 0 d_ref, 1 idx_ref, then per side (left at 2, right at 9): d_cg, d_vertex[4], hit, spare.  Lets a
// machine without a GPU check that the pruned search equals the exhaustive one (`exhaustive` = 0 / 1) on any map.
extern "C" int sgb_debug_scan_batch(const sgb_map_desc* map, int32_t n, const int32_t* path, const float* x, const float* y,
                                    const float* psi, const int32_t* hint_idx, float half_length, float half_width,
                                    int32_t exhaustive, float* out) {
    if (n < 0 || !path || !x || !y || !psi || !hint_idx || !out) return SGB_ERR_ARG;
    Packed pk;
    const int rc = pack_map(map, pk);
    if (rc != SGB_OK) return rc;
    const unsigned char* blob = pk.blob.data();
    const BlobHeader* h = reinterpret_cast<const BlobHeader*>(blob);
    const PathRec* paths = reinterpret_cast<const PathRec*>(blob + h->path_off);
    const float2* pts = reinterpret_cast<const float2*>(blob + h->pts_off);
    const float4* boxes = reinterpret_cast<const float4*>(blob + h->box_off);
    const __half2* cones = reinterpret_cast<const __half2*>(blob + h->cone_off);
    const float rect_radius = std::sqrt(half_length * half_length + half_width * half_width) * 1.0001f;   // launch_env
    const float near2 = (rect_radius + kFarMargin) * (rect_radius + kFarMargin);
    // bit 1 of `exhaustive`: run the scans the way the product kernels do when no debug buffer is bound — only the
    // minimum over the four vertices is exact then (all four outputs hold it), which allows a tighter chunk vote
    const bool want_dv = (exhaustive & 2) == 0;
    exhaustive &= 1;
    for (int i = 0; i < n; i++) {
        if (path[i] < 0 || path[i] >= h->n_paths) return SGB_ERR_ARG;
        const PathRec* prp = paths + path[i];
        const float px = x[i], py = y[i];
        const float sy = sinf(psi[i]), cy = cosf(psi[i]);
        float rvx[4], rvy[4];
        sgb::rect_of_pose(px, py, cy, sy, half_length, half_width, rvx, rvy);     // phase A's vertices
        const float psim = fmaf(-3.14159274f, floorf(psi[i] * 0.318309873f), psi[i]);   // phase A's heading mod pi
        float* o = out + 16 * (size_t)i;
        for (int k = 0; k < 16; k++) o[k] = 0.0f;
        float d_ref;
        int idx_ref;
        sgb::scan_center<1>(pts + prp->c_off, boxes + prp->cbox, prp->n_c, hint_idx[i] - 1, exhaustive != 0, px, py, 0, d_ref, idx_ref);
        o[0] = d_ref; o[1] = (float)idx_ref;
        const int h2 = idx_ref - 1;
        for (int side = 0; side < 2; side++) {
            float dc, dvv[4], m4;
            bool hit;
            sgb::scan_boundary<1>(pts + (side ? prp->r_off : prp->l_off), boxes + (side ? prp->rbox : prp->lbox),
                                  cones + (side ? prp->rcone : prp->lcone), side ? prp->n_r : prp->n_l, h2, exhaustive != 0, px, py,
                                  &cy, &sy, &psim, rvx, rvy, rect_radius, near2, half_length, half_width, half_length + 1e-3f,
                                  half_width + 1e-3f, want_dv, 0, dc, dvv, m4, hit);
            float* q = o + (side ? 9 : 2);
            q[0] = dc;
            for (int v = 0; v < 4; v++) q[1 + v] = dvv[v];
            q[5] = hit ? 1.0f : 0.0f;
        }
    }
    return SGB_OK;
}

// Host builds of the kernels' small helpers, against the reference's known-answer vectors (tests/test_abi_and_host.py):
// which 0 wrap_pi(in[0]) (angle_eliminate_two_pi), 1 dec_lin(in[0], in[1], in[2]) (decreasing_fcn, linear),
// 2 kth_nearest(in[1..N], N = n - 1, kk = (int)in[0]) -> out[0] index, out[1] distance (torch.topk order)
extern "C" int sgb_debug_helper(int32_t which, const float* in, int32_t n, float* out) {
    if (!in || !out) return SGB_ERR_ARG;
    if (which == 0) { out[0] = sgb::wrap_pi(in[0]); return SGB_OK; }
    if (which == 1) { out[0] = sgb::dec_lin(in[0], in[1], in[2]); return SGB_OK; }
    if (which == 2 && n >= 2 && n - 1 <= SGB_MAX_AGENTS) {
        float d = 0.0f;
        out[0] = (float)sgb::kth_nearest(in + 1, n - 1, (int)in[0], &d);
        out[1] = d;
        return SGB_OK;
    }
    return SGB_ERR_ARG;
}
// ... and of short_term() (get_short_term_reference_path with shift 1, interval 2) on a padded polyline [n_pts][2]
extern "C" int sgb_debug_short_term(const float* poly_xy, int32_t n_c, int32_t is_loop, int32_t idx, float* out6) {
    if (!poly_xy || !out6) return SGB_ERR_ARG;
    float2 st[3];
    sgb::short_term(reinterpret_cast<const float2*>(poly_xy), n_c, is_loop != 0, idx, st);
    for (int k = 0; k < 3; k++) { out6[2 * k] = st[k].x; out6[2 * k + 1] = st[k].y; }
    return SGB_OK;
}

// Rectangle-pair crossing (interX(vertices[lo], vertices[hi]), world_state_rt_sim.py:384-393) for n pose pairs
// (x, y, psi each): out[i] bit 0 = the kernels' rect_cross_rect (host build), bit 1 = "the pair gate would skip it".
// The gate itself is inline in env_step_kernel; the four lines below restate it (keep in sync) — what is checked is the
// certificate: a skipped pair never crosses.
extern "C" int sgb_debug_pair_batch(int32_t n, const float* lo3, const float* hi3, float half_length, float half_width,
                                    uint8_t* out) {
    if (n < 0 || !lo3 || !hi3 || !out) return SGB_ERR_ARG;
    const float rect_radius = std::sqrt(half_length * half_length + half_width * half_width) * 1.0001f;
    for (int i = 0; i < n; i++) {
        const float* a = lo3 + 3 * (size_t)i;
        const float* b = hi3 + 3 * (size_t)i;
        const float c1 = cosf(a[2]), s1 = sinf(a[2]), c2 = cosf(b[2]), s2 = sinf(b[2]);
        sgb::Rect rl;
        float hx[4], hy[4];
        sgb::rect_of_pose(a[0], a[1], c1, s1, half_length, half_width, rl.vx, rl.vy);
        sgb::rect_of_pose(b[0], b[1], c2, s2, half_length, half_width, hx, hy);
        rl.finish();
        const float ddx = b[0] - a[0], ddy = b[1] - a[1];
        const float cr = c1 * s2 - s1 * c2, dt = c1 * c2 + s1 * s2;
        const float reach = 2.0f * rect_radius + kFarMargin;
        const bool skip = ddx * ddx + ddy * ddy > reach * reach && fminf(fabsf(cr), fabsf(dt)) > kCollinear;
        out[i] = (uint8_t)((sgb::rect_cross_rect(rl, hx, hy) ? 1 : 0) | (skip ? 2 : 0));
    }
    return SGB_OK;
}

extern "C" void sgb_debug_scan_counters(int64_t* out8, int32_t reset) {
    for (int i = 0; i < 8; i++) { if (out8) out8[i] = sgb::g_scan_counters[i]; if (reset) sgb::g_scan_counters[i] = 0; }
}

extern "C" int sgb_debug_pack_map_blob(const sgb_map_desc* map, void* out, int64_t capacity) {
    Packed pk;
    const int rc = pack_map(map, pk);
    if (rc != SGB_OK) return rc;
    if (!out || capacity < (int64_t)pk.blob.size()) return SGB_ERR_ARG;
    std::memcpy(out, pk.blob.data(), pk.blob.size());
    return SGB_OK;
}

#endif   // SGB_TEST_HOOKS

extern "C" int sgb_create(sgb_ctx** out, int device, const sgb_map_desc* map, const sgb_config* cfg) {
    if (!out || !map || !cfg) return SGB_ERR_ARG;
    *out = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0 || device < 0 || device >= n_dev) {
        snprintf(g_err, sizeof g_err, "no CUDA device %d (found %d)", device, n_dev);
        return SGB_ERR_NO_DEVICE;
    }
    if (cfg->k_near < 0 || cfg->k_near >= SGB_MAX_AGENTS || cfg->max_steps < 2 || !(cfg->dt > 0.0f)) return SGB_ERR_ARG;
    constexpr uint32_t kObsKnown = SGB_OBS_BIRD_VIEW | SGB_OBS_CENTRES | SGB_OBS_STEERING | SGB_OBS_REF_OTHERS |
                                   SGB_OBS_NO_DIST_AGENTS | SGB_OBS_NO_DIST_CENTER | SGB_OBS_BOUNDARY_POINTS | SGB_OBS_APPLY_MASK |
                                   SGB_OBS_MASK_LANELETS;
    if (cfg->obs_flags & ~kObsKnown) {
        snprintf(g_err, sizeof g_err, "obs_flags 0x%x: unknown observation layout bits", cfg->obs_flags);
        return SGB_ERR_UNSUPPORTED;
    }
    if ((cfg->obs_flags & SGB_OBS_BIRD_VIEW) && !(cfg->norm_pos_world_x > 0.0f && cfg->norm_pos_world_y > 0.0f)) return SGB_ERR_ARG;
    if ((cfg->obs_flags & SGB_OBS_CENTRES) && !(cfg->norm_dist_agent > 0.0f)) return SGB_ERR_ARG;
    if ((cfg->obs_flags & SGB_OBS_APPLY_MASK) && !(cfg->mask_distance > 0.0f)) return SGB_ERR_ARG;
    if (!(cfg->obs_noise_level >= 0.0f) || cfg->reset_fixed_period < 0 || cfg->use_mtv_distance > 1u) return SGB_ERR_ARG;
    Packed pk;
    int rc = pack_map(map, pk);
    if (rc != SGB_OK) return rc;
    DeviceGuard guard(device);       // the caller's current device is restored on return
    if (guard.err != cudaSuccess) return cuda_fail(guard.err, "cudaSetDevice");
    sgb_ctx* c = new (std::nothrow) sgb_ctx();
    if (!c) return SGB_ERR_ARG;
    c->device = device;
    c->cfg = *cfg;
    c->n_paths = map->n_paths;
    c->max_center = pk.max_center;
    c->blob_bytes = (int32_t)pk.blob.size();
    if (const char* e = getenv("SGB_NO_PDL")) c->pdl = atoi(e) == 0;
    rc = init_device_state(c, pk);
    if (rc != SGB_OK) { sgb_destroy(c); return rc; }   // frees whatever was allocated before the failure
    *out = c;
    return SGB_OK;
}

extern "C" int sgb_destroy(sgb_ctx* c) {
    if (!c) return SGB_ERR_ARG;
    DeviceGuard guard(c->device);    // frees happen on the context's device; the caller's current device is restored
    cudaFree(c->d_blob);
    cudaFree(c->d_yaw);
    cudaFree(c->d_list);
    cudaFree(c->d_count);
    cudaFree(c->d_spawn);
    cudaFree(c->d_fresh);
    cudaFree(c->d_lanelet_xy);
    cudaFree(c->d_lanelet_off);
    cudaFree(c->d_lanelet_adj);
    if (c->pipe_ready) {
        for (int i = 0; i < 2; i++) { cudaStreamDestroy(c->pipe_stream[i]); cudaEventDestroy(c->pipe_event[i]); }
        cudaEventDestroy(c->pipe_start);
    }
    for (int i = 0; i < 2; i++) { cudaFree(c->pipe_list[i]); cudaFree(c->pipe_count[i]); }
    delete c;
    return SGB_OK;
}

extern "C" int sgb_set_lanelets(sgb_ctx* c, int32_t n, const float* xy, const int32_t* off, const uint8_t* adj) {
    if (!c || n <= 0 || n > 4096 || !xy || !off || !adj || off[0] != 0) return SGB_ERR_ARG;
    int max_len = 0;
    for (int l = 0; l < n; l++) {
        if (off[l + 1] <= off[l]) return SGB_ERR_ARG;       // every lanelet has at least one centre point
        max_len = std::max(max_len, off[l + 1] - off[l]);
    }
    GUARD(c);
    cudaFree(c->d_lanelet_xy); cudaFree(c->d_lanelet_off); cudaFree(c->d_lanelet_adj);
    c->d_lanelet_xy = nullptr; c->d_lanelet_off = nullptr; c->d_lanelet_adj = nullptr;
    c->n_lanelets = 0;
    CK(cudaMalloc(&c->d_lanelet_xy, sizeof(float) * 2 * (size_t)off[n]));
    CK(cudaMemcpy(c->d_lanelet_xy, xy, sizeof(float) * 2 * (size_t)off[n], cudaMemcpyHostToDevice));
    CK(cudaMalloc(&c->d_lanelet_off, sizeof(int32_t) * ((size_t)n + 1)));
    CK(cudaMemcpy(c->d_lanelet_off, off, sizeof(int32_t) * ((size_t)n + 1), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&c->d_lanelet_adj, (size_t)n * n));
    CK(cudaMemcpy(c->d_lanelet_adj, adj, (size_t)n * n, cudaMemcpyHostToDevice));
    c->n_lanelets = n;
    c->lanelet_max_len = max_len;
    return SGB_OK;
}

extern "C" int sgb_set_path_sets(sgb_ctx* c, int32_t n_sets, const int32_t* set_lo, const int32_t* set_hi,
                                 const float* probability) {
    if (!c || n_sets < 1 || n_sets > 4 || !set_lo || !set_hi || !probability) return SGB_ERR_ARG;
    double tot = 0.0;
    for (int i = 0; i < n_sets; i++) {
        if (set_lo[i] < 0 || set_hi[i] > c->n_paths || set_lo[i] >= set_hi[i] || !(probability[i] >= 0.0f)) return SGB_ERR_ARG;
        tot += probability[i];
    }
    if (!(tot > 0.0)) return SGB_ERR_ARG;
    int last_nz = 0;
    for (int i = 0; i < n_sets; i++)
        if (probability[i] > 0.0f) last_nz = i;
    double acc = 0.0;
    for (int i = 0; i < 4; i++) {
        const int k = std::min(i, n_sets - 1);
        if (i < n_sets) acc += probability[i] / tot;
        c->set_lo[i] = set_lo[k]; c->set_hi[i] = set_hi[k];
        // the last set with a non-zero weight takes whatever rounding leaves: a zero-weight set is never drawn
        c->set_cum[i] = (i >= last_nz) ? 2.0f : (float)acc;
    }
    c->n_sets = n_sets;
    return SGB_OK;
}

extern "C" int sgb_set_env_offset(sgb_ctx* c, int64_t env_offset) {
    if (!c || env_offset < 0) return SGB_ERR_ARG;
    c->env_offset = env_offset;
    return SGB_OK;
}

extern "C" int sgb_obs_dim(const sgb_ctx* c) { return c ? obs_dim_of(c->cfg.obs_flags, c->cfg.k_near) : SGB_ERR_ARG; }
extern "C" int sgb_max_ref_path_points(const sgb_ctx* c) { return c ? c->max_center + kExt + 2 : SGB_ERR_ARG; }
extern "C" int64_t sgb_launch_count(const sgb_ctx* c) { return c ? c->launches : 0; }
extern "C" int64_t sgb_map_bytes(const sgb_ctx* c) { return c ? c->blob_bytes : 0; }

extern "C" int sgb_step(sgb_ctx* c, int32_t B, int32_t N, const sgb_buffers* buf, void* stream) {
    if (!c || B <= 0 || N <= 0 || N > SGB_MAX_AGENTS || c->cfg.k_near > N - 1) return SGB_ERR_ARG;
    int rc = check_buffers(buf, 1);
    if (rc) return rc;
    GUARD(c);
    NvtxRange _nvtx("sgb_step");
    c->noise_epoch++;
    return launch_env(c, B, N, buf, 0, nullptr, nullptr, 1, (cudaStream_t)stream);
}

extern "C" int sgb_refresh(sgb_ctx* c, int32_t B, int32_t N, const sgb_buffers* buf, const uint8_t* env_mask,
                           int32_t write_obs, void* stream) {
    if (!c || B <= 0 || N <= 0 || N > SGB_MAX_AGENTS || c->cfg.k_near > N - 1) return SGB_ERR_ARG;
    int rc = check_buffers(buf, 0);
    if (rc) return rc;
    if (write_obs && !buf->obs) return SGB_ERR_ARG;
    GUARD(c);
    NvtxRange _nvtx("sgb_refresh");
    c->noise_epoch++;
    cudaStream_t st = (cudaStream_t)stream;
    if (!env_mask) return launch_env(c, B, N, buf, 1, nullptr, nullptr, write_obs, st);
    rc = ensure_list(c, B);
    if (rc) return rc;
    CK(cudaMemsetAsync(c->d_count + 2, 0, sizeof(int32_t), st));
    mask_to_list_kernel<<<(B + 255) / 256, 256, 0, st>>>(env_mask, B, c->d_list, c->d_count + 2);
    c->launches++;
    CK(cudaGetLastError());
    return launch_env(c, B, N, buf, 1, c->d_list, c->d_count + 2, write_obs, st);
}

extern "C" int sgb_place(sgb_ctx* c, int32_t B, int32_t N, const sgb_buffers* buf, const uint8_t* agent_mask,
                         const int32_t* path, const int32_t* point, const float* speed, void* stream) {
    if (!c || B <= 0 || N <= 0 || !path || !point || !speed) return SGB_ERR_ARG;
    int rc = check_buffers(buf, 0);
    if (rc) return rc;
    GUARD(c);
    NvtxRange _nvtx("sgb_place");
    PlaceParams p{};
    p.cfg = c->cfg; p.buf = *buf; p.blob = c->d_blob; p.yaw = c->d_yaw;
    p.agent_mask = agent_mask; p.path = path; p.point = point; p.speed = speed; p.B = B; p.N = N;
    const int n = B * N;
    place_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p);
    c->launches++;
    CK(cudaGetLastError());
    return SGB_OK;
}

// scratch of one reset: compacted list of the fully reset envs + its length, spawn-pose distances of the re-placed
// agents.  The context owns one set (ensure_list / ensure_fresh); the chunked host pipeline keeps one per stream.
struct ResetScratch {
    int32_t* list = nullptr;
    int32_t* count = nullptr;   // two counters used alternately; *phase says which one this reset appends to
    int* phase = nullptr;
    float* fresh = nullptr;
    int64_t env_first = 0;      // index of buf's env 0 within the batch the caller's env_offset refers to
};

static int reset_impl(sgb_ctx* c, int32_t B, int32_t N, const sgb_buffers* buf, int32_t path_lo, int32_t path_hi,
                      uint64_t seed, uint64_t epoch, int64_t env_offset, int32_t max_tries, int32_t write_obs,
                      int32_t* n_failed, int all, cudaStream_t st, int explicit_sel = 0,
                      const uint8_t* env_mask = nullptr, const uint8_t* agent_mask = nullptr,
                      const ResetScratch* scratch = nullptr) {
    if (!c || B <= 0 || N <= 0 || N > SGB_MAX_AGENTS || max_tries <= 0) return SGB_ERR_ARG;
    const bool use_sets = path_lo == -1;        // draw per env from the context's path sets
    if (use_sets) {
        if (c->n_sets <= 0 || !buf || !buf->scenario_id) {
            snprintf(g_err, sizeof g_err, "path_lo = -1 needs sgb_set_path_sets and buf->scenario_id");
            return SGB_ERR_ARG;
        }
    } else if (path_lo < 0 || path_hi > c->n_paths || path_lo >= path_hi) {
        return SGB_ERR_ARG;
    }
    int rc = check_buffers(buf, 0);
    if (rc) return rc;
    if (!buf->step_count || (!all && !explicit_sel && !buf->done)) return SGB_ERR_ARG;
    if (write_obs && !buf->obs) return SGB_ERR_ARG;
    GUARD(c);
    NvtxRange _nvtx("sgb_reset");
    ResetScratch own;
    if (!scratch) {
        c->noise_epoch++;
        c->env_offset = env_offset;      // the observation noise is keyed by the GLOBAL env index as well
        rc = ensure_list(c, B);
        if (rc) return rc;
        rc = ensure_fresh(c, (int64_t)B * N);
        if (rc) return rc;
        own.list = c->d_list; own.count = c->d_count; own.phase = &c->count_phase; own.fresh = c->d_fresh;
        scratch = &own;
    }
    // list length: two counters used alternately — this reset appends to one (cleared by the reset before it) and clears
    // the other for the reset after it, so no memset sits between the kernels of a step -> reset -> refresh -> step chain
    int32_t* const count_cur = scratch->count + *scratch->phase;
    int32_t* const count_nxt = scratch->count + (*scratch->phase ^ 1);
    *scratch->phase ^= 1;
    ResetParams p{};
    p.cfg = c->cfg; p.buf = *buf; p.blob = c->d_blob; p.yaw = c->d_yaw; p.list = scratch->list; p.count = count_cur;
    p.count_next = count_nxt;
    p.n_failed = n_failed; p.seed = seed; p.epoch = epoch; p.env_offset = env_offset + scratch->env_first;
    p.B = B; p.N = N; p.path_lo = path_lo; p.path_hi = path_hi; p.max_tries = max_tries; p.all = all;
    p.spawn_tab = c->d_spawn; p.fresh = scratch->fresh; p.list_full_only = 1;
    p.explicit_sel = explicit_sel; p.env_mask = env_mask; p.agent_mask = agent_mask;
    if (use_sets) {
        p.n_sets = c->n_sets;
        p.path_lo = c->set_lo[0]; p.path_hi = c->set_hi[0];
        for (int i = 0; i < 4; i++) { p.set_lo[i] = c->set_lo[i]; p.set_hi[i] = c->set_hi[i]; p.set_cum[i] = c->set_cum[i]; }
    }
    // envs per warp: a warp inspects `epw` consecutive envs and deals the touched ones to its sub-warps (32 / W at a time).
    // Many envs per warp fill the sub-warps (a quarter of the envs is touched per step under random actions), few keep
    // more warps in flight for the dependent loads: B / (32 warps per SM), at most 16 (measured at 65536 x 8, 26 % done,
    // reset only: epw 3 -> 0.0445, 8 -> 0.0377, 16 -> 0.0371, 32 -> 0.0459 ms)
    p.epw = std::max(1, std::min(16, B / (c->num_sms * 32)));
    // small batches: too few envs per warp to fill sub-warps — one warp per env (32 tries of an agent at a time, which
    // crowded maps need), a few envs per warp (8 192 envs: sub-warps 0.029 / 0.050 ms on cpm_entire x 8 / roundabout x 12,
    // warp per env 0.026 / 0.047)
    // (sub-warps of 16 lanes for 9 - 16 agents were measured too and are slower than a warp per env — 32 768 envs:
    // roundabout x 12 0.1105 vs 0.1060 ms, cpm_entire x 15 0.128 vs 0.117 — crowded envs want all 32 tries at once)
    const bool sub_warps = p.epw >= 4 && N <= 8;
    if (!sub_warps) p.epw = std::max(1, std::min(32, (B + c->num_sms * 192 - 1) / (c->num_sms * 192)));
    if (const char* e = getenv("SGB_RESET_EPW")) p.epw = std::max(1, std::min(32, atoi(e)));   // tuning knob (profiles/kbench.py)
    const int64_t n_warps = ((int64_t)B + p.epw - 1) / p.epw;
    // lanes per env: 8 (four envs in flight per warp) for large batches of up to 8 agents, a whole warp otherwise
    const dim3 rgrid((unsigned)((n_warps * 32 + 255) / 256));
    rc = sub_warps ? launch_chained(c, reset_kernel<8>, rgrid, dim3(256), 0, st, p)
                   : launch_chained(c, reset_kernel<32>, rgrid, dim3(256), 0, st, p);
    if (rc) return rc;
    c->launches++;
    CK(cudaGetLastError());
    // carry / aux / flags of every touched env are complete (spawn table); what is left is the all-fresh observation
    // (and info block) of the FULLY reset envs — respawned agents keep their step-time observation, like the
    // reference — which needs the other agents of the env but no polyline scan
    if (!write_obs && !buf->info) return SGB_OK;
    return launch_env(c, B, N, buf, 1, scratch->list, count_cur, write_obs, st, 1, nullptr, scratch->env_first,
                      scratch->fresh);
}

extern "C" int sgb_reset(sgb_ctx* c, int32_t B, int32_t N, const sgb_buffers* buf, int32_t path_lo, int32_t path_hi,
                         uint64_t seed, uint64_t epoch, int64_t env_offset, int32_t max_tries, int32_t write_obs,
                         int32_t* n_failed, void* stream) {
    return reset_impl(c, B, N, buf, path_lo, path_hi, seed, epoch, env_offset, max_tries, write_obs, n_failed, 0,
                      (cudaStream_t)stream);
}

extern "C" int sgb_reset_all(sgb_ctx* c, int32_t B, int32_t N, const sgb_buffers* buf, int32_t path_lo, int32_t path_hi,
                             uint64_t seed, uint64_t epoch, int64_t env_offset, int32_t max_tries, int32_t* n_failed,
                             void* stream) {
    return reset_impl(c, B, N, buf, path_lo, path_hi, seed, epoch, env_offset, max_tries, buf && buf->obs ? 1 : 0,
                      n_failed, 1, (cudaStream_t)stream);
}

extern "C" int sgb_reset_masked(sgb_ctx* c, int32_t B, int32_t N, const sgb_buffers* buf, const uint8_t* env_mask,
                                const uint8_t* agent_mask, int32_t path_lo, int32_t path_hi, uint64_t seed, uint64_t epoch,
                                int64_t env_offset, int32_t max_tries, int32_t write_obs, int32_t* n_failed, void* stream) {
    if (!env_mask && !agent_mask) return SGB_ERR_ARG;
    return reset_impl(c, B, N, buf, path_lo, path_hi, seed, epoch, env_offset, max_tries, write_obs, n_failed, 0,
                      (cudaStream_t)stream, 1, env_mask, agent_mask);
}

// Host-buffer step, pipelined: the batch is cut into chunks that alternate between two internal streams, so
// the H2D copy of chunk c+1, the kernels of chunk c and the D2H copy of chunk c-1 overlap (PCIe is full duplex).
// With `rs` the chunk also runs the masked device reset (+ fresh observations) before its observation goes out, so
// that the host receives what its policy acts on next.
struct HostReset {
    int32_t path_lo, path_hi, max_tries;
    uint64_t seed, epoch;
    int64_t env_offset;
    int32_t* n_failed;
};

static int ensure_pipe(sgb_ctx* c, int chunk_envs, int N) {
    if (!c->pipe_ready) {
        for (int i = 0; i < 2; i++) {
            CK(cudaStreamCreateWithFlags(&c->pipe_stream[i], cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&c->pipe_event[i], cudaEventDisableTiming));
        }
        CK(cudaEventCreateWithFlags(&c->pipe_start, cudaEventDisableTiming));
        c->pipe_ready = true;
    }
    if (c->pipe_list_cap < chunk_envs) {
        for (int i = 0; i < 2; i++) {
            cudaFree(c->pipe_list[i]);
            c->pipe_list[i] = nullptr;
            CK(cudaMalloc(&c->pipe_list[i], sizeof(int32_t) * (size_t)chunk_envs));
            if (!c->pipe_count[i]) {
                CK(cudaMalloc(&c->pipe_count[i], 2 * sizeof(int32_t)));
                CK(cudaMemset(c->pipe_count[i], 0, 2 * sizeof(int32_t)));
            }
        }
        c->pipe_list_cap = chunk_envs;
    }
    (void)N;
    return SGB_OK;
}

static int step_host_impl(sgb_ctx* c, int32_t B, int32_t N, const sgb_buffers* buf, const float* h_action, float* h_obs,
                          float* h_reward, uint8_t* h_done, const HostReset* rs, cudaStream_t st) {
    if (!c || !h_action || !h_obs || !h_reward || !h_done || B <= 0 || N <= 0 || N > SGB_MAX_AGENTS ||
        c->cfg.k_near > N - 1)
        return SGB_ERR_ARG;
    int rc = check_buffers(buf, 1);
    if (rc) return rc;
    if (rs && rs->path_lo != -1 && (rs->path_lo < 0 || rs->path_hi > c->n_paths || rs->path_lo >= rs->path_hi)) return SGB_ERR_ARG;
    if (rs && rs->max_tries <= 0) return SGB_ERR_ARG;
    GUARD(c);
    NvtxRange _nvtx("sgb_step_host");
    // noise key: the step and the reset count as the two API calls they replace (sgb_step, sgb_reset)
    const uint64_t epoch_step = ++c->noise_epoch;
    const uint64_t epoch_reset = rs ? ++c->noise_epoch : epoch_step;
    const int D = obs_dim_of(c->cfg.obs_flags, c->cfg.k_near);
    // Chunks of whole kernel waves: one wave = one env-tile per resident warp (num_sms CTAs x warps per CTA x envs per
    // warp), so a chunk of k waves keeps every SM busy for exactly k tile iterations and the map is staged once per SM
    // and chunk.  About four waves per chunk, at least two chunks once there is work for two (measured at 65536 x 8
    // on one B200: 14 / 7 / 5 / 4 / 2 / 1 chunks -> 1.76 / 1.62 / 1.59 / 1.57 / 1.72 / 1.72 ms per call); small
    // batches run as one launch.
    const int g = pick_group(N);
    const int wave = c->num_sms * (cta_threads(g) / 32) * std::max(1, 32 / (N * g));
    int waves = 4;
    if (const char* e = getenv("SGB_HOST_CHUNK_WAVES")) waves = std::max(1, std::min(64, atoi(e)));   // tuning knob
    int n_chunks = B >= 2 * wave ? std::max(2, (B + waves * wave - 1) / (waves * wave)) : 1;
    int chunk = (B + n_chunks - 1) / n_chunks;
    if (n_chunks > 1) chunk = (chunk + wave - 1) / wave * wave;
    n_chunks = (B + chunk - 1) / chunk;
    rc = ensure_pipe(c, chunk, N);
    if (rc) return rc;
    if (rs) {
        c->env_offset = rs->env_offset;
        rc = ensure_fresh(c, (int64_t)B * N);
        if (rc) return rc;
    }
    CK(cudaEventRecord(c->pipe_start, st));
    for (int i = 0; i < 2; i++) CK(cudaStreamWaitEvent(c->pipe_stream[i], c->pipe_start, 0));
    for (int k = 0; k < n_chunks; k++) {
        cudaStream_t s = c->pipe_stream[k & 1];
        const int e0 = k * chunk, e1 = std::min(B, e0 + chunk);
        const int nb = e1 - e0;
        if (nb <= 0) continue;
        const size_t a0 = (size_t)e0 * N;
        sgb_buffers sub = *buf;
        sub.pose += a0 * 4; sub.aux += a0 * 4; sub.path_id += a0; sub.carry += a0 * 4; sub.action += a0 * 2;
        sub.step_count += e0; sub.obs += a0 * D; sub.reward += a0; sub.done += e0; sub.agent_flags += a0;
        if (sub.collide_with) sub.collide_with += a0;
        if (sub.dbg) sub.dbg += a0 * 16;
        if (sub.info) sub.info += a0 * SGB_INFO_DIM;
        if (sub.task_tries) sub.task_tries += e0;
        if (sub.task_success) sub.task_success += e0;
        if (sub.scenario_id) sub.scenario_id += e0;
        CK(cudaMemcpyAsync(sub.action, h_action + a0 * 2, (size_t)nb * N * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
        c->noise_epoch = epoch_step;
        rc = launch_env(c, nb, N, &sub, 0, nullptr, nullptr, 1, s, 0, nullptr, e0);
        if (rc) return rc;
        CK(cudaMemcpyAsync(h_reward + a0, sub.reward, (size_t)nb * N * sizeof(float), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(h_done + e0, sub.done, (size_t)nb, cudaMemcpyDeviceToHost, s));
        if (rs) {
            // reset / respawn of this chunk's envs, keyed by the GLOBAL env index: same draws as an unchunked sgb_reset
            ResetScratch sc;
            sc.list = c->pipe_list[k & 1]; sc.count = c->pipe_count[k & 1]; sc.phase = &c->pipe_phase[k & 1];
            sc.fresh = c->d_fresh + a0 * 4; sc.env_first = e0;
            c->noise_epoch = epoch_reset;
            rc = reset_impl(c, nb, N, &sub, rs->path_lo, rs->path_hi, rs->seed, rs->epoch, rs->env_offset, rs->max_tries, 1,
                            rs->n_failed, 0, s, 0, nullptr, nullptr, &sc);
            if (rc) return rc;
        }
        CK(cudaMemcpyAsync(h_obs + a0 * D, sub.obs, (size_t)nb * N * D * sizeof(float), cudaMemcpyDeviceToHost, s));
    }
    for (int i = 0; i < 2; i++) {
        CK(cudaEventRecord(c->pipe_event[i], c->pipe_stream[i]));
        CK(cudaStreamWaitEvent(st, c->pipe_event[i], 0));
    }
    c->noise_epoch = epoch_reset;
    CK(cudaStreamSynchronize(st));
    return SGB_OK;
}

extern "C" int sgb_step_host(sgb_ctx* c, int32_t B, int32_t N, const sgb_buffers* buf, const float* h_action,
                             float* h_obs, float* h_reward, uint8_t* h_done, void* stream) {
    return step_host_impl(c, B, N, buf, h_action, h_obs, h_reward, h_done, nullptr, (cudaStream_t)stream);
}

extern "C" int sgb_step_reset_host(sgb_ctx* c, int32_t B, int32_t N, const sgb_buffers* buf, const float* h_action,
                                   float* h_obs, float* h_reward, uint8_t* h_done, int32_t path_lo, int32_t path_hi,
                                   uint64_t seed, uint64_t epoch, int64_t env_offset, int32_t max_tries, int32_t* n_failed,
                                   void* stream) {
    HostReset rs{path_lo, path_hi, max_tries, seed, epoch, env_offset, n_failed};
    return step_host_impl(c, B, N, buf, h_action, h_obs, h_reward, h_done, &rs, (cudaStream_t)stream);
}

extern "C" int sgb_gae(int32_t T, int32_t B, int32_t N, const float* reward, const float* value, const float* next_value,
                       const uint8_t* done, float gamma, float lmbda, float* adv, float* target, void* stream) {
    if (T <= 0 || B <= 0 || N <= 0 || !reward || !value || !next_value || !done || !adv || !target) return SGB_ERR_ARG;
    const int bn = B * N;
    // no context here: run on the device that owns the buffers, and leave the caller's current device alone
    int dev = 0;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, reward) == cudaSuccess && at.type == cudaMemoryTypeDevice) dev = at.device;
    else if (cudaGetDevice(&dev) != cudaSuccess) return cuda_fail(cudaGetLastError(), "cudaGetDevice");
    DeviceGuard guard(dev);
    if (guard.err != cudaSuccess) return cuda_fail(guard.err, "cudaSetDevice");
    gae_kernel<<<(bn + 255) / 256, 256, 0, (cudaStream_t)stream>>>(T, bn, N, reward, value, next_value, done, gamma, lmbda,
                                                                  adv, target);
    CK(cudaGetLastError());
    return SGB_OK;
}

extern "C" int sgb_gae_allgather(int32_t T, int32_t B, int32_t N, const float* reward, const float* value,
                                 const float* next_value, const uint8_t* done, float gamma, float lmbda, int32_t world,
                                 int32_t rank, float* const* adv_peers, float* const* target_peers, float* adv_multicast,
                                 float* target_multicast, void* stream) {
    if (T <= 0 || B <= 0 || N <= 0 || !reward || !value || !next_value || !done) return SGB_ERR_ARG;
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world || !adv_peers || !target_peers) return SGB_ERR_ARG;
    if ((adv_multicast == nullptr) != (target_multicast == nullptr)) return SGB_ERR_ARG;
    GaePeers peers{};
    for (int w = 0; w < world; w++) {
        if (!adv_peers[w] || !target_peers[w]) return SGB_ERR_ARG;
        peers.adv[w] = adv_peers[w];
        peers.tgt[w] = target_peers[w];
    }
    peers.adv_mc = adv_multicast;
    peers.tgt_mc = target_multicast;
    const int bn = B * N;
    int dev = 0;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, reward) == cudaSuccess && at.type == cudaMemoryTypeDevice) dev = at.device;
    else if (cudaGetDevice(&dev) != cudaSuccess) return cuda_fail(cudaGetLastError(), "cudaGetDevice");
    DeviceGuard guard(dev);
    if (guard.err != cudaSuccess) return cuda_fail(guard.err, "cudaSetDevice");
    // 16-byte path: four consecutive columns per thread belong to one env and every base is 16-byte aligned
    bool vec = (N % 4 == 0);
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    vec = vec && al16(reward) && al16(value) && al16(next_value);
    for (int w = 0; w < world; w++) vec = vec && al16(adv_peers[w]) && al16(target_peers[w]);
    cudaStream_t st = (cudaStream_t)stream;
    const int threads = vec ? bn / 4 : bn;
    const int grid = (threads + 255) / 256;
    if (adv_multicast)
        gae_allgather_kernel<true, 1><<<(bn + 255) / 256, 256, 0, st>>>(T, bn, N, reward, value, next_value, done, gamma, lmbda,
                                                                        peers, world, rank);
    else if (vec)
        gae_allgather_kernel<false, 4><<<grid, 256, 0, st>>>(T, bn, N, reward, value, next_value, done, gamma, lmbda, peers,
                                                             world, rank);
    else
        gae_allgather_kernel<false, 1><<<grid, 256, 0, st>>>(T, bn, N, reward, value, next_value, done, gamma, lmbda, peers,
                                                             world, rank);
    CK(cudaGetLastError());
    return SGB_OK;
}
