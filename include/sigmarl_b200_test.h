/*
 * sigmarl_b200_test.h — self-test hooks of the sigmarl_b200 sources.  TEST INFRASTRUCTURE: these symbols exist only in
 * libsigmarl_b200_test.so (the same translation unit compiled with -DSGB_TEST_HOOKS, `make -C sigmarl_b200/csrc test`),
 * never in the product library libsigmarl_b200.so, and nothing on the product path calls them.  They run HOST builds of
 * the very functions the kernels compile (polyline scans, MTV distance, small helpers, map packing), so that a machine
 * without a GPU can check the kernels' arithmetic and pruning certificates against the reference's vectors.
 */
#ifndef SIGMARL_B200_TEST_H
#define SIGMARL_B200_TEST_H

#include "sigmarl_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Arithmetic self-test hook (no device needed, nothing on the product path calls it): the MTV-based distance of two
 * rectangles given as [4][2] vertex arrays, evaluated by the HOST compilation of the same source function the MTV
 * kernels use (helper_scenario.py:1030-1138).  tests/test_abi_and_host.py replays the reference's known-answer
 * vectors through it bit-exactly. */
float sgb_debug_mtv_distance(const float* vertices_i, const float* vertices_j);

/* Host-only part of sgb_create (no device needed): validates and packs a map exactly as sgb_create would and reports
 * the size of the blob every CTA stages into shared memory (SGB_ERR_MAP for a degenerate polyline or one with more
 * than 256 segments).  Lets a build machine without a GPU check that every shipped map is accepted. */
int sgb_debug_pack_map(const sgb_map_desc* map, int64_t* blob_bytes);
/* ... and a copy of the packed blob itself (layout: BlobHeader / PathRec in sgb_kernels.cuh), so that the pruning
 * certificates stored in it (chunk boxes, direction cones) can be validated against the polylines on the host. */
int sgb_debug_pack_map_blob(const sgb_map_desc* map, void* out, int64_t capacity);
/* Host run of the kernels' own polyline scans (scan_center / scan_boundary with one lane per agent) for n independent
 * poses on a freshly packed blob: the pruned search (exhaustive = 0) must give exactly what the exhaustive one
 * (exhaustive = 1) gives.  out[16 * i]: d_ref, idx_ref, then per side (left at 2, right at 9) d_cg, 4 vertex
 * distances, crossing flag.  exhaustive | 2: as the product kernels run without a debug buffer — all four vertex slots
 * hold the minimum over the vertices, the only vertex quantity anything downstream consumes.  hint_idx is the carried
 * closest index (any value is valid).  The host compiler does not
 * contract a*b+c into FMAs, the device does: the certificates must (and do) hold under either rounding. */
int sgb_debug_scan_batch(const sgb_map_desc* map, int32_t n, const int32_t* path, const float* x, const float* y,
                         const float* psi, const int32_t* hint_idx, float half_length, float half_width,
                         int32_t exhaustive, float* out);
/* Work counters of sgb_debug_scan_batch on this thread since the last reset: 0 segment evaluations of the centre-line
 * scans, 1 of the boundary scans (each covers the centre + 4 vertices), 2 chunk boxes tested in the votes, 3 exact
 * crossing predicates, 4 centre scans, 5 boundary scans.  For sizing changes to the pruning logic without a GPU. */
void sgb_debug_scan_counters(int64_t* out8, int32_t reset);
/* Host builds of the kernels' small helpers: which 0 wrap_pi(in[0]); 1 dec_lin(in[0], in[1], in[2]); 2 kth_nearest over
 * in[1..n-1] with rank (int)in[0] -> out[0] index, out[1] distance.  And short_term() on a padded polyline. */
int sgb_debug_helper(int32_t which, const float* in, int32_t n, float* out);
int sgb_debug_short_term(const float* poly_xy, int32_t n_center, int32_t is_loop, int32_t idx, float* out6);
/* Rectangle-pair crossing for n pose pairs (x, y, psi): out[i] bit 0 = the kernels' rect_cross_rect (host build), bit 1 =
 * the far-and-not-collinear gate of the pair loop would skip the pair (a skipped pair must never cross). */
int sgb_debug_pair_batch(int32_t n, const float* lo_xyp, const float* hi_xyp, float half_length, float half_width,
                         uint8_t* out);
/* Host build of the kernels' current_lanelet() on a host-resident lanelet table (arithmetic self-test). */
int sgb_debug_current_lanelet(int32_t n_lanelets, const float* center_xy, const int32_t* center_off, float x, float y);

#ifdef __cplusplus
}
#endif
#endif /* SIGMARL_B200_TEST_H */
