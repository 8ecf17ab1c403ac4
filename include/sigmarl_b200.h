/*
 * sigmarl_b200.h — C-ABI of libsigmarl_b200.so: SigmaRL's vectorised road-traffic environment
 * step as hand-written CUDA for sm_100a (NVIDIA B200).
 *
 * The reference has no FFI — its hot path is Python/PyTorch behind the VMAS `BaseScenario` plug-in
 * API.  The entry points below are what a binding for that path has to expose; each one names the
 * reference interface it replaces (paths relative to /root/reference/sigmarl/).  INTEGRATION.md
 * shows the ctypes stub that plugs them into `ScenarioRoadTraffic`.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch types; every call returns 0 (SGB_OK) or a
 *     negative sgb_status; nothing throws.
 *   - The CALLER owns every state / output buffer (device memory; from PyTorch pass
 *     tensor.data_ptr()).  The library owns only the opaque context (map geometry + config on the
 *     device, a [B]-byte scratch mask).
 *   - All launches are asynchronous on the `stream` passed (a cudaStream_t cast to void*; NULL = the
 *     legacy default stream).  A context is bound to one device; it is not thread-safe.
 *   - The library's kernels are launched with programmatic stream serialization: a kernel may BEGIN
 *     (block scheduling, staging of the read-only map) while the library's previous kernel on the
 *     same stream drains, and waits (griddepcontrol.wait) before it reads or writes any caller
 *     buffer — stream order of all buffer accesses is kept.  SGB_NO_PDL=1 in the environment at
 *     sgb_create time launches them the plain way.
 *   - Layouts are row-major, env-major / agent-minor: index (b, a) -> b*N + a.
 *   - All arithmetic is fp32 with the reference's operation order (no FMA contraction on anything
 *     that feeds an argmin or a collision predicate); masks are bytes.
 */
#ifndef SIGMARL_B200_H
#define SIGMARL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGB_VERSION 129
#define SGB_MAX_AGENTS 32       /* agents per env (collide_with is a 32-bit mask) */
#define SGB_N_SHORT_TERM 3      /* n_points_short_term   (road_traffic.py:273-275) */

typedef enum {
    SGB_OK = 0,
    SGB_ERR_ARG = -1,        /* NULL / out-of-range argument */
    SGB_ERR_CUDA = -2,       /* a CUDA runtime call failed; see sgb_last_error() */
    SGB_ERR_NO_DEVICE = -3,  /* no CUDA device / wrong architecture: there is NO CPU fallback */
    SGB_ERR_MAP = -4,        /* map does not fit in shared memory / malformed polyline */
    SGB_ERR_UNSUPPORTED = -5 /* config flag outside the hot path (e.g. unknown observation layout bits) */
} sgb_status;

/* Flat map: the polylines of `map_manager.py:13-40` / `parse_xml.py:785-797` (one entry per reference
 * path).  Replaces the per-(env, agent) copies of `world_state_rt.py:155-175, 313-392`: the library
 * keeps ONE read-only copy and agents carry an int path index.  Host pointers; copied at create. */
typedef struct {
    int32_t n_paths;
    const float*   center_xy;   /* [center_off[n_paths]][2]  centre lines ("center_line")            */
    const int32_t* center_off;  /* [n_paths+1] offsets in points                                    */
    const float*   left_xy;     /* "left_boundary_shared"                                           */
    const int32_t* left_off;
    const float*   right_xy;    /* "right_boundary_shared"                                          */
    const int32_t* right_off;
    const float*   center_yaw;  /* [center_off[n_paths] - n_paths] "center_line_yaw" (n-1 per path)  */
    const uint8_t* is_loop;     /* [n_paths]                                                        */
} sgb_map_desc;

/* reward composition flags == the substring tests of road_traffic.py:1058-1112 */
#define SGB_REW_EXACT_SPARSE 1u /* rew_method == "sparse"        */
#define SGB_REW_TTC 2u          /* "ttc" in rew_method           */
#define SGB_REW_DISTANCE 4u     /* "distance" in rew_method      */
#define SGB_REW_SPARSE 8u       /* "sparse" in rew_method        */

/* Everything `road_traffic.py:_init_params` (:112-768) bakes into the scenario, as fp32 the way the
 * reference holds it (torch.tensor(..., dtype=float32)).  Replaces Thresholds / Penalties / Rewards /
 * Normalizers (helper_scenario.py:13-100) and the AGENTS table (constants.py:628-647). */
typedef struct {
    float dt;                 /* world.dt                                 helper_training.py:821 */
    float max_speed, max_steering, max_acc, max_steering_rate;         /* constants.py:638-644 */
    float l_wb, lr_over_lwb;  /* wheelbase, l_r / l_wb                    dynamics.py:102-110    */
    float half_length, half_width;                                     /* constants.py:630-631 */
    float diag;               /* sqrt(x_semidim^2 + y_semidim^2)          helper_scenario.py:1140 */
    float w_ref[SGB_N_SHORT_TERM]; /* weighting_ref_directions            road_traffic.py:536-543 */
    float speed_dt;           /* float32(max_speed * dt)                  road_traffic.py:986    */
    float reward_progress;
    float near_boundary_low, near_boundary_high;
    float near_agents_low, near_agents_high;
    float ttc_low, ttc_high;
    float penalty_near_boundary, penalty_near_agents;
    float penalty_collide_agents, penalty_collide_lane;
    float norm_pos, norm_v, norm_rot, norm_dist;                        /* road_traffic.py:587-608 */
    float dsafe_sq;           /* float32(d_safe * d_safe)                 road_traffic.py:1291   */
    float reset_min_dist_sq;  /* reset_agent_min_distance ** 2            world_state_rt_sim.py:305 */
    uint32_t rew_flags;       /* SGB_REW_*                                                        */
    int32_t k_near;           /* min(n_nearing_agents_observed, N-1)      road_traffic.py:441-443 */
    int32_t max_steps;        /* timer.step == max_steps-1 -> done        road_traffic.py:1413   */
    int32_t respawn_on_exit;  /* scenario_type != "cpm_entire"            road_traffic.py:1449   */
    int32_t exhaustive;       /* debug: 1 = scan every segment (no pruning); results must not change */
    float reward_reach_goal;  /* rewards.reach_goal (info["rew_reach_goal"]; added to the reward only in testing
                                 mode)                                    road_traffic.py:217-219, 996-1003 */
    int32_t testing_mode;     /* parameters.is_testing_mode: sparse-only reward (:1050-1055), env done only at the
                                 time limit and colliding / leaving agents respawned one by one (:1429-1447),
                                 spawn range growing with the try count (world_state_rt_sim.py:254-261) */
    uint32_t obs_flags;       /* SGB_OBS_*: observation layout, 0 = the reference's default flags
                                 observation_provider_rt.py:594-925 */
    float norm_pos_world_x, norm_pos_world_y; /* normalizers.pos_world (bird view)  road_traffic.py:593-595 */
    float norm_dist_agent;    /* normalizers.distance_agent (lengths / widths)      road_traffic.py:605-607 */
    float obs_noise_level;    /* is_obs_noise ? obs_noise_level : 0: obs += level * U[0,1) per element, fresh draws per
                                 call (observation_provider_rt.py:611-617); counter-based device generator keyed by
                                 (obs_noise_seed, API-call counter, global env index, agent, column):
                                 distribution-equivalent, independent of sharding */
    uint32_t obs_noise_seed;
    int32_t reset_fixed_period; /* reset_agent_fixed_duration as a step period, 0 = off: an env is also done at every
                                 step with timer.step % period == 0.  The reference tests float32(timer.step * dt) %
                                 duration == 0 (road_traffic.py:1388-1393); the host layer derives the period from
                                 (dt, duration) and refuses pairs for which that float test is not periodic
                                 (EnvConfig.fixed_period) */
    uint32_t use_mtv_distance; /* is_use_mtv_distance (0 / 1): distances.agents = MTV-based (SAT) distance between the
                                 agents' rectangles (helper_scenario.py:1030-1138) instead of the centre distance; taken,
                                 like the reference does, from the rectangles of the PRE-step poses
                                 (world_state_rt_sim.py:432-448), fresh ones after a reset; agents collide iff it is
                                 exactly 0 (:394-396).  near_agents_low / high then carry the MTV thresholds
                                 (road_traffic.py:264-270, 632-648) */
    float mask_distance;      /* thresholds.distance_mask_agents (5 agent lengths, road_traffic.py:663); read with
                                 SGB_OBS_APPLY_MASK */
    uint32_t reserved0, reserved1, reserved2; /* sizeof(sgb_config) == 200: grows in steps of 16 so that the kernel
                                 parameters behind it keep their alignment (and the tuned kernels their exact code) */
} sgb_config;

/* Observation layout flags == the reference's Parameters of the same meaning (helper_common.py:60-118;
 * observation_provider_rt.py:594-925).  Own part: [pos(2), rot] (bird view) | vel (1 ego, 2 bird) | [steering] |
 * short-term path (6) | [distance to centre line] | min distance to left, right boundary (or 5 + 5 boundary points).  Per observed
 * neighbour: 4 vertices (8) or pos(2), rot, length, width | vel (2) | [steering] | [distance] | [its
 * short-term path (6)].  Not offered (sgb_create returns SGB_ERR_UNSUPPORTED for unknown bits;
 * the Python host layer refuses the parameters in EnvConfig.validate): is_partial_observation = False (the
 * reference itself crashes there, observation_provider_rt.py:808). */
#define SGB_OBS_BIRD_VIEW 1u       /* is_ego_view = False: global coordinates / pos_world                 */
#define SGB_OBS_CENTRES 2u         /* is_observe_vertices = False: pos, rot, length, width of a neighbour  */
#define SGB_OBS_STEERING 4u        /* is_obs_steering: own and neighbours' steering angle / (2 pi)         */
#define SGB_OBS_REF_OTHERS 8u      /* is_observe_ref_path_other_agents                                     */
#define SGB_OBS_NO_DIST_AGENTS 16u /* is_observe_distance_to_agents = False                                */
#define SGB_OBS_NO_DIST_CENTER 32u /* is_observe_distance_to_center_line = False                           */
#define SGB_OBS_BOUNDARY_POINTS 64u /* is_observe_distance_to_boundaries = False: 5 points of each boundary
                                      around the closest one instead of the two distances.  carry.w then keeps, in
                                      bit 30, whether the pose was written by a reset (the reference samples the
                                      points with shift +1 there and -2 in a step, world_state_rt.py:531-576, :686-725) */
#define SGB_OBS_APPLY_MASK 128u    /* is_apply_mask: an observed neighbour whose distance is >= mask_distance shows
                                      constants instead of its state — positions, vertices, reference path and distance 1,
                                      heading, steering and velocity 0; length / width stay (observation_provider_rt.py:
                                      638-749) */
#define SGB_OBS_MASK_LANELETS 256u /* with SGB_OBS_APPLY_MASK: the reference's second criterion — a neighbour is also masked
                                      unless its current lanelet is the ego's or adjacent to it (:646-664; map_manager.py:
                                      39-119).  Live in the reference only in bird view (the lanelet assignment is computed
                                      there, :585-588) on OSM maps (parse_osm.py:257-262); needs sgb_set_lanelets */

/* Device buffers of one batch of B envs x N agents.  in = read, out = written, io = both. */
typedef struct {
    float*   pose;        /* io [B,N,4]  x, y, psi, v          Vehicle state helper_common.py:290-430 */
    float*   aux;         /* io [B,N,4]  delta, vx, vy, beta   (steering, vel, sideslip_angle)      */
    int32_t* path_id;     /* in [B,N]    index into the map's paths                                 */
    float*   carry;       /* io [B,N,4]  d_ref, min dL, min dR, (int) idx_ref of the pre-step pose — */
                          /*             the one-step-stale values the VMAS call order exposes      */
                          /*             (SURVEY.md A.6); (re)built by sgb_refresh                  */
    float*   action;      /* io [B,N,2]  target speed, target steering; clamped in place            */
                          /*             helper_training.py:807-818                                 */
    int32_t* step_count;  /* io [B]      timer.step                     road_traffic.py:449-460    */
    float*   obs;         /* out [B,N,D] D = sgb_obs_dim()              road_traffic.py:1334       */
    float*   reward;      /* out [B,N]                                  road_traffic.py:925        */
    uint8_t* done;        /* out [B]                                    road_traffic.py:1368       */
    uint8_t* agent_flags; /* out [B,N]   SGB_FLAG_* collision bits      world_state_rt_sim.py:36-55 */
    uint32_t* collide_with; /* out [B,N] bit j = collisions.with_agents[b,a,j]; may be NULL        */
    float*   info;        /* out [B,N,SGB_INFO_DIM] what info(agent) adds to the state (road_traffic.py:1547-1633); */
                          /*   may be NULL.  0..5 "ref" (fresh short-term path) 6 distance_ref 7 distance_left_b     */
                          /*   8 distance_right_b 9 rew_near_other_agents 10 rew_collide_other_agents              */
                          /*   11 rew_collide_lane 12 rew_reach_goal 13 rew_total 14 distances.boundaries 15 spare */
    int32_t* task_tries;  /* io [B] num_task_tries   (road_traffic.py:1029-1035); may be NULL                       */
    int32_t* task_success;/* io [B] task_success_times (:998-1002); may be NULL                                     */
    float*   dbg;         /* out [B,N,16] internals for parity tests; may be NULL:                  */
                          /*   0 d_ref 1 (int)idx_ref 2..6 dL[0..4] 7..11 dR[0..4] 12 d_bound       */
                          /*   13 nearest-agent index 14 second-nearest index 15 reserved           */
    int32_t* scenario_id; /* io [B] path set of the env (sgb_set_path_sets): written by a full reset, read by   */
                          /*   respawns — ref_paths_agent_related.scenario_id - 1 (world_state_rt_sim.py:313-358); */
                          /*   may be NULL unless a reset is asked to draw from the path sets (path_lo = -1)    */
    uint32_t* nan_flags;  /* io [1] sticky health word, may be NULL: bit 0 = a non-finite pose / observation /  */
                          /*   reward was produced by a step (the reference asserts on these, road_traffic.py:   */
                          /*   1245-1246); the caller clears it                                                  */
} sgb_buffers;

#define SGB_INFO_DIM 16

#define SGB_FLAG_COLLIDE_AGENT 1u /* any collisions.with_agents[b,a,:]  */
#define SGB_FLAG_COLLIDE_LANE 2u  /* collisions.with_lanelets[b,a]      */
#define SGB_FLAG_ENTRY 4u         /* collisions.with_entry_segments     */
#define SGB_FLAG_EXIT 8u          /* collisions.with_exit_segments (= info["is_reach_goal"]) */

typedef struct sgb_ctx sgb_ctx;

/* Build a context on `device`: packs the map (centre lines + 6 extension points
 * world_state_rt.py:279-311, boundaries, per-chunk bounding boxes) into one blob that every CTA
 * bulk-copies (TMA) into shared memory.  Replaces ScenarioRoadTraffic.make_world's map/constant
 * set-up (road_traffic.py:104-110, 112-768). */
int sgb_create(sgb_ctx** out, int device, const sgb_map_desc* map, const sgb_config* cfg);
int sgb_destroy(sgb_ctx* ctx);

/* Lanelet table for SGB_OBS_MASK_LANELETS: the centre lines of ALL lanelets of the map (MapManager.parser.lanelets_all,
 * concatenated, center_off[n_lanelets + 1]) and the adjacency matrix [n_lanelets][n_lanelets] (adjacency[i][j] != 0 iff
 * j is in parser.neighboring_lanelets_idx[i]).  Host pointers, copied to the device.  Replaces
 * MapManager.determine_current_lanelet / determine_masked_agents_by_lanelets (map_manager.py:39-119), which the
 * flag-driven observation writer then evaluates per observed neighbour. */
int sgb_set_lanelets(sgb_ctx* ctx, int32_t n_lanelets, const float* center_xy, const int32_t* center_off,
                     const uint8_t* adjacency);

/* Path sets of a map (cpm_mixed: intersection / merge-in / merge-out).  The reference draws ONE set per env at every
 * full reset — scenario_id ~ multinomial(cpm_scenario_probabilities) — places all agents of the env on paths of that
 * set, and keeps the set for single-agent respawns (world_state_rt_sim.py:313-358; road_traffic.py:333-334).
 * set_lo / set_hi: global path index ranges [lo, hi) of the n_sets <= 4 sets; probability: their weights (normalised
 * here).  A reset called with path_lo = -1 then draws per env from these sets (counter-based, keyed by the global env
 * index like every reset draw) and needs buf->scenario_id; any other path_lo keeps the explicit range. */
int sgb_set_path_sets(sgb_ctx* ctx, int32_t n_sets, const int32_t* set_lo, const int32_t* set_hi, const float* probability);

/* Global index of this context's env 0 when a batch is sharded over several contexts / GPUs (default 0).  The reset
 * entry points take it as an argument (and remember it); the observation noise is keyed by it as well, so that noisy
 * observations do not depend on how envs are sharded.  Has no counterpart in the reference (single process). */
int sgb_set_env_offset(sgb_ctx* ctx, int64_t env_offset);

/* Observation width D for the configured layout: 10 + 11*k_near with the default flags
 * (observation_provider_rt.py:594-925). */
int sgb_obs_dim(const sgb_ctx* ctx);
/* max_ref_path_points the reference would use for this map (road_traffic.py:505-530). */
int sgb_max_ref_path_points(const sgb_ctx* ctx);

/* ONE fused kernel per rollout step.  Replaces vmas Environment.step's scenario work:
 * WorldCustom.step (helper_training.py:797-861) + KinematicBicycleModel.step (dynamics.py:120-192)
 * + for every agent reward() (road_traffic.py:925-1253), observation() (:1334-1366) + done()
 * (:1368-1487, without the respawn side effect — see sgb_reset).  Pure function of
 * (pose, aux, path_id, carry, step_count, action). */
int sgb_step(sgb_ctx* ctx, int32_t B, int32_t N, const sgb_buffers* buf, void* stream);

/* Recompute everything derived from the current pose for the envs with env_mask[b] != 0
 * (NULL = all): aux.vx/vy/beta from (psi, v, delta), carry, collision flags cleared; if write_obs,
 * also the all-fresh observation the reference returns right after a reset.  Replaces
 * reset_world_at's tail (road_traffic.py:897-923; world_state_rt.py:422-529). */
int sgb_refresh(sgb_ctx* ctx, int32_t B, int32_t N, const sgb_buffers* buf, const uint8_t* env_mask,
                int32_t write_obs, void* stream);

/* Put agents at given path points: pose = (center[path][point], center_yaw[path][point], speed),
 * delta = 0 — for the (b,a) with agent_mask != 0 (NULL = all).  Replaces _reset_init_state
 * (world_state_rt_sim.py:143-213) when the caller supplies the draws (parity mode).  Follow with
 * sgb_refresh. */
int sgb_place(sgb_ctx* ctx, int32_t B, int32_t N, const sgb_buffers* buf, const uint8_t* agent_mask,
              const int32_t* path, const int32_t* point, const float* speed, void* stream);

/* Device-side masked reset + respawn after a step.  For envs with done[b]: re-place all N agents by
 * bounded rejection sampling (uniform path in [path_lo, path_hi), uniform point in [3, n/2), every
 * pair >= reset_min_dist apart) and zero step_count; for not-done envs (when cfg.respawn_on_exit, or
 * cfg.testing_mode) respawn the agents whose flags ask for it.  Carry, aux and cleared collision flags
 * of every touched env come from the SPAWN TABLE — what sgb_refresh derives from a pose on a centre
 * point, computed once at sgb_create by the refresh kernel itself (bit-identical) — so a reset runs no
 * polyline scan.  If write_obs (or buf->info), the FULLY reset envs additionally get the all-fresh
 * observation (info block) the reference returns after reset_at; respawned agents of not-done envs
 * keep their step-time observation, like the reference (road_traffic.py:1462-1472).
 * Counter-based RNG keyed by (seed, step counter `epoch`, global
 * env index env_offset + b, agent, try) so results do not depend on how envs are sharded over GPUs.
 * Replaces reset_world_at (road_traffic.py:816-923) and _generate_feasible_initial_positions
 * (world_state_rt_sim.py:215-311); distribution-equivalent, not stream-equivalent, to the reference's
 * global torch CPU generator.  n_failed (device int32, may be NULL) counts agents for which no
 * feasible point was found in `max_tries` (the reference would spin forever). */
int sgb_reset(sgb_ctx* ctx, int32_t B, int32_t N, const sgb_buffers* buf, int32_t path_lo, int32_t path_hi,
              uint64_t seed, uint64_t epoch, int64_t env_offset, int32_t max_tries, int32_t write_obs,
              int32_t* n_failed, void* stream);

/* Same machinery with an EXPLICIT selection instead of done / collision flags: envs with env_mask[b] != 0 are
 * reset fully, agents with agent_mask[b,a] != 0 (in envs not reset fully) are respawned, whatever the map or
 * mode.  Either mask may be NULL.  Replaces reset_world_at(env_index, agent_index) called from outside the
 * step (road_traffic.py:816-923; evaluation code, vmas Environment.reset_at). */
int sgb_reset_masked(sgb_ctx* ctx, int32_t B, int32_t N, const sgb_buffers* buf, const uint8_t* env_mask,
                     const uint8_t* agent_mask, int32_t path_lo, int32_t path_hi, uint64_t seed, uint64_t epoch,
                     int64_t env_offset, int32_t max_tries, int32_t write_obs, int32_t* n_failed, void* stream);

/* Same as sgb_reset but for ALL envs regardless of `done` (Environment.reset()). */
int sgb_reset_all(sgb_ctx* ctx, int32_t B, int32_t N, const sgb_buffers* buf, int32_t path_lo, int32_t path_hi,
                  uint64_t seed, uint64_t epoch, int64_t env_offset, int32_t max_tries, int32_t* n_failed,
                  void* stream);

/* Host-buffer convenience for callers whose policy lives on the host (the reference's default device
 * is the CPU, config.json:5): copies `h_action` [B,N,2] to buf->action, runs sgb_step, copies
 * obs / reward / done back into the host pointers (pinned memory recommended) and synchronises the
 * stream before returning. */
int sgb_step_host(sgb_ctx* ctx, int32_t B, int32_t N, const sgb_buffers* buf, const float* h_action,
                  float* h_obs, float* h_reward, uint8_t* h_done, void* stream);

/* sgb_step_host followed, chunk by chunk inside the same copy / compute pipeline, by the masked device reset of
 * sgb_reset (same arguments, same draws as an unchunked call): `h_obs` then holds what a host-resident policy acts on
 * NEXT — the all-fresh post-reset observation for the envs that finished in this step, the step-time observation for
 * the others — while `h_reward` / `h_done` are the step's.  One call = one iteration of the collector loop
 * (helper_training.py:745-770: env.step, then reset of the done envs). */
int sgb_step_reset_host(sgb_ctx* ctx, int32_t B, int32_t N, const sgb_buffers* buf, const float* h_action,
                        float* h_obs, float* h_reward, uint8_t* h_done, int32_t path_lo, int32_t path_hi,
                        uint64_t seed, uint64_t epoch, int64_t env_offset, int32_t max_tries, int32_t* n_failed,
                        void* stream);

/* "Next" row: generalised advantage estimation over a finished rollout, written straight
 * into caller-provided buffers (e.g. this rank's slot of the all-gather buffer).  All pointers are device
 * memory; reward / value / next_value / adv / target are [T,B,N] fp32, done is [T,B] bytes (terminated ==
 * done).  Replaces TorchRL GAE as configured in optimization_module.py:62-67 (gamma 0.99, lambda 0.9, no
 * advantage normalisation) and the done expansion of mappo_cavs.py:342-355. */
int sgb_gae(int32_t T, int32_t B, int32_t N, const float* reward, const float* value, const float* next_value,
            const uint8_t* done, float gamma, float lmbda, float* adv, float* target, void* stream);

/* The same scan fused with the design's one collective (mappo_cavs.py:357-378 consumes the result): the
 * all-gather of advantage / value target at PPO-update time.  Every rank holds [world,T,B,N] gather buffers in memory that
 * is mapped into its peers (CUDA IPC / cuMem symmetric memory; one process per GPU of ONE node); `adv_peers[w]` /
 * `target_peers[w]` (host arrays of `world` device pointers) are the bases of rank w's buffers AS MAPPED ON THIS DEVICE
 * (w == rank: the local ones).  Each value is stored, as it is computed, into slot `rank` of every rank's buffer over
 * NVLink; with `adv_multicast` / `target_multicast` (NVLS multicast mappings of the same buffers, both or neither) ONE
 * multimem store per value is replicated by the switch instead.  No staging copy and no separate collective.
 * Ordering is the caller's: a cross-rank barrier on `stream` BEFORE the call (every rank is done reading the previous
 * contents) and AFTER it (all ranks' stores have landed) — e.g. torch symmetric memory's `handle.barrier()`.
 * world == 1 degenerates to sgb_gae into slot 0. */
int sgb_gae_allgather(int32_t T, int32_t B, int32_t N, const float* reward, const float* value, const float* next_value,
                      const uint8_t* done, float gamma, float lmbda, int32_t world, int32_t rank,
                      float* const* adv_peers, float* const* target_peers, float* adv_multicast, float* target_multicast,
                      void* stream);

/* Number of kernels this context has launched since creation (for launch accounting in benchmarks). */
int64_t sgb_launch_count(const sgb_ctx* ctx);
/* Bytes of the packed map blob each CTA stages into shared memory. */
int64_t sgb_map_bytes(const sgb_ctx* ctx);

/* Host-side self-test hooks (sgb_debug_*) are NOT part of this library: they are declared in sigmarl_b200_test.h and
 * exist only in libsigmarl_b200_test.so, which the test suite builds from the same sources with -DSGB_TEST_HOOKS. */

const char* sgb_status_string(int status);
const char* sgb_last_error(void); /* text of the last CUDA error seen by this thread */
int sgb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SIGMARL_B200_H */
