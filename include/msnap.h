/* include/msnap.h -- C ABI of the B200-native batched minimum-snap trajectory solver.
 *
 * This is the drop-in boundary for the reference's TrajectoryGeneratorTool
 *   /root/reference/math_util/minimum_snap.hpp:36-63   (class)
 *   /root/reference/math_util/minimum_snap.cpp:22-206   GenerateTrajectoryMatrix
 *   /root/reference/math_util/minimum_snap.cpp:227-649  SolveQPClosedForm
 * as called by UavPathPlanner::Minisnap_3D / Minisnap_EN (/root/reference/uavPathPlanning.cpp:4401-4474).
 * Plain pointers and sizes only; no C++/torch types.  INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *   - A batch holds B independent trajectories.  Trajectory b has ns_b >= 1 segments and ns_b + 1 waypoints.
 *     Segments are indexed CSR-style by seg_offset[B+1] (seg_offset[0] = 0); the waypoints of trajectory b start at
 *     point index seg_offset[b] + b.  If every trajectory has the same segment count pass seg_offset = NULL and
 *     ns_uniform = that count.
 *   - waypoints : [sum(ns_b) + B][3] row-major (east, north, up) == std::vector<ENUPoint> (uavPathPlanning.hpp:152-156)
 *   - times     : [sum(ns_b)] segment durations in seconds
 *   - coeff     : [sum(ns_b)][3][2*order]  = the reference's PolyCoeff rows (ms.cpp:220-225): per segment,
 *                 x | y | z blocks, highest power first, local time t in [0, T_k] (no time scaling).  DEVICE coefficient
 *                 buffers must be 16-byte aligned (128-bit stores); cudaMalloc / torch allocations are.
 *   - samples   : [rows][3] row-major; trajectory b owns rows [sample_offset[b], sample_offset[b+1])
 *   - Functions suffixed _dev take DEVICE pointers for every array argument and only enqueue work on the handle's
 *     stream (no host synchronisation).  Functions suffixed _host take HOST pointers (pinned memory makes the copies
 *     asynchronous and lets them overlap the kernels), use device mirrors owned by the handle, and return after the
 *     results are in the caller's buffers.
 *   - Every function returns an msnap_status.  There is no CPU fallback: without a usable CUDA device
 *     msnap_create fails with MSNAP_ERR_NO_DEVICE and nothing else can be called.
 *   - A handle is bound to one device and one stream and is not thread-safe: one handle per host thread / GPU.
 */
#ifndef MSNAP_H
#define MSNAP_H

#ifdef __cplusplus
extern "C" {
#endif

#define MSNAP_VERSION 100 /* 0.1.0 */

typedef enum msnap_status {
    MSNAP_OK = 0,
    MSNAP_ERR_INVALID_ARG = 1, /* NULL where data is required, order outside 2..5, B < 0, ns < 1, ...            */
    MSNAP_ERR_CUDA = 2,        /* a CUDA call failed; text via msnap_last_error                                  */
    MSNAP_ERR_NO_DEVICE = 3,   /* no CUDA device / device index out of range / not an sm_100 part                */
    MSNAP_ERR_CAPACITY = 4,    /* sample buffer too small; sample_offset still holds the exact required layout   */
    MSNAP_ERR_ALLOC = 5,       /* host or device allocation failed                                               */
    MSNAP_ERR_IO = 6           /* msnap_config_load_yaml: file unreadable                                        */
} msnap_status;

/* Field-for-field mirror of struct MinimumSnapConfig (minimum_snap.hpp:9-33). */
typedef struct msnap_config {
    int order;              /* derivative order minimised: 2 = acceleration (cubic), 3 = jerk, 4 = snap, 5 = crackle */
    double path_weight;     /* straight-line deviation penalty (two-pass, ms.cpp:347-469)                        */
    double vel_zero_weight; /* waypoint velocity penalty (ms.cpp:474-509); start value of the reweighting loop   */
    double V_avg;           /* cruise speed for time allocation, m/s (ms.cpp:63-72)                              */
    double min_time_s;      /* lower bound of a segment's duration, s                                            */
    double sample_distance; /* sampler spacing threshold, m (ms.cpp:145)                                         */
    double start_vel[3], end_vel[3], start_acc[3], end_acc[3];
} msnap_config;

/* per-trajectory flag bits written to flags_out */
#define MSNAP_FLAG_NONFINITE 1u /* a non-finite value or a non-positive pivot appeared in the solve */
#define MSNAP_FLAG_TRUNCATED 2u /* samples of this trajectory did not fit the caller's buffer       */
/* Longest segment duration the sampler walks (seconds; 1e7 candidates at 10 Hz).  A segment whose allocated time is not a
 * positive finite number <= this (an inf / NaN / absurd waypoint) gets no sampled candidates and its trajectory
 * MSNAP_FLAG_NONFINITE: in the reference such an input keeps ms.cpp:140's `t += dt` loop spinning for that one call; in a
 * batch it must not take the launch of every other trajectory down with it. */
#define MSNAP_MAX_SEGMENT_TIME 1.0e6

typedef struct msnap_context *msnap_handle;

/* ---- library / handle ------------------------------------------------------------------------------------ */
int msnap_version(void);
const char *msnap_status_string(int status);
/* Text of the last CUDA failure seen by this handle ("" if none). */
const char *msnap_last_error(msnap_handle h);

/* Bind a solver instance to CUDA device `device` (creates a stream, uploads the constant tables). */
int msnap_create(int device, msnap_handle *out);
int msnap_destroy(msnap_handle h);
/* Use the caller's cudaStream_t (passed as void*) for all subsequent _dev work; NULL restores the handle's own.  The new
 * stream is made to wait (event, no host synchronisation) for everything the handle has enqueued on the old one: a handle
 * owns ONE workspace that every call reuses, so calls of one handle are always ordered.  Use one handle per concurrent
 * stream.  Host synchronisation points of the _dev entry points: the first call of a larger problem size grows the
 * workspace (cudaStreamSynchronize + cudaFree + cudaMalloc); ragged batches (seg_offset given) read seg_offset[B] back
 * (8 bytes, one synchronisation) -- neither is CUDA-graph capturable; uniform batches at a warmed-up size only enqueue. */
int msnap_set_stream(msnap_handle h, void *cuda_stream);
/* Block until everything enqueued on the handle's stream has finished. */
int msnap_synchronize(msnap_handle h);
/* Execution policy of the reweighting loop (ms.cpp:76-90): 0 = automatic, 1 = sequential per trajectory,
 * 2 = speculative (all 11 velocity weights solved concurrently, the first admissible one selected). Results are
 * identical; this is a throughput/latency knob only. */
int msnap_set_reweight_policy(msnap_handle h, int policy);
/* Host-pointer entry points cut big batches into chunks whose kernels overlap the previous chunk's device-to-host
 * copies (two internal streams).  n_chunks = 0: automatic (one chunk per 8 192 trajectories, at most 8); 1: no
 * pipelining.  Results do not depend on the chunking. */
int msnap_set_host_chunks(msnap_handle h, int n_chunks);
/* Host-pointer entry points: when the caller's coefficient / sample buffers are PINNED host memory (cudaHostAlloc,
 * cudaHostRegister, torch pin_memory) the kernels store the results straight into them over PCIe, overlapping transfer
 * and computation (samples: only when no statistics are requested and the batch is a single chunk).  enable = 0 uses
 * device buffers + cudaMemcpyAsync.  Default: DISABLED -- on B200 / PCIe 5 the SMs' stores to host memory reach about
 * 24 GB/s while the copy engine moves the same bytes at about 53 GB/s (scripts/e2e_chunks.py).  Results are identical. */
int msnap_set_zero_copy(msnap_handle h, int enable);
/* Number of kernels this handle has launched since creation (monotonic; used by bench.py's gpu_launches). */
long long msnap_launch_count(msnap_handle h);

/* ---- configuration (minimum_snap.hpp:9-33, minimum_snap_config.yaml, uavPathPlanning.cpp:851-879) --------- */
/* Struct defaults: order 3, weights 0, V_avg 5, min_time 0.1, sample_distance 1, zero boundary vel/acc. */
void msnap_config_default(msnap_config *cfg);
/* Read the ten min-snap keys from a YAML file, flat or wrapped under `minimum_snap:`; keys that are absent or
 * malformed keep their current value (yamlAssignIfPresent semantics, uavPathPlanning.cpp:35-59, 859-872). */
int msnap_config_load_yaml(const char *path, msnap_config *cfg);

/* ---- SolveQPClosedForm, batched (minimum_snap.hpp:45-53; ms.cpp:227-649) ----------------------------------
 * One closed-form solve per trajectory with caller-supplied segment times: no time allocation, no reweighting.
 *   vel, acc : [B][2][3] (row 0 = start, row 1 = end) or NULL for zeros
 *   coeff_out: [sum ns][3][2*order];  max_dev_out: [B] or NULL (ms.cpp:594-624)
 *   best_s_out: [sum ns] (int) or NULL -- index s in 0..16 of the worst-deviation sample t* = T s/16 chosen per
 *               segment (ms.cpp:408-439); all zero when path_weight <= 0 */
int msnap_solve_qp_batch_dev(msnap_handle h, int order, double path_weight, double vel_zero_weight, long long B,
                             int ns_uniform, const long long *seg_offset, const double *waypoints, const double *vel,
                             const double *acc, const double *times, double *coeff_out, double *max_dev_out,
                             int *best_s_out, unsigned *flags_out);
int msnap_solve_qp_batch_host(msnap_handle h, int order, double path_weight, double vel_zero_weight, long long B,
                              int ns_uniform, const long long *seg_offset, const double *waypoints, const double *vel,
                              const double *acc, const double *times, double *coeff_out, double *max_dev_out,
                              int *best_s_out, unsigned *flags_out);

/* ---- GenerateTrajectoryMatrix, batched (minimum_snap.hpp:60-61; ms.cpp:22-206) -----------------------------
 * Time allocation (ms.cpp:63-72), the reweighting loop around the closed-form solve (ms.cpp:76-90), the
 * distance-thresholded sampler (ms.cpp:97-161) and the climb/turn statistics (ms.cpp:163-195), for B trajectories.
 *   sample_distance_override / v_avg_override: used iff > 0 (ms.cpp:42-48)
 * MSNAP_ERR_INVALID_ARG for parameters the sampler's candidate loop cannot terminate with: min_time_s <= 0 (coincident
 * waypoints would give T = 0, dt = 0) or > MSNAP_MAX_SEGMENT_TIME, non-finite V_avg / sample_distance / weights.
 * Outputs (any optional pointer may be NULL):
 *   times_out     [sum ns]            optional   allocated segment times
 *   coeff_out     [sum ns][3][2o]     optional   final polynomial coefficients
 *   max_dev_out   [B]                 optional   final max deviation ratio
 *   iters_out     [B] (int)           optional   reweighting iterations performed (0..10)
 *   vw_final_out  [B]                 optional   final vel_zero_weight
 *   best_s_out    [sum ns] (int)      optional   worst-deviation sample index per segment (see above)
 *   sample_offset_out [B+1] (int64)   required   exact CSR layout of the samples (always complete)
 *   samples_out   [sample_capacity][3] required  rows beyond sample_capacity are dropped (MSNAP_FLAG_TRUNCATED,
 *                                                 and the _host variant returns MSNAP_ERR_CAPACITY)
 *   stats_out     [B][2]              optional   max |dz|/dxy and min turn radius (ms.cpp:163-195; 1e12 if none)
 *   flags_out     [B] (unsigned)      optional   MSNAP_FLAG_* */
int msnap_generate_batch_dev(msnap_handle h, const msnap_config *cfg, double sample_distance_override,
                             double v_avg_override, long long B, int ns_uniform, const long long *seg_offset,
                             const double *waypoints, double *times_out, double *coeff_out, double *max_dev_out,
                             int *iters_out, double *vw_final_out, int *best_s_out, long long sample_capacity,
                             long long *sample_offset_out, double *samples_out, double *stats_out,
                             unsigned *flags_out);
int msnap_generate_batch_host(msnap_handle h, const msnap_config *cfg, double sample_distance_override,
                              double v_avg_override, long long B, int ns_uniform, const long long *seg_offset,
                              const double *waypoints, double *times_out, double *coeff_out, double *max_dev_out,
                              int *iters_out, double *vw_final_out, int *best_s_out, long long sample_capacity,
                              long long *sample_offset_out, double *samples_out, double *stats_out,
                              unsigned *flags_out);

/* Upper bound on the number of sample rows msnap_generate_batch can produce for this input (every candidate of
 * ms.cpp:140 accepted, plus first and last point), computed on the device from the waypoints alone.
 * waypoints/seg_offset are DEVICE pointers for _dev (result written to *rows_out_dev, a device int64) and HOST
 * pointers for _host (result returned in *rows_out). */
int msnap_sample_bound_dev(msnap_handle h, const msnap_config *cfg, double v_avg_override, long long B, int ns_uniform,
                           const long long *seg_offset, const double *waypoints, long long *rows_out_dev);
int msnap_sample_bound_host(msnap_handle h, const msnap_config *cfg, double v_avg_override, long long B,
                            int ns_uniform, const long long *seg_offset, const double *waypoints, long long *rows_out);

/* ---- single-trajectory convenience (what Minisnap_3D / Minisnap_EN need; batch of one) ---------------------
 * Returns the sample count in *n_samples_out; writes at most sample_capacity rows.  n_points < 2 is
 * MSNAP_ERR_INVALID_ARG with *n_samples_out = 0 (the reference returns an empty matrix, ms.cpp:54-57). */
int msnap_generate_one_host(msnap_handle h, const msnap_config *cfg, double sample_distance_override,
                            double v_avg_override, int n_points, const double *waypoints, long long sample_capacity,
                            double *samples_out, long long *n_samples_out);

/* ---- WGS84 <-> ENU, batched -----------------------------------------------------
 * Drop-in for UavPathPlanner::wgs84ToENU_Batch / enuToWGS84_Batch (/root/reference/uavPathPlanning.cpp:1085-1108;
 * per point wgs84ToENU cpp:1047-1063 and enuToWGS84 cpp:1066-1083 over wgs84ToECEF cpp:894-910, ecefToWGS84
 * cpp:926-968, ecefToENU / enuToECEF cpp:1023-1044; constants uavPathPlanning.hpp:134-173): the map every waypoint
 * takes right before the minimum-snap path (cpp:2640, 3217) and every sampled point right after it (cpp:3699, 3806,
 * 4909-4913).
 *   reference_lla : HOST pointer to {lon_deg, lat_deg, alt_m} == struct WGS84Point (hpp:145-149), the ENU origin
 *   lla rows      : [n][3] {lon_deg, lat_deg, alt_m}  == std::vector<WGS84Point>
 *   enu rows      : [n][3] {east, north, up} metres    == std::vector<ENUPoint> (hpp:152-156)
 * _dev: device rows, enqueued on the handle's stream; _host: host rows, returns when the output is complete.
 * In place (output == input) is allowed.  Same formulas in the same operation order as the reference, including the
 * <= 10-step fixed-point iteration with its 1e-12 rad stopping rule; results agree with the reference's to ~1e-9 m
 * (tests: 1e-6 m), the difference being CUDA's sin/cos/atan2 against the host libm's. */
int msnap_wgs84_to_enu_dev(msnap_handle h, const double *reference_lla, long long n, const double *lla, double *enu_out);
int msnap_wgs84_to_enu_host(msnap_handle h, const double *reference_lla, long long n, const double *lla, double *enu_out);
int msnap_enu_to_wgs84_dev(msnap_handle h, const double *reference_lla, long long n, const double *enu, double *lla_out);
int msnap_enu_to_wgs84_host(msnap_handle h, const double *reference_lla, long long n, const double *enu, double *lla_out);
/* msnap_enu_to_wgs84_dev for a buffer whose fill level is only known on the device: converts the first
 * min(*n_rows_dev, n_rows_cap) rows (n_rows_dev = e.g. sample_offset_out + B of msnap_generate_batch_dev). */
int msnap_enu_to_wgs84_counted_dev(msnap_handle h, const double *reference_lla, long long n_rows_cap,
                                   const long long *n_rows_dev, const double *enu, double *lla_out);
/* Frame of the rows msnap_generate_batch_* / msnap_generate_one_host write to samples_out: 0 (default) = ENU, as
 * GenerateTrajectoryMatrix returns them; 1 = WGS84 {lon, lat, alt} about reference_lla, i.e. getPlan's
 * `enuToWGS84_Batch(Trajectory_ENU, origin_)` (cpp:3699) applied on the device before the rows leave it.  The
 * statistics in stats_out are those of the ENU rows either way.  reference_lla may be NULL for frame 0. */
int msnap_set_sample_frame(msnap_handle h, int frame, const double *reference_lla);
/* Frame of the waypoints msnap_generate_batch_* / msnap_generate_one_host / msnap_sample_bound_* read: 0 (default) = ENU;
 * 1 = WGS84 {lon, lat, alt} rows, converted about reference_lla on the device before anything else, i.e. prepareWaypoints'
 * `Enu_waypoint = wgs84ToENU_Batch(wgs84_points, origin_)` (cpp:2640) in front of Minisnap_3D (cpp:3684).  With both
 * frames set to 1 a call is getPlan's leader chain WGS84 waypoints -> ENU -> minimum snap -> sampled ENU -> WGS84 rows
 * (cpp:2640, 3684, 3699) without the ENU data ever leaving the device.  msnap_solve_qp_batch_* is not affected. */
int msnap_set_waypoint_frame(msnap_handle h, int frame, const double *reference_lla);
/* Execution form of ecefToWGS84's fixed-point iteration (cpp:926-968) inside every ENU -> WGS84 call of this handle:
 * 0 (default) = the iteration carried on direction vectors (one sqrt per step, the angle formed once at the end);
 * 1 = statement by statement as the reference writes it (sin, cos, sqrt, two divisions and an atan2 per step).  Same start
 * value, step, stopping rule and step limit; results differ by rounding only (~1e-16 rad), about 4x in speed. */
int msnap_set_geo_exact_trig(msnap_handle h, int enable);
/* Test hook: msnap_enu_to_wgs84_dev that also writes the number of fixed-point steps taken per point (cpp:939-949). */
int msnap_debug_geo_steps_dev(msnap_handle h, const double *reference_lla, long long n, const double *enu, double *lla_out,
                              int *steps_out);

/* ---- altitude optimisation of sampled trajectories, batched ----------------------
 * Drop-in for UavPathPlanner::optimizeSegmentAltitudeENU (/root/reference/uavPathPlanning.cpp:1329-1364) = optimizeHeights
 * (cpp:1575-1712) followed by optimizeHeightsGlobalSmooth with lambda_smooth x 10 and max_climb_rate x 0.5 (cpp:1714-1827):
 * the step getPlan runs on Minisnap_3D's sampled trajectory before converting it to WGS84 (cpp:3712-3729, 1535-1573).  One
 * independent problem per trajectory; the reference's Eigen::SimplicialLDLT on the SPD pentadiagonal Hessian is replaced
 * by a banded LDL' recurrence per trajectory.
 *   params      : struct AltitudeParams (uavPathPlanning.hpp:415-421; config.yaml:1-8 ships 1.0, 1.0, 0.3, 2.0, 10.0)
 *   row_offset  : [B+1] (int64) CSR layout of the rows, e.g. the sample_offset_out of msnap_generate_batch_*
 *   rows_inout  : [rows][3] {east, north, up}; the `up` column is replaced by the optimised heights (cpp:1357-1359)
 *   elev        : [rows] terrain elevation the cost map returns at each row (getCostAt, elevation_cost_map.cpp:373-380),
 *                 NaN where it has no value; NULL = no terrain anywhere.  msnap_cost_map_lookup_dev fills it from a grid.
 *   z_pass1_out : [rows] optional, the heights after optimizeHeights (before the global smoothing pass)
 *   solves_out  : [B] optional, solves the active-set loop of pass 2 took (1..10, cpp:1733-1814)
 *   flags_out   : [B] optional, bit 0 = a non-positive pivot appeared (the reference's "decomposition failed"); bit 2 =
 *                 it appeared in pass 2; bit 1 (== MSNAP_FLAG_TRUNCATED) = the trajectory's rows reach beyond n_rows_cap
 * Failure behaviour is the reference's (cpp:1342-1344, 1356): a trajectory whose pass 1 fails keeps its input `up` values,
 * one whose pass 2 fails keeps the pass-1 heights.
 * _dev: device pointers, n_rows_cap = allocated rows of rows_inout / elev / z_pass1_out (the exact count is read from
 * row_offset[B] on the device), enqueued on the handle's stream.  No row at or beyond n_rows_cap is read or written: when
 * row_offset comes from a msnap_generate_batch_dev call that ran out of sample capacity (MSNAP_FLAG_TRUNCATED), the
 * trajectories whose rows do not fit entirely are skipped (bit 1 of flags_out) and keep whatever rows they have.
 * _host: host pointers, returns when the outputs are complete. */
typedef struct msnap_altitude_params {
    double lambda_smooth, lambda_follow, max_climb_rate, uav_R, safe_distance;
} msnap_altitude_params;
void msnap_altitude_params_default(msnap_altitude_params *p);
int msnap_altitude_optimize_batch_dev(msnap_handle h, const msnap_altitude_params *params, long long B,
                                      const long long *row_offset, long long n_rows_cap, double *rows_inout,
                                      const double *elev, double *z_pass1_out, int *solves_out, unsigned *flags_out);
int msnap_altitude_optimize_batch_host(msnap_handle h, const msnap_altitude_params *params, long long B,
                                       const long long *row_offset, double *rows_inout, const double *elev,
                                       double *z_pass1_out, int *solves_out, unsigned *flags_out);
/* Execution form of the per-trajectory banded solves: 2 (default) = partitioned: a trajectory's rows are cut into up to
 * 8 chunks (32 for trajectories of more than 262 rows) eliminated by one lane each, the 2-row separators between them are
 * solved across the lanes, and the whole stage (edge weights, both passes, write-back) is one launch; 0 = a lane pair per
 * trajectory (two-sided elimination meeting at the two middle rows), 1 = one lane per trajectory (plain downward LDL').
 * Same systems, same active-set decisions; heights differ by rounding (~1e-9 m). */
int msnap_set_altitude_policy(msnap_handle h, int policy);
/* ElevationCostMap::getCostAt (elevation_cost_map.cpp:373-380) for every row: grid is a DEVICE float array
 * [height][width], row-major, top-left origin at (origin_x, origin_y) in ENU metres, square cells of `resolution` metres;
 * elev_out[i] = the cell containing (east_i, north_i), NaN outside the grid.  n_rows_dev (device int64, may be NULL)
 * bounds the rows actually looked up, e.g. sample_offset + B. */
int msnap_cost_map_lookup_dev(msnap_handle h, const float *grid, int width, int height, double resolution, double origin_x,
                              double origin_y, long long n_rows_cap, const long long *n_rows_dev, const double *rows,
                              double *elev_out);

/* ---- Bezier generator, batched -----------------------------------------------------
 * Drop-in for math_util::Bezier::GenerateTrajectoryMatrix (/root/reference/math_util/bezier.hpp:98-120, bezier.cpp:127-189
 * around Bezier::GeneratePath, bezier.cpp:29-118): the reference's alternative trajectory generator, selected by
 * getPlan(algorithm == "bezier") through UavPathPlanner::Bezier_3D (uavPathPlanning.cpp:3691-3692, 4477-4505).  Same batch
 * layout as msnap_generate_batch_*: one call = B independent GenerateTrajectoryMatrix calls.
 *   sample_distance_override : the path resolution in metres; <= 0 means 1.0 (bezier.cpp:133-136)
 *   min_radius               : BezierConfig::min_radius (bezier.hpp:91-96; default 1.0 = no curvature constraint; Bezier_3D
 *                              sets 300 whenever its own min_radius argument is > 0, cpp:4491-4494)
 *   sample_offset_out [B+1]  : exact CSR layout of the rows (always complete, also with sample_capacity = 0: the sizing call)
 *   samples_out [capacity][3]: rows beyond sample_capacity are dropped (MSNAP_FLAG_TRUNCATED; _host: MSNAP_ERR_CAPACITY)
 *   flags_out [B]            : optional; MSNAP_FLAG_NONFINITE = a non-finite or absurdly long segment (more than 1e7 samples)
 *                              was replaced by its end waypoint instead of spinning in bezier.cpp:109
 * Row counts are the reference's (the number of accumulated `t += resolution / dis` steps <= 1, bezier.cpp:109), rows agree
 * to the rounding of atan2 / cos / sin / hypot (1e-12 m; tests: 1e-9 m).  yaml_path and v_avg_override of the reference
 * signature are unused there (bezier.cpp:127) and have no counterpart here. */
int msnap_bezier_generate_batch_dev(msnap_handle h, double sample_distance_override, double min_radius, long long B,
                                    int ns_uniform, const long long *seg_offset, const double *waypoints,
                                    long long sample_capacity, long long *sample_offset_out, double *samples_out,
                                    unsigned *flags_out);
int msnap_bezier_generate_batch_host(msnap_handle h, double sample_distance_override, double min_radius, long long B,
                                     int ns_uniform, const long long *seg_offset, const double *waypoints,
                                     long long sample_capacity, long long *sample_offset_out, double *samples_out,
                                     unsigned *flags_out);

/* ---- single-loop patrol post-processing on the sampled rows ----------------------------
 * What UavPathPlanner::gen_single_patrol does with Minisnap_3D's output (/root/reference/uavPathPlanning.cpp:1829-1906,
 * helpers cpp:118-206).  The caller closes each patrol polygon P0..Pn-1 into the waypoints P0..Pn-1, P0, P1 (cpp:1841-1847)
 * and runs msnap_generate_batch_* on them with getPlan's overrides (cpp:1849); this call then, per trajectory: trims the loop
 * at the sample closest to the second P0 (cpp:1857-1879), sets `up` to keep_up and closes the loop with its first sample
 * (cpp:1885-1892), tests it for a horizontal self-intersection (hasSelfIntersection2D, cpp:152-177) and, if there is one,
 * replaces it by the polygon boundary sampled every `distance` metres (sampleClosedPolygonBoundary, cpp:179-206, 1897-1903).
 *   waypoints / ns_uniform / seg_offset : the CLOSED waypoint lists the generate call took (>= 5 points per trajectory)
 *   sample_offset [B+1], samples        : that call's outputs; rows at or beyond sample_capacity are treated as absent
 *   keep_up [B] or NULL                 : the height the loop is flown at (cpp:1839: the last `up` of the trajectory flown
 *                                         before it); NULL = the polygon's first vertex
 *   out_offset [B+1]                    : exact CSR layout of the result (always complete, also with out_capacity = 0)
 *   out_rows [out_capacity][3]          : the patrol loops; rows beyond the capacity are dropped (MSNAP_FLAG_TRUNCATED; _host
 *                                         returns MSNAP_ERR_CAPACITY)
 *   flags_out [B], optional             : bit 0 = the generator had produced no rows (cpp:1850-1855: empty result), bit 2 =
 *                                         self-intersection, boundary sampling used
 * Decisions (trim index, intersection verdict, fallback row count) are the reference's bit for bit: they are comparisons of
 * sums and products of the input doubles evaluated in the reference's order. */
int msnap_patrol_postprocess_dev(msnap_handle h, double distance, long long B, int ns_uniform, const long long *seg_offset,
                                 const double *waypoints, const long long *sample_offset, const double *samples,
                                 long long sample_capacity, const double *keep_up, long long out_capacity,
                                 long long *out_offset, double *out_rows, unsigned *flags_out);
int msnap_patrol_postprocess_host(msnap_handle h, double distance, long long B, int ns_uniform, const long long *seg_offset,
                                  const double *waypoints, const long long *sample_offset, const double *samples,
                                  const double *keep_up, long long out_capacity, long long *out_offset, double *out_rows,
                                  unsigned *flags_out);

/* ---- follower formation trajectories on the leader's sampled rows ---------------------
 * Drop-in for the numeric core of UavPathPlanner::generateFollowerTrajectories (/root/reference/uavPathPlanning.cpp:3931-4074)
 * and its four formation generators (cpp:4076-4398): smoothed leader headings (central differences, +-10-sample circular
 * mean when the trajectory has more than 5 rows), one fixed body-frame offset per follower from the formation model, the
 * offset rotated into the leader's heading at every row, the result converted to WGS84.  One call = B leader trajectories
 * (e.g. the sample_offset_out / samples_out of msnap_generate_batch_*, after msnap_altitude_optimize_batch_*).
 *   formation_model        : 1 V shape, 2 line abreast, 3 trail (columns of uav_formation_max_row), 4 triangle; other = 1
 *   formation_distance     : metres, AFTER the reference's lower bound -- msnap_formation_distance() applies it (cpp:4044-4051)
 *   frame                  : 0 = rows {east, north, up}; 1 = rows {lon, lat, alt} about reference_lla (cpp:4141), with the
 *                            reference's rule for models 2-4 that a follower's first row is {its start lon, its start lat,
 *                            the leader's first up} (cpp:4188-4195) when starts_wgs84 is given
 *   starts_wgs84 [n_followers][3] : uav_start_point_wgs84, may be NULL (frame 0 ignores it)
 *   row_offset [B+1], leader_rows : CSR rows of the leaders; rows at or beyond n_rows_cap are treated as absent
 *   out_rows [n_followers * rows][3] : trajectory b owns the block starting at n_followers * row_offset[b], follower-major
 *                            ([follower][row]); rows beyond out_capacity are not written
 * _dev: device pointers (reference_lla is a HOST pointer), enqueued on the handle's stream; _host: host pointers. */
double msnap_formation_distance(double formation_distance, double position_misalignment, double uav_R);
int msnap_followers_dev(msnap_handle h, int formation_model, double formation_distance, int uav_formation_max_row,
                        int n_followers, int frame, const double *reference_lla, const double *starts_wgs84_dev, long long B,
                        const long long *row_offset, const double *leader_rows, long long n_rows_cap, long long out_capacity,
                        double *out_rows);
int msnap_followers_host(msnap_handle h, int formation_model, double formation_distance, int uav_formation_max_row,
                         int n_followers, int frame, const double *reference_lla, const double *starts_wgs84, long long B,
                         const long long *row_offset, const double *leader_rows, double *out_rows);

/* ---- per-kernel timing (bench.py's roofline pass) ------------------------------------------------------------
 * Between begin and end every kernel the handle launches is bracketed by a CUDA event pair on the launching stream.
 * msnap_profile_end synchronises and writes a JSON object {"<kernel>": {"launches": n, "total_ms": t}, ...}. */
int msnap_profile_begin(msnap_handle h);
int msnap_profile_end(msnap_handle h, char *json_out, long long capacity);

/* Developer instrumentation (not needed by an integrator): per-CTA clock64() stamps after each phase of the fused
 * kernel's first tile (rows 0..4095) and of the sampler's tiles (rows 4096..8191).  enable != 0 arms a device buffer
 * [8192][16]; out != NULL copies the last stamps to the host. */
int msnap_debug_phase_clocks(msnap_handle h, int enable, long long *out);

/* ---- micro-benchmarks used for the roofline denominators (bench.py) --------------------------------------- */
/* Sustained DFMA rate of this GPU in TFLOP/s (2 flops per DFMA), measured with CUDA events. */
int msnap_measure_fp64_peak(msnap_handle h, double *tflops_out);

#ifdef __cplusplus
}
#endif
#endif /* MSNAP_H */
