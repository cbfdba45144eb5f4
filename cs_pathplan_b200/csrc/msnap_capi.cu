// msnap_capi.cu -- handle, workspace, kernel launch sequences and the extern "C" boundary of include/msnap.h.
// Compiled for sm_100a only; there is no CPU code path behind any compute entry point.
#include "../../include/msnap.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "msnap_device.cuh"
#include "msnap_generic.cuh"
#include "msnap_fused.cuh"
#include "msnap_geo.cuh"
#include "msnap_alt.cuh"
#include "msnap_alt_part.cuh"
#include "msnap_bezier.cuh"
#include "msnap_patrol.cuh"
#include "msnap_follow.cuh"

static const MsnapOrderTab h_tab[MSNAP_MAX_ORDER - MSNAP_MIN_ORDER + 1] = MSNAP_ORDER_TABLES;

using namespace msnap;

static constexpr int T_TABLE_N = SAMPLE_TTAB_BIG;  // candidate-time table of the sampler (>= SAMPLE_TTAB_N; long segments index all of it)

// ------------------------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------------------------
struct Arena {  // grow-only device buffer with bump allocation, reset per call
    char *base = nullptr;
    size_t cap = 0, used = 0;
};

// Launch plan of the fused kernel for a uniform batch, or tpc == 0 if the batch does not qualify.
struct FusedPlan {
    int tpc = 0, nit = 1, lane_stride = 32, traj_stride = 0, grid = 0;
    int variant = 0;  // 0: FUSED_THREADS x 2 CTAs per SM, 1: FUSED_THREADS_3 x 3 CTAs per SM (many-wave batches)
    long long n_tiles = 0;
    size_t smem = 0, state_bytes = 0;
    // cache key
    int order = 0, ns = 0;
    long long B = 0;
};

struct msnap_context {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    MsnapOrderTab *d_tab = nullptr;  // global-memory copy of the tables (lane-divergent lookups)
    double *d_ttab = nullptr;        // t_table[i] = the i-th candidate time of the sampler's `t += 0.1` accumulation
    Arena ws;                        // solver workspace
    Arena io;                        // device mirrors of host buffers (_host entry points)
    long long launches = 0;
    int policy = 0;
    bool scan_coef_smem = false;      // sampler: stage the tile's coefficients in shared memory (MSNAP_SCAN_COEF_SMEM=1)
    std::vector<FusedPlan> plans;     // cached launch plans
    // optional per-kernel timing (msnap_profile_begin/end): one event pair per launch on the launching stream
    long long *phase_clocks = nullptr;  // dev instrumentation (msnap_debug_phase_clocks)
    bool profiling = false;
    struct Timed { const char *name; cudaEvent_t e0, e1; };
    std::vector<Timed> timed;
    int sm_count = 0;
    std::string last_error;
    // host-pointer entry points: big batches are cut into chunks that alternate between two child contexts (own
    // stream, own arenas), so that a chunk's kernels overlap the previous chunk's device-to-host copies
    std::vector<msnap_context *> kids;
    int host_chunks = 0;              // 0 = automatic (MSNAP_HOST_CHUNKS)
    bool discard_state = true;        // fused solve: discard dead sweep state from L2 (MSNAP_DISCARD_STATE)
    bool zero_copy = false;           // host path: store big results straight into pinned caller buffers (MSNAP_ZERO_COPY);
                                      // measured on B200/PCIe 5: SM stores reach ~24 GB/s, the copy engine ~53 GB/s => off
    long long *h_off = nullptr;       // pinned staging for a chunk's sample offsets
    size_t h_off_cap = 0;
    cudaStream_t aux = nullptr;       // host path: copies the solve's results out while the sampler is still running
    cudaEvent_t ev_solved = nullptr;  // recorded on `stream` between the solve and the sampler (host path only)
    bool mark_solved = false;
    // sampler output frame (msnap_set_sample_frame): 0 = ENU rows as the reference's sampler returns them, 1 = the rows
    // leave as [lon, lat, alt] = enuToWGS84_Batch of the sampled trajectory (uavPathPlanning.cpp:3699)
    int frame = 0;
    GeoFrame geo{};
    int wp_frame = 0;       // msnap_set_waypoint_frame: 1 = generate / sample_bound take WGS84 waypoints (cpp:2640)
    GeoFrame wp_geo{};
    bool alt_smem_opted = false;
    int alt_policy = 2;     // msnap_set_altitude_policy: 0 = lane pairs (two-sided elimination), 1 = one lane per trajectory,
                            // 2 = partitioned over 8 / 32 lanes, one launch (default)
    bool geo_trig = false;  // msnap_set_geo_exact_trig: ENU -> WGS84 with the reference's per-step sin/cos/atan2
};

namespace {

#define MS_CUDA(h, expr)                                                                         \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            (h)->last_error = std::string(#expr) + ": " + cudaGetErrorString(e__);               \
            return MSNAP_ERR_CUDA;                                                               \
        }                                                                                        \
    } while (0)

struct DeviceGuard {  // make the handle's device current for the duration of a call
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int arena_reserve(msnap_context *h, Arena &a, size_t bytes) {
    a.used = 0;
    if (bytes <= a.cap) return MSNAP_OK;
    if (a.base) {
        MS_CUDA(h, cudaStreamSynchronize(h->stream));
        MS_CUDA(h, cudaFree(a.base));
        a.base = nullptr;
        a.cap = 0;
    }
    const size_t want = bytes + bytes / 8 + (1u << 20);
    cudaError_t e = cudaMalloc(&a.base, want);
    if (e != cudaSuccess) {
        h->last_error = std::string("cudaMalloc(workspace): ") + cudaGetErrorString(e);
        cudaGetLastError();
        return MSNAP_ERR_ALLOC;
    }
    a.cap = want;
    return MSNAP_OK;
}

template <class T>
T *arena_take(Arena &a, size_t n) {
    const size_t bytes = (n * sizeof(T) + 255) & ~size_t(255);
    T *p = reinterpret_cast<T *>(a.base + a.used);
    a.used += bytes;
    return p;
}
inline size_t padded(size_t bytes) { return (bytes + 255) & ~size_t(255); }

inline unsigned grid_for(long long n, int block) { return (unsigned)((n + block - 1) / block); }

inline void prof_before(msnap_context *h, const char *name) {
    if (!h->profiling) return;
    msnap_context::Timed t{name, nullptr, nullptr};
    cudaEventCreate(&t.e0);
    cudaEventCreate(&t.e1);
    cudaEventRecord(t.e0, h->stream);
    h->timed.push_back(t);
}
inline void prof_after(msnap_context *h) {
    if (h->profiling) cudaEventRecord(h->timed.back().e1, h->stream);
}

#define MS_LAUNCH(h, kernel, grid, block, ...)                                                   \
    do {                                                                                         \
        prof_before((h), #kernel);                                                               \
        kernel<<<(grid), (block), 0, (h)->stream>>>(__VA_ARGS__);                                \
        prof_after((h));                                                                         \
        ++(h)->launches;                                                                         \
        cudaError_t e__ = cudaPeekAtLastError();                                                 \
        if (e__ != cudaSuccess) {                                                                \
            (h)->last_error = std::string(#kernel) + ": " + cudaGetErrorString(e__);             \
            cudaGetLastError();                                                                  \
            return MSNAP_ERR_CUDA;                                                               \
        }                                                                                        \
    } while (0)

// ------------------------------------------------------------------------------------------------------------
// generic (per-phase) pipeline
// ------------------------------------------------------------------------------------------------------------
struct SolveIO {
    const double *wp = nullptr;   // device
    const double *times_in = nullptr;  // device, or nullptr => allocate (v_avg, min_time)
    double v_avg = 0, min_time = 0;
    double *times_out = nullptr, *coeff_out = nullptr, *max_dev_out = nullptr, *vw_final_out = nullptr;
    int *iters_out = nullptr, *best_s_out = nullptr;
    unsigned *flags_out = nullptr;
    // zero-copy outputs of the host-pointer path: device-visible addresses of the caller's PINNED host buffers
    double *coeff_mirror = nullptr;   // the fused kernel stores the coefficients there as well (honoured: mirror_done)
    bool *mirror_done = nullptr;
};


template <int O>
FusedPlan plan_fused(msnap_context *h, const BatchIdx &bi, const SolveParams &sp) {
    using D = Dim<O>;
    FusedPlan f;
    if (bi.ns_uniform <= 0 || h->policy == 1) return f;
    const int ns = bi.ns_uniform;
    const int nit = sp.pw > 0.0 ? sp.max_iter + 1 : 1;
    for (const FusedPlan &c : h->plans)
        if (c.order == O && c.ns == ns && c.nit == nit && c.B == bi.B) return c;
    f.order = O;
    f.ns = ns;
    f.nit = nit;
    f.B = bi.B;
    if (nit > FUSED_THREADS) return f;
    const FusedSmem<O> L(ns);
    const size_t blk = (size_t)L.size * sizeof(double);
    const size_t st1 = (size_t)(ns - 1) * D::NSTATE * FUSED_SMEM_LANES * sizeof(double);  // shared-memory state rows
    auto smem_for = [&](int tpc) { return tpc * blk + st1 + (size_t)tpc * nit * 12 + (size_t)tpc * 8 + 16; };
    const size_t hard = 220 * 1024;
    // lanes: 2 * tpc lanes (one pair per trajectory, within ONE warp) for the last iteration, then, from the next warp
    // boundary, tpc * (nit - 1) speculative lanes
    int tmax = FUSED_SMEM_LANES < 16 ? FUSED_SMEM_LANES : 16;
    while (tmax > 1 && (32 + tmax * (nit - 1) > FUSED_THREADS || tmax * (nit - 1) > FUSED_SLOT_LANES)) --tmax;
    if (32 + tmax * (nit - 1) > FUSED_THREADS || tmax * (nit - 1) > FUSED_SLOT_LANES) return f;
    if ((long long)tmax > bi.B) tmax = (int)bi.B;
    // Pick the tile size that needs the fewest waves of resident CTAs (the kernel is latency-bound per tile, so a
    // partial second wave costs a whole tile latency); among those, the largest tile (best lane utilisation).
    long long best_waves = -1;
    int best_occ = 0;
    for (int tpc = tmax; tpc >= 1; --tpc) {
        const size_t sm = smem_for(tpc);
        if (sm > hard) continue;
        int occ = 0;
        cudaFuncSetAttribute(k_fused_solve<O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_fused_solve<O>, FUSED_THREADS, sm) != cudaSuccess ||
            occ < 1) {
            cudaGetLastError();
            continue;
        }
        const long long tiles = (bi.B + tpc - 1) / tpc, resident = (long long)occ * h->sm_count;
        const long long waves = (tiles + resident - 1) / resident;
        // beyond a few waves the persistent loop amortises everything: prefer big tiles then
        const long long score = waves > 4 ? 4 : waves;
        // single wave: the smallest such tile (more CTAs per SM overlap each other's phases)
        if (best_waves < 0 || score < best_waves || (score == 1 && best_waves == 1)) {
            best_waves = score;
            best_occ = occ;
            f.tpc = tpc;
        }
    }
    // Many waves deep: what counts is how many chains an SM keeps in flight.  The 192-thread build of the kernel fits three
    // CTAs per SM when the tile's shared memory allows it (short trajectories): half again as many warps.
    static const bool allow3 = !(std::getenv("MSNAP_FUSED_3CTA") && std::atoi(std::getenv("MSNAP_FUSED_3CTA")) == 0);  // (A/B knob)
    // (measured at cfg3, 2^20 x 8: 6.05 -> 5.50 ms with the reweighting lanes; a single solve per trajectory is 7 % slower with it)
    if (allow3 && nit > 1 && f.tpc > 0 && best_waves >= 4 && 32 + f.tpc * (nit - 1) <= FUSED_THREADS_3) {
        const size_t sm = smem_for(f.tpc);
        int occ3 = 0;
        cudaFuncSetAttribute((k_fused_solve<O, FUSED_THREADS_3, 3>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hard);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ3, (k_fused_solve<O, FUSED_THREADS_3, 3>), FUSED_THREADS_3, sm) ==
                cudaSuccess && occ3 > best_occ) {
            f.variant = 1;
            best_occ = occ3;
        } else {
            cudaGetLastError();
        }
    }
    if (f.tpc == 0) return f;  // a single trajectory does not fit: generic path
    f.traj_stride = L.size;
    f.smem = smem_for(f.tpc);
    cudaFuncSetAttribute(k_fused_solve<O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hard);
    f.lane_stride = FUSED_SLOT_LANES;
    f.n_tiles = (bi.B + f.tpc - 1) / f.tpc;
    const long long resident = (long long)best_occ * h->sm_count;
    f.grid = (int)(f.n_tiles < resident ? f.n_tiles : resident);
    f.state_bytes = nit > 1 ? (size_t)f.grid * (ns - 1) * D::NSTATE * f.lane_stride * sizeof(double) : 0;
    if (h->plans.size() > 64) h->plans.clear();
    h->plans.push_back(f);
    return f;
}

struct SolveWs {
    double *T, *base, *state, *segx, *coeff;
    int *s_star;
    FusedPlan fused;
    // speculative reweighting loop of the generic path (nullptr: sequential loop)
    double *spec_state = nullptr, *md_ws = nullptr, *md_last = nullptr;
    int *ok_ws = nullptr;
    unsigned *flag_last = nullptr;
};
constexpr int SPEC_NIT1 = 10;                         // speculative iterations of the generic path (max_iter = 10)
constexpr size_t SPEC_WS_LIMIT = (size_t)24 << 30;    // beyond this the generic path falls back to the sequential loop

template <int O>
bool use_generic_spec(const msnap_context *h, long long n_seg, const SolveParams &sp) {
    return h->policy != 1 && sp.pw > 0.0 && sp.max_iter == SPEC_NIT1 &&
           (size_t)(n_seg + 1) * Dim<O>::NSTATE * SPEC_NIT1 * sizeof(double) <= SPEC_WS_LIMIT;
}

template <int O>
size_t solve_ws_bytes(long long n_seg, long long B, bool need_T, bool need_coeff, const FusedPlan &f, bool spec) {
    using D = Dim<O>;
    size_t b = 0;
    if (f.tpc == 0 && spec)
        b += padded((size_t)(n_seg + 1) * D::NSTATE * SPEC_NIT1 * sizeof(double)) +
             padded((size_t)B * SPEC_NIT1 * sizeof(double)) + padded((size_t)B * SPEC_NIT1 * sizeof(int)) +
             padded(B * sizeof(double)) + padded(B * sizeof(unsigned));
    if (need_T) b += padded(n_seg * sizeof(double));
    if (f.tpc > 0) {  // fused path: only the per-CTA state slots (+ coefficients if the caller keeps none)
        b += padded(f.state_bytes + 256);
        if (need_coeff) b += padded((size_t)n_seg * 3 * D::M * sizeof(double));
        return b;
    }
    b += padded((size_t)n_seg * D::NBASE * sizeof(double));
    b += padded((size_t)(n_seg + 1) * D::NSTATE * sizeof(double));
    b += padded((size_t)n_seg * D::NSEGX * sizeof(double));
    b += padded((size_t)n_seg * sizeof(int));
    if (need_coeff) b += padded((size_t)n_seg * 3 * D::M * sizeof(double));
    return b;
}

// Enqueue the closed-form solve (with or without the reweighting loop) for the whole batch.
template <int O>
int run_solve(msnap_context *h, const BatchIdx &bi, const SolveParams &sp, const SolveIO &io, SolveWs &w) {
    using D = Dim<O>;
    const int blk = 128;
    // Thread-per-trajectory sweeps touch 32 different cache lines per warp access, so they are bound by the L1 of the
    // SMs they run on: small batches use one warp per CTA to spread over as many SMs as possible.
    const int blk_b = bi.B >= 128LL * h->sm_count ? 128 : 32;
    // k_thomas_pair: two lanes per trajectory.  Every access of a warp touches as many cache lines as it has lanes, so
    // small batches run with FEWER lanes per warp (down to 4) to get ~4 warps per SM instead of a few full ones.
    int blk_p = 64;
    while (blk_p > 4 && 2 * bi.B < (long long)blk_p * 4 * h->sm_count) blk_p >>= 1;
    const unsigned gs = grid_for(bi.n_seg, blk), gb = grid_for(bi.B, blk_b);
    const double *ht = &h->d_tab[O - MSNAP_MIN_ORDER].HT[0][0];
    if (w.fused.tpc > 0) {  // uniform batch: one persistent launch for the whole closed-form solve
        const FusedPlan &f = w.fused;
        FusedParams fp{};
        fp.B = bi.B;
        fp.ns = bi.ns_uniform;
        fp.tpc = f.tpc;
        fp.nit = f.nit;
        fp.traj_stride = f.traj_stride;
        fp.n_tiles = f.n_tiles;
        fp.wp = io.wp;
        fp.times_in = io.times_in;
        fp.v_avg = io.v_avg;
        fp.min_time = io.min_time;
        fp.sp = sp;
        fp.ht = ht;
        fp.times_out = io.times_in ? nullptr : w.T;  // allocated times go to the workspace (the sampler reads them)
        fp.coeff_out = w.coeff;
        fp.discard_state = h->discard_state ? 1 : 0;
        fp.coeff_mirror = io.coeff_mirror;
        if (io.coeff_mirror && io.mirror_done) *io.mirror_done = true;
        fp.max_dev_out = io.max_dev_out;
        fp.vw_final_out = io.vw_final_out;
        fp.iters_out = io.iters_out;
        fp.best_s_out = io.best_s_out;
        fp.flags = io.flags_out;
        fp.state_ws = w.state;
        fp.phase_clocks = h->phase_clocks;
        prof_before(h, "k_fused_solve");
        if (f.variant == 1)
            k_fused_solve<O, FUSED_THREADS_3, 3><<<f.grid, FUSED_THREADS_3, f.smem, h->stream>>>(fp);
        else
            k_fused_solve<O><<<f.grid, FUSED_THREADS, f.smem, h->stream>>>(fp);
        prof_after(h);
        ++h->launches;
        cudaError_t e = cudaPeekAtLastError();
        if (e != cudaSuccess) {
            h->last_error = std::string("k_fused_solve: ") + cudaGetErrorString(e);
            cudaGetLastError();
            return MSNAP_ERR_CUDA;
        }
        if (!io.times_in && io.times_out)
            MS_CUDA(h, cudaMemcpyAsync(io.times_out, w.T, bi.n_seg * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        return MSNAP_OK;
    }
    const double *T = io.times_in;
    if (!T) {
        MS_LAUNCH(h, k_times, gs, blk, bi, io.wp, io.v_avg, io.min_time, w.T);
        T = w.T;
    }
    if (io.flags_out) MS_CUDA(h, cudaMemsetAsync(io.flags_out, 0, bi.B * sizeof(unsigned), h->stream));
    const bool use_pw = sp.pw > 0.0;
    if (io.best_s_out) {
        w.s_star = io.best_s_out;  // the caller's buffer doubles as the decision workspace
        if (!use_pw) MS_CUDA(h, cudaMemsetAsync(io.best_s_out, 0, bi.n_seg * sizeof(int), h->stream));
    }
    if (use_pw) {
        // pass 1: snap cost only (no path penalty, no velocity penalty) -> worst-deviation sample per segment
        MS_LAUNCH(h, (k_rows<O>), gs, blk, bi, sp, io.wp, T, false, (const int *)nullptr, ht, w.base, w.segx);
        SolveParams sp1 = sp;
        sp1.max_iter = 0;
        if (h->policy != 1) {
            MS_LAUNCH(h, (k_thomas_pair<O>), grid_for(2 * bi.B, blk_p), blk_p, bi, sp1, io.wp, w.base, w.state, w.segx,
                      false, 0.0, (double *)nullptr, io.flags_out);
        } else {
            MS_LAUNCH(h, (k_thomas<O>), gb, blk_b, bi, sp1, io.wp, w.base, w.state, w.segx, false, false,
                      (double *)nullptr, (int *)nullptr, (double *)nullptr, io.flags_out);
        }
        MS_LAUNCH(h, (k_search<O>), gs, blk, bi, sp, io.wp, T, w.state, w.s_star);
    }
    MS_LAUNCH(h, (k_rows<O>), gs, blk, bi, sp, io.wp, T, use_pw, w.s_star, ht, w.base, w.segx);
    if (w.spec_state) {
        // all reweighting iterations at once: iterations 0..9 as speculative lanes, the last one into the final state
        MS_LAUNCH(h, (k_thomas_spec<O, SPEC_NIT1>), grid_for(bi.B * SPEC_NIT1, blk), blk, bi, sp, io.wp, w.base,
                  w.spec_state, w.segx, w.md_ws, w.ok_ws);
        MS_CUDA(h, cudaMemsetAsync(w.flag_last, 0, bi.B * sizeof(unsigned), h->stream));
        MS_LAUNCH(h, (k_thomas_pair<O>), grid_for(2 * bi.B, blk_p), blk_p, bi, sp, io.wp, w.base, w.state, w.segx, true,
                  reweighted_vw(sp.vw0, SPEC_NIT1), w.md_last, w.flag_last);
        MS_LAUNCH(h, (k_spec_select<O, SPEC_NIT1>), gb, blk_b, bi, sp, io.wp, w.spec_state, w.state, w.segx, w.md_ws,
                  w.ok_ws, w.md_last, w.flag_last, io.max_dev_out, io.iters_out, io.vw_final_out, io.flags_out);
    } else if ((!use_pw || sp.max_iter == 0) && h->policy != 1) {
        // no path penalty (the reweighting loop ends after its first solve, max_dev = 0) or a bare SolveQPClosedForm:
        // exactly one solve -- one lane pair per trajectory
        MS_LAUNCH(h, (k_thomas_pair<O>), grid_for(2 * bi.B, blk_p), blk_p, bi, sp, io.wp, w.base, w.state, w.segx, use_pw,
                  sp.vw0, io.max_dev_out, io.flags_out, io.iters_out, io.vw_final_out);
    } else {
        MS_LAUNCH(h, (k_thomas<O>), gb, blk_b, bi, sp, io.wp, w.base, w.state, w.segx, use_pw, true, io.max_dev_out,
                  io.iters_out, io.vw_final_out, io.flags_out);
    }
    MS_LAUNCH(h, (k_coeff<O>), gs, blk, bi, sp, io.wp, T, w.state, w.coeff, io.flags_out);
    if (io.times_out && io.times_out != T)
        MS_CUDA(h, cudaMemcpyAsync(io.times_out, T, bi.n_seg * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    (void)D::M;
    return MSNAP_OK;
}

template <int O>
void carve_solve_ws(Arena &a, long long n_seg, long long B, bool need_T, double *coeff_out, const FusedPlan &f, bool spec,
                    SolveWs &w) {
    using D = Dim<O>;
    w.fused = f;
    if (f.tpc == 0 && spec) {
        w.spec_state = arena_take<double>(a, (size_t)(n_seg + 1) * D::NSTATE * SPEC_NIT1);
        w.md_ws = arena_take<double>(a, (size_t)B * SPEC_NIT1);
        w.ok_ws = arena_take<int>(a, (size_t)B * SPEC_NIT1);
        w.md_last = arena_take<double>(a, B);
        w.flag_last = arena_take<unsigned>(a, B);
    }
    w.T = need_T ? arena_take<double>(a, n_seg) : nullptr;
    if (f.tpc > 0) {
        w.base = w.segx = nullptr;
        w.s_star = nullptr;
        w.state = arena_take<double>(a, f.state_bytes / sizeof(double) + 32);
        w.coeff = coeff_out ? coeff_out : arena_take<double>(a, (size_t)n_seg * 3 * D::M);
        return;
    }
    w.base = arena_take<double>(a, (size_t)n_seg * D::NBASE);
    w.state = arena_take<double>(a, (size_t)(n_seg + 1) * D::NSTATE);
    w.segx = arena_take<double>(a, (size_t)n_seg * D::NSEGX);
    w.s_star = arena_take<int>(a, n_seg);
    w.coeff = coeff_out ? coeff_out : arena_take<double>(a, (size_t)n_seg * 3 * D::M);
}

struct SampleWs {
    int *seg_count, *append_end;
    unsigned long long *seg_mask;
    double *seg_last;
    long long *seg_start, *traj_count, *partial;
    // single-launch sampler (uniform batches)
    unsigned long long *status = nullptr;
    unsigned int *ticket = nullptr;
    int *sm_ctr = nullptr;  // [512] arrival order of the CTAs on each SM
    long long n_tiles = 0;
    int tpt = 0;
};
// Tile size of the single-launch sampler.  One segment per thread while that still gives every CTA slot of the GPU a
// tile; two segments per thread for big batches (half as many tiles to look back over).
// `capacity` (the caller's row capacity, normally msnap_sample_bound = the number of candidates) is the only hint the host
// has about the segment durations: a batch with thousands of candidates per trajectory (kilometre-long legs: the
// reference's shipped mission) has segments that a whole warp walks (warp_sample_long), and gets one trajectory per tile so
// that those warps spread over all SMs.
int scan_tpt(int ns, long long B, int sm_count, long long capacity) {
    if (capacity / (B > 0 ? B : 1) > 16LL * SAMPLE_MASK_BITS) return 1;
    const int one = ns >= SCAN_THREADS ? 1 : SCAN_THREADS / ns;
    const int two = ns >= 2 * SCAN_THREADS ? 1 : 2 * SCAN_THREADS / ns;
    const long long tiles_two = (B + two - 1) / two;
    return tiles_two >= 2LL * 4 * sm_count ? two : one;
}

// The single-launch sampler keeps a tile's per-segment bookkeeping in shared memory: uniform batches only, and only
// while a one-trajectory tile fits (very long trajectories take the per-pass kernels).
bool use_scan_sampler(int ns_uniform, long long B, int policy, int sm_count, long long capacity) {
    if (ns_uniform <= 0 || policy == 1) return false;
    return scan_smem_bytes(scan_tpt(ns_uniform, B, sm_count, capacity), ns_uniform, 4, false) <= 96 * 1024;
}

size_t sample_ws_bytes(long long n_seg, long long B, int ns_uniform, int policy, int sm_count, long long capacity) {
    if (use_scan_sampler(ns_uniform, B, policy, sm_count, capacity)) {
        const long long n_tiles = B;  // upper bound on the tile count (tpt >= 1)
        return padded((size_t)(n_tiles + 1 + 256) * sizeof(unsigned long long)) + 256;  // + ticket + per-SM counters
    }
    return padded(n_seg * sizeof(int)) + padded(B * sizeof(int)) + padded((size_t)n_seg * 2 * sizeof(unsigned long long)) +
           padded((size_t)n_seg * 3 * sizeof(double)) + padded(n_seg * sizeof(long long)) +
           padded(B * sizeof(long long)) + padded((size_t)(B / SCAN_BLOCK + 2) * sizeof(long long));
}
void carve_sample_ws(Arena &a, long long n_seg, long long B, int ns_uniform, int policy, int sm_count, long long capacity,
                     SampleWs &s) {
    if (use_scan_sampler(ns_uniform, B, policy, sm_count, capacity)) {
        s.tpt = scan_tpt(ns_uniform, B, sm_count, capacity);
        s.n_tiles = (B + s.tpt - 1) / s.tpt;
        s.status = arena_take<unsigned long long>(a, s.n_tiles + 1 + 256);  // + ticket + per-SM arrival counters
        s.ticket = reinterpret_cast<unsigned int *>(s.status + s.n_tiles);
        s.sm_ctr = reinterpret_cast<int *>(s.status + s.n_tiles + 1);
        return;
    }
    s.seg_count = arena_take<int>(a, n_seg);
    s.append_end = arena_take<int>(a, B);
    s.seg_mask = arena_take<unsigned long long>(a, (size_t)n_seg * 2);
    s.seg_last = arena_take<double>(a, (size_t)n_seg * 3);
    s.seg_start = arena_take<long long>(a, n_seg);
    s.traj_count = arena_take<long long>(a, B);
    s.partial = arena_take<long long>(a, B / SCAN_BLOCK + 2);
}

template <int O>
int run_sample(msnap_context *h, const BatchIdx &bi, const double *coeff, const double *T, double sd,
               long long capacity, long long *sample_offset, double *samples, double *stats, unsigned *flags,
               SampleWs &s) {
    if (s.status) {  // uniform batch: count + scan + write in one launch
        const int ns = bi.ns_uniform;
        // stage the tile's coefficients in shared memory while four CTAs still fit an SM
        const bool coef_smem = h->scan_coef_smem && scan_smem_bytes(s.tpt, ns, O, true) <= 52 * 1024;
        const size_t smem = scan_smem_bytes(s.tpt, ns, O, coef_smem);
        MS_CUDA(h, cudaMemsetAsync(s.status, 0, (size_t)(s.n_tiles + 1 + 256) * sizeof(unsigned long long), h->stream));
        // a batch of long legs (the row capacity, normally the number of candidates, averages more than twice what an
        // acceptance mask describes PER SEGMENT: most of the work is in warp-walked segments) walks them four blocks at a time
        const bool long_legs = capacity / ((bi.B > 0 ? bi.B : 1) * (long long)ns) > 2LL * SAMPLE_MASK_BITS;
        if (smem > (long_legs ? 24 : 40) * 1024) {  // the kernel also has ~4 KB (long legs: ~20 KB) of static shared memory
            if (long_legs)
                cudaFuncSetAttribute(k_sample_scan<O, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            else
                cudaFuncSetAttribute(k_sample_scan<O, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        const long long resident = (long long)h->sm_count * 4;
        const unsigned grid = (unsigned)(s.n_tiles < resident ? s.n_tiles : resident);
        prof_before(h, "k_sample_scan");
        if (long_legs)
            k_sample_scan<O, true><<<grid, SCAN_THREADS, smem, h->stream>>>(bi.B, ns, s.tpt, s.n_tiles, coeff, T, sd, h->d_ttab,
                                                                            s.status, s.ticket, capacity, sample_offset, samples,
                                                                            flags, coef_smem ? 1 : 0, s.sm_ctr, h->phase_clocks);
        else
            k_sample_scan<O, false><<<grid, SCAN_THREADS, smem, h->stream>>>(bi.B, ns, s.tpt, s.n_tiles, coeff, T, sd, h->d_ttab,
                                                                             s.status, s.ticket, capacity, sample_offset, samples,
                                                                             flags, coef_smem ? 1 : 0, s.sm_ctr, h->phase_clocks);
        prof_after(h);
        ++h->launches;
        cudaError_t e = cudaPeekAtLastError();
        if (e != cudaSuccess) {
            h->last_error = std::string("k_sample_scan: ") + cudaGetErrorString(e);
            cudaGetLastError();
            return MSNAP_ERR_CUDA;
        }
        if (stats) MS_LAUNCH(h, k_stats, grid_for(bi.B * 32, 256), 256, bi.B, sample_offset, samples, capacity, stats);
        return MSNAP_OK;
    }
    const int blk = 128;
    const unsigned gs = grid_for(bi.n_seg, blk), gb = grid_for(bi.B, blk);
    MS_LAUNCH(h, (k_count<O>), gs, blk, bi, coeff, T, sd, s.seg_count, s.seg_mask, s.seg_last, flags);
    MS_LAUNCH(h, (k_traj_count<O>), gb, blk, bi, coeff, T, s.seg_count, s.seg_last, s.seg_start, s.append_end,
              s.traj_count);
    const int nblk = (int)grid_for(bi.B, SCAN_BLOCK);
    MS_LAUNCH(h, k_scan_reduce, nblk, SCAN_BLOCK, s.traj_count, bi.B, s.partial);
    MS_LAUNCH(h, k_scan_partials, 1, SCAN_BLOCK, s.partial, nblk, sample_offset + bi.B);
    MS_LAUNCH(h, k_scan_apply, nblk, SCAN_BLOCK, s.traj_count, bi.B, s.partial, sample_offset);
    MS_LAUNCH(h, (k_write<O>), gs, blk, bi, coeff, T, sd, h->d_ttab, s.seg_count, s.seg_mask, s.seg_start,
              sample_offset, s.append_end, capacity, samples, flags);
    if (stats) MS_LAUNCH(h, k_stats, grid_for(bi.B * 32, 256), 256, bi.B, sample_offset, samples, capacity, stats);
    return MSNAP_OK;
}

// Parameters the sampler cannot run with (its candidate loop would not terminate): see SAMPLE_T_MAX in msnap_generic.cuh.
bool valid_generate_config(const msnap_config *cfg, double sd, double va) {
    return cfg->min_time_s > 0.0 && cfg->min_time_s <= SAMPLE_T_MAX && std::isfinite(va) && std::isfinite(sd) &&
           std::isfinite(cfg->path_weight) && std::isfinite(cfg->vel_zero_weight);
}

int check_batch(long long B, int ns_uniform, const long long *seg_offset, const double *waypoints) {
    if (B < 0 || !waypoints) return MSNAP_ERR_INVALID_ARG;
    if (ns_uniform <= 0 && !seg_offset) return MSNAP_ERR_INVALID_ARG;
    return MSNAP_OK;
}

template <int O>
int solve_qp_dev(msnap_context *h, double pw, double vw, long long B, int ns_uniform, const long long *seg_offset,
                 long long n_seg, const double *wp, const double *vel, const double *acc, const double *times,
                 double *coeff_out, double *max_dev_out, int *best_s_out, unsigned *flags_out) {
    BatchIdx bi{B, n_seg, ns_uniform > 0 ? ns_uniform : 0, ns_uniform > 0 ? nullptr : seg_offset};
    SolveParams sp{};
    sp.pw = pw;
    sp.vw0 = vw;
    sp.max_iter = 0;
    sp.vel = vel;
    sp.acc = acc;
    const FusedPlan f = plan_fused<O>(h, bi, sp);
    int rc = arena_reserve(h, h->ws, solve_ws_bytes<O>(n_seg, B, false, false, f, false));
    if (rc) return rc;
    SolveWs w;
    carve_solve_ws<O>(h->ws, n_seg, B, false, coeff_out, f, false, w);
    SolveIO io;
    io.wp = wp;
    io.times_in = times;
    io.coeff_out = coeff_out;
    io.max_dev_out = max_dev_out;
    io.best_s_out = best_s_out;
    io.flags_out = flags_out;
    return run_solve<O>(h, bi, sp, io, w);
}

inline unsigned geo_grid(const msnap_context *h, long long n) {  // grid-stride over warps of 32 rows
    const long long want = (n + GEO_BLOCK - 1) / GEO_BLOCK, cap = (long long)h->sm_count * 8;
    return (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
}

int launch_wgs84_to_enu(msnap_context *h, const GeoFrame &f, long long n, const double *lla, double *enu) {
    if (h->geo_trig)
        MS_LAUNCH(h, k_wgs84_to_enu<true>, geo_grid(h, n), GEO_BLOCK, f, n, lla, enu);
    else
        MS_LAUNCH(h, k_wgs84_to_enu<false>, geo_grid(h, n), GEO_BLOCK, f, n, lla, enu);
    return MSNAP_OK;
}

int launch_enu_to_wgs84(msnap_context *h, const GeoFrame &f, long long n_cap, const long long *n_dev, const double *enu,
                        double *lla, int *steps) {
    if (h->geo_trig)
        MS_LAUNCH(h, k_enu_to_wgs84<true>, geo_grid(h, n_cap), GEO_BLOCK, f, n_cap, n_dev, enu, lla, steps);
    else
        MS_LAUNCH(h, k_enu_to_wgs84<false>, geo_grid(h, n_cap), GEO_BLOCK, f, n_cap, n_dev, enu, lla, steps);
    return MSNAP_OK;
}

template <int O>
int generate_dev(msnap_context *h, const msnap_config *cfg, double sd, double v_avg, long long B, int ns_uniform,
                 const long long *seg_offset, long long n_seg, const double *wp, double *times_out, double *coeff_out,
                 double *max_dev_out, int *iters_out, double *vw_final_out, int *best_s_out, long long capacity,
                 long long *sample_offset, double *samples, double *stats, unsigned *flags,
                 double *coeff_mirror = nullptr, bool *mirror_done = nullptr) {
    BatchIdx bi{B, n_seg, ns_uniform > 0 ? ns_uniform : 0, ns_uniform > 0 ? nullptr : seg_offset};
    SolveParams sp{};
    sp.pw = cfg->path_weight;
    sp.vw0 = cfg->vel_zero_weight;
    sp.max_iter = 10;
    for (int a = 0; a < 3; ++a) {
        sp.bc[a] = cfg->start_vel[a];
        sp.bc[3 + a] = cfg->end_vel[a];
        sp.bc[6 + a] = cfg->start_acc[a];
        sp.bc[9 + a] = cfg->end_acc[a];
    }
    const FusedPlan f = plan_fused<O>(h, bi, sp);
    const bool spec = use_generic_spec<O>(h, n_seg, sp);
    const bool ragged = bi.ns_uniform <= 0 && B < 2000000000LL;
    const size_t n_pts = (size_t)(n_seg + B);
    int rc = arena_reserve(h, h->ws, solve_ws_bytes<O>(n_seg, B, true, coeff_out == nullptr, f, spec) +
                                        sample_ws_bytes(n_seg, B, bi.ns_uniform, h->policy, h->sm_count, capacity) +
                                        (ragged ? padded(n_seg * sizeof(int)) : 0) +
                                        (h->wp_frame ? padded(n_pts * 3 * sizeof(double)) : 0));
    if (rc) return rc;
    if (h->wp_frame) {  // the waypoints arrive as WGS84 rows: wgs84ToENU_Batch (cpp:2640) into the workspace first
        double *enu = arena_take<double>(h->ws, n_pts * 3);
        rc = launch_wgs84_to_enu(h, h->wp_geo, (long long)n_pts, wp, enu);
        if (rc) return rc;
        wp = enu;
    }
    if (ragged) {  // segment -> trajectory map, so that the per-segment kernels need no binary search
        int *st = arena_take<int>(h->ws, n_seg);
        MS_LAUNCH(h, k_seg_traj, grid_for(B * 32, 256), 256, B, seg_offset, st);
        bi.seg_traj = st;
    }
    SolveWs w;
    carve_solve_ws<O>(h->ws, n_seg, B, true, coeff_out, f, spec, w);
    SampleWs s;
    carve_sample_ws(h->ws, n_seg, B, bi.ns_uniform, h->policy, h->sm_count, capacity, s);
    SolveIO io;
    io.wp = wp;
    io.v_avg = v_avg;
    io.min_time = cfg->min_time_s;
    io.times_out = times_out;
    io.coeff_out = coeff_out;
    io.max_dev_out = max_dev_out;
    io.iters_out = iters_out;
    io.vw_final_out = vw_final_out;
    io.best_s_out = best_s_out;
    io.flags_out = flags;
    io.coeff_mirror = coeff_mirror;
    io.mirror_done = mirror_done;
    rc = run_solve<O>(h, bi, sp, io, w);
    if (rc) return rc;
    if (h->mark_solved) MS_CUDA(h, cudaEventRecord(h->ev_solved, h->stream));
    rc = run_sample<O>(h, bi, w.coeff, w.T, sd, capacity, sample_offset, samples, stats, flags, s);
    if (rc || h->frame == 0 || capacity <= 0) return rc;
    // the rows leave as WGS84: in place, after the statistics (which are defined on the ENU rows, ms.cpp:163-195); the
    // row count stays on the device (sample_offset[B])
    return launch_enu_to_wgs84(h, h->geo, capacity, sample_offset + B, samples, samples, nullptr);
}

#define MS_DISPATCH_ORDER(order, CALL)            \
    switch (order) {                              \
        case 2: { constexpr int O = 2; CALL; } break; \
        case 3: { constexpr int O = 3; CALL; } break; \
        case 4: { constexpr int O = 4; CALL; } break; \
        case 5: { constexpr int O = 5; CALL; } break; \
        default: return MSNAP_ERR_INVALID_ARG;    \
    }

// total segments of a batch whose seg_offset lives on the host
long long host_total_segments(long long B, int ns_uniform, const long long *seg_offset) {
    return ns_uniform > 0 ? B * (long long)ns_uniform : seg_offset[B];
}

// total segments of a batch whose seg_offset lives on the device (one 8-byte read-back for ragged batches)
int device_total_segments(msnap_context *h, long long B, int ns_uniform, const long long *seg_offset, long long *out) {
    if (ns_uniform > 0) {
        *out = B * (long long)ns_uniform;
        return MSNAP_OK;
    }
    MS_CUDA(h, cudaMemcpyAsync(out, seg_offset + B, sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    MS_CUDA(h, cudaStreamSynchronize(h->stream));
    return *out >= B ? MSNAP_OK : MSNAP_ERR_INVALID_ARG;
}

bool valid_host_offsets(long long B, int ns_uniform, const long long *seg_offset) {
    if (ns_uniform > 0) return true;
    if (seg_offset[0] != 0) return false;
    for (long long b = 0; b < B; ++b)
        if (seg_offset[b + 1] - seg_offset[b] < 1) return false;
    return true;
}

// FP64 peak micro-benchmark: 8 independent DFMA chains per thread.
__global__ void k_dfma_peak(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456) out[0] = s;
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------
// extern "C"
// ------------------------------------------------------------------------------------------------------------
extern "C" {

int msnap_version(void) { return MSNAP_VERSION; }

const char *msnap_status_string(int status) {
    switch (status) {
        case MSNAP_OK: return "ok";
        case MSNAP_ERR_INVALID_ARG: return "invalid argument";
        case MSNAP_ERR_CUDA: return "CUDA error";
        case MSNAP_ERR_NO_DEVICE: return "no usable CUDA device";
        case MSNAP_ERR_CAPACITY: return "sample buffer too small";
        case MSNAP_ERR_ALLOC: return "allocation failed";
        case MSNAP_ERR_IO: return "I/O error";
        default: return "unknown status";
    }
}

const char *msnap_last_error(msnap_handle h) { return h ? h->last_error.c_str() : ""; }

int msnap_create(int device, msnap_handle *out) {
    if (!out) return MSNAP_ERR_INVALID_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) {
        cudaGetLastError();
        return MSNAP_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
        cudaGetLastError();
        return MSNAP_ERR_NO_DEVICE;  // the kernels are built for sm_100a only
    }
    msnap_context *h = new (std::nothrow) msnap_context();
    if (!h) return MSNAP_ERR_ALLOC;
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    DeviceGuard guard(device);
    // candidate times exactly as ms.cpp:140 accumulates them: t_1 = 0.1, t_{i+1} = fl(t_i + 0.1)
    std::vector<double> ttab(T_TABLE_N);
    ttab[0] = 0.0;
    {
        volatile double t = 0.1;
        for (int i = 1; i < T_TABLE_N; ++i) {
            ttab[i] = t;
            t = t + 0.1;
        }
    }
    if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&h->d_ttab, T_TABLE_N * sizeof(double)) != cudaSuccess ||
        cudaMemcpy(h->d_ttab, ttab.data(), T_TABLE_N * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMalloc(&h->d_tab, sizeof(h_tab)) != cudaSuccess ||
        cudaMemcpy(h->d_tab, h_tab, sizeof(h_tab), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaGetLastError();
        if (h->d_tab) cudaFree(h->d_tab);
        if (h->d_ttab) cudaFree(h->d_ttab);
        if (h->own_stream) cudaStreamDestroy(h->own_stream);
        delete h;
        return MSNAP_ERR_CUDA;
    }
    h->stream = h->own_stream;
    if (cudaStreamCreateWithFlags(&h->aux, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_solved, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        msnap_destroy(h);
        return MSNAP_ERR_CUDA;
    }
    if (const char *e = std::getenv("MSNAP_SCAN_COEF_SMEM")) h->scan_coef_smem = std::atoi(e) != 0;
    if (const char *e = std::getenv("MSNAP_HOST_CHUNKS")) h->host_chunks = std::atoi(e);
    if (const char *e = std::getenv("MSNAP_ZERO_COPY")) h->zero_copy = std::atoi(e) != 0;
    if (const char *e = std::getenv("MSNAP_DISCARD_STATE")) h->discard_state = std::atoi(e) != 0;
    *out = h;
    return MSNAP_OK;
}

int msnap_destroy(msnap_handle h) {
    if (!h) return MSNAP_ERR_INVALID_ARG;
    DeviceGuard guard(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->ws.base) cudaFree(h->ws.base);
    if (h->io.base) cudaFree(h->io.base);
    if (h->d_tab) cudaFree(h->d_tab);
    if (h->d_ttab) cudaFree(h->d_ttab);
    if (h->phase_clocks) cudaFree(h->phase_clocks);
    if (h->h_off) cudaFreeHost(h->h_off);
    if (h->aux) {
        cudaStreamSynchronize(h->aux);
        cudaStreamDestroy(h->aux);
    }
    if (h->ev_solved) cudaEventDestroy(h->ev_solved);
    for (msnap_context *k : h->kids) msnap_destroy(k);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return MSNAP_OK;
}

int msnap_set_stream(msnap_handle h, void *cuda_stream) {
    if (!h) return MSNAP_ERR_INVALID_ARG;
    cudaStream_t next = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->own_stream;
    if (next != h->stream) {
        // The handle has ONE workspace arena, reset by every call: work still queued on the old stream must not be
        // overtaken by calls on the new one.  The new stream waits for an event recorded behind everything enqueued so far
        // (no host synchronisation).
        DeviceGuard guard(h->device);
        cudaEvent_t ev = nullptr;
        MS_CUDA(h, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        cudaError_t e = cudaEventRecord(ev, h->stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(next, ev, 0);
        cudaEventDestroy(ev);
        if (e != cudaSuccess) {
            h->last_error = std::string("msnap_set_stream: ") + cudaGetErrorString(e);
            return MSNAP_ERR_CUDA;
        }
    }
    h->stream = next;
    return MSNAP_OK;
}

int msnap_synchronize(msnap_handle h) {
    if (!h) return MSNAP_ERR_INVALID_ARG;
    DeviceGuard guard(h->device);
    MS_CUDA(h, cudaStreamSynchronize(h->stream));
    return MSNAP_OK;
}

int msnap_set_reweight_policy(msnap_handle h, int policy) {
    if (!h || policy < 0 || policy > 2) return MSNAP_ERR_INVALID_ARG;
    h->policy = policy;
    return MSNAP_OK;
}

int msnap_set_zero_copy(msnap_handle h, int enable) {
    if (!h) return MSNAP_ERR_INVALID_ARG;
    h->zero_copy = enable != 0;
    return MSNAP_OK;
}

int msnap_set_host_chunks(msnap_handle h, int n_chunks) {
    if (!h || n_chunks < 0) return MSNAP_ERR_INVALID_ARG;
    h->host_chunks = n_chunks;
    return MSNAP_OK;
}

long long msnap_launch_count(msnap_handle h) {
    if (!h) return 0;
    long long n = h->launches;
    for (msnap_context *k : h->kids) n += k->launches;
    return n;
}

void msnap_config_default(msnap_config *cfg) {
    if (!cfg) return;
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->order = 3;
    cfg->V_avg = 5.0;
    cfg->min_time_s = 0.1;
    cfg->sample_distance = 1.0;
}

// --- YAML subset: `key: value` lines, optional `minimum_snap:` wrapper, flow or block sequences for the vec3s ---
static bool parse_double_strict(const std::string &s, double &out) {
    if (s.empty()) return false;
    char *end = nullptr;
    const double v = std::strtod(s.c_str(), &end);
    while (end && *end == ' ') ++end;
    if (!end || *end != '\0') return false;
    out = v;
    return true;
}
static bool parse_int_strict(const std::string &s, int &out) {
    if (s.empty()) return false;
    char *end = nullptr;
    const long v = std::strtol(s.c_str(), &end, 10);
    while (end && *end == ' ') ++end;
    if (!end || *end != '\0') return false;  // "2.0" is not an int for yaml-cpp either
    out = (int)v;
    return true;
}
static std::string trim(const std::string &s) {
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

int msnap_config_load_yaml(const char *path, msnap_config *cfg) {
    if (!path || !cfg) return MSNAP_ERR_INVALID_ARG;
    std::ifstream f(path);
    if (!f) return MSNAP_ERR_IO;
    struct Line { int indent; std::string key, val; std::vector<std::string> items; };
    std::vector<Line> lines;
    std::string raw;
    while (std::getline(f, raw)) {
        bool in_q = false;
        size_t cut = std::string::npos;
        for (size_t i = 0; i < raw.size(); ++i) {
            if (raw[i] == '"' || raw[i] == '\'') in_q = !in_q;
            if (raw[i] == '#' && !in_q && (i == 0 || raw[i - 1] == ' ' || raw[i - 1] == '\t')) { cut = i; break; }
        }
        if (cut != std::string::npos) raw = raw.substr(0, cut);
        if (trim(raw).empty()) continue;
        const int indent = (int)raw.find_first_not_of(" \t");
        const std::string body = trim(raw);
        if (body[0] == '-') {  // block sequence item of the previous key
            if (!lines.empty()) lines.back().items.push_back(trim(body.substr(1)));
            continue;
        }
        const size_t colon = body.find(':');
        if (colon == std::string::npos) continue;
        Line L;
        L.indent = indent;
        L.key = trim(body.substr(0, colon));
        L.val = trim(body.substr(colon + 1));
        if (!L.val.empty() && L.val.front() == '[' && L.val.back() == ']') {
            std::stringstream ss(L.val.substr(1, L.val.size() - 2));
            std::string item;
            while (std::getline(ss, item, ',')) L.items.push_back(trim(item));
        }
        lines.push_back(L);
    }
    // wrapper style: use the keys nested under a top-level `minimum_snap:`
    size_t lo = 0, hi = lines.size();
    int want_indent = lines.empty() ? 0 : lines[0].indent;
    for (size_t i = 0; i < lines.size(); ++i)
        if (lines[i].key == "minimum_snap" && lines[i].val.empty() && lines[i].indent == want_indent) {
            lo = i + 1;
            hi = lo;
            while (hi < lines.size() && lines[hi].indent > lines[i].indent) ++hi;
            want_indent = lo < lines.size() ? lines[lo].indent : 0;
            break;
        }
    auto scalar = [&](const char *key, double &out) {
        for (size_t i = lo; i < hi; ++i)
            if (lines[i].indent == want_indent && lines[i].key == key) {
                double v;
                if (parse_double_strict(lines[i].val, v)) out = v;
            }
    };
    auto vec3 = [&](const char *key, double *out) {
        for (size_t i = lo; i < hi; ++i)
            if (lines[i].indent == want_indent && lines[i].key == key && lines[i].items.size() >= 3) {
                double v[3];
                if (parse_double_strict(lines[i].items[0], v[0]) && parse_double_strict(lines[i].items[1], v[1]) &&
                    parse_double_strict(lines[i].items[2], v[2])) {
                    out[0] = v[0]; out[1] = v[1]; out[2] = v[2];
                }
            }
    };
    for (size_t i = lo; i < hi; ++i)
        if (lines[i].indent == want_indent && lines[i].key == "order") {
            int v;
            if (parse_int_strict(lines[i].val, v)) cfg->order = v;
        }
    scalar("path_weight", cfg->path_weight);
    scalar("vel_zero_weight", cfg->vel_zero_weight);
    scalar("V_avg", cfg->V_avg);
    scalar("min_time_s", cfg->min_time_s);
    scalar("sample_distance", cfg->sample_distance);
    vec3("start_vel", cfg->start_vel);
    vec3("end_vel", cfg->end_vel);
    vec3("start_acc", cfg->start_acc);
    vec3("end_acc", cfg->end_acc);
    return MSNAP_OK;
}

// ---------------------------------------------------------------------------------------------- solve_qp
int msnap_solve_qp_batch_dev(msnap_handle h, int order, double path_weight, double vel_zero_weight, long long B,
                             int ns_uniform, const long long *seg_offset, const double *waypoints, const double *vel,
                             const double *acc, const double *times, double *coeff_out, double *max_dev_out,
                             int *best_s_out, unsigned *flags_out) {
    if (!h) return MSNAP_ERR_INVALID_ARG;
    int rc = check_batch(B, ns_uniform, seg_offset, waypoints);
    if (rc) return rc;
    if (!times || !coeff_out) return MSNAP_ERR_INVALID_ARG;
    if (reinterpret_cast<uintptr_t>(coeff_out) & 15) return MSNAP_ERR_INVALID_ARG;  // 128-bit stores
    if (B == 0) return MSNAP_OK;
    DeviceGuard guard(h->device);
    long long n_seg = 0;
    rc = device_total_segments(h, B, ns_uniform, seg_offset, &n_seg);
    if (rc) return rc;
    MS_DISPATCH_ORDER(order, return solve_qp_dev<O>(h, path_weight, vel_zero_weight, B, ns_uniform, seg_offset, n_seg,
                                                     waypoints, vel, acc, times, coeff_out, max_dev_out, best_s_out,
                                                     flags_out));
    return MSNAP_OK;
}

int msnap_solve_qp_batch_host(msnap_handle h, int order, double path_weight, double vel_zero_weight, long long B,
                              int ns_uniform, const long long *seg_offset, const double *waypoints, const double *vel,
                              const double *acc, const double *times, double *coeff_out, double *max_dev_out,
                              int *best_s_out, unsigned *flags_out) {
    if (!h) return MSNAP_ERR_INVALID_ARG;
    int rc = check_batch(B, ns_uniform, seg_offset, waypoints);
    if (rc) return rc;
    if (!times || !coeff_out || order < MSNAP_MIN_ORDER || order > MSNAP_MAX_ORDER) return MSNAP_ERR_INVALID_ARG;
    if (B == 0) return MSNAP_OK;
    if (!valid_host_offsets(B, ns_uniform, seg_offset)) return MSNAP_ERR_INVALID_ARG;
    DeviceGuard guard(h->device);
    const long long n_seg = host_total_segments(B, ns_uniform, seg_offset);
    const size_t n_pts = (size_t)(n_seg + B), m3 = (size_t)3 * 2 * order;
    size_t bytes = padded((B + 1) * sizeof(long long)) + padded(n_pts * 3 * sizeof(double)) +
                   2 * padded((size_t)B * 6 * sizeof(double)) + padded(n_seg * sizeof(double)) +
                   padded((size_t)n_seg * m3 * sizeof(double)) + padded(B * sizeof(double)) +
                   padded(B * sizeof(unsigned)) + padded(n_seg * sizeof(int));
    rc = arena_reserve(h, h->io, bytes);
    if (rc) return rc;
    long long *d_off = arena_take<long long>(h->io, B + 1);
    double *d_wp = arena_take<double>(h->io, n_pts * 3);
    double *d_vel = arena_take<double>(h->io, (size_t)B * 6);
    double *d_acc = arena_take<double>(h->io, (size_t)B * 6);
    double *d_t = arena_take<double>(h->io, n_seg);
    double *d_c = arena_take<double>(h->io, (size_t)n_seg * m3);
    double *d_md = arena_take<double>(h->io, B);
    unsigned *d_fl = arena_take<unsigned>(h->io, B);
    int *d_bs = arena_take<int>(h->io, n_seg);
    cudaStream_t st = h->stream;
    if (ns_uniform <= 0)
        MS_CUDA(h, cudaMemcpyAsync(d_off, seg_offset, (B + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
    MS_CUDA(h, cudaMemcpyAsync(d_wp, waypoints, n_pts * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    if (vel) MS_CUDA(h, cudaMemcpyAsync(d_vel, vel, (size_t)B * 6 * sizeof(double), cudaMemcpyHostToDevice, st));
    if (acc) MS_CUDA(h, cudaMemcpyAsync(d_acc, acc, (size_t)B * 6 * sizeof(double), cudaMemcpyHostToDevice, st));
    MS_CUDA(h, cudaMemcpyAsync(d_t, times, n_seg * sizeof(double), cudaMemcpyHostToDevice, st));
    MS_DISPATCH_ORDER(order, rc = solve_qp_dev<O>(h, path_weight, vel_zero_weight, B, ns_uniform,
                                                  ns_uniform > 0 ? nullptr : d_off, n_seg, d_wp, vel ? d_vel : nullptr,
                                                  acc ? d_acc : nullptr, d_t, d_c, d_md, best_s_out ? d_bs : nullptr, d_fl));
    if (rc) return rc;
    MS_CUDA(h, cudaMemcpyAsync(coeff_out, d_c, (size_t)n_seg * m3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (max_dev_out) MS_CUDA(h, cudaMemcpyAsync(max_dev_out, d_md, B * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (flags_out) MS_CUDA(h, cudaMemcpyAsync(flags_out, d_fl, B * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    if (best_s_out) MS_CUDA(h, cudaMemcpyAsync(best_s_out, d_bs, n_seg * sizeof(int), cudaMemcpyDeviceToHost, st));
    MS_CUDA(h, cudaStreamSynchronize(st));
    return MSNAP_OK;
}

// ---------------------------------------------------------------------------------------------- generate
int msnap_generate_batch_dev(msnap_handle h, const msnap_config *cfg, double sample_distance_override,
                             double v_avg_override, long long B, int ns_uniform, const long long *seg_offset,
                             const double *waypoints, double *times_out, double *coeff_out, double *max_dev_out,
                             int *iters_out, double *vw_final_out, int *best_s_out, long long sample_capacity,
                             long long *sample_offset_out, double *samples_out, double *stats_out,
                             unsigned *flags_out) {
    if (!h || !cfg) return MSNAP_ERR_INVALID_ARG;
    int rc = check_batch(B, ns_uniform, seg_offset, waypoints);
    if (rc) return rc;
    if (!sample_offset_out || (!samples_out && sample_capacity > 0) || sample_capacity < 0) return MSNAP_ERR_INVALID_ARG;
    if (reinterpret_cast<uintptr_t>(coeff_out) & 15) return MSNAP_ERR_INVALID_ARG;  // 128-bit stores
    if (B == 0) return MSNAP_OK;
    DeviceGuard guard(h->device);
    long long n_seg = 0;
    rc = device_total_segments(h, B, ns_uniform, seg_offset, &n_seg);
    if (rc) return rc;
    const double sd = sample_distance_override > 0.0 ? sample_distance_override : cfg->sample_distance;  // ms.cpp:42-44
    const double va = v_avg_override > 0.0 ? v_avg_override : cfg->V_avg;                                // ms.cpp:46-48
    if (!valid_generate_config(cfg, sd, va)) return MSNAP_ERR_INVALID_ARG;
    MS_DISPATCH_ORDER(cfg->order,
                      return generate_dev<O>(h, cfg, sd, va, B, ns_uniform, seg_offset, n_seg, waypoints, times_out,
                                             coeff_out, max_dev_out, iters_out, vw_final_out, best_s_out, sample_capacity,
                                             sample_offset_out, samples_out, stats_out, flags_out));
    return MSNAP_OK;
}

// One chunk of a host-pointer generate call: trajectories [b0, b0 + B) of the caller's batch on context k.
struct HostChunk {
    long long B = 0, n_seg = 0, b0 = 0, g0 = 0, p0 = 0;  // sizes; first trajectory / segment / waypoint of the chunk
    std::vector<long long> off_local;                       // ragged batches: seg_offset of the chunk, rebased to 0
    long long *d_so = nullptr;
    double *d_s = nullptr;
    bool single = false;               // the only chunk of the call
    long long capacity = 0;            // rows of the chunk's device mirror: min(caller's capacity, the chunk's own row bound)
    double *samples_direct = nullptr;  // zero-copy: the kernels wrote the rows straight into the caller's pinned buffer
};

// Enqueue H2D, the kernels and the D2H of everything but the samples on k's stream; no host synchronisation.
// Device-visible address of a pinned (page-locked, mapped) host allocation, or nullptr for anything else.
static void *device_view_of_pinned(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

// Upper bound of the rows a chunk can produce, from its waypoints on the host: k_bound's formula (candidates from the
// segment duration alone, + first and end point per trajectory) with one spare candidate per segment, so that a last-bit
// difference between the host's and the device's duration cannot make it too small.
static long long host_rows_bound(const msnap_config *cfg, double va, const double *wp, long long p0, long long B,
                                 int ns_uniform, const std::vector<long long> &off_local) {
    unsigned long long n = 0;
    for (long long b = 0; b < B; ++b) {
        const long long g0 = ns_uniform > 0 ? b * ns_uniform : off_local[(size_t)b];
        const long long g1 = ns_uniform > 0 ? g0 + ns_uniform : off_local[(size_t)b + 1];
        const double *p = wp + 3 * (p0 + g0 + b);
        n += 2;
        for (long long g = g0; g < g1; ++g, p += 3) {
            const double dx = p[3] - p[0], dy = p[4] - p[1], dz = p[5] - p[2];
            double t = va > 1e-6 ? std::sqrt(dx * dx + dy * dy + dz * dz) / va : cfg->min_time_s;
            if (t < cfg->min_time_s) t = cfg->min_time_s;
            if (!(t > 0.0 && t <= SAMPLE_T_MAX)) continue;  // a failed segment has no candidates (sample_time_ok)
            n += t >= 1.0 ? (unsigned long long)(t / 0.1) + 3ull : 13ull;
        }
    }
    return (long long)n;
}

static int host_chunk_enqueue(msnap_context *k, const msnap_config *cfg, double sd, double va, HostChunk &j,
                              int ns_uniform, const double *waypoints, double *times_out, double *coeff_out,
                              double *max_dev_out, int *iters_out, double *vw_final_out, int *best_s_out,
                              long long sample_capacity, double *samples_out, double *stats_out, unsigned *flags_out) {
    const long long B = j.B, n_seg = j.n_seg;
    const size_t n_pts = (size_t)(n_seg + B), m3 = (size_t)3 * 2 * cfg->order;
    size_t bytes = padded((B + 1) * sizeof(long long)) + padded(n_pts * 3 * sizeof(double)) +
                   padded(n_seg * sizeof(double)) + padded((size_t)n_seg * m3 * sizeof(double)) +
                   2 * padded(B * sizeof(double)) + padded(B * sizeof(int)) + padded((B + 1) * sizeof(long long)) +
                   padded((size_t)j.capacity * 3 * sizeof(double)) + padded((size_t)B * 2 * sizeof(double)) +
                   padded(B * sizeof(unsigned)) + padded(n_seg * sizeof(int));
    MS_CUDA(k, cudaStreamSynchronize(k->aux));  // copies of this context's previous chunk still read its device buffers
    int rc = arena_reserve(k, k->io, bytes);
    if (rc) return rc;
    if (k->h_off_cap < (size_t)(B + 1)) {
        if (k->h_off) cudaFreeHost(k->h_off);
        k->h_off = nullptr;
        k->h_off_cap = 0;
        if (cudaMallocHost(&k->h_off, (size_t)(B + 1) * sizeof(long long)) != cudaSuccess) {
            cudaGetLastError();
            return MSNAP_ERR_ALLOC;
        }
        k->h_off_cap = (size_t)(B + 1);
    }
    long long *d_off = arena_take<long long>(k->io, B + 1);
    double *d_wp = arena_take<double>(k->io, n_pts * 3);
    double *d_t = arena_take<double>(k->io, n_seg);
    double *d_c = arena_take<double>(k->io, (size_t)n_seg * m3);
    double *d_md = arena_take<double>(k->io, B);
    double *d_vw = arena_take<double>(k->io, B);
    int *d_it = arena_take<int>(k->io, B);
    j.d_so = arena_take<long long>(k->io, B + 1);
    j.d_s = arena_take<double>(k->io, (size_t)j.capacity * 3);
    double *d_st = arena_take<double>(k->io, (size_t)B * 2);
    unsigned *d_fl = arena_take<unsigned>(k->io, B);
    int *d_bs = arena_take<int>(k->io, n_seg);
    cudaStream_t st = k->stream;
    if (ns_uniform <= 0)
        MS_CUDA(k, cudaMemcpyAsync(d_off, j.off_local.data(), (B + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
    MS_CUDA(k, cudaMemcpyAsync(d_wp, waypoints + 3 * j.p0, n_pts * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    // Zero-copy outputs: when the caller's big result buffers are pinned (device-visible) host memory the kernels store
    // straight into them over PCIe -- coalesced 64-byte coefficient rows and 768-byte sample row groups -- so the
    // transfer overlaps the computation and no separate device-to-host copy (nor a round trip for the row count) is
    // needed.  Samples only without statistics (k_stats would read them back over PCIe) and for a single chunk.
    double *coeff_zc = nullptr;
    bool coeff_direct = false;
    if (k->zero_copy && coeff_out && (reinterpret_cast<uintptr_t>(coeff_out) & 15) == 0)
        coeff_zc = static_cast<double *>(device_view_of_pinned(coeff_out));
    j.samples_direct = nullptr;
    if (k->zero_copy && j.single && !stats_out && samples_out && k->frame == 0)
        j.samples_direct = static_cast<double *>(device_view_of_pinned(samples_out));
    k->mark_solved = true;  // ev_solved: the solve's outputs are final, the sampler has not started yet
    MS_DISPATCH_ORDER(cfg->order,
                      rc = generate_dev<O>(k, cfg, sd, va, B, ns_uniform, ns_uniform > 0 ? nullptr : d_off, n_seg, d_wp,
                                           times_out ? d_t : nullptr, coeff_out ? d_c : nullptr, d_md, d_it, d_vw,
                                           best_s_out ? d_bs : nullptr, j.capacity, j.d_so,
                                           j.samples_direct ? j.samples_direct : j.d_s, stats_out ? d_st : nullptr, d_fl,
                                           coeff_zc ? coeff_zc + j.g0 * m3 : nullptr, &coeff_direct));
    k->mark_solved = false;
    if (rc) return rc;
    // the solve's results leave on the auxiliary stream while the sampler is still running on the main one
    cudaStream_t ax = k->aux;
    MS_CUDA(k, cudaStreamWaitEvent(ax, k->ev_solved, 0));
    if (coeff_out && !coeff_direct)
        MS_CUDA(k, cudaMemcpyAsync(coeff_out + j.g0 * m3, d_c, (size_t)n_seg * m3 * sizeof(double), cudaMemcpyDeviceToHost, ax));
    if (times_out)
        MS_CUDA(k, cudaMemcpyAsync(times_out + j.g0, d_t, n_seg * sizeof(double), cudaMemcpyDeviceToHost, ax));
    if (max_dev_out) MS_CUDA(k, cudaMemcpyAsync(max_dev_out + j.b0, d_md, B * sizeof(double), cudaMemcpyDeviceToHost, ax));
    if (vw_final_out) MS_CUDA(k, cudaMemcpyAsync(vw_final_out + j.b0, d_vw, B * sizeof(double), cudaMemcpyDeviceToHost, ax));
    if (iters_out) MS_CUDA(k, cudaMemcpyAsync(iters_out + j.b0, d_it, B * sizeof(int), cudaMemcpyDeviceToHost, ax));
    if (best_s_out) MS_CUDA(k, cudaMemcpyAsync(best_s_out + j.g0, d_bs, n_seg * sizeof(int), cudaMemcpyDeviceToHost, ax));
    MS_CUDA(k, cudaMemcpyAsync(k->h_off, j.d_so, (B + 1) * sizeof(long long), cudaMemcpyDeviceToHost, st));
    if (stats_out)
        MS_CUDA(k, cudaMemcpyAsync(stats_out + 2 * j.b0, d_st, (size_t)B * 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (flags_out) MS_CUDA(k, cudaMemcpyAsync(flags_out + j.b0, d_fl, B * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    return MSNAP_OK;
}

// Wait for the chunk's kernels, place its rows behind those of the previous chunks (rows_base) and start their copy.
static int host_chunk_finish(msnap_context *k, HostChunk &j, long long &rows_base, long long sample_capacity,
                             long long *sample_offset_out, double *samples_out, unsigned *flags_out) {
    MS_CUDA(k, cudaStreamSynchronize(k->stream));  // h_off[B] = exact row count of the chunk
    const long long total = k->h_off[j.B];
    long long room = sample_capacity - rows_base;
    if (room < 0) room = 0;
    long long rows = total < room ? total : room;
    if (rows > j.capacity) rows = j.capacity;  // (the mirror never holds more; the kernels flagged what they dropped)
    if (rows > 0 && !j.samples_direct)
        MS_CUDA(k, cudaMemcpyAsync(samples_out + 3 * rows_base, j.d_s, (size_t)rows * 3 * sizeof(double),
                                   cudaMemcpyDeviceToHost, k->stream));
    for (long long b = 0; b < j.B; ++b) sample_offset_out[j.b0 + b] = rows_base + k->h_off[b];
    if (total > room && flags_out)  // trajectories whose rows do not fit the caller's buffer any more
        for (long long b = 0; b < j.B; ++b)
            if (rows_base + k->h_off[b + 1] > sample_capacity) flags_out[j.b0 + b] |= MSNAP_FLAG_TRUNCATED;
    rows_base += total;
    return MSNAP_OK;
}

int msnap_generate_batch_host(msnap_handle h, const msnap_config *cfg, double sample_distance_override,
                              double v_avg_override, long long B, int ns_uniform, const long long *seg_offset,
                              const double *waypoints, double *times_out, double *coeff_out, double *max_dev_out,
                              int *iters_out, double *vw_final_out, int *best_s_out, long long sample_capacity,
                              long long *sample_offset_out, double *samples_out, double *stats_out,
                              unsigned *flags_out) {
    if (!h || !cfg) return MSNAP_ERR_INVALID_ARG;
    int rc = check_batch(B, ns_uniform, seg_offset, waypoints);
    if (rc) return rc;
    if (!sample_offset_out || (!samples_out && sample_capacity > 0) || sample_capacity < 0 ||
        cfg->order < MSNAP_MIN_ORDER || cfg->order > MSNAP_MAX_ORDER)
        return MSNAP_ERR_INVALID_ARG;
    if (B == 0) { sample_offset_out[0] = 0; return MSNAP_OK; }
    if (!valid_host_offsets(B, ns_uniform, seg_offset)) return MSNAP_ERR_INVALID_ARG;
    DeviceGuard guard(h->device);
    const double sd = sample_distance_override > 0.0 ? sample_distance_override : cfg->sample_distance;
    const double va = v_avg_override > 0.0 ? v_avg_override : cfg->V_avg;
    if (!valid_generate_config(cfg, sd, va)) return MSNAP_ERR_INVALID_ARG;
    // Chunks: results cross PCIe at ~50 GB/s while the kernels take a fraction of that time, so a big batch is cut into
    // chunks whose kernels run (on a second stream) while the previous chunk's results are still being copied out.
    // (Measured on B200/PCIe 5: at 4 096 trajectories the extra API calls of chunking cost more than the overlap wins;
    // from ~16 k trajectories on it gains 5-6 %.  The copies themselves run at ~53 GB/s either way.)
    long long n_chunks = h->host_chunks > 0 ? h->host_chunks : B / 8192;
    if (n_chunks > 8) n_chunks = 8;
    if (n_chunks < 1) n_chunks = 1;
    if (n_chunks > B) n_chunks = B;
    std::vector<msnap_context *> ctx(1, h);
    if (n_chunks > 1) {
        while (h->kids.size() < 2) {
            msnap_handle k = nullptr;
            rc = msnap_create(h->device, &k);
            if (rc) return rc;
            k->policy = h->policy;
            h->kids.push_back(k);
        }
        for (msnap_context *k : h->kids) {
            k->policy = h->policy;
            k->scan_coef_smem = h->scan_coef_smem;
            k->discard_state = h->discard_state;
            k->zero_copy = h->zero_copy;
            k->frame = h->frame;
            k->geo = h->geo;
            k->geo_trig = h->geo_trig;
            k->wp_frame = h->wp_frame;
            k->wp_geo = h->wp_geo;
        }
        ctx = h->kids;
    }
    std::vector<HostChunk> jobs((size_t)n_chunks);
    for (long long c = 0; c < n_chunks; ++c) {
        HostChunk &j = jobs[(size_t)c];
        const long long b0 = B * c / n_chunks, b1 = B * (c + 1) / n_chunks;
        j.b0 = b0;
        j.B = b1 - b0;
        if (ns_uniform > 0) {
            j.g0 = b0 * ns_uniform;
            j.n_seg = j.B * ns_uniform;
        } else {
            j.g0 = seg_offset[b0];
            j.n_seg = seg_offset[b1] - seg_offset[b0];
            j.off_local.resize((size_t)j.B + 1);
            for (long long b = 0; b <= j.B; ++b) j.off_local[(size_t)b] = seg_offset[b0 + b] - j.g0;
        }
        j.p0 = j.g0 + b0;
    }
    auto enqueue = [&](long long c) {
        HostChunk &jc = jobs[(size_t)c];
        jc.single = n_chunks == 1;
        // a chunk's device mirror holds the chunk's own rows, not the whole batch's (the caller's capacity bounds it too)
        jc.capacity = sample_capacity;
        if (n_chunks > 1) {
            const long long bound = host_rows_bound(cfg, va, waypoints, jc.p0, jc.B,
                                                    ns_uniform, jc.off_local);
            if (bound < jc.capacity) jc.capacity = bound;
        }
        return host_chunk_enqueue(ctx[(size_t)c % ctx.size()], cfg, sd, va, jobs[(size_t)c], ns_uniform, waypoints,
                                  times_out, coeff_out, max_dev_out, iters_out, vw_final_out, best_s_out,
                                  sample_capacity, samples_out, stats_out, flags_out);
    };
    auto fail = [&](int code, msnap_context *k) {
        if (k != h) h->last_error = k->last_error;
        for (msnap_context *q : ctx) {
            cudaStreamSynchronize(q->aux);
            cudaStreamSynchronize(q->stream);
        }
        return code;
    };
    long long rows_base = 0;
    rc = enqueue(0);
    if (rc) return fail(rc, ctx[0]);
    for (long long c = 0; c < n_chunks; ++c) {
        msnap_context *k = ctx[(size_t)c % ctx.size()];
        if (c + 1 < n_chunks) {
            // the next chunk's context is free: its previous chunk (c - 1) was finished in the last trip, and its
            // sample copy is ordered before the new work on the same stream
            rc = enqueue(c + 1);
            if (rc) return fail(rc, ctx[(size_t)(c + 1) % ctx.size()]);
        }
        rc = host_chunk_finish(k, jobs[(size_t)c], rows_base, sample_capacity, sample_offset_out, samples_out, flags_out);
        if (rc) return fail(rc, k);
    }
    sample_offset_out[B] = rows_base;
    for (msnap_context *q : ctx) {
        if (cudaStreamSynchronize(q->aux) != cudaSuccess || cudaStreamSynchronize(q->stream) != cudaSuccess) {
            h->last_error = "cudaStreamSynchronize(host chunk)";
            cudaGetLastError();
            return MSNAP_ERR_CUDA;
        }
    }
    return rows_base > sample_capacity ? MSNAP_ERR_CAPACITY : MSNAP_OK;
}

// ---------------------------------------------------------------------------------------------- WGS84 <-> ENU
static bool geo_reference_ok(const double *r) {
    return r && std::isfinite(r[0]) && std::isfinite(r[1]) && std::isfinite(r[2]);
}

int msnap_set_sample_frame(msnap_handle h, int frame, const double *reference_lla) {
    if (!h || frame < 0 || frame > 1 || (frame == 1 && !geo_reference_ok(reference_lla))) return MSNAP_ERR_INVALID_ARG;
    h->frame = frame;
    if (frame == 1) geo_make_frame(reference_lla, h->geo);
    return MSNAP_OK;
}

int msnap_set_waypoint_frame(msnap_handle h, int frame, const double *reference_lla) {
    if (!h || frame < 0 || frame > 1 || (frame == 1 && !geo_reference_ok(reference_lla))) return MSNAP_ERR_INVALID_ARG;
    h->wp_frame = frame;
    if (frame == 1) geo_make_frame(reference_lla, h->wp_geo);
    return MSNAP_OK;
}

int msnap_set_geo_exact_trig(msnap_handle h, int enable) {
    if (!h) return MSNAP_ERR_INVALID_ARG;
    h->geo_trig = enable != 0;
    return MSNAP_OK;
}

int msnap_enu_to_wgs84_dev(msnap_handle h, const double *reference_lla, long long n, const double *enu, double *lla_out) {
    if (!h || !geo_reference_ok(reference_lla) || n < 0 || (n > 0 && (!enu || !lla_out))) return MSNAP_ERR_INVALID_ARG;
    if (n == 0) return MSNAP_OK;
    DeviceGuard guard(h->device);
    GeoFrame f;
    geo_make_frame(reference_lla, f);
    return launch_enu_to_wgs84(h, f, n, nullptr, enu, lla_out, nullptr);
}

int msnap_enu_to_wgs84_counted_dev(msnap_handle h, const double *reference_lla, long long n_rows_cap,
                                   const long long *n_rows_dev, const double *enu, double *lla_out) {
    if (!h || !geo_reference_ok(reference_lla) || n_rows_cap < 0 || !n_rows_dev || (n_rows_cap > 0 && (!enu || !lla_out)))
        return MSNAP_ERR_INVALID_ARG;
    if (n_rows_cap == 0) return MSNAP_OK;
    DeviceGuard guard(h->device);
    GeoFrame f;
    geo_make_frame(reference_lla, f);
    return launch_enu_to_wgs84(h, f, n_rows_cap, n_rows_dev, enu, lla_out, nullptr);
}

int msnap_wgs84_to_enu_dev(msnap_handle h, const double *reference_lla, long long n, const double *lla, double *enu_out) {
    if (!h || !geo_reference_ok(reference_lla) || n < 0 || (n > 0 && (!lla || !enu_out))) return MSNAP_ERR_INVALID_ARG;
    if (n == 0) return MSNAP_OK;
    DeviceGuard guard(h->device);
    GeoFrame f;
    geo_make_frame(reference_lla, f);
    return launch_wgs84_to_enu(h, f, n, lla, enu_out);
}

// Host rows in, host rows out: the batch is cut into pieces that alternate between the handle's two streams, so that a
// piece's kernel and download overlap the next piece's upload (PCIe is full duplex).
static int geo_host(msnap_handle h, const double *reference_lla, long long n, const double *in, double *out, bool to_wgs) {
    if (!h || !geo_reference_ok(reference_lla) || n < 0 || (n > 0 && (!in || !out))) return MSNAP_ERR_INVALID_ARG;
    if (n == 0) return MSNAP_OK;
    DeviceGuard guard(h->device);
    GeoFrame f;
    geo_make_frame(reference_lla, f);
    MS_CUDA(h, cudaStreamSynchronize(h->aux));
    int rc = arena_reserve(h, h->io, padded((size_t)n * 3 * sizeof(double)));
    if (rc) return rc;
    double *d = arena_take<double>(h->io, (size_t)n * 3);
    const long long piece = 1 << 18;  // 262 144 rows = 6.3 MB per copy
    const long long n_pieces = (n + piece - 1) / piece;
    cudaStream_t own = h->stream, lanes[2] = {h->stream, n_pieces > 1 ? h->aux : h->stream};
    int err = MSNAP_OK;
    for (long long c = 0; c < n_pieces && !err; ++c) {
        const long long r0 = c * piece, m = (n - r0) < piece ? (n - r0) : piece;
        h->stream = lanes[c & 1];
        err = [&]() -> int {
            MS_CUDA(h, cudaMemcpyAsync(d + 3 * r0, in + 3 * r0, (size_t)m * 3 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            if (to_wgs) {
                const int rc2 = launch_enu_to_wgs84(h, f, m, nullptr, d + 3 * r0, d + 3 * r0, nullptr);
                if (rc2) return rc2;
            } else {
                const int rc2 = launch_wgs84_to_enu(h, f, m, d + 3 * r0, d + 3 * r0);
                if (rc2) return rc2;
            }
            MS_CUDA(h, cudaMemcpyAsync(out + 3 * r0, d + 3 * r0, (size_t)m * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            return MSNAP_OK;
        }();
    }
    h->stream = own;
    cudaError_t e1 = cudaStreamSynchronize(h->aux), e2 = cudaStreamSynchronize(h->stream);
    if (err) return err;
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        h->last_error = std::string("geo_host: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2);
        cudaGetLastError();
        return MSNAP_ERR_CUDA;
    }
    return MSNAP_OK;
}

int msnap_enu_to_wgs84_host(msnap_handle h, const double *reference_lla, long long n, const double *enu, double *lla_out) {
    return geo_host(h, reference_lla, n, enu, lla_out, true);
}
int msnap_wgs84_to_enu_host(msnap_handle h, const double *reference_lla, long long n, const double *lla, double *enu_out) {
    return geo_host(h, reference_lla, n, lla, enu_out, false);
}

// Developer/test hook: ENU -> WGS84 on device rows, also returning the number of fixed-point steps per point (cpp:939-949).
int msnap_debug_geo_steps_dev(msnap_handle h, const double *reference_lla, long long n, const double *enu, double *lla_out,
                              int *steps_out) {
    if (!h || !geo_reference_ok(reference_lla) || n < 0 || (n > 0 && (!enu || !lla_out || !steps_out))) return MSNAP_ERR_INVALID_ARG;
    if (n == 0) return MSNAP_OK;
    DeviceGuard guard(h->device);
    GeoFrame f;
    geo_make_frame(reference_lla, f);
    return launch_enu_to_wgs84(h, f, n, nullptr, enu, lla_out, steps_out);
}

// ---------------------------------------------------------------------------------------------- altitude optimisation
void msnap_altitude_params_default(msnap_altitude_params *p) {  // uavPathPlanning.hpp:415-421
    if (!p) return;
    p->lambda_smooth = 1.0;
    p->lambda_follow = 0.0;
    p->max_climb_rate = 2.0;
    p->uav_R = 2.0;
    p->safe_distance = 50.0;
}

int msnap_set_altitude_policy(msnap_handle h, int policy) {
    if (!h || policy < 0 || policy > 2) return MSNAP_ERR_INVALID_ARG;
    h->alt_policy = policy;
    return MSNAP_OK;
}

int msnap_cost_map_lookup_dev(msnap_handle h, const float *grid, int width, int height, double resolution, double origin_x,
                              double origin_y, long long n_rows_cap, const long long *n_rows_dev, const double *rows,
                              double *elev_out) {
    if (!h || !grid || width <= 0 || height <= 0 || !(resolution > 0.0) || n_rows_cap < 0 ||
        (n_rows_cap > 0 && (!rows || !elev_out)))
        return MSNAP_ERR_INVALID_ARG;
    if (n_rows_cap == 0) return MSNAP_OK;
    DeviceGuard guard(h->device);
    const long long want = (n_rows_cap + 255) / 256, cap = (long long)h->sm_count * 8;
    MS_LAUNCH(h, k_cost_lookup, (unsigned)(want < cap ? want : cap), 256, grid, width, height, resolution, origin_x, origin_y,
              n_rows_cap, n_rows_dev, rows, elev_out);
    return MSNAP_OK;
}

int msnap_altitude_optimize_batch_dev(msnap_handle h, const msnap_altitude_params *params, long long B,
                                      const long long *row_offset, long long n_rows_cap, double *rows_inout,
                                      const double *elev, double *z_pass1_out, int *solves_out, unsigned *flags_out) {
    if (!h || !params || B < 0 || n_rows_cap < 0 || (B > 0 && (!row_offset || (n_rows_cap > 0 && !rows_inout))))
        return MSNAP_ERR_INVALID_ARG;
    if (B == 0 || n_rows_cap == 0) return MSNAP_OK;
    DeviceGuard guard(h->device);
    const size_t n = (size_t)n_rows_cap;
    const bool part = h->alt_policy == 2;  // (its seven arrays are only touched by trajectories beyond ALTP_NMAX32 rows)
    int rc = arena_reserve(h, h->ws, (part ? AF_COUNT : 9) * padded(n * sizeof(double)) + padded((size_t)B * sizeof(int)));
    if (rc) return rc;
    double *w1 = arena_take<double>(h->ws, n), *w2 = arena_take<double>(h->ws, n);
    double *l1 = arena_take<double>(h->ws, n), *l2 = arena_take<double>(h->ws, n), *yd = arena_take<double>(h->ws, n);
    double *zin = arena_take<double>(h->ws, n), *cur = arena_take<double>(h->ws, n);
    double *tgt = part ? nullptr : arena_take<double>(h->ws, n);
    double *act = part ? nullptr : arena_take<double>(h->ws, n);  // active-set marks of pass 2 (0.0 / 1.0: staged like the other fields)
    int *st = arena_take<int>(h->ws, (size_t)B);  // per-trajectory outcome (ALT_ST_*)
    const AltParams p{params->lambda_smooth, params->lambda_follow, params->max_climb_rate, params->uav_R,
                      params->safe_distance};
    if (!h->alt_smem_opted) {  // > 48 KB of dynamic shared memory needs the opt-in (per device; once per handle)
        MS_CUDA(h, cudaFuncSetAttribute(k_alt_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ALT_SMEM_BYTES));
        MS_CUDA(h, cudaFuncSetAttribute(k_alt_solve_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ALT_SMEM_BYTES));
        MS_CUDA(h, cudaFuncSetAttribute(k_alt_part, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ALTP_SMEM_BYTES));
        h->alt_smem_opted = true;
    }
    if (h->alt_policy == 2) {  // partitioned solves, the whole stage in one launch (msnap_alt_part.cuh)
        const AltPartScratch sc{{w1, zin, yd, l1, l2, w2, cur}};  // W, E, YD, L1, L2, V1, V2
        prof_before(h, "k_alt_part");
        k_alt_part<<<grid_for(B, 32 / ALTP_GROUP), 32, ALTP_SMEM_BYTES, h->stream>>>(p, B, row_offset, rows_inout, elev, z_pass1_out,
                                                                                   solves_out, flags_out, n_rows_cap, sc);
        prof_after(h);
        ++h->launches;
        MS_CUDA(h, cudaPeekAtLastError());
        return MSNAP_OK;
    }
    const long long want = (n_rows_cap + 255) / 256, cap = (long long)h->sm_count * 8;
    MS_LAUNCH(h, k_alt_prep, (unsigned)(want < cap ? want : cap), 256, p, B, row_offset, rows_inout, elev, w1, w2, tgt, act,
              n_rows_cap);
    MS_LAUNCH(h, k_alt_ends, grid_for(B, 256), 256, B, row_offset, w1, w2, n_rows_cap);
    if (h->alt_policy == 1) {  // one lane per trajectory
        prof_before(h, "k_alt_solve");
        k_alt_solve<<<grid_for(B, 32), 32, ALT_SMEM_BYTES, h->stream>>>(p, B, row_offset, elev, w1, w2, tgt, l1, l2, yd, zin,
                                                                        cur, act, z_pass1_out, solves_out, flags_out,
                                                                        n_rows_cap, st);
    } else {  // lane pairs (two-sided elimination)
        prof_before(h, "k_alt_solve_pair");
        k_alt_solve_pair<<<grid_for(B, 16), 32, ALT_SMEM_BYTES, h->stream>>>(p, B, row_offset, elev, w1, w2, tgt, l1, l2, yd,
                                                                             zin, cur, act, z_pass1_out, solves_out, flags_out,
                                                                             n_rows_cap, st);
    }
    prof_after(h);
    ++h->launches;
    MS_CUDA(h, cudaPeekAtLastError());
    MS_LAUNCH(h, k_alt_mark, grid_for(B, 256), 256, B, row_offset, n_rows_cap, st, zin, cur);
    MS_LAUNCH(h, k_alt_finish, (unsigned)(want < cap ? want : cap), 256, B, row_offset, cur, zin, rows_inout, n_rows_cap);
    return MSNAP_OK;
}

int msnap_altitude_optimize_batch_host(msnap_handle h, const msnap_altitude_params *params, long long B,
                                       const long long *row_offset, double *rows_inout, const double *elev,
                                       double *z_pass1_out, int *solves_out, unsigned *flags_out) {
    if (!h || !params || B < 0 || (B > 0 && !row_offset)) return MSNAP_ERR_INVALID_ARG;
    if (B == 0) return MSNAP_OK;
    if (row_offset[0] != 0) return MSNAP_ERR_INVALID_ARG;
    for (long long b = 0; b < B; ++b)
        if (row_offset[b + 1] < row_offset[b]) return MSNAP_ERR_INVALID_ARG;
    const long long n = row_offset[B];
    if (n == 0) {
        for (long long b = 0; b < B; ++b) {
            if (solves_out) solves_out[b] = 0;
            if (flags_out) flags_out[b] = 0;
        }
        return MSNAP_OK;
    }
    if (!rows_inout) return MSNAP_ERR_INVALID_ARG;
    DeviceGuard guard(h->device);
    MS_CUDA(h, cudaStreamSynchronize(h->aux));
    int rc = arena_reserve(h, h->io, padded((B + 1) * sizeof(long long)) + padded((size_t)n * 3 * sizeof(double)) +
                                         2 * padded((size_t)n * sizeof(double)) + padded(B * sizeof(int)) +
                                         padded(B * sizeof(unsigned)));
    if (rc) return rc;
    long long *d_off = arena_take<long long>(h->io, B + 1);
    double *d_rows = arena_take<double>(h->io, (size_t)n * 3);
    double *d_elev = arena_take<double>(h->io, (size_t)n), *d_z1 = arena_take<double>(h->io, (size_t)n);
    int *d_sol = arena_take<int>(h->io, B);
    unsigned *d_fl = arena_take<unsigned>(h->io, B);
    cudaStream_t st = h->stream;
    MS_CUDA(h, cudaMemcpyAsync(d_off, row_offset, (B + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
    MS_CUDA(h, cudaMemcpyAsync(d_rows, rows_inout, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    if (elev) MS_CUDA(h, cudaMemcpyAsync(d_elev, elev, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
    rc = msnap_altitude_optimize_batch_dev(h, params, B, d_off, n, d_rows, elev ? d_elev : nullptr,
                                           z_pass1_out ? d_z1 : nullptr, d_sol, d_fl);
    if (rc) return rc;
    MS_CUDA(h, cudaMemcpyAsync(rows_inout, d_rows, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (z_pass1_out) MS_CUDA(h, cudaMemcpyAsync(z_pass1_out, d_z1, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (solves_out) MS_CUDA(h, cudaMemcpyAsync(solves_out, d_sol, B * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (flags_out) MS_CUDA(h, cudaMemcpyAsync(flags_out, d_fl, B * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    MS_CUDA(h, cudaStreamSynchronize(st));
    return MSNAP_OK;
}

// ---------------------------------------------------------------------------------------------- bound
int msnap_sample_bound_dev(msnap_handle h, const msnap_config *cfg, double v_avg_override, long long B, int ns_uniform,
                           const long long *seg_offset, const double *waypoints, long long *rows_out_dev) {
    if (!h || !cfg || !rows_out_dev) return MSNAP_ERR_INVALID_ARG;
    int rc = check_batch(B, ns_uniform, seg_offset, waypoints);
    if (rc) return rc;
    DeviceGuard guard(h->device);
    MS_CUDA(h, cudaMemsetAsync(rows_out_dev, 0, sizeof(long long), h->stream));
    if (B == 0) return MSNAP_OK;
    long long n_seg = 0;
    rc = device_total_segments(h, B, ns_uniform, seg_offset, &n_seg);
    if (rc) return rc;
    BatchIdx bi{B, n_seg, ns_uniform > 0 ? ns_uniform : 0, ns_uniform > 0 ? nullptr : seg_offset};
    const double va = v_avg_override > 0.0 ? v_avg_override : cfg->V_avg;
    if (h->wp_frame) {
        const long long n_pts = n_seg + B;
        rc = arena_reserve(h, h->ws, padded((size_t)n_pts * 3 * sizeof(double)));
        if (rc) return rc;
        double *enu = arena_take<double>(h->ws, (size_t)n_pts * 3);
        rc = launch_wgs84_to_enu(h, h->wp_geo, n_pts, waypoints, enu);
        if (rc) return rc;
        waypoints = enu;
    }
    MS_LAUNCH(h, k_bound, grid_for(n_seg, 256), 256, bi, waypoints, va, cfg->min_time_s,
              reinterpret_cast<unsigned long long *>(rows_out_dev));
    return MSNAP_OK;
}

int msnap_sample_bound_host(msnap_handle h, const msnap_config *cfg, double v_avg_override, long long B,
                            int ns_uniform, const long long *seg_offset, const double *waypoints, long long *rows_out) {
    if (!h || !cfg || !rows_out) return MSNAP_ERR_INVALID_ARG;
    int rc = check_batch(B, ns_uniform, seg_offset, waypoints);
    if (rc) return rc;
    *rows_out = 0;
    if (B == 0) return MSNAP_OK;
    if (!valid_host_offsets(B, ns_uniform, seg_offset)) return MSNAP_ERR_INVALID_ARG;
    DeviceGuard guard(h->device);
    const long long n_seg = host_total_segments(B, ns_uniform, seg_offset);
    const size_t n_pts = (size_t)(n_seg + B);
    rc = arena_reserve(h, h->io, padded((B + 1) * sizeof(long long)) + padded(n_pts * 3 * sizeof(double)) + 256);
    if (rc) return rc;
    long long *d_off = arena_take<long long>(h->io, B + 1);
    double *d_wp = arena_take<double>(h->io, n_pts * 3);
    long long *d_rows = arena_take<long long>(h->io, 1);
    cudaStream_t st = h->stream;
    if (ns_uniform <= 0)
        MS_CUDA(h, cudaMemcpyAsync(d_off, seg_offset, (B + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
    MS_CUDA(h, cudaMemcpyAsync(d_wp, waypoints, n_pts * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    rc = msnap_sample_bound_dev(h, cfg, v_avg_override, B, ns_uniform, ns_uniform > 0 ? nullptr : d_off, d_wp, d_rows);
    if (rc) return rc;
    MS_CUDA(h, cudaMemcpyAsync(rows_out, d_rows, sizeof(long long), cudaMemcpyDeviceToHost, st));
    MS_CUDA(h, cudaStreamSynchronize(st));
    return MSNAP_OK;
}

// ---------------------------------------------------------------------------------------------- Bezier generator
// math_util::Bezier::GenerateTrajectoryMatrix for B trajectories (bezier.cpp:127-189); see msnap_bezier.cuh.
static int bezier_dev(msnap_context *h, double resolution, double min_radius, long long B, int ns_uniform,
                      const long long *seg_offset, long long n_seg, const double *wp, long long capacity,
                      long long *sample_offset, double *samples, unsigned *flags) {
    BatchIdx bi{B, n_seg, ns_uniform > 0 ? ns_uniform : 0, ns_uniform > 0 ? nullptr : seg_offset};
    const bool ragged = bi.ns_uniform <= 0 && B < 2000000000LL;
    int rc = arena_reserve(h, h->ws, padded((size_t)n_seg * BEZ_WS * sizeof(double)) + padded(n_seg * sizeof(long long)) +
                                         padded(B * sizeof(long long)) + padded((size_t)(B / SCAN_BLOCK + 2) * sizeof(long long)) +
                                         (ragged ? padded(n_seg * sizeof(int)) : 0));
    if (rc) return rc;
    double *ws = arena_take<double>(h->ws, (size_t)n_seg * BEZ_WS);
    long long *seg_start = arena_take<long long>(h->ws, n_seg);
    long long *traj_count = arena_take<long long>(h->ws, B);
    long long *partial = arena_take<long long>(h->ws, B / SCAN_BLOCK + 2);
    if (ragged) {
        int *st = arena_take<int>(h->ws, n_seg);
        MS_LAUNCH(h, k_seg_traj, grid_for(B * 32, 256), 256, B, seg_offset, st);
        bi.seg_traj = st;
    }
    if (flags) MS_CUDA(h, cudaMemsetAsync(flags, 0, B * sizeof(unsigned), h->stream));
    MS_LAUNCH(h, k_bezier_prep, grid_for(n_seg, 128), 128, bi, wp, resolution, min_radius, ws, flags);
    MS_LAUNCH(h, k_bezier_rows, grid_for(B, 128), 128, bi, ws, seg_start, traj_count);
    const int nblk = (int)grid_for(B, SCAN_BLOCK);
    MS_LAUNCH(h, k_scan_reduce, nblk, SCAN_BLOCK, traj_count, B, partial);
    MS_LAUNCH(h, k_scan_partials, 1, SCAN_BLOCK, partial, nblk, sample_offset + B);
    MS_LAUNCH(h, k_scan_apply, nblk, SCAN_BLOCK, traj_count, B, partial, sample_offset);
    if (capacity > 0)
        MS_LAUNCH(h, k_bezier_write, grid_for(n_seg, 128), 128, bi, wp, ws, seg_start, sample_offset, capacity, samples, flags);
    return MSNAP_OK;
}

extern "C" int msnap_bezier_generate_batch_dev(msnap_handle h, double sample_distance_override, double min_radius, long long B,
                                               int ns_uniform, const long long *seg_offset, const double *waypoints,
                                               long long sample_capacity, long long *sample_offset_out, double *samples_out,
                                               unsigned *flags_out) {
    if (!h) return MSNAP_ERR_INVALID_ARG;
    int rc = check_batch(B, ns_uniform, seg_offset, waypoints);
    if (rc) return rc;
    if (!sample_offset_out || (!samples_out && sample_capacity > 0) || sample_capacity < 0 || !std::isfinite(sample_distance_override) ||
        !(min_radius == min_radius))
        return MSNAP_ERR_INVALID_ARG;
    if (B == 0) return MSNAP_OK;
    DeviceGuard guard(h->device);
    long long n_seg = 0;
    rc = device_total_segments(h, B, ns_uniform, seg_offset, &n_seg);
    if (rc) return rc;
    const double resolution = sample_distance_override > 0.0 ? sample_distance_override : 1.0;  // bezier.cpp:133-136
    return bezier_dev(h, resolution, min_radius, B, ns_uniform, seg_offset, n_seg, waypoints, sample_capacity, sample_offset_out,
                      samples_out, flags_out);
}

extern "C" int msnap_bezier_generate_batch_host(msnap_handle h, double sample_distance_override, double min_radius, long long B,
                                                int ns_uniform, const long long *seg_offset, const double *waypoints,
                                                long long sample_capacity, long long *sample_offset_out, double *samples_out,
                                                unsigned *flags_out) {
    if (!h) return MSNAP_ERR_INVALID_ARG;
    int rc = check_batch(B, ns_uniform, seg_offset, waypoints);
    if (rc) return rc;
    if (!sample_offset_out || (!samples_out && sample_capacity > 0) || sample_capacity < 0 || !std::isfinite(sample_distance_override) ||
        !(min_radius == min_radius))
        return MSNAP_ERR_INVALID_ARG;
    if (B == 0) { sample_offset_out[0] = 0; return MSNAP_OK; }
    if (!valid_host_offsets(B, ns_uniform, seg_offset)) return MSNAP_ERR_INVALID_ARG;
    DeviceGuard guard(h->device);
    const long long n_seg = host_total_segments(B, ns_uniform, seg_offset);
    const size_t n_pts = (size_t)(n_seg + B);
    MS_CUDA(h, cudaStreamSynchronize(h->aux));
    rc = arena_reserve(h, h->io, 2 * padded((B + 1) * sizeof(long long)) + padded(n_pts * 3 * sizeof(double)) +
                                     padded(B * sizeof(unsigned)) + padded((size_t)sample_capacity * 3 * sizeof(double)));
    if (rc) return rc;
    long long *d_off = arena_take<long long>(h->io, B + 1), *d_so = arena_take<long long>(h->io, B + 1);
    double *d_wp = arena_take<double>(h->io, n_pts * 3);
    unsigned *d_fl = arena_take<unsigned>(h->io, B);
    double *d_s = arena_take<double>(h->io, (size_t)sample_capacity * 3);
    cudaStream_t st = h->stream;
    if (ns_uniform <= 0) MS_CUDA(h, cudaMemcpyAsync(d_off, seg_offset, (B + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
    MS_CUDA(h, cudaMemcpyAsync(d_wp, waypoints, n_pts * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    const double resolution = sample_distance_override > 0.0 ? sample_distance_override : 1.0;
    rc = bezier_dev(h, resolution, min_radius, B, ns_uniform, ns_uniform > 0 ? nullptr : d_off, n_seg, d_wp, sample_capacity, d_so,
                    d_s, d_fl);
    if (rc) return rc;
    MS_CUDA(h, cudaMemcpyAsync(sample_offset_out, d_so, (B + 1) * sizeof(long long), cudaMemcpyDeviceToHost, st));
    if (flags_out) MS_CUDA(h, cudaMemcpyAsync(flags_out, d_fl, B * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    MS_CUDA(h, cudaStreamSynchronize(st));
    const long long total = sample_offset_out[B], rows = total < sample_capacity ? total : sample_capacity;
    if (rows > 0) {
        MS_CUDA(h, cudaMemcpyAsync(samples_out, d_s, (size_t)rows * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
        MS_CUDA(h, cudaStreamSynchronize(st));
    }
    return total > sample_capacity ? MSNAP_ERR_CAPACITY : MSNAP_OK;
}

// ---------------------------------------------------------------------------------------------- patrol post-processing
// gen_single_patrol after Minisnap_3D (uavPathPlanning.cpp:1857-1903); see msnap_patrol.cuh.
extern "C" int msnap_patrol_postprocess_dev(msnap_handle h, double distance, long long B, int ns_uniform,
                                            const long long *seg_offset, const double *waypoints,
                                            const long long *sample_offset, const double *samples, long long sample_capacity,
                                            const double *keep_up, long long out_capacity, long long *out_offset,
                                            double *out_rows, unsigned *flags_out) {
    if (!h) return MSNAP_ERR_INVALID_ARG;
    int rc = check_batch(B, ns_uniform, seg_offset, waypoints);
    if (rc) return rc;
    if (!sample_offset || (!samples && sample_capacity > 0) || sample_capacity < 0 || !out_offset || out_capacity < 0 ||
        (!out_rows && out_capacity > 0) || !(distance == distance) || (ns_uniform > 0 && ns_uniform < 4))
        return MSNAP_ERR_INVALID_ARG;  // a patrol polygon has >= 3 vertices: >= 5 closed waypoints, >= 4 segments
    if (B == 0) return MSNAP_OK;
    DeviceGuard guard(h->device);
    long long n_seg = 0;
    rc = device_total_segments(h, B, ns_uniform, seg_offset, &n_seg);
    if (rc) return rc;
    BatchIdx bi{B, n_seg, ns_uniform > 0 ? ns_uniform : 0, ns_uniform > 0 ? nullptr : seg_offset};
    rc = arena_reserve(h, h->ws, 2 * padded(B * sizeof(long long)) + padded((size_t)(B / SCAN_BLOCK + 2) * sizeof(long long)) +
                                     padded(B * sizeof(unsigned)));
    if (rc) return rc;
    long long *best = arena_take<long long>(h->ws, B), *count = arena_take<long long>(h->ws, B);
    long long *partial = arena_take<long long>(h->ws, B / SCAN_BLOCK + 2);
    unsigned *fl = flags_out ? flags_out : arena_take<unsigned>(h->ws, B);
    MS_LAUNCH(h, k_patrol_count, (unsigned)B, PATROL_THREADS, bi, waypoints, sample_offset, samples, sample_capacity, keep_up,
              distance, best, count, fl);
    const int nblk = (int)grid_for(B, SCAN_BLOCK);
    MS_LAUNCH(h, k_scan_reduce, nblk, SCAN_BLOCK, count, B, partial);
    MS_LAUNCH(h, k_scan_partials, 1, SCAN_BLOCK, partial, nblk, out_offset + B);
    MS_LAUNCH(h, k_scan_apply, nblk, SCAN_BLOCK, count, B, partial, out_offset);
    if (out_capacity > 0)
        MS_LAUNCH(h, k_patrol_write, (unsigned)B, PATROL_THREADS, bi, waypoints, sample_offset, samples, keep_up, distance, best,
                  out_offset, out_capacity, out_rows, fl);
    return MSNAP_OK;
}

extern "C" int msnap_patrol_postprocess_host(msnap_handle h, double distance, long long B, int ns_uniform,
                                             const long long *seg_offset, const double *waypoints,
                                             const long long *sample_offset, const double *samples, const double *keep_up,
                                             long long out_capacity, long long *out_offset, double *out_rows,
                                             unsigned *flags_out) {
    if (!h) return MSNAP_ERR_INVALID_ARG;
    int rc = check_batch(B, ns_uniform, seg_offset, waypoints);
    if (rc) return rc;
    if (!sample_offset || !out_offset || out_capacity < 0 || (!out_rows && out_capacity > 0)) return MSNAP_ERR_INVALID_ARG;
    if (B == 0) { out_offset[0] = 0; return MSNAP_OK; }
    if (!valid_host_offsets(B, ns_uniform, seg_offset) || sample_offset[0] != 0) return MSNAP_ERR_INVALID_ARG;
    for (long long b = 0; b < B; ++b)
        if (sample_offset[b + 1] < sample_offset[b]) return MSNAP_ERR_INVALID_ARG;
    const long long n_rows = sample_offset[B];
    if (n_rows > 0 && !samples) return MSNAP_ERR_INVALID_ARG;
    DeviceGuard guard(h->device);
    const long long n_seg = host_total_segments(B, ns_uniform, seg_offset);
    const size_t n_pts = (size_t)(n_seg + B);
    MS_CUDA(h, cudaStreamSynchronize(h->aux));
    rc = arena_reserve(h, h->io, 3 * padded((B + 1) * sizeof(long long)) + padded(n_pts * 3 * sizeof(double)) +
                                     padded((size_t)n_rows * 3 * sizeof(double)) + padded(B * sizeof(double)) +
                                     padded(B * sizeof(unsigned)) + padded((size_t)out_capacity * 3 * sizeof(double)));
    if (rc) return rc;
    long long *d_off = arena_take<long long>(h->io, B + 1), *d_so = arena_take<long long>(h->io, B + 1),
              *d_oo = arena_take<long long>(h->io, B + 1);
    double *d_wp = arena_take<double>(h->io, n_pts * 3), *d_s = arena_take<double>(h->io, (size_t)n_rows * 3);
    double *d_up = arena_take<double>(h->io, B);
    unsigned *d_fl = arena_take<unsigned>(h->io, B);
    double *d_out = arena_take<double>(h->io, (size_t)out_capacity * 3);
    cudaStream_t st = h->stream;
    if (ns_uniform <= 0) MS_CUDA(h, cudaMemcpyAsync(d_off, seg_offset, (B + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
    MS_CUDA(h, cudaMemcpyAsync(d_so, sample_offset, (B + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
    MS_CUDA(h, cudaMemcpyAsync(d_wp, waypoints, n_pts * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    if (n_rows > 0) MS_CUDA(h, cudaMemcpyAsync(d_s, samples, (size_t)n_rows * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    if (keep_up) MS_CUDA(h, cudaMemcpyAsync(d_up, keep_up, B * sizeof(double), cudaMemcpyHostToDevice, st));
    rc = msnap_patrol_postprocess_dev(h, distance, B, ns_uniform, ns_uniform > 0 ? nullptr : d_off, d_wp, d_so, d_s, n_rows,
                                      keep_up ? d_up : nullptr, out_capacity, d_oo, d_out, d_fl);
    if (rc) return rc;
    MS_CUDA(h, cudaMemcpyAsync(out_offset, d_oo, (B + 1) * sizeof(long long), cudaMemcpyDeviceToHost, st));
    if (flags_out) MS_CUDA(h, cudaMemcpyAsync(flags_out, d_fl, B * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    MS_CUDA(h, cudaStreamSynchronize(st));
    const long long total = out_offset[B], rows = total < out_capacity ? total : out_capacity;
    if (rows > 0) {
        MS_CUDA(h, cudaMemcpyAsync(out_rows, d_out, (size_t)rows * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
        MS_CUDA(h, cudaStreamSynchronize(st));
    }
    return total > out_capacity ? MSNAP_ERR_CAPACITY : MSNAP_OK;
}

// ---------------------------------------------------------------------------------------------- follower formations
// generateFollowerTrajectories (uavPathPlanning.cpp:3931-4398); see msnap_follow.cuh.
extern "C" double msnap_formation_distance(double formation_distance, double position_misalignment, double uav_R) {
    const double min_d = (2.0 * position_misalignment + uav_R) * 1.41421;  // cpp:4044-4051
    return formation_distance < min_d ? min_d : formation_distance;
}

extern "C" int msnap_followers_dev(msnap_handle h, int formation_model, double formation_distance, int uav_formation_max_row,
                                   int n_followers, int frame, const double *reference_lla, const double *starts_wgs84_dev,
                                   long long B, const long long *row_offset, const double *leader_rows, long long n_rows_cap,
                                   long long out_capacity, double *out_rows) {
    if (!h || B < 0 || n_followers < 0 || n_rows_cap < 0 || out_capacity < 0 || (frame != 0 && frame != 1) ||
        !(formation_distance == formation_distance) || (B > 0 && (!row_offset || (n_rows_cap > 0 && !leader_rows))) ||
        (out_capacity > 0 && !out_rows) || (frame == 1 && !geo_reference_ok(reference_lla)))
        return MSNAP_ERR_INVALID_ARG;
    if (B == 0 || n_followers == 0 || n_rows_cap == 0 || out_capacity == 0) return MSNAP_OK;
    DeviceGuard guard(h->device);
    int rc = arena_reserve(h, h->ws, 4 * padded((size_t)n_rows_cap * sizeof(double)) + 256);
    if (rc) return rc;
    double *ws_sin = arena_take<double>(h->ws, (size_t)n_rows_cap), *ws_cos = arena_take<double>(h->ws, (size_t)n_rows_cap);
    FollowParams p{formation_model, n_followers, uav_formation_max_row < 1 ? 1 : uav_formation_max_row, formation_distance};
    if (frame == 0) {
        MS_LAUNCH(h, k_follow_enu, (unsigned)B, FOLLOW_THREADS, p, B, row_offset, leader_rows, n_rows_cap, ws_sin, ws_cos, out_rows,
                  out_capacity);
        return MSNAP_OK;
    }
    // WGS84 rows: formation offsets and enuToWGS84 (cpp:4141) in one pass over the output
    double *ws_sin2 = arena_take<double>(h->ws, (size_t)n_rows_cap), *ws_cos2 = arena_take<double>(h->ws, (size_t)n_rows_cap);
    GeoFrame f;
    geo_make_frame(reference_lla, f);
    if (h->geo_trig)
        MS_LAUNCH(h, k_follow_lla<true>, (unsigned)B, FOLLOW_THREADS, p, f, B, row_offset, leader_rows, n_rows_cap, ws_sin, ws_cos,
                  ws_sin2, ws_cos2, out_rows, out_capacity);
    else
        MS_LAUNCH(h, k_follow_lla<false>, (unsigned)B, FOLLOW_THREADS, p, f, B, row_offset, leader_rows, n_rows_cap, ws_sin, ws_cos,
                  ws_sin2, ws_cos2, out_rows, out_capacity);
    if (starts_wgs84_dev && formation_model >= 2 && formation_model <= 4)
        MS_LAUNCH(h, k_follow_starts, grid_for(B * n_followers, 128), 128, n_followers, B, row_offset, leader_rows, n_rows_cap,
                  starts_wgs84_dev, out_rows, out_capacity);
    return MSNAP_OK;
}

extern "C" int msnap_followers_host(msnap_handle h, int formation_model, double formation_distance, int uav_formation_max_row,
                                    int n_followers, int frame, const double *reference_lla, const double *starts_wgs84,
                                    long long B, const long long *row_offset, const double *leader_rows, double *out_rows) {
    if (!h || B < 0 || n_followers < 0 || (B > 0 && !row_offset)) return MSNAP_ERR_INVALID_ARG;
    if (B == 0 || n_followers == 0) return MSNAP_OK;
    if (row_offset[0] != 0) return MSNAP_ERR_INVALID_ARG;
    for (long long b = 0; b < B; ++b)
        if (row_offset[b + 1] < row_offset[b]) return MSNAP_ERR_INVALID_ARG;
    const long long n = row_offset[B];
    if (n == 0) return MSNAP_OK;
    if (!leader_rows || !out_rows) return MSNAP_ERR_INVALID_ARG;
    DeviceGuard guard(h->device);
    MS_CUDA(h, cudaStreamSynchronize(h->aux));
    const size_t n_out = (size_t)n * n_followers;
    int rc = arena_reserve(h, h->io, padded((B + 1) * sizeof(long long)) + padded((size_t)n * 3 * sizeof(double)) +
                                         padded((size_t)n_followers * 3 * sizeof(double)) + padded(n_out * 3 * sizeof(double)));
    if (rc) return rc;
    long long *d_off = arena_take<long long>(h->io, B + 1);
    double *d_rows = arena_take<double>(h->io, (size_t)n * 3), *d_st = arena_take<double>(h->io, (size_t)n_followers * 3);
    double *d_out = arena_take<double>(h->io, n_out * 3);
    cudaStream_t st = h->stream;
    MS_CUDA(h, cudaMemcpyAsync(d_off, row_offset, (B + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
    MS_CUDA(h, cudaMemcpyAsync(d_rows, leader_rows, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    if (starts_wgs84) MS_CUDA(h, cudaMemcpyAsync(d_st, starts_wgs84, (size_t)n_followers * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    rc = msnap_followers_dev(h, formation_model, formation_distance, uav_formation_max_row, n_followers, frame, reference_lla,
                             starts_wgs84 ? d_st : nullptr, B, d_off, d_rows, n, (long long)n_out, d_out);
    if (rc) return rc;
    MS_CUDA(h, cudaMemcpyAsync(out_rows, d_out, n_out * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    MS_CUDA(h, cudaStreamSynchronize(st));
    return MSNAP_OK;
}

// ---------------------------------------------------------------------------------------------- single
int msnap_generate_one_host(msnap_handle h, const msnap_config *cfg, double sample_distance_override,
                            double v_avg_override, int n_points, const double *waypoints, long long sample_capacity,
                            double *samples_out, long long *n_samples_out) {
    if (n_samples_out) *n_samples_out = 0;
    if (!h || !cfg || !n_samples_out || n_points < 2) return MSNAP_ERR_INVALID_ARG;
    long long off[2] = {0, 0};
    int rc = msnap_generate_batch_host(h, cfg, sample_distance_override, v_avg_override, 1, n_points - 1, nullptr,
                                       waypoints, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, sample_capacity,
                                       off,
                                       samples_out, nullptr, nullptr);
    *n_samples_out = off[1];
    return rc;
}

// ---------------------------------------------------------------------------------------------- profiling
int msnap_profile_begin(msnap_handle h) {
    if (!h) return MSNAP_ERR_INVALID_ARG;
    DeviceGuard guard(h->device);
    for (auto &t : h->timed) {
        cudaEventDestroy(t.e0);
        cudaEventDestroy(t.e1);
    }
    h->timed.clear();
    h->profiling = true;
    return MSNAP_OK;
}

int msnap_profile_end(msnap_handle h, char *json_out, long long capacity) {
    if (!h || !json_out || capacity < 3) return MSNAP_ERR_INVALID_ARG;
    DeviceGuard guard(h->device);
    h->profiling = false;
    MS_CUDA(h, cudaStreamSynchronize(h->stream));
    struct Acc { std::string name; long long n; double ms; };
    std::vector<Acc> acc;
    for (auto &t : h->timed) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t.e0, t.e1);
        cudaEventDestroy(t.e0);
        cudaEventDestroy(t.e1);
        std::string nm(t.name);
        for (char &c : nm)
            if (c == '"' || c == '\\') c = '_';
        bool found = false;
        for (auto &a : acc)
            if (a.name == nm) { a.n++; a.ms += ms; found = true; break; }
        if (!found) acc.push_back({nm, 1, (double)ms});
    }
    h->timed.clear();
    std::string js = "{";
    for (size_t i = 0; i < acc.size(); ++i) {
        char buf[96];
        std::snprintf(buf, sizeof buf, "\": {\"launches\": %lld, \"total_ms\": %.6f}", acc[i].n, acc[i].ms);
        js += (i ? ", \"" : "\"") + acc[i].name + buf;
    }
    js += "}";
    if ((long long)js.size() + 1 > capacity) return MSNAP_ERR_CAPACITY;
    std::memcpy(json_out, js.c_str(), js.size() + 1);
    return MSNAP_OK;
}

// ---------------------------------------------------------------------------------------------- dev instrumentation
// Not part of the drop-in boundary: lets the developer see where a fused-kernel CTA spends its cycles.
// enable != 0 allocates a [8192][16] clock buffer that the next launches fill; out (host, 8192*16 int64) reads it.
int msnap_debug_phase_clocks(msnap_handle h, int enable, long long *out) {
    if (!h) return MSNAP_ERR_INVALID_ARG;
    DeviceGuard guard(h->device);
    const size_t bytes = 8192 * 16 * sizeof(long long);  // rows 0..4095: fused-solve CTAs, 4096..8191: sampler tiles
    if (out && h->phase_clocks) {
        MS_CUDA(h, cudaStreamSynchronize(h->stream));
        MS_CUDA(h, cudaMemcpy(out, h->phase_clocks, bytes, cudaMemcpyDeviceToHost));
    }
    if (enable && !h->phase_clocks) {
        MS_CUDA(h, cudaMalloc(&h->phase_clocks, bytes));
    }
    if (enable) MS_CUDA(h, cudaMemset(h->phase_clocks, 0, bytes));
    if (!enable && h->phase_clocks) {
        MS_CUDA(h, cudaStreamSynchronize(h->stream));
        cudaFree(h->phase_clocks);
        h->phase_clocks = nullptr;
    }
    return MSNAP_OK;
}

// ---------------------------------------------------------------------------------------------- fp64 peak
int msnap_measure_fp64_peak(msnap_handle h, double *tflops_out) {
    if (!h || !tflops_out) return MSNAP_ERR_INVALID_ARG;
    DeviceGuard guard(h->device);
    double *d = nullptr;
    MS_CUDA(h, cudaMalloc(&d, 64));
    cudaEvent_t e0, e1;
    MS_CUDA(h, cudaEventCreate(&e0));
    MS_CUDA(h, cudaEventCreate(&e1));
    const int iters = 1 << 14, block = 256, grid = h->sm_count * 8;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        MS_CUDA(h, cudaEventRecord(e0, h->stream));
        MS_LAUNCH(h, k_dfma_peak, grid, block, d, iters, 1.0000001, 1e-9);
        MS_CUDA(h, cudaEventRecord(e1, h->stream));
        MS_CUDA(h, cudaEventSynchronize(e1));
        float ms = 0.f;
        MS_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
        const double tf = 2.0 * 8.0 * (double)iters * block * grid / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *tflops_out = best;
    return MSNAP_OK;
}

}  // extern "C"
