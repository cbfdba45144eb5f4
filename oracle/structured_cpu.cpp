// CPU BASELINE (test / bench infrastructure, never part of the product path): the library's own structured O(ns)
// algorithm -- block-tridiagonal rows, twisted block elimination, the reference's two-pass path penalty and reweighting
// loop, the distance-thresholded sampler -- compiled for the host.  It is the "good CPU" line SURVEY.md section 8(d)
// asks for next to the reference's dense O(n^3) implementation: same formulation as the GPU kernels, plain C++, OpenMP
// over trajectories.
//
// How: cs_pathplan_b200/csrc/msnap_generic.cuh holds the sequential ("policy 1") kernel set, one thread per segment or
// per trajectory with no shared memory and no warp collectives: k_times, k_rows, k_thomas (the reweighting loop as the
// reference writes it), k_search, k_coeff, k_count, k_traj_count, k_write.  This file defines the handful of CUDA
// keywords and intrinsics those kernels use as ordinary C++ (thread_local threadIdx / blockIdx, IEEE arithmetic with
// -ffp-contract=off, std::fma), includes the header with MSNAP_HOST_EMULATION and runs every "kernel" as a loop over
// its thread indices.  Only bench.py's cpu_baseline leg and tests/ load the resulting library; the product
// (libmsnap_b200.so) has no CPU path and fails loudly without a GPU.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MSNAP_HOST_EMULATION 1
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__  // (a block is one thread here: its shared arrays are that thread's locals)
#define __grid_constant__

struct EmuDim3 { unsigned x = 0, y = 0, z = 0; };
static thread_local EmuDim3 threadIdx, blockIdx, blockDim, gridDim;

struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct uint4 { unsigned x, y, z, w; };

// IEEE operations (the translation unit is compiled with -ffp-contract=off: no contraction behind these)
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }
using std::fma;
using std::fabs;
using std::fmax;
using std::fmin;
using std::sqrt;
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
    return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
static inline unsigned long long atomicExch(unsigned long long *p, unsigned long long v) {
    return __atomic_exchange_n(p, v, __ATOMIC_RELAXED);
}
static inline void __syncthreads() {}
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
static inline long long clock64() { return 0; }
// Warp collectives: only the kernels this file never runs use them (lane pairs, speculative lanes, single-launch sampler).
[[noreturn]] static void emu_unsupported(const char *what) {
    std::fprintf(stderr, "structured_cpu: %s is not emulated\n", what);
    std::abort();
}
template <class T> static inline T __shfl_sync(unsigned, T, int) { emu_unsupported("__shfl_sync"); }
template <class T> static inline T __shfl_down_sync(unsigned, T, int) { emu_unsupported("__shfl_down_sync"); }
template <class T> static inline T __shfl_up_sync(unsigned, T, int) { emu_unsupported("__shfl_up_sync"); }
template <class T> static inline T __shfl_xor_sync(unsigned, T, int) { emu_unsupported("__shfl_xor_sync"); }
static inline unsigned __ballot_sync(unsigned, bool) { emu_unsupported("__ballot_sync"); }
static inline bool __any_sync(unsigned, bool) { emu_unsupported("__any_sync"); }
static inline bool __all_sync(unsigned, bool) { emu_unsupported("__all_sync"); }
static inline int __syncthreads_or(int) { emu_unsupported("__syncthreads_or"); }
static inline unsigned __fns(unsigned, unsigned, int) { emu_unsupported("__fns"); }

#include "../cs_pathplan_b200/csrc/msnap_generic.cuh"

using namespace msnap;

static const MsnapOrderTab g_tab[MSNAP_MAX_ORDER - MSNAP_MIN_ORDER + 1] = MSNAP_ORDER_TABLES;

// Runs `kernel(args...)` for thread indices 0 .. n-1 in blocks of `block` threads, blocks in parallel.
template <class K, class... A>
static void launch(long long n, int block, int threads, K kernel, A... args) {
    const long long n_blocks = (n + block - 1) / block;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (long long blk = 0; blk < n_blocks; ++blk) {
        blockDim.x = (unsigned)block;
        gridDim.x = (unsigned)n_blocks;
        blockIdx.x = (unsigned)blk;
        for (int t = 0; t < block; ++t) {
            threadIdx.x = (unsigned)t;
            kernel(args...);
        }
    }
}

struct CpuConfig {  // field-for-field msnap_config (include/msnap.h)
    int order;
    double path_weight, vel_zero_weight, V_avg, min_time_s, sample_distance;
    double start_vel[3], end_vel[3], start_acc[3], end_acc[3];
};

template <int O>
static int generate(const CpuConfig &cfg, double sd, double va, long long B, int ns_uniform, const long long *seg_offset,
                    const double *wp, double *times_out, double *coeff_out, double *max_dev_out, int *iters_out,
                    double *vw_final_out, long long capacity, long long *sample_offset, double *samples, unsigned *flags_out,
                    int threads) {
    using D = Dim<O>;
    const long long n_seg = ns_uniform > 0 ? B * ns_uniform : seg_offset[B];
    BatchIdx bi{B, n_seg, ns_uniform > 0 ? ns_uniform : 0, ns_uniform > 0 ? nullptr : seg_offset};
    std::vector<int> seg_traj;
    if (ns_uniform <= 0) {
        seg_traj.resize((size_t)n_seg);
        for (long long b = 0; b < B; ++b)
            for (long long g = seg_offset[b]; g < seg_offset[b + 1]; ++g) seg_traj[(size_t)g] = (int)b;
        bi.seg_traj = seg_traj.data();
    }
    SolveParams sp{};
    sp.pw = cfg.path_weight;
    sp.vw0 = cfg.vel_zero_weight;
    sp.max_iter = 10;
    for (int a = 0; a < 3; ++a) {
        sp.bc[a] = cfg.start_vel[a];
        sp.bc[3 + a] = cfg.end_vel[a];
        sp.bc[6 + a] = cfg.start_acc[a];
        sp.bc[9 + a] = cfg.end_acc[a];
    }
    std::vector<double> T((size_t)n_seg), base((size_t)n_seg * D::NBASE), state((size_t)(n_seg + 1) * D::NSTATE),
        segx((size_t)n_seg * D::NSEGX), coeff_own;
    std::vector<int> s_star((size_t)n_seg, 0);
    std::vector<unsigned> flags_own((size_t)B, 0u);
    unsigned *flags = flags_out ? flags_out : flags_own.data();
    std::fill(flags, flags + B, 0u);
    double *coeff = coeff_out;
    if (!coeff) {
        coeff_own.resize((size_t)n_seg * 3 * D::M);
        coeff = coeff_own.data();
    }
    const double *ht = &g_tab[O - MSNAP_MIN_ORDER].HT[0][0];
    const int blk = 64;
    const bool use_pw = sp.pw > 0.0;
    // the sequence of run_solve (csrc/msnap_capi.cu), policy 1
    launch(n_seg, blk, threads, k_times, bi, wp, va, cfg.min_time_s, T.data());
    if (use_pw) {
        launch(n_seg, blk, threads, k_rows<O>, bi, sp, wp, (const double *)T.data(), false, (const int *)nullptr, ht, base.data(),
               segx.data());
        SolveParams sp1 = sp;
        sp1.max_iter = 0;
        launch(B, 8, threads, k_thomas<O>, bi, sp1, wp, base.data(), state.data(), segx.data(), false, false, (double *)nullptr,
               (int *)nullptr, (double *)nullptr, flags);
        launch(n_seg, blk, threads, k_search<O>, bi, sp, wp, (const double *)T.data(), (const double *)state.data(), s_star.data());
    }
    launch(n_seg, blk, threads, k_rows<O>, bi, sp, wp, (const double *)T.data(), use_pw, (const int *)s_star.data(), ht, base.data(),
           segx.data());
    launch(B, 8, threads, k_thomas<O>, bi, sp, wp, base.data(), state.data(), segx.data(), use_pw, true, max_dev_out, iters_out,
           vw_final_out, flags);
    launch(n_seg, blk, threads, k_coeff<O>, bi, sp, wp, (const double *)T.data(), (const double *)state.data(), coeff, flags);
    if (times_out) std::memcpy(times_out, T.data(), (size_t)n_seg * sizeof(double));
    if (!sample_offset) return 0;
    // the sequence of run_sample (generic branch): count, rows per trajectory, exclusive scan, write
    std::vector<int> seg_count((size_t)n_seg), append_end((size_t)B);
    std::vector<unsigned long long> seg_mask((size_t)n_seg * 2);
    std::vector<double> seg_last((size_t)n_seg * 3), ttab(SAMPLE_TTAB_BIG);
    std::vector<long long> seg_start((size_t)n_seg), traj_count((size_t)B);
    {
        ttab[0] = 0.0;
        volatile double t = 0.1;  // candidate times exactly as ms.cpp:140 accumulates them
        for (int i = 1; i < SAMPLE_TTAB_BIG; ++i) {
            ttab[i] = t;
            t = t + 0.1;
        }
    }
    launch(n_seg, blk, threads, k_count<O>, bi, (const double *)coeff, (const double *)T.data(), sd, seg_count.data(),
           seg_mask.data(), seg_last.data(), flags);
    launch(B, blk, threads, k_traj_count<O>, bi, (const double *)coeff, (const double *)T.data(), (const int *)seg_count.data(),
           (const double *)seg_last.data(), seg_start.data(), append_end.data(), traj_count.data());
    sample_offset[0] = 0;
    for (long long b = 0; b < B; ++b) sample_offset[b + 1] = sample_offset[b] + traj_count[(size_t)b];
    if (samples && capacity > 0)  // one "thread" per block: k_write fills its shared candidate-time table cooperatively
        launch(n_seg, 1, threads, k_write<O>, bi, (const double *)coeff, (const double *)T.data(), sd, (const double *)ttab.data(),
               (const int *)seg_count.data(), (const unsigned long long *)seg_mask.data(), (const long long *)seg_start.data(),
               (const long long *)sample_offset, (const int *)append_end.data(), capacity, samples, flags);
    return 0;
}

extern "C" int msnap_structured_cpu_generate(const CpuConfig *cfg, double sample_distance_override, double v_avg_override,
                                             long long B, int ns_uniform, const long long *seg_offset, const double *wp,
                                             double *times_out, double *coeff_out, double *max_dev_out, int *iters_out,
                                             double *vw_final_out, long long capacity, long long *sample_offset,
                                             double *samples, unsigned *flags_out, int threads) {
    if (!cfg || !wp || B < 0 || (ns_uniform <= 0 && !seg_offset)) return 1;
    const double sd = sample_distance_override > 0.0 ? sample_distance_override : cfg->sample_distance;
    const double va = v_avg_override > 0.0 ? v_avg_override : cfg->V_avg;
    if (threads <= 0) {
#ifdef _OPENMP
        threads = omp_get_max_threads();
#else
        threads = 1;
#endif
    }
    switch (cfg->order) {
        case 2: return generate<2>(*cfg, sd, va, B, ns_uniform, seg_offset, wp, times_out, coeff_out, max_dev_out, iters_out,
                                   vw_final_out, capacity, sample_offset, samples, flags_out, threads);
        case 3: return generate<3>(*cfg, sd, va, B, ns_uniform, seg_offset, wp, times_out, coeff_out, max_dev_out, iters_out,
                                   vw_final_out, capacity, sample_offset, samples, flags_out, threads);
        case 4: return generate<4>(*cfg, sd, va, B, ns_uniform, seg_offset, wp, times_out, coeff_out, max_dev_out, iters_out,
                                   vw_final_out, capacity, sample_offset, samples, flags_out, threads);
        case 5: return generate<5>(*cfg, sd, va, B, ns_uniform, seg_offset, wp, times_out, coeff_out, max_dev_out, iters_out,
                                   vw_final_out, capacity, sample_offset, samples, flags_out, threads);
        default: return 1;
    }
}
