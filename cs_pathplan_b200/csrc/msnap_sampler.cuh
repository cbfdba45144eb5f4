// msnap_sampler.cuh -- single-launch sampler for uniform batches that evaluates every candidate ONCE (sm_100a, fp64).
//
// The distance-thresholded sampler of GenerateTrajectoryMatrix (ms.cpp:97-161) is sequential per segment (the acceptance
// test depends on the last accepted point) and its output is a CSR array whose row positions depend on the counts of every
// trajectory before it.  k_sample_scan (msnap_generic.cuh) therefore counts first -- remembering WHICH candidates were
// accepted in a 128-bit mask -- scans, and then evaluates the accepted candidates a second time at their final position,
// after expanding the masks into one descriptor per row and re-loading the coefficients.  Three quarters of its issued
// instructions are that bookkeeping, not arithmetic.
//
// k_sample_stage keeps what the count pass computes: an accepted point is stored straight away into the CTA's private
// staging slot in global memory -- a per-segment region laid out from the candidate-count upper bounds, which depend on
// the segment times only -- and the write pass is a copy: staging rows -> final rows, a warp per segment, coalesced on both
// sides.  The slot is reused by every tile the (persistent) CTA processes, so it lives in L2: the extra traffic never
// reaches HBM, whatever the batch size.  No masks, no descriptors, no second evaluation, no second read of the
// coefficients.  Same candidate times, same evaluation code, same tests in the same order as k_sample_scan / the
// per-pass kernels, hence the same bits.
//
//   A  count+stage  thread per segment (longest first), accepted rows -> staging region of the segment;
//                   segments with more than 128 candidates: a whole warp (warp_sample_long), count only
//   B  rows         thread per trajectory: start row of every segment, end-point rule (ms.cpp:157-160), row count
//   C  scan         tile-local scan, publish the aggregate, decoupled look-back for the tile's first row
//   D  copy         warp per segment: staging rows -> final rows; first / end points; long segments walked again (writing)
#ifndef MSNAP_SAMPLER_CUH
#define MSNAP_SAMPLER_CUH

#include "msnap_generic.cuh"

namespace msnap {

constexpr int STAGE_THREADS = 128;
constexpr int STAGE_ROWS_PER_SEG = SAMPLE_MASK_BITS + 2;  // staging rows a (non-long) segment can need

// candidates a non-long segment can have, from its duration alone (an upper bound; exact up to the rounding of T * 10)
__device__ __forceinline__ int stage_cand_ub(double Tk) {
    if (!sample_time_ok(Tk)) return 0;
    if (sample_dt(Tk) != 0.1) return 12;  // T < 1 s: dt = T / 10, 10 or 11 candidates
    const double n = Tk * 10.0 + 2.0;
    return n < (double)STAGE_ROWS_PER_SEG ? (int)n : STAGE_ROWS_PER_SEG;
}

__host__ __device__ inline size_t stage_smem_bytes(int tpt, int ns) {
    const size_t seg_cap = (size_t)tpt * ns;
    return SAMPLE_TTAB_N * sizeof(double) + seg_cap * 3 * sizeof(double) /*last*/ + seg_cap * sizeof(double) /*segT*/ +
           (size_t)tpt * 6 * sizeof(double) /*first / end point*/ + (size_t)(tpt + 1) * sizeof(long long) +
           seg_cap * 5 * sizeof(int) /*cnt, seg_start, perm, long_list, cand_off*/ + (size_t)tpt * sizeof(int) + 32;
}

template <int O>
__global__ void __launch_bounds__(STAGE_THREADS, 4) k_sample_stage(
    long long B, int ns, int tpt, long long n_tiles, const double *__restrict__ coeff, const double *__restrict__ T,
    double sample_distance, const double *__restrict__ t_table, unsigned long long *status, unsigned int *ticket,
    long long capacity, long long *__restrict__ sample_offset, double *__restrict__ samples, unsigned *__restrict__ flags,
    int *sm_ctr, double *stage_ws, long long slot_rows, long long *phase_clocks) {
    extern __shared__ unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = STAGE_THREADS / 32;
    const int seg_cap = tpt * ns;
    double *ttab = reinterpret_cast<double *>(smem_raw);                     // [SAMPLE_TTAB_N]
    double *last = ttab + SAMPLE_TTAB_N;                                      // [seg_cap][3] last accepted point
    double *segT = last + 3 * (size_t)seg_cap;                                // [seg_cap]
    double *ends = segT + seg_cap;                                            // [tpt][6] first point, end point
    long long *traj_base = reinterpret_cast<long long *>(ends + 6 * (size_t)tpt);  // [tpt + 1]
    int *cnt = reinterpret_cast<int *>(traj_base + tpt + 1);                  // [seg_cap] accepted (long segments: -n-1)
    int *seg_start = cnt + seg_cap;                                           // [seg_cap] row in trajectory
    int *perm = seg_start + seg_cap;                                          // [seg_cap] longest first
    int *long_list = perm + seg_cap;                                          // [seg_cap]
    int *cand_off = long_list + seg_cap;                                      // [seg_cap] first staging row of the segment
    int *append = cand_off + seg_cap;                                         // [tpt]
    const AcceptTest accept(sample_distance);
    __shared__ int hist[SAMPLE_MASK_BITS + 1];
    __shared__ long long sh_tile, sh_base, sh_part[NW];
    __shared__ int sh_first[NW], sh_slot, sh_sched[NW], sh_n_long, sh_wsum[NW];
    double *slot = stage_ws + (size_t)blockIdx.x * slot_rows * 3;
    for (int i = tid; i < SAMPLE_TTAB_N; i += STAGE_THREADS) ttab[i] = t_table[i];
    // chunk of the ranking a warp takes in phase A: a Latin square over (scheduler, co-resident CTA), as in k_sample_scan
    if (tid == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        sh_slot = atomicAdd(sm_ctr + smid, 1);
    }
    if (lane == 0) {
        unsigned wslot;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(wslot));
        sh_sched[wid] = (int)(wslot & 3u);
    }
    __syncthreads();
    int chunk = wid;
    {
        unsigned seen = 0;
        for (int w = 0; w < NW; ++w) seen |= 1u << sh_sched[w];
        if (seen == (1u << NW) - 1u) chunk = (sh_sched[wid] + sh_slot) & (NW - 1);
    }

    while (true) {
        if (tid == 0) {
            sh_tile = (long long)atomicAdd(ticket, 1u);
            sh_n_long = 0;
        }
        for (int i = tid; i <= SAMPLE_MASK_BITS; i += STAGE_THREADS) hist[i] = 0;
        __syncthreads();
        const long long tile = sh_tile;
        if (tile >= n_tiles) break;
        int stamp = 0;
#define STAGE_STAMP()                                                                  \
    do {                                                                               \
        if (phase_clocks && tid == 0 && tile < 4096 && stamp < 8)                      \
            phase_clocks[(4096 + tile) * 16 + stamp++] = clock64();                    \
    } while (0)
        STAGE_STAMP();
        const long long b0 = tile * tpt;
        const int nt = (int)min((long long)tpt, B - b0);
        const long long g0 = b0 * ns;
        const int nseg = nt * ns;
        // ---- segment times, counting sort by candidate-count estimate (longest first), staging regions
        for (int i = tid; i < nseg; i += STAGE_THREADS) {
            const double Tk = T[g0 + i];
            segT[i] = Tk;
            const int key = (int)fmax(0.0, fmin(Tk * 10.0, (double)SAMPLE_MASK_BITS));
            seg_start[i] = key;                     // (scratch until phase B)
            cnt[i] = atomicAdd(&hist[key], 1);      // rank inside the bin (scratch until phase A)
        }
        __syncthreads();
        if (tid < 32) {  // hist[k] := number of segments with a larger key
            int run = 0;
            for (int base = SAMPLE_MASK_BITS - (SAMPLE_MASK_BITS % 32); base >= 0; base -= 32) {
                const int k = base + lane;
                const int h = k <= SAMPLE_MASK_BITS ? hist[k] : 0;
                int v = h;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int u = __shfl_down_sync(0xffffffffu, v, o);
                    if (lane + o < 32) v += u;
                }
                if (k <= SAMPLE_MASK_BITS) hist[k] = run + v - h;
                run += __shfl_sync(0xffffffffu, v, 0);
            }
        }
        // exclusive prefix of the staging-row bounds over the segments in natural order (block scan, one or two per thread)
        {
            int carry = 0;
            for (int i0 = 0; i0 < nseg; i0 += STAGE_THREADS) {
                const int i = i0 + tid;
                const bool is_long = i < nseg && sample_is_long(segT[i], ttab);
                const int ub = (i < nseg && !is_long) ? stage_cand_ub(segT[i]) : 0;
                int v = ub;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int u = __shfl_up_sync(0xffffffffu, v, o);
                    if (lane >= o) v += u;
                }
                if (lane == 31) sh_wsum[wid] = v;
                __syncthreads();
                int wbase = carry;
                for (int w = 0; w < wid; ++w) wbase += sh_wsum[w];
                if (i < nseg) cand_off[i] = wbase + v - ub;
                int tot = 0;
                for (int w = 0; w < NW; ++w) tot += sh_wsum[w];
                carry += tot;
                __syncthreads();
            }
        }
        for (int i = tid; i < nseg; i += STAGE_THREADS) perm[hist[seg_start[i]] + cnt[i]] = i;
        __syncthreads();
        // ---- A: walk the candidates of every segment once; accepted points go to the segment's staging region
        const int rot_tid = (chunk << 5) | lane;
        for (int r0 = 0; r0 < nseg; r0 += 2 * STAGE_THREADS) {  // snake order over the ranking
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const int rk = half == 0 ? r0 + rot_tid : r0 + 2 * STAGE_THREADS - 1 - rot_tid;
                if (rk >= nseg) continue;
                const int i = perm[rk];
                const double Tk = segT[i];
                if (sample_is_long(Tk, ttab)) {
                    long_list[atomicAdd(&sh_n_long, 1)] = i;
                    continue;
                }
                double c[3][2 * O];
                load_coeff<O>(coeff, g0 + i, c);
                double *dst = slot + 3 * (size_t)cand_off[i];
                const double dt = sample_dt(Tk);
                const double tmax = sample_time_ok(Tk) ? Tk + 1e-12 : -1.0;
                if (!sample_time_ok(Tk) && flags) atomicOr(flags + b0 + i / ns, 1u);
                double prev[3], ca[3], cb[3];
                eval_xyz<O>(c, 0.0, prev);
                int n = 0;
                double t = dt;
                while (t <= tmax) {  // two candidates per trip: their positions do not depend on the decisions
                    const double t2 = t + dt;
                    eval_xyz<O>(c, fmin(t, Tk), ca);
                    eval_xyz<O>(c, fmin(t2, Tk), cb);
                    if (accept(ca, prev)) {
                        prev[0] = ca[0]; prev[1] = ca[1]; prev[2] = ca[2];
                        dst[3 * n] = ca[0]; dst[3 * n + 1] = ca[1]; dst[3 * n + 2] = ca[2];
                        ++n;
                    }
                    if (t2 <= tmax && accept(cb, prev)) {
                        prev[0] = cb[0]; prev[1] = cb[1]; prev[2] = cb[2];
                        dst[3 * n] = cb[0]; dst[3 * n + 1] = cb[1]; dst[3 * n + 2] = cb[2];
                        ++n;
                    }
                    t = t2 + dt;
                }
                cnt[i] = n;
                last[3 * i] = prev[0]; last[3 * i + 1] = prev[1]; last[3 * i + 2] = prev[2];
            }
        }
        __syncthreads();
        const int n_long = sh_n_long;
        for (int q = wid; q < n_long; q += NW) {  // long segments: a warp each, count only (rows: phase D)
            const int i = long_list[q];
            double c[3][2 * O], lp[3];
            load_coeff<O>(coeff, g0 + i, c);
            bool dropped = false;
            const int n = warp_sample_long<O, false>(c, segT[i], accept, t_table, 0, 0, nullptr, dropped, lp);
            if (lane == 0) {
                cnt[i] = -n - 1;
                last[3 * i] = lp[0]; last[3 * i + 1] = lp[1]; last[3 * i + 2] = lp[2];
            }
        }
        __syncthreads();
        STAGE_STAMP();
        // ---- B: per trajectory: segment start rows, first point, end-point rule (ms.cpp:157-160), row count
        if (tid < nt) {
            int total = 1;  // the first point
            int last_seg = -1;
            for (int k = 0; k < ns; ++k) {
                const int i = tid * ns + k;
                const int n = cnt[i] >= 0 ? cnt[i] : -cnt[i] - 1;
                seg_start[i] = total;
                total += n;
                if (n > 0) last_seg = i;
            }
            double first[3], back[3], endp[3], c[3][2 * O];
            load_coeff<O>(coeff, g0 + tid * ns, c);
            eval_xyz<O>(c, 0.0, first);
            if (last_seg >= 0) {
                back[0] = last[3 * last_seg]; back[1] = last[3 * last_seg + 1]; back[2] = last[3 * last_seg + 2];
            } else {
                back[0] = first[0]; back[1] = first[1]; back[2] = first[2];
            }
            const int il = tid * ns + ns - 1;
            load_coeff<O>(coeff, g0 + il, c);
            eval_xyz<O>(c, segT[il], endp);
            const int app = dist3(back, endp) > 1e-6 ? 1 : 0;
            append[tid] = app;
            traj_base[tid + 1] = total + app;
            double *e = ends + 6 * tid;
            e[0] = first[0]; e[1] = first[1]; e[2] = first[2];
            e[3] = endp[0]; e[4] = endp[1]; e[5] = endp[2];
        }
        __syncthreads();
        STAGE_STAMP();
        // ---- C: scan inside the tile, publish the aggregate, look back for the exclusive prefix of the tile
        if (tid == 0) {
            long long run = 0;
            traj_base[0] = 0;
            for (int t = 0; t < nt; ++t) {
                const long long c = traj_base[t + 1];
                traj_base[t] = run;
                run += c;
            }
            traj_base[nt] = run;
            const unsigned long long word = ((unsigned long long)run << 2) | (tile == 0 ? 2ull : 1ull);
            __threadfence();
            atomicExch(status + tile, word);
            if (tile == 0) sh_base = 0;
        }
        __syncthreads();
        if (tile > 0) {
            long long base = 0;
            long long hi = tile;
            bool done = false;
            while (!done) {
                const long long idx = hi - 1 - tid;
                unsigned long long w = 2ull;
                if (idx >= 0) {
                    const volatile unsigned long long *sp = status + idx;
                    do {
                        w = *sp;
                    } while ((w & 3ull) == 0ull);
                }
                const unsigned ball = __ballot_sync(0xffffffffu, (w & 3ull) == 2ull);
                if (lane == 0) sh_first[wid] = ball ? (tid & ~31) + __ffs(ball) - 1 : -1;
                __syncthreads();
                int first = -1;
                for (int wi = 0; wi < NW; ++wi)
                    if (sh_first[wi] >= 0) { first = sh_first[wi]; break; }
                long long v = (first < 0 || tid <= first) ? (long long)(w >> 2) : 0;
                if (idx < 0) v = 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
                if (lane == 0) sh_part[wid] = v;
                __syncthreads();
                for (int wi = 0; wi < NW; ++wi) base += sh_part[wi];
                done = first >= 0;
                hi -= STAGE_THREADS;
                __syncthreads();
            }
            if (tid == 0) {
                sh_base = base;
                __threadfence();
                atomicExch(status + tile, ((unsigned long long)(base + traj_base[nt]) << 2) | 2ull);
            }
        }
        __syncthreads();
        STAGE_STAMP();
        const long long tile_base = sh_base;
        if (tid < nt) sample_offset[b0 + tid] = tile_base + traj_base[tid];
        if (tile == n_tiles - 1 && tid == 0) sample_offset[B] = tile_base + traj_base[nt];
        // ---- D: copy.  First / end point of every trajectory, then the staged rows of every segment (a warp per segment:
        // 3 n contiguous doubles on both sides), then the long segments (walked again, writing).
        if (tid < nt) {
            const double *e = ends + 6 * tid;
            const long long r_first = tile_base + traj_base[tid], r_end = tile_base + traj_base[tid + 1] - 1;
            bool drop = false;
            if (r_first < capacity) {
                samples[3 * r_first] = e[0]; samples[3 * r_first + 1] = e[1]; samples[3 * r_first + 2] = e[2];
            } else {
                drop = true;
            }
            if (append[tid]) {
                if (r_end < capacity) {
                    samples[3 * r_end] = e[3]; samples[3 * r_end + 1] = e[4]; samples[3 * r_end + 2] = e[5];
                } else {
                    drop = true;
                }
            }
            if (drop && flags) atomicOr(flags + b0 + tid, 2u);
        }
        for (int i = wid; i < nseg; i += NW) {
            const int n = cnt[i];
            if (n <= 0) continue;  // nothing accepted, or a long segment
            const int t = i / ns;
            const long long row0 = tile_base + traj_base[t] + seg_start[i];
            long long room = capacity - row0;  // rows of this segment that fit the caller's buffer
            const int n_fit = room >= n ? n : (room > 0 ? (int)room : 0);
            const double *src = slot + 3 * (size_t)cand_off[i];
            double *dst = samples + 3 * row0;
            for (int e = lane; e < 3 * n_fit; e += 32) dst[e] = __ldcg(src + e);  // (L1 may hold the slot's lines of an earlier tile)
            if (n_fit < n && lane == 0 && flags) atomicOr(flags + b0 + t, 2u);
        }
        for (int q = wid; q < n_long; q += NW) {
            const int i = long_list[q];
            const int t = i / ns;
            double c[3][2 * O], lp[3];
            load_coeff<O>(coeff, g0 + i, c);
            bool dropped = false;
            warp_sample_long<O, true>(c, segT[i], accept, t_table, tile_base + traj_base[t] + seg_start[i], capacity, samples, dropped,
                                      lp);
            if (__any_sync(0xffffffffu, dropped) && lane == 0 && flags) atomicOr(flags + b0 + t, 2u);
        }
        __syncthreads();  // smem and the staging slot are reused by the next tile
        STAGE_STAMP();
    }
#undef STAGE_STAMP
}

}  // namespace msnap

#endif  // MSNAP_SAMPLER_CUH
