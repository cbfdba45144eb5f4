// msnap_fused.cuh -- fused, persistent solve kernel for uniform batches (every trajectory has `ns` segments).
//
// One launch replaces k_times .. k_coeff of the generic path.  A CTA of FUSED_THREADS threads owns a tile of `tpc`
// consecutive trajectories whose waypoints, time powers, block rows and deviation probes live in shared memory; it
// walks the phases below with __syncthreads() between them and loops over tiles (persistent grid).
//
//   phase          parallel over            work                                                   reference lines
//   load           elements (coalesced)     waypoints + boundary derivatives -> smem
//   times          (traj, seg)              T_k, T_k^-e, T_k^r                                       ms.cpp:63-72
//   rows (pass 1)  (traj, row)              D_j, U_j, r_j without penalties                          ms.cpp:247-330, 350
//   thomas pass 1  traj                     block-tridiagonal solve, 3 axes (state in smem)          ms.cpp:357-405
//   search         (traj, seg)              17-sample arg-max of the deviation -> s*                 ms.cpp:408-439
//   rows (pass 2)  (traj, row) / (traj,seg) rows with pw*h h' and g; deviation probes                ms.cpp:441-469, 511-522
//   thomas spec    (traj, reweight iter)    ALL velocity weights of the reweighting loop at once     ms.cpp:76-90, 474-509, 524-592
//                                           + max deviation at the recorded t*                       ms.cpp:594-624
//   select         traj                     first iteration whose max_dev <= 0.2 (or the 11th)       ms.cpp:82
//   coeff          (traj, seg, axis)        c = M^-1 d, PolyCoeff layout, straight to HBM            ms.cpp:584-591, 626-646
//
// Why speculate on the reweighting loop: its velocity weights are a fixed sequence (vw0, then 0.01 or 2x, ...) and
// only the *decision* to continue is data dependent, so the <= 11 solves of a trajectory are independent problems.
// At the headline size (4 096 x 16) one Thomas chain per trajectory would leave 4 096 threads on a 148-SM GPU;
// 11 chains per trajectory give 45 056.  Results are bitwise identical to the sequential loop (same arithmetic per
// iteration, same selection rule).
//
// The sweeps are issue-bound, so every array in the hot loops has COMPILE-TIME strides (addresses are base +
// immediate): shared-memory rows are [row][fields] with an odd row pitch, sweep state is [row][field][lanes] with
// a fixed lane count (14 in shared memory, 160 in the per-CTA L2-resident scratch slot of the speculative lanes), two
// consecutive fields of a lane side by side (RowPair): 128-bit state loads / stores.
#ifndef MSNAP_FUSED_CUH
#define MSNAP_FUSED_CUH

#include "msnap_generic.cuh"

namespace msnap {

constexpr int FUSED_THREADS = 224;     // default variant: 2 CTAs per SM (<= 144 registers per thread); 14 x 16 (trajectory, segment)
                                       // items = one round of the segment-parallel phases
constexpr int FUSED_THREADS_3 = 192;   // big-batch variant: 3 CTAs per SM (<= 112 registers; the row sweeps stay spill-free, the
                                       // segment-parallel phases spill a little): half again as many warps per SM when the
                                       // batch is many waves deep and the shared-memory tile is small enough
constexpr int FUSED_SLOT_LANES = 160;  // lanes per state row of the global scratch slot (>= 16 trajectories x 10 speculative iterations)
constexpr int FUSED_SMEM_LANES = 14;   // lanes per state row in shared memory (= max trajectories per tile)

struct FusedParams {
    long long B;
    int ns;
    int tpc;           // trajectories per tile (<= FUSED_SMEM_LANES)
    int nit;           // solves per trajectory in the speculative phase: max_iter + 1 if pw > 0 else 1
    int traj_stride;   // doubles per trajectory block in shared memory (FusedSmem::size: even, 2 mod 4)
    long long n_tiles;
    const double *wp;
    const double *times_in;  // nullptr => allocate from v_avg / min_time
    double v_avg, min_time;
    SolveParams sp;
    const double *ht;        // HT table of this order in global memory
    double *times_out, *coeff_out, *max_dev_out, *vw_final_out;
    int discard_state;       // drop the dead scratch lines from L2 at tile end (less write-back, a few discards per thread)
    double *coeff_mirror;    // optional second destination of the coefficients (the caller's pinned host buffer)
    int *iters_out, *best_s_out;
    unsigned *flags;
    double *state_ws;        // [gridDim.x][ns-1][NSTATE][FUSED_SLOT_LANES]
    long long *phase_clocks; // optional [gridDim.x][16]: clock64() of thread 0 after each phase of the CTA's first tile
                             // (columns 0..9) and (forward end, backward end) of speculative lanes 0, 64, 128 (10..15)
};

// Shared-memory layout of one trajectory block (offsets in doubles).  Row pitches are odd so that the
// (traj, row)-parallel writers hit distinct banks; readers in the sweeps broadcast.
template <int O>
struct FusedSmem {
    using D = Dim<O>;
    static constexpr int ST = 3;                 // per segment: T, 1/T (powers are re-multiplied, not stored)
    // Base rows (D, U, r) are read by the sweeps with 128-bit loads: even row pitch, even offsets, and a pitch that is
    // 2 mod 4 doubles apart from a multiple of 32 banks so that the (traj, row)-parallel writers spread over the banks
    // (24 doubles = 48 words would put all rows of a trajectory on two bank groups).
    static constexpr int BS = ((D::NBASE + 1) & ~1) + ((((D::NBASE + 1) & ~1) % 4 == 0) ? 2 : 0);
    static constexpr int XS = D::NSEGX | 1;      // per segment: deviation probe (scalar reads: odd pitch)
    int oSeg, oP, oBC, oBase, oSegx, oS, size;
    __host__ __device__ static constexpr int even(int x) { return (x + 1) & ~1; }
    __host__ __device__ FusedSmem(int ns) {
        const int nr = ns - 1;
        oSeg = 0;
        oP = oSeg + ST * ns;              // [w][3]
        oBC = oP + 3 * (ns + 1);          // d0[NR], dN[NR]: fixed boundary derivatives, [r-1][axis]
        oBase = even(oBC + 2 * D::NR);    // 16-byte aligned rows
        oSegx = oBase + BS * nr;
        oS = oSegx + XS * ns;             // ns ints
        size = even(oS + (ns + 1) / 2);
        if (size % 4 == 0) size += 2;     // even (alignment) and 2 mod 4: the blocks of the ~4 trajectories a warp of
                                          // sweep lanes reads at the same offsets start 4 banks apart, not on top of each other
    }
};

template <int O>
struct FBaseRows {
    const double *p;
    static constexpr int FS = 1;
    static constexpr bool VEC = true;  // rows are 16-byte aligned (FusedSmem): 128-bit shared-memory loads
    using Mem = PlainMem;
    __device__ __forceinline__ const double *operator()(int j) const { return p + j * FusedSmem<O>::BS; }
    __device__ __forceinline__ void prefetch(int) const {}
};
template <int O>
struct FSegxRows {
    const double *p;
    static constexpr int FS = 1;
    static constexpr bool VEC = false;
    using Mem = PlainMem;
    __device__ __forceinline__ const double *operator()(int k) const { return p + k * FusedSmem<O>::XS; }
    __device__ __forceinline__ void prefetch(int) const {}
};
struct FPos {
    const double *p;
    __device__ __forceinline__ void operator()(int w, double (&out)[3]) const {
        out[0] = p[3 * w];
        out[1] = p[3 * w + 1];
        out[2] = p[3 * w + 2];
    }
};
template <int O, int LANES, class MemT = PlainMem>
struct FStateRows {
    double *p;  // this lane's column: base + 2 * lane (rows are [field / 2][lane][2], see RowPair)
    static constexpr int FS = 1;
    static constexpr bool VEC = false;
    static constexpr bool ENABLED = true;  // usable as the solution sink of thomas_backward
    static constexpr int PAIR = LANES;
    using Mem = MemT;
    __device__ __forceinline__ double *operator()(int j) const { return p + (size_t)j * (Dim<O>::NSTATE * LANES); }
    __device__ __forceinline__ void prefetch(int) const {}
};

// Sweep state of the speculative lanes in the CTA's L2-resident scratch slot: rows [row][field / 2][lane][2] (RowPair).
template <int O, int LANES>
struct FSlotRows {
    double *p;  // slot + 2 * lane
    static constexpr int FS = 1;
    static constexpr bool VEC = false;
    static constexpr bool ENABLED = true;
    static constexpr int PAIR = LANES;
    using Mem = L2KeepMem;
    __device__ __forceinline__ double *operator()(int j) const { return p + (size_t)j * (Dim<O>::NSTATE * LANES); }
    __device__ __forceinline__ void prefetch(int) const {}
};

// field f of state row j in a paired-layout state array of LANES lanes, relative to the lane's column pointer
template <int O, int LANES>
__device__ __forceinline__ constexpr int state_at_idx(int j, int f) {
    return j * (Dim<O>::NSTATE * LANES) + (f >> 1) * (2 * LANES) + (f & 1);
}

// segment time powers from smem: ip[e] = T^-e (e = 0..2o-1), pT[r] = T^r (r = 0..o-1)
template <int O>
__device__ __forceinline__ void fused_powers(const double *seg, double (&ip)[2 * O], double (&pT)[O]) {
    const double T = seg[0], inv = seg[1];  // same multiplication sequence as time_powers(): bitwise identical
    ip[0] = 1.0;
#pragma unroll
    for (int e = 1; e < 2 * O; ++e) ip[e] = ip[e - 1] * inv;
    pT[0] = 1.0;
#pragma unroll
    for (int r = 1; r < O; ++r) pT[r] = pT[r - 1] * T;
}

// Endpoint derivative vector ([r-1][axis]) of waypoint w from the fixed boundary data or the solved state rows.
template <int O>
__device__ __forceinline__ void fused_derivs(const double *bcs, const double *state_lane, int ns, int w,
                                             double (&d)[3 * (O - 1)]) {
    using D = Dim<O>;
#pragma unroll
    for (int i = 0; i < D::NR; ++i) {
        if (w == 0) d[i] = bcs[i];
        else if (w == ns) d[i] = bcs[D::NR + i];
        else d[i] = state_lane[state_at_idx<O, FUSED_SMEM_LANES>(w - 1, D::SX + i)];
    }
}

// All phase bodies are inlined: the order tables are constexpr (literal operands), so nothing is hoisted out of
// the persistent loop, and there is no call ABI (callee-saved spills go to local memory, and with two ~110 KB CTAs
// per SM there is almost no L1 left to catch them).

// rows: item (t, j), j = 1..ns-1
template <int O>
__device__ __forceinline__ void fused_row_item(const FusedParams &p, double *blk, int ns, int j, bool with_pw) {
    using D = Dim<O>;
    const FusedSmem<O> L(ns);
    const double *P = blk + L.oP + 3 * j;
    const int *ss = reinterpret_cast<const int *>(blk + L.oS);
    Boundary<O> bc;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        bc.y0[a][0] = bc.yN[a][0] = 0.0;
#pragma unroll
        for (int r = 1; r < O; ++r) {
            bc.y0[a][r] = blk[L.oBC + (r - 1) * 3 + a];
            bc.yN[a][r] = blk[L.oBC + D::NR + (r - 1) * 3 + a];
        }
    }
    double Pm[3], P0[3], Pp[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        Pm[a] = P[a - 3];
        P0[a] = P[a];
        Pp[a] = P[a + 3];
    }
    double ipa[2 * O], ipc[2 * O], pTa[O], pTc[O];
    fused_powers<O>(blk + L.oSeg + (j - 1) * L.ST, ipa, pTa);
    fused_powers<O>(blk + L.oSeg + j * L.ST, ipc, pTc);
    assemble_row_p<O>(ipa, pTa, ipc, pTc, Pm, P0, Pp, j == 1, j == ns - 1, bc, with_pw, p.sp.pw,
                      with_pw ? ss[j - 1] : 0, with_pw ? ss[j] : 0, p.ht, blk + L.oBase + (j - 1) * L.BS, 1);
}

// search: item (t, k) -> index of the worst-deviation sample (first strict maximum, ms.cpp:435)
template <int O>
__device__ __forceinline__ int fused_search_item(const double *blk, const double *state_lane, int ns, int k) {
    const FusedSmem<O> L(ns);
    double dk[3 * (O - 1)], dk1[3 * (O - 1)];
    fused_derivs<O>(blk + L.oBC, state_lane, ns, k, dk);
    fused_derivs<O>(blk + L.oBC, state_lane, ns, k + 1, dk1);
    const double *P = blk + L.oP + 3 * k;
    double ip[2 * O], pT[O];
    fused_powers<O>(blk + L.oSeg + k * L.ST, ip, pT);
    double dh[3][2 * O];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        dh[a][0] = P[a];
        dh[a][O] = P[3 + a];
#pragma unroll
        for (int q = 1; q < O; ++q) {
            dh[a][q] = pT[q] * dk[(q - 1) * 3 + a];
            dh[a][O + q] = pT[q] * dk1[(q - 1) * 3 + a];
        }
    }
    double best = -1.0;
    int best_s = 0;
#pragma unroll
    for (int s = 0; s <= 16; ++s) {
        const double tau = (double)s * 0.0625;
        double d2 = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            double v = 0.0;
#pragma unroll
            for (int q = 0; q < 2 * O; ++q) v = fma(Tab<O>::HT(s, q), dh[a][q], v);
            const double dd = v - fma(tau, P[3 + a] - P[a], P[a]);
            d2 = fma(dd, dd, d2);
        }
        if (d2 > best) {
            best = d2;
            best_s = s;
        }
    }
    return best_s;
}

// deviation probe of segment (t, k): h, L(t*), 1/len^2
template <int O>
__device__ __forceinline__ void fused_probe_item(const FusedParams &p, double *blk, int ns, int k) {
    const FusedSmem<O> L(ns);
    const double *P = blk + L.oP + 3 * k;
    double ip[2 * O], pT[O], h[2 * O];
    fused_powers<O>(blk + L.oSeg + k * L.ST, ip, pT);
    const int s = reinterpret_cast<const int *>(blk + L.oS)[k];
    hermite_at<O>(p.ht, s, pT, h);
    double *x = blk + L.oSegx + k * L.XS;
#pragma unroll
    for (int q = 0; q < 2 * O; ++q) x[q] = h[q];
    const double tau = (double)s * 0.0625;
    double l2 = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double d = P[3 + a] - P[a];
        x[2 * O + a] = fma(tau, d, P[a]);
        l2 = fma(d, d, l2);
    }
    const double len = sqrt(l2);
    x[2 * O + 3] = len > 1e-6 ? 1.0 / l2 : 0.0;
}

// The chains below are NOT inlined on purpose: as separate functions the row sweeps get the whole register budget to
// themselves (inlined next to the other phases they spill into local memory inside the row loop); a call is made once
// per chain, so its ABI cost is noise.
//
// One speculative chain of trajectory block `blk` with diagonal shift add00 by ONE lane (split row n-1): elimination
// into `state_at`, back-substitution with the deviation probes, leaving x nowhere.  Returns the pivot status.
struct ChainResult {
    double max_dev;
    bool ok;
};

template <int O, class StateAt>
__device__ __forceinline__ ChainResult fused_chain_spec(const double *blk, int ns, double add00, const StateAt state_at,
                                                        long long *clk = nullptr) {
    using D = Dim<O>;
    const FusedSmem<O> L(ns);
    const FBaseRows<O> base_at{blk + L.oBase};
    const int n = ns - 1, m = split_row(n, false);
    ChainResult r;
    r.ok = thomas_forward<O>(n, m, add00, base_at, state_at);
    if (clk) clk[0] = clock64();
    const FSegxRows<O> segx_at{blk + L.oSegx};
    const FPos pos{blk + L.oP};
    r.max_dev = thomas_backward<O, true, true>(n, m, state_at, NoOut{}, segx_at, pos, blk + L.oBC, blk + L.oBC + D::NR);
    if (clk) clk[1] = clock64();
    return r;
}

// One latency-critical chain (pass 1, the last reweighting iteration, a bare solve) by TWO adjacent lanes of a warp
// (balanced split): side 0 eliminates the top half, solves the split row and substitutes back upwards; side 1 takes the
// bottom half.  `pair_mask` = the lanes of the warp that call this function (all of them must).  The solution is left
// in `state_at` (shared memory).  Both lanes return the combined pivot status and the max deviation (0 when !EVAL).
template <int O, bool EVAL, class StateAt>
__device__ __forceinline__ bool fused_chain_pair(const double *blk, int ns, double add00, const StateAt state_at, int side,
                                              unsigned pair_mask, double *max_dev_out) {
    using D = Dim<O>;
    const FusedSmem<O> L(ns);
    const FBaseRows<O> base_at{blk + L.oBase};
    const FSegxRows<O> segx_at{blk + L.oSegx};
    const FPos pos{blk + L.oP};
    const double *d0 = blk + L.oBC, *dN = blk + L.oBC + D::NR;
    const int n = ns - 1;
    bool ok = true;
    double m2 = 0.0;
    if (n <= 0) {
        if (EVAL && side == 0) m2 = single_segment_dev2<O>(segx_at, pos, d0, dN);
    } else {
        const int m = split_row(n, true);
        ok = elim_half<O>(n, side ? n - 1 - m : m, side != 0, add00, base_at, state_at);
        __syncwarp(pair_mask);
        if (side == 0) ok = elim_middle<O>(n, m, add00, base_at, state_at) && ok;
        __syncwarp(pair_mask);
        m2 = back_half<O, EVAL, false>(n, m, side != 0, state_at, state_at, segx_at, pos, side ? dN : d0, 0.0);
    }
    m2 = fmax(m2, __shfl_xor_sync(pair_mask, m2, 1));
    ok = __shfl_xor_sync(pair_mask, ok ? 1 : 0, 1) != 0 && ok;
    *max_dev_out = EVAL ? sqrt(m2) : 0.0;
    return ok;
}

// Back-substitution only: replay a finished speculative elimination (`state_at`) and leave the solution in `xout`.
template <int O, class StateAt, class XOut>
__device__ __noinline__ void fused_replay(const double *blk, int ns, const StateAt state_at, const XOut xout) {
    using D = Dim<O>;
    const FusedSmem<O> L(ns);
    const FSegxRows<O> segx_at{blk + L.oSegx};
    const FPos pos{blk + L.oP};
    thomas_backward<O, false, false>(ns - 1, split_row(ns - 1, false), state_at, xout, segx_at, pos, blk + L.oBC,
                                     blk + L.oBC + D::NR);
}

// coefficients of (t, k, axis) from the final solution in the shared-memory state rows; returns finiteness
template <int O>
__device__ __forceinline__ bool fused_coeff_item(const double *blk, const double *state_lane, int ns, int k, int a,
                                              double *dst_row, double *mirror_row) {
    using D = Dim<O>;
    constexpr int M = D::M;
    const FusedSmem<O> L(ns);
    const double *P = blk + L.oP + 3 * k;
    const double *bcs = blk + L.oBC;
    double yk[O], yk1[O];
    yk[0] = P[a];
    yk1[0] = P[3 + a];
#pragma unroll
    for (int d = 1; d < O; ++d) {
        const int f = (d - 1) * 3 + a;
        yk[d] = (k == 0) ? bcs[f] : state_lane[state_at_idx<O, FUSED_SMEM_LANES>(k - 1, D::SX + f)];
        yk1[d] = (k == ns - 1) ? bcs[D::NR + f] : state_lane[state_at_idx<O, FUSED_SMEM_LANES>(k, D::SX + f)];
    }
    double ip[2 * O], pT[O], co[M];
    fused_powers<O>(blk + L.oSeg + k * L.ST, ip, pT);
    hermite_coeffs<O>(yk, yk1, ip, pT, co);
    double2 *dst = reinterpret_cast<double2 *>(dst_row);
    bool finite = true;
#pragma unroll
    for (int q = 0; q < M / 2; ++q) {
        dst[q] = make_double2(co[2 * q], co[2 * q + 1]);
        if (mirror_row) reinterpret_cast<double2 *>(mirror_row)[q] = make_double2(co[2 * q], co[2 * q + 1]);
        finite = finite && (fabs(co[2 * q]) <= 1.7976931348623157e308) && (fabs(co[2 * q + 1]) <= 1.7976931348623157e308);
    }
    return finite;
}

// Pointers into a CTA's dynamic shared memory behind the trajectory blocks.
template <int O>
struct FusedTail {
    double *state1, *md;
    int *okf, *sel, *ok1;
    __device__ __forceinline__ FusedTail(double *smem, int ns, int tpc, int nit, int tstride) {
        state1 = smem + (size_t)tpc * tstride;                            // [nr][NSTATE][SL]: pass-1 sweep, then the final x
        md = state1 + (size_t)(ns - 1) * Dim<O>::NSTATE * FUSED_SMEM_LANES;  // [tpc][nit] max deviation of every solve
        okf = reinterpret_cast<int *>(md + tpc * nit);                    // [tpc][nit] pivot status
        sel = okf + tpc * nit;                                            // [tpc] selected iteration
        ok1 = sel + tpc;                                                  // [tpc] pass-1 pivot status
    }
};

// The chain phases are functions of their own that derive everything from (p, tile, threadIdx): a callee may not touch
// the registers its caller keeps live across the call, and the kernel's tile loop kept 54 of the 168 there -- enough to
// make the row sweeps spill six doubles of W per row into local memory, whose reloads (26 % L1 misses: the L1 is what two
// 112 KB CTAs leave of it) sat on the critical path of every row.  Called like this, only `tile` survives the call.
//
// pass 1 (ms.cpp:357-405): SIX lanes per trajectory -- two half chains (balanced twisted split) x three axes, the one-axis forms
// of msnap_device.cuh -- five trajectories per warp; solution left in the shared-memory state rows.  The other lanes of the CTA
// have nothing to do in this phase, and the per-row instruction stream of the lanes that do is what the phase lasts.
template <int O>
__device__ __noinline__ void fused_phase_pass1(const FusedParams &p, int nt) {
    extern __shared__ double smem[];  // (declared here, not passed in: the compiler then knows every pointer below is shared)
    const int tid = threadIdx.x, l = tid & 31;
    const int slot = l / 6, t = (tid >> 5) * 5 + slot;
    const bool valid = l < 30 && t < nt;
    const unsigned mask = __ballot_sync(0xffffffffu, valid);
    if (!valid) return;
    const FusedTail<O> T(smem, p.ns, p.tpc, p.nit, p.traj_stride);
    const int side = (l % 6) / 3, ax = l % 3;
    const FStateRows<O, FUSED_SMEM_LANES> st{T.state1 + 2 * t};
    const FusedSmem<O> L(p.ns);
    const FBaseRows<O> base_at{smem + t * p.traj_stride + L.oBase};
    const int n = p.ns - 1;
    bool ok = true;
    if (n > 0) {
        const int m = split_row(n, true);
        ok = elim_half_ax<O>(n, side ? n - 1 - m : m, side != 0, 0.0, ax, base_at, st);
        __syncwarp(mask);
        if (side == 0) ok = elim_middle_ax<O>(n, m, 0.0, ax, base_at, st) && ok;
        __syncwarp(mask);
        back_half_ax<O>(n, m, side != 0, ax, st);
    }
    const unsigned bad = __ballot_sync(mask, !ok);
    if (l % 6 == 0) T.ok1[t] = ((bad >> (6 * slot)) & 0x3fu) ? 0 : 1;
}

// pass 1 (ms.cpp:357-405), lane-pair form: one lane pair per trajectory, solution left in the shared-memory state rows.  Used by the
// many-wave build (three CTAs per SM), where the FP64 pipe is the bound and the six-lane form's threefold factorisation costs
// more than its shorter rows save (cfg3: 8.18 -> 8.30 ms per batch with it).
template <int O>
__device__ __noinline__ void fused_phase_pass1_pair(const FusedParams &p, int nt) {
    extern __shared__ double smem[];  // (declared here, not passed in: the compiler then knows every pointer below is shared)
    const int tid = threadIdx.x;
    if (tid >= 2 * nt) return;
    const FusedTail<O> T(smem, p.ns, p.tpc, p.nit, p.traj_stride);
    double unused;
    const int t = tid >> 1;
    const FStateRows<O, FUSED_SMEM_LANES> st{T.state1 + 2 * t};
    const unsigned pm = 2 * nt >= 32 ? 0xffffffffu : (1u << (2 * nt)) - 1u;
    const bool ok = fused_chain_pair<O, false>(smem + t * p.traj_stride, p.ns, 0.0, st, tid & 1, pm, &unused);
    if ((tid & 1) == 0) T.ok1[t] = ok ? 1 : 0;
}

// Speculative Thomas over (trajectory, reweighting iteration), see the kernel.
template <int O>
__device__ __noinline__ void fused_phase_spec(const FusedParams &p, long long tile, int nt) {
    extern __shared__ double smem[];
    constexpr int SL = FUSED_SMEM_LANES, GL = FUSED_SLOT_LANES;
    using D = Dim<O>;
    const int tid = threadIdx.x;
    const int ns = p.ns, nit = p.nit, tstride = p.traj_stride;
    const FusedTail<O> T(smem, ns, p.tpc, nit, tstride);
    const bool use_pw = p.sp.pw > 0.0;
    const int n_glob = nt * (nit - 1);
    const int g0l = (2 * p.tpc + 31) & ~31;  // first speculative lane
    if (tid >= g0l && tid < g0l + n_glob) {
        const int l = tid - g0l;
        const int t = l / (nit - 1), q = l - t * (nit - 1);
        const double vw = reweighted_vw(p.sp.vw0, q);
        const double add00 = vw > 0.0 ? 2.0 * vw : 0.0;
        double *slot = p.state_ws + (size_t)blockIdx.x * (ns - 1) * D::NSTATE * GL;
        const FSlotRows<O, GL> st{slot + 2 * l};
        long long *clk = nullptr;  // dev instrumentation: speculative lanes 0, 64 and 128 of the CTA's first tile
        if (p.phase_clocks && tile == blockIdx.x && (l & 63) == 0) clk = p.phase_clocks + blockIdx.x * 16 + 10 + 2 * (l >> 6);
        const ChainResult r = fused_chain_spec<O>(smem + t * tstride, ns, add00, st, clk);
        T.md[t * nit + q] = r.max_dev;
        T.okf[t * nit + q] = r.ok ? 1 : 0;
    } else if (tid < 2 * nt) {  // two lanes per trajectory (fused_chain_pair)
        const int t = tid >> 1, q = nit - 1;
        const double vw = reweighted_vw(p.sp.vw0, q);
        const double add00 = vw > 0.0 ? 2.0 * vw : 0.0;
        const FStateRows<O, SL> st{T.state1 + 2 * t};
        const unsigned pm = 2 * nt >= 32 ? 0xffffffffu : (1u << (2 * nt)) - 1u;
        double mdv;
        const bool ok = use_pw ? fused_chain_pair<O, true>(smem + t * tstride, ns, add00, st, tid & 1, pm, &mdv)
                               : fused_chain_pair<O, false>(smem + t * tstride, ns, add00, st, tid & 1, pm, &mdv);
        if ((tid & 1) == 0) {
            T.md[t * nit + q] = mdv;
            T.okf[t * nit + q] = ok ? 1 : 0;
        }
        if (p.phase_clocks && tile == blockIdx.x && tid == 0) p.phase_clocks[blockIdx.x * 16 + 9] = clock64();  // pair chain done
    }
#ifdef MSNAP_WARP_END_STAMPS  // dev instrumentation: when does every warp leave the phase (replaces the lane stamps)
    __syncwarp();
    if (p.phase_clocks && tile == blockIdx.x && (tid & 31) == 0 && (tid >> 5) >= 1 && (tid >> 5) <= 6)
        p.phase_clocks[blockIdx.x * 16 + 9 + (tid >> 5)] = clock64();
#endif
}

template <int O, int NT = FUSED_THREADS, int MINB = 2>
__global__ void __launch_bounds__(NT, MINB) k_fused_solve(const __grid_constant__ FusedParams p) {
    using D = Dim<O>;
    constexpr int SL = FUSED_SMEM_LANES, GL = FUSED_SLOT_LANES;
    extern __shared__ double smem[];
    const int tid = threadIdx.x;
    const int ns = p.ns, nr = ns - 1, tpc = p.tpc, nit = p.nit, tstride = p.traj_stride;
    const FusedSmem<O> L(ns);
    double *state1 = smem + (size_t)tpc * tstride;            // [nr][NSTATE][SL]: pass-1 sweep, then the final x
    double *md = state1 + (size_t)nr * D::NSTATE * SL;        // [tpc][nit] max deviation of every speculative solve
    int *okf = reinterpret_cast<int *>(md + tpc * nit);       // [tpc][nit] pivot status
    int *sel = okf + tpc * nit;                               // [tpc] selected iteration
    int *ok1 = sel + tpc;                                     // [tpc] pass-1 pivot status
    double *slot = p.state_ws + (size_t)blockIdx.x * nr * D::NSTATE * GL;
    const bool use_pw = p.sp.pw > 0.0;
    int stamp = 0;
#define MSNAP_STAMP()                                                                              \
    do {                                                                                           \
        if (p.phase_clocks && tid == 0 && tile == blockIdx.x && stamp < 10)                        \
            p.phase_clocks[blockIdx.x * 16 + stamp++] = clock64();                                 \
    } while (0)

    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const long long b0 = tile * tpc;
        const int nt = (int)min((long long)tpc, p.B - b0);  // trajectories in this tile
        const long long g0 = b0 * ns;                       // first global segment of the tile
        MSNAP_STAMP();
        // ---- load: waypoints of the tile (contiguous in HBM, [w][3] per trajectory) and boundary derivatives
        {
            const double *src = p.wp + 3 * (g0 + b0);
            const int per = 3 * (ns + 1), n = nt * per;
            for (int i = tid; i < n; i += NT) {
                const int t = i / per;
                smem[t * tstride + L.oP + (i - t * per)] = src[i];
            }
            // fixed boundary derivatives (velocity, acceleration; higher ones zero -- ms.cpp:526-555), [r-1][axis]
            for (int i = tid; i < nt * 2 * D::NR; i += NT) {
                const int t = i / (2 * D::NR), r = i - t * 2 * D::NR;
                const int end = r / D::NR, d = (r % D::NR) / 3 + 1, a = r % 3;  // derivative order d
                double v = 0.0;
                if (d == 1) v = p.sp.vel ? p.sp.vel[6 * (b0 + t) + 3 * end + a] : p.sp.bc[3 * end + a];
                if (d == 2) v = p.sp.acc ? p.sp.acc[6 * (b0 + t) + 3 * end + a] : p.sp.bc[6 + 3 * end + a];
                smem[t * tstride + L.oBC + r] = v;
            }
        }
        __syncthreads();
        MSNAP_STAMP();
        // ---- times (plain IEEE mul/add: bit-identical to the reference's allocation, ms.cpp:63-72) and powers
        for (int i = tid; i < nt * ns; i += NT) {
            const int t = i / ns, k = i - t * ns;
            double *blk = smem + t * tstride;
            double Tk;
            if (p.times_in) {
                Tk = p.times_in[g0 + i];
            } else {
                const double *P = blk + L.oP + 3 * k;
                const double dx = __dsub_rn(P[3], P[0]), dy = __dsub_rn(P[4], P[1]), dz = __dsub_rn(P[5], P[2]);
                const double len =
                    __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
                Tk = (p.v_avg > 1e-6) ? __ddiv_rn(len, p.v_avg) : p.min_time;
                if (Tk < p.min_time) Tk = p.min_time;
                if (p.times_out) p.times_out[g0 + i] = Tk;
            }
            double *seg = blk + L.oSeg + k * L.ST;
            seg[0] = Tk;
            seg[1] = 1.0 / Tk;
            if (!use_pw) {
                reinterpret_cast<int *>(blk + L.oS)[k] = 0;
                if (p.best_s_out) p.best_s_out[g0 + i] = 0;
            }
        }
        __syncthreads();
        MSNAP_STAMP();

        if (use_pw) {
            // ---- pass 1: snap cost only -> worst-deviation sample per segment
            for (int i = tid; i < nt * nr; i += NT)
                fused_row_item<O>(p, smem + (i / nr) * tstride, ns, i % nr + 1, false);
            __syncthreads();
            MSNAP_STAMP();
            if constexpr (MINB >= 3) fused_phase_pass1_pair<O>(p, nt);
            else fused_phase_pass1<O>(p, nt);
            __syncthreads();
            MSNAP_STAMP();
            for (int i = tid; i < nt * ns; i += NT) {
                const int t = i / ns, k = i - t * ns;
                const int s = fused_search_item<O>(smem + t * tstride, state1 + 2 * t, ns, k);
                reinterpret_cast<int *>(smem + t * tstride + L.oS)[k] = s;
                if (p.best_s_out) p.best_s_out[g0 + i] = s;
            }
            __syncthreads();
            MSNAP_STAMP();
            for (int i = tid; i < nt * ns; i += NT) fused_probe_item<O>(p, smem + (i / ns) * tstride, ns, i % ns);
        }
        // ---- rows of the final system
        for (int i = tid; i < nt * nr; i += NT)
            fused_row_item<O>(p, smem + (i / nr) * tstride, ns, i % nr + 1, use_pw);
        __syncthreads();
        MSNAP_STAMP();
        // ---- speculative Thomas over (trajectory, reweighting iteration).
        // nt*(nit-1) speculative lanes solve iterations 0..nit-2 with their sweep state in the CTA's global scratch slot
        // and keep only the max deviation.  The LAST iteration -- the one most trajectories end on -- is solved by a
        // separate warp-aligned group of nt lanes in the shared-memory state rows (idle since the search) and leaves
        // its x there, ready for the coefficient phase.  With a single solve per trajectory (nit == 1: no penalty, or
        // a bare SolveQPClosedForm) only that group runs.  Keeping the two groups in different warps avoids
        // executing both code paths in every warp.
        {
            // The last-iteration group sits in warp 0 .. (the CTA's oldest warps, which the scheduler favours): its
            // full backward sweep is the longest chain of this phase.  Speculative lanes start at the next warp boundary.
            fused_phase_spec<O>(p, tile, nt);
            __syncthreads();
            // the iteration the sequential loop would have stopped at (ms.cpp:82)
            if (tid < nt) {
                int q = 0;
                while (md[tid * nit + q] > 0.2 && q < nit - 1) ++q;
                sel[tid] = q;
            }
            __syncthreads();
            // a trajectory that stopped early: the lane that solved that iteration replays its backward sweep, this
            // time leaving the solution in the shared-memory state rows
            const int n_glob = nt * (nit - 1), g0l = (2 * tpc + 31) & ~31;
            if (tid >= g0l && tid < g0l + n_glob) {
                const int l = tid - g0l;
                const int t = l / (nit - 1), q = l - t * (nit - 1);
                if (sel[t] == q) {
                    const FSlotRows<O, GL> st{slot + 2 * l};
                    const FStateRows<O, SL> xo{state1 + 2 * t};
                    fused_replay<O>(smem + t * tstride, ns, st, xo);
                }
            }
            __syncthreads();
        }
        MSNAP_STAMP();
        if (tid < nt) {
            const int q = sel[tid];
            double vw = p.sp.vw0;
            for (int i = 0; i < q; ++i) vw = (vw < 1e-6) ? 0.01 : vw * 2.0;
            const double mdv = md[tid * nit + q];
            const long long b = b0 + tid;
            if (p.max_dev_out) p.max_dev_out[b] = mdv;
            if (p.iters_out) p.iters_out[b] = q;
            if (p.vw_final_out) p.vw_final_out[b] = vw;
            if (p.flags) {
                const bool bad = !okf[tid * nit + q] || !(mdv == mdv) || (use_pw && !ok1[tid]);
                p.flags[b] = bad ? 1u : 0u;
            }
        }
        __syncthreads();  // flags[b] is OR-ed below
        // The speculative sweep state in this CTA's scratch slot is dead from here on: tell L2 to drop the lines instead
        // of writing them back to HBM when they are evicted (the slot is rewritten before it is read again).
        if (nit > 1 && p.discard_state) {
            char *sb = reinterpret_cast<char *>(slot);
            const int n_lines = (int)(((size_t)nr * D::NSTATE * GL * sizeof(double)) / 128);
            for (int i = tid; i < n_lines; i += NT)
                asm volatile("discard.global.L2 [%0], 128;" ::"l"(sb + (size_t)i * 128) : "memory");
        }
        // ---- coefficients: items (t, k, axis), 64-byte rows straight to HBM
        for (int i = tid; i < nt * ns * 3; i += NT) {
            const int t = i / (ns * 3), r = i - t * ns * 3;
            const int k = r / 3, a = r - 3 * k;
            const long long row = ((g0 + (long long)t * ns + k) * 3 + a) * D::M;
            const bool finite = fused_coeff_item<O>(smem + t * tstride, state1 + 2 * t, ns, k, a, p.coeff_out + row,
                                                    p.coeff_mirror ? p.coeff_mirror + row : nullptr);
            if (!finite && p.flags) atomicOr(p.flags + b0 + t, 1u);
        }
        __syncthreads();  // the tile's smem and state slot are reused by the next tile
        MSNAP_STAMP();
    }
#undef MSNAP_STAMP
}

}  // namespace msnap

#endif  // MSNAP_FUSED_CUH
