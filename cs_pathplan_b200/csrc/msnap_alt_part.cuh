// msnap_alt_part.cuh -- altitude optimisation, partitioned form (msnap_set_altitude_policy 2, the default): the same SPD
// pentadiagonal systems as msnap_alt.cuh (optimizeHeights cpp:1575-1712, optimizeHeightsGlobalSmooth cpp:1714-1827), every
// solve spread over up to 16 lanes (32 for long trajectories) and the whole stage -- edge weights, follow targets, both
// passes, the active-set loop, the write-back -- in ONE launch.
//
// Why: k_alt_solve_pair walks ~100 dependent rows per lane and sweep and spends most of its ~75 instructions per row on
// moving rows between global memory and its staging tiles; with one warp per scheduler the kernel's time is that
// instruction stream.  Here a trajectory's rows are cut into P contiguous chunks
//     chunk_0 | sep_0 | chunk_1 | sep_1 | ... | chunk_{P-1}          (sep_p = 2 rows: the half bandwidth)
// Lane p eliminates the interior rows of chunk p by the plain downward LDL' recurrence (alt_fwd_row's arithmetic), carrying
// two extra columns V = L^-1 H[chunk_p, sep_{p-1}] (the chunk's coupling to the separator above it), and continues the
// recurrence into sep_p without eliminating it.  What is left is an SPD block-tridiagonal system in the separators
// (2 x 2 blocks, P-1 of them):
//     A_p = H[sep_p, sep_p] - (from chunk p above: the continued recurrence) - (from chunk p+1 below: V' D^-1 V)
//     C_p = coupling sep_p / sep_{p-1} = the V columns continued into sep_p
// solved across the lanes of the group by shuffles (block Thomas from both ends, ~P/2 dependent steps each way), after which every lane substitutes
// back through its chunk:  z_k = y_k/D_k - V_k/D_k . z[sep_{p-1}] - L[k+1,k] z_{k+1} - L[k+2,k] z_{k+2}.
// This is Gaussian elimination of the same SPD matrix in a nested-dissection order (the reference's SimplicialLDLT uses a
// fill-reducing order of its own), so the heights agree with the other two forms to rounding and the active-set decisions
// are the same (tests/test_gpu_alt.py).  Dependent steps per solve: ~n/16 + 16 instead of n/2, and ~50 instructions per
// row because every per-row field lives in shared memory, laid out [field][slot][lane] -- a lane's rows are private to it,
// addresses are base + immediate, nothing is staged and nothing synchronises inside a sweep.
//
// Shared memory: 7 fields; the three a separator row needs (W, E, YD) have 17 slots per lane, the four only interior rows
// use (L1, L2, V1, V2) have 15: 111 x 32 lanes x 8 B = 27.75 KB per warp, one warp per CTA, SEVEN CTAs per SM; a group of
// 16 lanes holds trajectories of up to 16 * 15 + 30 = 270 rows (measured: 8 lanes x 31 rows, four CTAs per SM, is 8 % slower
// -- fewer warps per scheduler).  Longer trajectories are taken one at a time by all 32 lanes of the warp with the same code:
// up to 32 * 15 + 62 = 542 rows in the same shared memory, beyond that with the fields in the caller's global scratch
// arrays (natural row order) instead.
#ifndef MSNAP_ALT_PART_CUH
#define MSNAP_ALT_PART_CUH

#include "msnap_alt.cuh"

namespace msnap {

constexpr int ALTP_CMAX = 15;              // interior rows of a chunk that fit the shared-memory layout
constexpr int ALTP_SLOTS = ALTP_CMAX + 2;  // + the lane's two separator rows
constexpr int ALTP_GROUP = 16;             // lanes per trajectory in the shared-memory form
constexpr int ALTP_NMAX = ALTP_GROUP * ALTP_CMAX + 2 * (ALTP_GROUP - 1);  // 270 rows
enum AltPartField { AF_W = 0, AF_E, AF_YD, AF_L1, AF_L2, AF_V1, AF_V2, AF_COUNT };
// first slot of a field: W, E, YD hold every row of the lane (ALTP_SLOTS), the factor fields interior rows only (ALTP_CMAX)
__host__ __device__ constexpr int altp_off(int f) { return f <= AF_L1 ? f * ALTP_SLOTS : AF_L1 * ALTP_SLOTS + (f - AF_L1) * ALTP_CMAX; }
constexpr int ALTP_SLOT_ROWS = altp_off(AF_COUNT);  // 111
constexpr size_t ALTP_SMEM_BYTES = (size_t)ALTP_SLOT_ROWS * 32 * sizeof(double);
// staging homes of the raw x / y columns during the load (ALTP_SLOTS + 1 entries each, over fields that are still unused)
constexpr int ALTP_STAGE_X = altp_off(AF_L1), ALTP_STAGE_Y = altp_off(AF_V1);
// bulk-copy staging (doubles per group of lanes): rows [3 n + 1] in the L1 .. V2 slots, elevations [n + 1] in the YD slots
__host__ __device__ constexpr int altp_stage_r(int gw) { return ((ALTP_SLOT_ROWS - ALTP_STAGE_X) * 32 / (32 / gw)) & ~1; }
__host__ __device__ constexpr int altp_stage_e(int gw) { return (ALTP_SLOTS * 32 / (32 / gw)) & ~1; }
constexpr int ALTP_NMAX32 = 32 * ALTP_CMAX + 2 * 31;  // 542 rows: what all 32 lanes hold in shared memory for ONE trajectory
static_assert(3 * ALTP_NMAX + 2 <= altp_stage_r(ALTP_GROUP) && ALTP_NMAX + 2 <= altp_stage_e(ALTP_GROUP), "bulk staging does not fit");
static_assert(3 * ALTP_NMAX32 + 2 <= altp_stage_r(32) && ALTP_NMAX32 + 2 <= altp_stage_e(32), "bulk staging does not fit");
static_assert(ALTP_STAGE_X + ALTP_SLOTS + 1 <= ALTP_STAGE_Y && ALTP_STAGE_Y + ALTP_SLOTS + 1 <= ALTP_SLOT_ROWS, "staging overlaps");

// Per-row fields of a lane's chunk (local row i = 0 .. len-1; the separator rows come last):
//   W   pass 1: climb weight of the edge (k, k+1), negative when pass 2 has no such edge (see alt_part_weight);
//       pass 2: that pass's climb weight, sign bit set once the row has joined the active set
//   E   terrain elevation (NaN: none); from the back substitution of pass 1 on: the pass-1 height z1
//   YD  before a row's elimination in pass 1: its follow target (NaN: the map has no value here); then y_k / D_k; z_k after
//       a back substitution
//   L1, L2       L[k,k-1], L[k,k-2]
//   V1, V2       the chunk's coupling columns, V_k / D_k
template <bool GLOBAL>
struct AltPartFld;
template <>
struct AltPartFld<false> {  // shared memory, [field][slot][lane]
    double *q;              // alt_sm + lane
    __device__ __forceinline__ AltPartFld at(long long) const { return *this; }
    __device__ __forceinline__ double ld(int f, int i) const { return q[(altp_off(f) + i) * 32]; }
    __device__ __forceinline__ void st(int f, int i, double v) const { q[(altp_off(f) + i) * 32] = v; }
};
template <>
struct AltPartFld<true> {  // global scratch arrays, natural row order: field f of row k at g[f][k]
    double *g[AF_COUNT];
    __device__ __forceinline__ AltPartFld at(long long first_row) const {
        AltPartFld r;
#pragma unroll
        for (int f = 0; f < AF_COUNT; ++f) r.g[f] = g[f] + first_row;
        return r;
    }
    __device__ __forceinline__ double ld(int f, int i) const { return g[f][i]; }
    __device__ __forceinline__ void st(int f, int i, double v) const { g[f][i] = v; }
};

// The climb weights of both passes from one stored number.  Pass 1 uses 1 / (dist * r)^2, pass 2 1 / (dist * r/2)^2
// (cpp:1649-1665, 1759-1775 with max_climb_rate * 0.5, cpp:1354): halving r and squaring are exact, so the second is
// 4 x the first bit for bit.  The two passes test their own `dist * rate > 1e-12` guard; an edge that passes the first
// and fails the second is stored negated.
template <int PASS>
__device__ __forceinline__ double alt_part_weight(double enc) {
    return PASS == 1 ? fabs(enc) : (enc > 0.0 ? 4.0 * enc : 0.0);
}

// optimizeSegmentAltitudeENU for the trajectories of one warp: lane (g, pp) = (lane / GW, lane % GW) works on the
// trajectory of group g (`valid`, `base`, `n` are group-uniform).  Returns group-uniform results.
template <int GW, bool GLOBAL>
__device__ __forceinline__ void alt_part_set(const AltParams &p, const AltPartFld<GLOBAL> fields, int lane, bool valid,
                                             long long base, int n, double *rows, const double *elev, double *z_pass1_out,
                                             long long n_cap, unsigned mbar, unsigned &phase, int &solves_ret, bool &ok_ret,
                                             bool &ok2_ret) {
    constexpr unsigned FULL = 0xffffffffu;
    const int pp = lane & (GW - 1);
    const unsigned gmask = GW == 32 ? FULL : (((1u << GW) - 1u) << (lane & ~(GW - 1)));
    auto group_any = [&](bool v) { return (__ballot_sync(FULL, v) & gmask) != 0u; };
    // chunk geometry: P partitions, every interior at least 2 rows long
    if (!valid || n <= 0) n = 0;
    const int P = n >= 8 ? (n / 4 < GW ? n / 4 : GW) : 1;
    const int m = n - 2 * (P - 1), bs = m / P, rem = m - bs * P;
    const bool mine = n > 0 && pp < P;
    const int cnt = mine ? bs + (pp < rem ? 1 : 0) : 0;                   // interior rows of this lane
    const int start = pp * (bs + 2) + (pp < rem ? pp : rem);              // trajectory row of local row 0
    const bool hr = mine && pp < P - 1, hl = mine && pp > 0;              // a separator below / above this chunk
    const int len = cnt + (hr ? 2 : 0);
    const int Pmax = __reduce_max_sync(FULL, mine ? P : 0);
    const AltPartFld<GLOBAL> F = fields.at(base + start);
    const double lf = p.lambda_follow, safe = p.safe_distance;

    // ---- load: edge weights (k_alt_prep's arithmetic), follow targets, elevations of this lane's rows.
    // Shared-memory form: a trajectory's rows (24 n contiguous bytes) and elevations (8 n) come in by ONE bulk copy each
    // (cp.async.bulk, issued by the group's first lane, completion counted on the warp's mbarrier) into the still unused
    // factor / YD slots in natural order; every lane then picks its own rows out of shared memory.  A bulk copy moves
    // 16-byte units between 16-byte aligned addresses: the window is widened to the next boundaries (never beyond the
    // caller's arrays: a last odd element is fetched by an ordinary load instead).  Arrays that are not 16-byte aligned
    // themselves take the per-lane path (cp.async, 8 bytes a copy, every copy of the lane in flight at once).
    double wlast = 0.0;
    {
        const double *r = rows + 3 * (base + start);
        const double *sx = r, *sy = r + 1, *su = r + 2;  // x, y, up of local row i at s?[i * sstride]
        int sstride = 3;
        if constexpr (!GLOBAL) {
            double *const sm0 = F.q - lane;
            const bool bulk = ((reinterpret_cast<unsigned long long>(rows) | reinterpret_cast<unsigned long long>(elev)) & 15ull) == 0ull;
            if (bulk) {
                double *Rg = sm0 + ALTP_STAGE_X * 32 + (lane / GW) * altp_stage_r(GW);     // rows of this group's trajectory
                double *Eg = sm0 + altp_off(AF_YD) * 32 + (lane / GW) * altp_stage_e(GW);  // its elevations
                const long long e0 = 3 * base, e1 = 3 * (base + n), c0 = base, c1 = base + n;
                const long long f0 = e0 & ~1ll, g0 = c0 & ~1ll;
                long long f1 = (e1 + 1) & ~1ll, g1 = (c1 + 1) & ~1ll;
                if (f1 > 3 * n_cap) f1 = e1 & ~1ll;
                if (g1 > n_cap) g1 = c1 & ~1ll;
                const int shr = (int)(e0 - f0), she = (int)(c0 - g0);
                // (a later use of the tile: order the warp's ordinary accesses before the copy engine's writes)
                __syncwarp();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                if (lane % ALTP_GROUP == 0) {  // the barrier counts 32 / ALTP_GROUP arrivals per use, whatever GW is
                    const unsigned bytes_r = (pp == 0 && n > 0) ? (unsigned)((f1 - f0) * 8) : 0u;
                    const unsigned bytes_e = (pp == 0 && n > 0 && elev) ? (unsigned)((g1 - g0) * 8) : 0u;
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes_r + bytes_e) : "memory");
                    if (bytes_r)
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                         (unsigned)__cvta_generic_to_shared(Rg)),
                                     "l"(rows + f0), "r"(bytes_r), "r"(mbar)
                                     : "memory");
                    if (bytes_e)
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                         (unsigned)__cvta_generic_to_shared(Eg)),
                                     "l"(elev + g0), "r"(bytes_e), "r"(mbar)
                                     : "memory");
                }
                {
                    unsigned done = 0;
                    while (!done)
                        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                     : "=r"(done)
                                     : "r"(mbar), "r"(phase & 1u)
                                     : "memory");
                    ++phase;
                }
                if (pp == 0 && n > 0) {  // an odd last element the window could not cover
                    if (f1 < e1) Rg[shr + 3 * n - 1] = rows[e1 - 1];
                    if (elev && g1 < c1) Eg[she + n - 1] = elev[c1 - 1];
                }
                __syncwarp();
                for (int i = 0; i < len; ++i) F.st(AF_E, i, elev ? Eg[she + start + i] : NAN);
                __syncwarp();  // the elevations' staging area is YD: free it before the follow targets land there
                sx = Rg + shr + 3 * start;
                sy = sx + 1;
                su = sx + 2;
            } else {
                const bool more = mine && start + len < n;  // the row after this lane's last one (its x, y close the last edge)
                for (int i = 0; i < len; ++i) {
                    __pipeline_memcpy_async(F.q + (ALTP_STAGE_X + i) * 32, r + 3 * i, sizeof(double));
                    __pipeline_memcpy_async(F.q + (ALTP_STAGE_Y + i) * 32, r + 3 * i + 1, sizeof(double));
                    __pipeline_memcpy_async(F.q + (altp_off(AF_YD) + i) * 32, r + 3 * i + 2, sizeof(double));
                    if (elev) __pipeline_memcpy_async(F.q + (altp_off(AF_E) + i) * 32, elev + base + start + i, sizeof(double));
                }
                if (more) {
                    __pipeline_memcpy_async(F.q + (ALTP_STAGE_X + len) * 32, r + 3 * len, sizeof(double));
                    __pipeline_memcpy_async(F.q + (ALTP_STAGE_Y + len) * 32, r + 3 * len + 1, sizeof(double));
                }
                __pipeline_commit();
                __pipeline_wait_prior(0);
                if (!elev)
                    for (int i = 0; i < len; ++i) F.st(AF_E, i, NAN);
                sx = F.q + ALTP_STAGE_X * 32;
                sy = F.q + ALTP_STAGE_Y * 32;
                su = F.q + altp_off(AF_YD) * 32;
                sstride = 32;
            }
        }
        auto rx = [&](int i) { return sx[i * sstride]; };
        auto ry = [&](int i) { return sy[i * sstride]; };
        double x0 = 0.0, y0 = 0.0;
        if (len > 0) {
            x0 = rx(0);
            y0 = ry(0);
        }
#pragma unroll 2
        for (int i = 0; i < len; ++i) {
            const int k = start + i;
            const double up = su[i * sstride];
            double enc = 0.0, x1 = 0.0, y1 = 0.0;
            if (k + 1 < n) {
                x1 = rx(i + 1);
                y1 = ry(i + 1);
                // std::hypot (cpp:1650): a square root of the sum of squares unless that could over- or underflow
                const double dx = x1 - x0, dy = y1 - y0, q2 = fma(dx, dx, dy * dy);
                const double dist = (q2 > 1e-280 && q2 < 1e280) ? sqrt(q2) : hypot(dx, dy);
                if (dist > 1e-9) {  // cpp:1655, 1765
                    const double d1 = dist * p.max_climb_rate, d2 = dist * (p.max_climb_rate * 0.5);
                    if (p.max_climb_rate > 0.0 && d1 > 1e-12) {
                        const double dd = d1 * d1;
                        const double a1 = (dd > 1e-280 && dd < 1e280) ? alt_rcp(dd) : 1.0 / dd;
                        enc = (p.max_climb_rate * 0.5 > 0.0 && d2 > 1e-12) ? a1 : -a1;
                    }
                }
            }
            double el = NAN;
            if (elev) {
                if constexpr (GLOBAL) el = elev[base + k];
                else el = F.ld(AF_E, i);
            }
            F.st(AF_W, i, enc);
            F.st(AF_YD, i, el == el ? fmax(up, el + safe) : NAN);  // the follow target, cpp:1637-1638
            if (GLOBAL) F.st(AF_E, i, el);
            x0 = x1;
            y0 = y1;
            wlast = enc;
        }
        if constexpr (!GLOBAL) __syncwarp();  // the staged rows lie in the factor slots the sweeps are about to fill
    }
    const double wprev = __shfl_up_sync(FULL, wlast, 1, GW);  // edge (start-1, start): the last row of the lane above

    bool violation = false;
    // One solve.  PASS 1: optimizeHeights' system; PASS 2: optimizeHeightsGlobalSmooth's with the current active set.
    // Returns this lane's pivot status (combine with group_any).
    auto solve = [&](auto pass_tag, double s, bool active) -> bool {
        constexpr int PASS = decltype(pass_tag)::value;
        const double s_eff = (n >= 3 && s > 0.0) ? s : 0.0;
        const double x_end = n == 1 ? 2.0 * ALT_FIX_WEIGHT : ALT_FIX_WEIGHT;  // (pass 2) penalty weight of rows 0 and n-1
        // s * [row j is interior: 1 <= j <= n-2], cpp:1588-1604 (one unsigned compare)
        const unsigned n_int = n > 2 ? (unsigned)(n - 2) : 0u;
        auto in = [&](int j) { return (unsigned)(j - 1) < n_int ? s_eff : 0.0; };
        auto coef = [&](int k, int i, double &extra, double &rhs) {  // what the pass adds to H[k,k] and b[k]
            if constexpr (PASS == 1) {
                const double t = F.ld(AF_YD, i);
                const bool has = t == t;
                extra = has ? lf : 0.0;
                rhs = has ? lf * t : 0.0;
            } else {
                // 1e10 per end the row is (cpp:1779-1784), 1e8 on interior rows of the active set (cpp:1787-1793)
                const bool interior = (unsigned)(k - 1) < n_int;
                const double x = interior ? (signbit(F.ld(AF_W, i)) ? ALT_CON_WEIGHT : 0.0) : x_end;
                extra = x;
                rhs = x * F.ld(AF_E, i);
            }
        };
        auto sink = [&](int k, int i, double z) {  // what the pass does with z_k
            if constexpr (PASS == 1) {
                const double el = F.ld(AF_E, i);
                if (el == el && z < el + safe) z = el + safe;  // cpp:1705-1707
                F.st(AF_E, i, z);
                if (z_pass1_out) z_pass1_out[base + k] = z;
            } else {
                F.st(AF_YD, i, z);
                if (z < F.ld(AF_E, i) - ALT_VIOLATION) {  // cpp:1805-1810
                    const double w = F.ld(AF_W, i);
                    if (!signbit(w)) {  // not in the active set yet
                        F.st(AF_W, i, -w);
                        violation = true;
                    }
                }
            }
        };
        const int cn = active ? cnt : 0;
        const bool a_hr = active && hr, a_hl = active && hl;
        // recurrence state: a1 = L[k,k-1], a2 = L[k,k-2], cc = L[k+1,k-1], D and y (= L^-1 b) of the two rows above,
        // the two coupling columns of the two rows above, the stencil flags of rows k-1 and k, the edge weight above
        double a1 = 0.0, a2 = 0.0, cc = 0.0, Dm1 = 0.0, Dm2 = 0.0, ym1 = 0.0, ym2 = 0.0;
        double v1m1 = 0.0, v1m2 = 0.0, v2m1 = 0.0, v2m2 = 0.0;
        double sm_ = in(start - 1), s0_ = in(start);
        double wm1 = a_hl ? alt_part_weight<PASS>(wprev) : 0.0;  // (wprev stays encoded; the stored W is converted after pass 1)
        // H[start, start-2], H[start, start-1], H[start+1, start-1]: the chunk's coupling to the separator above
        const double hL0 = a_hl ? sm_ : 0.0, eL1 = a_hl ? -2.0 * (sm_ + s0_) - wm1 : 0.0, hL1 = a_hl ? s0_ : 0.0;
        double q11 = 0.0, q12 = 0.0, q22 = 0.0, r1 = 0.0, r2 = 0.0;  // V' D^-1 V and V' D^-1 y
        double H1 = hL0, H2 = eL1, H2n = hL1;  // direct couplings of local rows 0 and 1 to the separator above
        bool ok = true;
#pragma unroll 4
        for (int i = 0; i < cn; ++i) {
            const int k = start + i;
            const double wk = fabs(F.ld(AF_W, i));
            double extra, rhs;
            coef(k, i, extra, rhs);
            const double sp = in(k + 1);
            const double d = fma(4.0, s0_, sp + sm_) + (wm1 + wk) + extra + ALT_REG;  // alt_fwd_row's expressions
            const double e = -2.0 * (s0_ + sp) - wk;
            const double D = fma(-a2 * a2, Dm2, fma(-a1 * a1, Dm1, d));
            const double y = fma(-a2, ym2, fma(-a1, ym1, rhs));
            const double v1 = fma(-a2, v1m2, fma(-a1, v1m1, H1)), v2 = fma(-a2, v2m2, fma(-a1, v2m1, H2));
            H1 = 0.0;
            H2 = H2n;
            H2n = 0.0;
            ok = ok && D > 0.0 && D < 1e300;
            const double inv = alt_rcp(D);
            const double n1 = fma(-cc * Dm1, a1, e) * inv, n2 = sp * inv;
            const double yd = y * inv, vd1 = v1 * inv, vd2 = v2 * inv;
            q11 = fma(v1, vd1, q11);
            q12 = fma(v1, vd2, q12);
            q22 = fma(v2, vd2, q22);
            r1 = fma(v1, yd, r1);
            r2 = fma(v2, yd, r2);
            F.st(AF_L1, i, a1);
            F.st(AF_L2, i, a2);
            F.st(AF_YD, i, yd);
            F.st(AF_V1, i, vd1);
            F.st(AF_V2, i, vd2);
            a2 = cc;
            a1 = n1;
            cc = n2;
            Dm2 = Dm1;
            Dm1 = D;
            ym2 = ym1;
            ym1 = y;
            v1m2 = v1m1;
            v1m1 = v1;
            v2m2 = v2m1;
            v2m1 = v2;
            wm1 = wk;
            sm_ = s0_;
            s0_ = sp;
        }
        // the separator below: its two rows with this chunk's eliminations applied, not eliminated themselves
        double A00 = 1.0, A01 = 0.0, A11 = 1.0, g0 = 0.0, g1 = 0.0, C00 = 0.0, C01 = 0.0, C10 = 0.0, C11 = 0.0;
        if (a_hr) {
            int i = cn, k = start + cn;
            double extra, rhs;
            const double wk0 = fabs(F.ld(AF_W, i));
            coef(k, i, extra, rhs);
            const double sp0 = in(k + 1);
            const double d0 = fma(4.0, s0_, sp0 + sm_) + (wm1 + wk0) + extra + ALT_REG;
            const double e0 = -2.0 * (s0_ + sp0) - wk0;
            A00 = fma(-a2 * a2, Dm2, fma(-a1 * a1, Dm1, d0));
            g0 = fma(-a2, ym2, fma(-a1, ym1, rhs));
            C00 = -fma(a2, v1m2, a1 * v1m1);
            C01 = -fma(a2, v2m2, a1 * v2m1);
            A01 = e0 - cc * Dm1 * a1;
            ++i, ++k;
            const double wk1 = fabs(F.ld(AF_W, i));
            coef(k, i, extra, rhs);
            const double sp1 = in(k + 1);
            const double d1 = fma(4.0, sp0, sp1 + s0_) + (wk0 + wk1) + extra + ALT_REG;
            A11 = fma(-cc * cc, Dm1, d1);
            g1 = fma(-cc, ym1, rhs);
            C10 = -cc * v1m1;
            C11 = -cc * v2m1;
        }
        // ... and the contribution of the chunk below it
        {
            const double q11n = __shfl_down_sync(FULL, q11, 1, GW), q12n = __shfl_down_sync(FULL, q12, 1, GW),
                         q22n = __shfl_down_sync(FULL, q22, 1, GW), r1n = __shfl_down_sync(FULL, r1, 1, GW),
                         r2n = __shfl_down_sync(FULL, r2, 1, GW);
            if (a_hr) {
                A00 -= q11n;
                A01 -= q12n;
                A11 -= q22n;
                g0 -= r1n;
                g1 -= r2n;
            }
        }
        // Block Thomas over the separators of the group (block t on lane t, nb = P - 1 blocks), eliminated from BOTH ends
        // towards the middle block m = nb / 2 and substituted back outwards: nb / 2 dependent steps each way instead of nb.
        // A lane above the middle couples to its upper neighbour through its own C, a lane below it to its lower neighbour
        // through the transpose of that neighbour's C (Kb); the back substitution uses the other one of the two.
        const int nb = P - 1, mb = nb / 2, nbmax = Pmax - 1, mbmax = nbmax / 2;
        const bool top = pp < mb, bot = pp > mb, mid = a_hr && pp == mb;
        const double Kb00 = __shfl_down_sync(FULL, C00, 1, GW), Kb01 = __shfl_down_sync(FULL, C10, 1, GW),
                     Kb10 = __shfl_down_sync(FULL, C01, 1, GW), Kb11 = __shfl_down_sync(FULL, C11, 1, GW);
        // A -= K B^-1 K', g -= K B^-1 bg: B, bg = the eliminated neighbour's block and right-hand side, K = the coupling
        auto absorb = [&](double K00, double K01, double K10, double K11, double B00, double B01, double B11, double bg0,
                          double bg1) {
            const double det = fma(B00, B11, -(B01 * B01));
            const double idet = alt_rcp(det);
            const double M00 = fma(K00, B11, -(K01 * B01)) * idet, M01 = fma(K01, B00, -(K00 * B01)) * idet;
            const double M10 = fma(K10, B11, -(K11 * B01)) * idet, M11 = fma(K11, B00, -(K10 * B01)) * idet;
            A00 -= fma(M00, K00, M01 * K01);
            A01 -= fma(M00, K10, M01 * K11);
            A11 -= fma(M10, K10, M11 * K11);
            g0 -= fma(M00, bg0, M01 * bg1);
            g1 -= fma(M10, bg0, M11 * bg1);
        };
        {
            const int src = top ? (pp > 0 ? pp - 1 : 0) : (pp + 1 < GW ? pp + 1 : GW - 1);  // the neighbour away from the middle
            const double K00 = top ? C00 : Kb00, K01 = top ? C01 : Kb01, K10 = top ? C10 : Kb10, K11 = top ? C11 : Kb11;
            const int smax = mbmax - 1 > nbmax - 2 - mbmax ? mbmax - 1 : nbmax - 2 - mbmax;
            for (int st = 1; st <= smax; ++st) {
                const double B00 = __shfl_sync(FULL, A00, src, GW), B01 = __shfl_sync(FULL, A01, src, GW),
                             B11 = __shfl_sync(FULL, A11, src, GW), bg0 = __shfl_sync(FULL, g0, src, GW),
                             bg1 = __shfl_sync(FULL, g1, src, GW);
                if (a_hr && ((top && pp == st) || (bot && pp == nb - 1 - st))) absorb(K00, K01, K10, K11, B00, B01, B11, bg0, bg1);
            }
        }
        {  // the middle block takes both sides
            const int up = pp > 0 ? pp - 1 : 0, dn = pp + 1 < GW ? pp + 1 : GW - 1;
            double B00 = __shfl_sync(FULL, A00, up, GW), B01 = __shfl_sync(FULL, A01, up, GW), B11 = __shfl_sync(FULL, A11, up, GW),
                   bg0 = __shfl_sync(FULL, g0, up, GW), bg1 = __shfl_sync(FULL, g1, up, GW);
            if (mid && mb >= 1) absorb(C00, C01, C10, C11, B00, B01, B11, bg0, bg1);
            B00 = __shfl_sync(FULL, A00, dn, GW), B01 = __shfl_sync(FULL, A01, dn, GW), B11 = __shfl_sync(FULL, A11, dn, GW);
            bg0 = __shfl_sync(FULL, g0, dn, GW), bg1 = __shfl_sync(FULL, g1, dn, GW);
            if (mid && mb + 1 <= nb - 1) absorb(Kb00, Kb01, Kb10, Kb11, B00, B01, B11, bg0, bg1);
        }
        double s0v = 0.0, s1v = 0.0;  // z of this lane's separator
        {
            double idet = 0.0;
            if (a_hr) {  // every block is inverted once, by its own lane
                const double det = fma(A00, A11, -(A01 * A01));
                ok = ok && A00 > 0.0 && det > 0.0 && det < 1e300;
                idet = alt_rcp(det);
            }
            if (mid) {
                s0v = fma(A11, g0, -(A01 * g1)) * idet;
                s1v = fma(A00, g1, -(A01 * g0)) * idet;
            }
            const int src = top ? pp + 1 : (pp > 0 ? pp - 1 : 0);  // the neighbour towards the middle
            // coupling to that neighbour: (C_{t+1}' x_{t+1}) above the middle, (C_t x_{t-1}) below it
            const double K00 = top ? Kb00 : C00, K01 = top ? Kb01 : C01, K10 = top ? Kb10 : C10, K11 = top ? Kb11 : C11;
            const int smax = mbmax > nbmax - 1 - mbmax ? mbmax : nbmax - 1 - mbmax;
            for (int st = 1; st <= smax; ++st) {
                const double x0 = __shfl_sync(FULL, s0v, src, GW), x1 = __shfl_sync(FULL, s1v, src, GW);
                if (a_hr && ((top && pp == mb - st) || (bot && pp == mb + st))) {
                    const double h0 = g0 - fma(K00, x0, K01 * x1), h1 = g1 - fma(K10, x0, K11 * x1);
                    s0v = fma(A11, h0, -(A01 * h1)) * idet;
                    s1v = fma(A00, h1, -(A01 * h0)) * idet;
                }
            }
        }
        const double sL0 = __shfl_up_sync(FULL, s0v, 1, GW), sL1 = __shfl_up_sync(FULL, s1v, 1, GW);
        // back substitution through the chunk
        AltBwd r;
        if (a_hr) {
            r.z1 = s0v;
            r.z2 = s1v;
            r.b1 = a1;   // L[sep row 0, last interior row]
            r.b2 = cc;   // L[sep row 1, last interior row]
            r.b2n = a2;  // L[sep row 0, the row above it]
            sink(start + cn + 1, cn + 1, s1v);
            sink(start + cn, cn, s0v);
        }
        const double uL0 = a_hl ? sL0 : 0.0, uL1 = a_hl ? sL1 : 0.0;
#pragma unroll 2
        for (int i = cn - 1; i >= 0; --i) {
            const double ydk = fma(-F.ld(AF_V2, i), uL1, fma(-F.ld(AF_V1, i), uL0, F.ld(AF_YD, i)));
            const double z = alt_bwd_row(r, F.ld(AF_L1, i), F.ld(AF_L2, i), ydk);
            sink(start + i, i, z);
        }
        return ok;
    };

    // ---- pass 1 (cpp:1575-1712), then the active-set loop of pass 2 with lambda_smooth * 10, max_climb_rate * 0.5
    const bool ok = !group_any(!solve(std::integral_constant<int, 1>{}, p.lambda_smooth, mine));
    // W becomes the climb weight of pass 2 (non-negative: its sign bit is that pass's active-set mark).
    for (int i = 0; i < len; ++i) F.st(AF_W, i, alt_part_weight<2>(F.ld(AF_W, i)));
    int solves = 0;
    bool ok2 = true, running = n > 0;
    for (int iter = 0; iter < ALT_MAX_ITER; ++iter) {
        if (!__any_sync(FULL, running)) break;
        violation = false;
        const bool good = solve(std::integral_constant<int, 2>{}, p.lambda_smooth * 10.0, mine && running);
        const bool bad = group_any(!good), viol = group_any(violation);
        if (running) {
            ok2 = ok2 && !bad;
            ++solves;
            if (!viol) running = false;  // converged (cpp:1814)
        }
    }
    // ---- write-back (cpp:1817-1821, 1357-1359; failure semantics cpp:1342-1344, 1356): max(z, z1), or z1 alone when
    // pass 2 failed, or nothing when pass 1 did
    if (ok) {
        double *r = rows + 3 * (base + start);
        for (int i = 0; i < len; ++i) {
            const double zi = F.ld(AF_E, i), z = F.ld(AF_YD, i);
            r[3 * i + 2] = (ok2 && z >= zi) ? z : zi;
        }
    }
    solves_ret = solves;
    ok_ret = ok;
    ok2_ret = ok2;
}

struct AltPartScratch {
    double *g[AF_COUNT];  // n_rows_cap doubles each (the long-trajectory form)
};

__global__ void __maxnreg__(192) k_alt_part(AltParams p, long long B, const long long *__restrict__ row_offset,
                                                 double *rows, const double *elev, double *z_pass1_out,
                                                 int *__restrict__ solves_out, unsigned *__restrict__ flags_out,
                                                 long long n_cap, AltPartScratch scratch) {
    extern __shared__ __align__(16) double altp_sm[];
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int TPW = 32 / ALTP_GROUP;  // trajectories per warp
    const int lane = threadIdx.x;
    const long long b = (long long)blockIdx.x * TPW + lane / ALTP_GROUP;
    const long long base = b < B ? row_offset[b] : 0;
    const bool truncated = b < B && row_offset[b + 1] > n_cap;  // rows missing from the caller's buffers: skip the trajectory
    const long long n_ll = (b < B && !truncated) ? row_offset[b + 1] - base : 0;
    const int n = n_ll > 0 ? (int)(n_ll < 0x7fffffff ? n_ll : 0x7fffffff) : 0;
    int solves = 0;
    bool ok = true, ok2 = true;
    __shared__ __align__(8) unsigned long long alt_mbar;  // completion of the bulk copies: one arrival per group of lanes
    const unsigned mbar = (unsigned)__cvta_generic_to_shared(&alt_mbar);
    unsigned phase = 0;  // uses of the barrier so far (its parity)
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(TPW) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    {
        const AltPartFld<false> F{altp_sm + lane};
        alt_part_set<ALTP_GROUP, false>(p, F, lane, n > 0 && n <= ALTP_NMAX, base, n, rows, elev, z_pass1_out, n_cap, mbar, phase,
                                        solves, ok, ok2);
    }
    // trajectories too long for a group of lanes: one at a time, all 32 lanes -- in shared memory up to ALTP_NMAX32 rows,
    // with the fields in the global scratch beyond
    unsigned longer = __ballot_sync(FULL, lane % ALTP_GROUP == 0 && n > ALTP_NMAX);
    while (longer) {
        const int src = __ffs(longer) - 1;
        longer &= longer - 1;
        const long long lbase = __shfl_sync(FULL, base, src);
        const int ln = __shfl_sync(FULL, n, src);
        int s2;
        bool o1, o2;
        __syncwarp();
        if (ln <= ALTP_NMAX32) {
            const AltPartFld<false> F{altp_sm + lane};
            alt_part_set<32, false>(p, F, lane, true, lbase, ln, rows, elev, z_pass1_out, n_cap, mbar, phase, s2, o1, o2);
        } else {
            AltPartFld<true> G;
#pragma unroll
            for (int f = 0; f < AF_COUNT; ++f) G.g[f] = scratch.g[f];
            alt_part_set<32, true>(p, G, lane, true, lbase, ln, rows, elev, z_pass1_out, n_cap, mbar, phase, s2, o1, o2);
        }
        if (lane / ALTP_GROUP == src / ALTP_GROUP) {
            solves = s2;
            ok = o1;
            ok2 = o2;
        }
    }
    if (b < B && lane % ALTP_GROUP == 0) {
        if (solves_out) solves_out[b] = solves;
        if (flags_out)
            flags_out[b] = truncated ? ALT_FLAG_TRUNCATED : (!ok ? ALT_FLAG_PIVOT : (!ok2 ? (ALT_FLAG_PIVOT | ALT_FLAG_PASS2) : 0u));
    }
}

}  // namespace msnap
#endif  // MSNAP_ALT_PART_CUH
