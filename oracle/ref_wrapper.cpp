// oracle/ref_wrapper.cpp -- TEST INFRASTRUCTURE, not product code.
//
// extern "C" door onto the UNMODIFIED reference class TrajectoryGeneratorTool
// (/root/reference/math_util/minimum_snap.hpp:36-63, minimum_snap.cpp:22-649), compiled against
// oracle/shim/Eigen/Dense.  Built by oracle/Makefile into oracle/_ref/ (git-ignored); only tests/,
// __graft_entry__.smoke() and bench.py's CPU-baseline legs may load it.
//
// All matrices cross this boundary ROW-major ([point][xyz], [segment][axis][power hi->lo]); the wrapper
// converts to the reference's column-major Eigen values.
#include "minimum_snap.hpp"

#include <algorithm>
#include <iostream>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct Quiet {  // the reference prints per segment per solve (ms.cpp:85,194,239,472,507,621): mute it
    Quiet() {
        std::cout.setstate(std::ios_base::failbit);
        std::cerr.setstate(std::ios_base::failbit);
    }
};

Eigen::MatrixXd to_path(const double *wp, int n_pts) {
    Eigen::MatrixXd P(n_pts, 3);
    for (int i = 0; i < n_pts; ++i)
        for (int a = 0; a < 3; ++a) P(i, a) = wp[3 * i + a];
    return P;
}

MinimumSnapConfig to_cfg(int order, double pw, double vw, double v_avg, double min_time, double sd,
                         const double *bc /* start_vel,end_vel,start_acc,end_acc : 12 doubles */) {
    MinimumSnapConfig c;
    c.order = order;
    c.path_weight = pw;
    c.vel_zero_weight = vw;
    c.V_avg = v_avg;
    c.min_time_s = min_time;
    c.sample_distance = sd;
    for (int a = 0; a < 3; ++a) {
        c.start_vel(a) = bc[a];
        c.end_vel(a) = bc[3 + a];
        c.start_acc(a) = bc[6 + a];
        c.end_acc(a) = bc[9 + a];
    }
    return c;
}

}  // namespace

extern "C" {

int msnap_ref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// One call of TrajectoryGeneratorTool::SolveQPClosedForm (ms.cpp:227-649).
//   wp[n_pts*3] row-major, vel[2*3] / acc[2*3] row-major (row 0 start, row 1 end), time[n_pts-1]
//   coeff_out[(n_pts-1) * 3 * 2*order] row-major == PolyCoeff rows.
int msnap_ref_solve_qp(int order, int n_pts, const double *wp, const double *vel, const double *acc,
                       const double *time, double path_weight, double vel_zero_weight, double *coeff_out,
                       double *max_dev_out) {
    static Quiet quiet;
    const int ns = n_pts - 1;
    Eigen::MatrixXd P = to_path(wp, n_pts), V(2, 3), A(2, 3);
    for (int r = 0; r < 2; ++r)
        for (int a = 0; a < 3; ++a) {
            V(r, a) = vel[3 * r + a];
            A(r, a) = acc[3 * r + a];
        }
    Eigen::VectorXd T(ns);
    for (int i = 0; i < ns; ++i) T(i) = time[i];
    TrajectoryGeneratorTool tool;
    double md = 0.0;
    Eigen::MatrixXd C = tool.SolveQPClosedForm(order, P, V, A, T, path_weight, vel_zero_weight, &md);
    const int w = 3 * 2 * order;
    for (int s = 0; s < ns; ++s)
        for (int j = 0; j < w; ++j) coeff_out[s * w + j] = C(s, j);
    if (max_dev_out) *max_dev_out = md;
    return 0;
}

// One call of TrajectoryGeneratorTool::GenerateTrajectoryMatrix (ms.cpp:22-206).
// Returns the number of samples S (rows of the reference's output); writes min(S, cap) rows row-major.
int msnap_ref_generate(int order, double pw, double vw, double v_avg, double min_time, double sd, const double *bc,
                       double sd_override, double v_avg_override, int n_pts, const double *wp, int cap,
                       double *samples_out) {
    static Quiet quiet;
    MinimumSnapConfig cfg = to_cfg(order, pw, vw, v_avg, min_time, sd, bc);
    TrajectoryGeneratorTool tool;
    Eigen::MatrixXd S = tool.GenerateTrajectoryMatrix(to_path(wp, n_pts), cfg, sd_override, v_avg_override);
    const int n = static_cast<int>(S.rows());
    for (int i = 0; i < std::min(n, cap); ++i)
        for (int a = 0; a < 3; ++a) samples_out[3 * i + a] = S(i, a);
    return n;
}

// The pieces GenerateTrajectoryMatrix computes but does not return -- segment times (ms.cpp:63-72) and the
// reweighting loop's final state (ms.cpp:76-90) -- obtained by driving the reference's own public
// SolveQPClosedForm with the same loop.  coeff_out as in msnap_ref_solve_qp; time_out[n_pts-1].
int msnap_ref_reweighted_solve(int order, double pw, double vw, double v_avg, double min_time, const double *bc,
                               double v_avg_override, int n_pts, const double *wp, double *time_out,
                               double *coeff_out, double *max_dev_out, int *iters_out, double *vw_final_out) {
    static Quiet quiet;
    const int ns = n_pts - 1;
    if (v_avg_override > 0.0) v_avg = v_avg_override;
    std::vector<double> T(static_cast<size_t>(ns));
    for (int i = 0; i < ns; ++i) {
        // not contracted into FMAs: the reference as shipped is built without -march flags (CMakeLists.txt)
        volatile double dx = wp[3 * (i + 1) + 0] - wp[3 * i + 0];
        volatile double dy = wp[3 * (i + 1) + 1] - wp[3 * i + 1];
        volatile double dz = wp[3 * (i + 1) + 2] - wp[3 * i + 2];
        volatile double xx = dx * dx, yy = dy * dy, zz = dz * dz;
        volatile double s1 = xx + yy;
        double len = std::sqrt(s1 + zz);
        double t = (v_avg > 1e-6) ? (len / v_avg) : min_time;
        if (t < min_time) t = min_time;
        T[static_cast<size_t>(i)] = t;
        if (time_out) time_out[i] = t;
    }
    double vel[6], acc[6];
    for (int a = 0; a < 3; ++a) {
        vel[a] = bc[a];
        vel[3 + a] = bc[3 + a];
        acc[a] = bc[6 + a];
        acc[3 + a] = bc[9 + a];
    }
    double md = 0.0;
    int iter = 0;
    while (true) {
        msnap_ref_solve_qp(order, n_pts, wp, vel, acc, T.data(), pw, vw, coeff_out, &md);
        if (md > 0.2 && iter < 10) {
            vw = (vw < 1e-6) ? 0.01 : vw * 2.0;
            ++iter;
        } else {
            break;
        }
    }
    if (max_dev_out) *max_dev_out = md;
    if (iters_out) *iters_out = iter;
    if (vw_final_out) *vw_final_out = vw;
    return 0;
}

// CPU-baseline driver: B independent GenerateTrajectoryMatrix calls, one TrajectoryGeneratorTool per call
// (the object is not re-entrant: ms.cpp:38 writes a member), optionally OpenMP-parallel over trajectories.
//   pt_offset[B+1] indexes wp rows (CSR); count_out[B] receives the sample counts.  Samples are discarded
//   unless samples_out != NULL, in which case trajectory b writes at most cap rows at samples_out + 3*cap*b.
// Returns the number of threads used.
int msnap_ref_generate_batch(int order, double pw, double vw, double v_avg, double min_time, double sd,
                             const double *bc, double sd_override, double v_avg_override, int B,
                             const long long *pt_offset, const double *wp, int nthreads, int *count_out, int cap,
                             double *samples_out) {
    static Quiet quiet;
    MinimumSnapConfig cfg = to_cfg(order, pw, vw, v_avg, min_time, sd, bc);
    int used = 1;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    used = nthreads;
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
#else
    (void)nthreads;
#endif
    for (int b = 0; b < B; ++b) {
        const int n_pts = static_cast<int>(pt_offset[b + 1] - pt_offset[b]);
        TrajectoryGeneratorTool tool;
        Eigen::MatrixXd S =
            tool.GenerateTrajectoryMatrix(to_path(wp + 3 * pt_offset[b], n_pts), cfg, sd_override, v_avg_override);
        const int n = static_cast<int>(S.rows());
        if (count_out) count_out[b] = n;
        if (samples_out)
            for (int i = 0; i < std::min(n, cap); ++i)
                for (int a = 0; a < 3; ++a) samples_out[(static_cast<size_t>(b) * cap + i) * 3 + a] = S(i, a);
    }
    return used;
}

// Batched msnap_ref_reweighted_solve: B independent trajectories (CSR pt_offset[B+1] into wp rows), OpenMP over
// trajectories.  time_out / coeff_out are laid out per segment in batch order (segment offset of trajectory b =
// pt_offset[b] - b).  Returns the number of threads used.
int msnap_ref_reweighted_solve_batch(int order, double pw, double vw, double v_avg, double min_time, const double *bc,
                                     double v_avg_override, int B, const long long *pt_offset, const double *wp,
                                     int nthreads, double *time_out, double *coeff_out, double *max_dev_out,
                                     int *iters_out, double *vw_final_out) {
    static Quiet quiet;
    int used = 1;
    const int w = 3 * 2 * order;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    used = nthreads;
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
#else
    (void)nthreads;
#endif
    for (int b = 0; b < B; ++b) {
        const long long p0 = pt_offset[b], g0 = p0 - b;
        const int n_pts = static_cast<int>(pt_offset[b + 1] - p0);
        msnap_ref_reweighted_solve(order, pw, vw, v_avg, min_time, bc, v_avg_override, n_pts, wp + 3 * p0,
                                   time_out + g0, coeff_out + g0 * w, max_dev_out + b, iters_out + b, vw_final_out + b);
    }
    return used;
}

}  // extern "C"
