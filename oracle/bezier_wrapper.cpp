// oracle/bezier_wrapper.cpp -- TEST INFRASTRUCTURE, not product code.
//
// extern "C" door onto the UNMODIFIED reference class math_util::Bezier
// (/root/reference/math_util/bezier.hpp:98-120, bezier.cpp:127-189: GenerateTrajectoryMatrix = heading estimation,
// one cubic Bezier per waypoint pair with heading-derived control points, resolution-stepped sampling), compiled
// against oracle/shim/Eigen/Dense, and -- when built with -DPATROL_REF_INC=<file> -- onto the planner's free helper
// functions of the single-patrol post-processing (uavPathPlanning.cpp:118-206: sameXYPoint, cross2D, onSegment2D,
// segmentsIntersect2D, hasSelfIntersection2D, sampleClosedPolygonBoundary).  The planner's translation unit itself cannot
// be compiled here (yaml-cpp, GDAL, out-of-tree json.hpp / elog.h), so oracle/Makefile cuts exactly those lines out of the
// reference file where it lies into the git-ignored oracle/_ref/patrol_helpers.inc at build time; the statements executed
// are the reference's own text, nothing of it is committed.  gen_single_patrol's trim loop (cpp:1857-1895) is a member
// function that calls Minisnap_3D; it is restated in oracle/patrol_port.py on top of these helpers.
//
// Matrices cross this boundary ROW-major ([point][xyz]).
#include "bezier.hpp"

#include <algorithm>
#include <cmath>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef PATROL_REF_INC
struct ENUPoint {  // uavPathPlanning.hpp:152-156
    double east;
    double north;
    double up;
};
#include PATROL_REF_INC
#endif

extern "C" {

// One call of math_util::Bezier::GenerateTrajectoryMatrix as UavPathPlanner::Bezier_3D drives it (uavPathPlanning.cpp:
// 4477-4505): config.min_radius = 300 iff min_radius_arg > 0, else the struct default 1.0.
// Returns the number of rows; writes min(rows, cap) of them.
int bezier_ref_generate(int n_pts, const double *wp, double sample_distance_override, double min_radius_arg, int cap,
                        double *out) {
    Eigen::MatrixXd P(n_pts, 3);
    for (int i = 0; i < n_pts; ++i)
        for (int a = 0; a < 3; ++a) P(i, a) = wp[3 * i + a];
    math_util::Bezier bezier;
    math_util::BezierConfig config;
    if (min_radius_arg > 0) config.min_radius = 300;
    bezier.SetConfig(config);
    Eigen::MatrixXd S = bezier.GenerateTrajectoryMatrix(P, "", sample_distance_override, -1.0);
    const int n = static_cast<int>(S.rows());
    for (int i = 0; i < std::min(n, cap); ++i)
        for (int a = 0; a < 3; ++a) out[3 * i + a] = S(i, a);
    return n;
}

// B independent calls (CSR pt_offset[B+1] into wp rows), OpenMP over trajectories.  count_out[b] = rows of trajectory b;
// trajectory b writes at most cap rows at out + 3*cap*b when out != NULL.  Returns the threads used.
int bezier_ref_generate_batch(int B, const long long *pt_offset, const double *wp, double sample_distance_override,
                              double min_radius_arg, int nthreads, int *count_out, int cap, double *out) {
    int used = 1;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    used = nthreads;
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
#endif
    for (int b = 0; b < B; ++b) {
        const int n_pts = static_cast<int>(pt_offset[b + 1] - pt_offset[b]);
        std::vector<double> tmp;
        double *dst = out ? out + static_cast<size_t>(b) * cap * 3 : nullptr;
        count_out[b] = bezier_ref_generate(n_pts, wp + 3 * pt_offset[b], sample_distance_override, min_radius_arg,
                                           dst ? cap : 0, dst);
    }
    return used;
}

#ifdef PATROL_REF_INC
static std::vector<ENUPoint> to_points(int n, const double *rows) {
    std::vector<ENUPoint> v(static_cast<size_t>(n));
    for (int i = 0; i < n; ++i) v[static_cast<size_t>(i)] = ENUPoint{rows[3 * i], rows[3 * i + 1], rows[3 * i + 2]};
    return v;
}

int patrol_ref_has_self_intersection(int n, const double *rows, int closed) {
    return hasSelfIntersection2D(to_points(n, rows), closed != 0) ? 1 : 0;
}

int patrol_ref_segments_intersect(const double *a1, const double *a2, const double *b1, const double *b2) {
    return segmentsIntersect2D(ENUPoint{a1[0], a1[1], a1[2]}, ENUPoint{a2[0], a2[1], a2[2]}, ENUPoint{b1[0], b1[1], b1[2]},
                               ENUPoint{b2[0], b2[1], b2[2]}) ? 1 : 0;
}

// sampleClosedPolygonBoundary (cpp:179-206): returns the row count, writes min(count, cap) rows.
int patrol_ref_sample_boundary(int n, const double *polygon, double spacing, int cap, double *out) {
    const std::vector<ENUPoint> s = sampleClosedPolygonBoundary(to_points(n, polygon), spacing);
    const int m = static_cast<int>(s.size());
    for (int i = 0; i < std::min(m, cap); ++i) {
        out[3 * i] = s[static_cast<size_t>(i)].east;
        out[3 * i + 1] = s[static_cast<size_t>(i)].north;
        out[3 * i + 2] = s[static_cast<size_t>(i)].up;
    }
    return m;
}
#endif

}  // extern "C"
