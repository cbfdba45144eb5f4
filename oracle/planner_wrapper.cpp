// oracle/planner_wrapper.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Executes the reference's OWN statements for three stages around the minimum-snap path that live inside the planner's
// translation unit /root/reference/uavPathPlanning.cpp, which cannot be compiled as a whole in this image (it needs
// yaml-cpp, GDAL, nlohmann json.hpp and elog.h from outside the reference tree):
//   * WGS84 <-> ECEF <-> ENU                 UavPathPlanner::wgs84ToECEF .. enuToWGS84_Batch        cpp:894-910, 926-1108
//   * altitude optimisation                  makeAltitudeParams, optimizeSegmentAltitudeENU,        cpp:1311-1364,
//                                            optimizeHeights, optimizeHeightsGlobalSmooth            1575-1827
//   * follower formation trajectories        generateFollowerTrajectories + the four formation      cpp:3931-4398
//                                            generators
//   * ElevationCostMap::getCostAt                                                                   elevation_cost_map.cpp:373-380
// oracle/Makefile cuts exactly those function definitions (and the structs / constants they need from
// uavPathPlanning.hpp) out of the reference files WHERE THEY LIE into git-ignored oracle/_ref/planner_*.inc at build time;
// this file supplies the class shell they are members of and stand-ins for what is absent from the image:
//   - <Eigen/Dense>, <Eigen/Sparse>: oracle/shim (SimplicialLDLT = banded LDL' in natural order; see that header)
//   - Eigen::Matrix2d / Vector2d: the 2 x 2 pieces below
//   - nlohmann::json: the array-of-numbers subset the formation generators build their result with
// Nothing of the reference's text is committed.  Only tests/, smoke() and bench.py's CPU legs may load the result.
#include <Eigen/Dense>
#include <Eigen/Sparse>

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <iostream>
#include <limits>
#include <memory>
#include <string>
#include <utility>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "math_util/coordinate_transform.hpp"  // WGS84Coord (reference header, included where it lies)
#include "math_util/minimum_snap.hpp"          // MinimumSnapConfig (member of PlannerConfig)

using namespace std;
using namespace math_util;

// ---- stand-ins -----------------------------------------------------------------------------------------------------
namespace Eigen {
class Vector2d {
public:
    Vector2d() : x_(0), y_(0) {}
    Vector2d(double x, double y) : x_(x), y_(y) {}
    double x() const { return x_; }
    double y() const { return y_; }

private:
    double x_, y_;
};
class Matrix2d {  // row-major fill through `m << a, b, c, d;`; product = coefficient-wise dot products like Eigen's 2 x 2 kernel
public:
    struct Init {
        Matrix2d *m;
        int k;
        Init operator,(double v) {
            m->a_[k] = v;
            return Init{m, k + 1};
        }
    };
    Init operator<<(double v) {
        a_[0] = v;
        return Init{this, 1};
    }
    Vector2d operator*(const Vector2d &v) const { return Vector2d(a_[0] * v.x() + a_[1] * v.y(), a_[2] * v.x() + a_[3] * v.y()); }

private:
    double a_[4] = {0, 0, 0, 0};
};
}  // namespace Eigen

class json {  // arrays of numbers / arrays, as generateFollowerTrajectories builds them
public:
    json() : num_(0.0) {}
    json(int v) : num_(v) {}
    json(double v) : num_(v) {}
    static json array() { return json(); }
    void push_back(const json &j) { arr_.push_back(j); }
    size_t size() const { return arr_.size(); }
    const json &operator[](size_t i) const { return arr_[i]; }
    template <class T>
    T get() const { return static_cast<T>(num_); }

private:
    double num_;
    std::vector<json> arr_;
};

class ElevationCostMap {  // elevation_cost_map.hpp:38-55: the cost-map half; the GDAL elevation half is absent (no raster here)
public:
    bool getCostAt(double x, double y, float &val) const;
    bool getElevationAt(double, double, double &) const { return false; }
    void createCostMap(int width, int height, double resolution, double origin_x, double origin_y, const float *data) {
        cost_width_ = width; cost_height_ = height; cost_resolution_ = resolution; cost_origin_x_ = origin_x; cost_origin_y_ = origin_y;
        cost_data_.assign(data, data + static_cast<size_t>(width) * height);
    }

private:
    int cost_width_ = 0;
    int cost_height_ = 0;
    double cost_resolution_ = 0.0;
    double cost_origin_x_ = 0.0;
    double cost_origin_y_ = 0.0;
    std::vector<float> cost_data_;
};
#include "_ref/planner_costmap.inc"  // bool ElevationCostMap::getCostAt(...)   elevation_cost_map.cpp:373-380

#include "_ref/planner_hpp_types.inc"   // ProhibitedZone .. InputData                         uavPathPlanning.hpp:26-95
#include "_ref/planner_hpp_geo.inc"     // WGS84_A .. rad2deg, WGS84Point, ENUPoint, ECEFPoint uavPathPlanning.hpp:134-173

class UavPathPlanner {
public:
#include "_ref/planner_hpp_config.inc"  // struct PlannerConfig                                  uavPathPlanning.hpp:178-215
#include "_ref/planner_hpp_alt.inc"     // struct AltitudeParams                                 uavPathPlanning.hpp:415-421
    std::unique_ptr<ElevationCostMap> elev_cost_map_;
    WGS84Point origin_{0.0, 0.0, 0.0};
    PlannerConfig config_;
    InputData input_data_;

    ECEFPoint wgs84ToECEF(const WGS84Point &lla);
    WGS84Point ecefToWGS84(const ECEFPoint &ecef);
    std::array<std::array<double, 3>, 3> computeENURotationMatrix(double lat_rad, double lon_rad);
    std::array<std::array<double, 3>, 3> computeENURotationMatrixInverse(double lat_rad, double lon_rad);
    ENUPoint ecefToENU(const ECEFPoint &delta_ecef, double ref_lat_rad, double ref_lon_rad);
    ECEFPoint enuToECEF(const ENUPoint &enu, double ref_lat_rad, double ref_lon_rad);
    ENUPoint wgs84ToENU(const WGS84Point &target, const WGS84Point &reference);
    WGS84Point enuToWGS84(const ENUPoint &enu, const WGS84Point &reference);
    std::vector<ENUPoint> wgs84ToENU_Batch(const std::vector<WGS84Point> &targets, const WGS84Point &reference);
    std::vector<WGS84Point> enuToWGS84_Batch(const std::vector<ENUPoint> &targets, const WGS84Point &reference);

    AltitudeParams makeAltitudeParams() const;
    bool optimizeSegmentAltitudeENU(std::vector<ENUPoint> &segment_enu);
    bool optimizeHeights(const std::vector<Eigen::Vector3d> &waypoints, const AltitudeParams &p, std::vector<double> &out_z);
    bool optimizeHeightsGlobalSmooth(const std::vector<double> &input_z, const std::vector<Eigen::Vector3d> &waypoints,
                                     const AltitudeParams &p, std::vector<double> &out_z);

    json generateFollowerTrajectories(const InputData &input_data, const std::vector<ENUPoint> &Trajectory_ENU,
                                      const std::vector<WGS84Point> &Trajectory_WGS84);
    json generateVShapeTrajectories(const json &uavs_ids, const json &uav_starts, const ENUPoint &leader_start_enu,
                                    const Eigen::Matrix2d &R0, const std::vector<Eigen::Vector2d> &leader_xy,
                                    const std::vector<double> &leader_headings, const std::vector<ENUPoint> &Trajectory_ENU,
                                    double safety_distance);
    json generateLineShapeTrajectories(const json &uavs_ids, const json &uav_starts, const ENUPoint &leader_start_enu,
                                       const Eigen::Matrix2d &R0, const std::vector<Eigen::Vector2d> &leader_xy,
                                       const std::vector<double> &leader_headings, const std::vector<ENUPoint> &Trajectory_ENU,
                                       double safety_distance);
    json generateVerticalLineShapeTrajectories(const json &uavs_ids, const json &uav_starts, const ENUPoint &leader_start_enu,
                                               const Eigen::Matrix2d &R0, const std::vector<Eigen::Vector2d> &leader_xy,
                                               const std::vector<double> &leader_headings,
                                               const std::vector<ENUPoint> &Trajectory_ENU, int uav_formation_max_row,
                                               double safety_distance);
    json generateTriangleShapeTrajectories(const json &uavs_ids, const json &uav_starts, const ENUPoint &leader_start_enu,
                                           const Eigen::Matrix2d &R0, const std::vector<Eigen::Vector2d> &leader_xy,
                                           const std::vector<double> &leader_headings,
                                           const std::vector<ENUPoint> &Trajectory_ENU, double safety_distance);
};

#include "_ref/planner_cpp_geo.inc"     // cpp:894-910, 926-1108
#include "_ref/planner_cpp_alt.inc"     // cpp:1311-1364, 1575-1827
#include "_ref/planner_cpp_follow.inc"  // cpp:3931-4398

// ---- extern "C" ------------------------------------------------------------------------------------------------------
namespace {
struct Quiet {  // the extracted functions print progress to cout / cerr
    Quiet() {
        std::cout.setstate(std::ios_base::failbit);
        std::cerr.setstate(std::ios_base::failbit);
    }
};
std::vector<ENUPoint> enu_rows(long long n, const double *r) {
    std::vector<ENUPoint> v(static_cast<size_t>(n));
    for (long long i = 0; i < n; ++i) v[static_cast<size_t>(i)] = ENUPoint{r[3 * i], r[3 * i + 1], r[3 * i + 2]};
    return v;
}
}  // namespace

extern "C" {

// wgs84ToENU_Batch / enuToWGS84_Batch (cpp:1085-1108).  origin = {lon, lat, alt}; rows [n][3].
void planner_ref_wgs84_to_enu(const double *origin, long long n, const double *lla, double *enu) {
    UavPathPlanner p;
    std::vector<WGS84Point> t(static_cast<size_t>(n));
    for (long long i = 0; i < n; ++i) t[static_cast<size_t>(i)] = WGS84Point{lla[3 * i], lla[3 * i + 1], lla[3 * i + 2]};
    const std::vector<ENUPoint> r = p.wgs84ToENU_Batch(t, WGS84Point{origin[0], origin[1], origin[2]});
    for (long long i = 0; i < n; ++i) {
        enu[3 * i] = r[static_cast<size_t>(i)].east; enu[3 * i + 1] = r[static_cast<size_t>(i)].north; enu[3 * i + 2] = r[static_cast<size_t>(i)].up;
    }
}
void planner_ref_enu_to_wgs84(const double *origin, long long n, const double *enu, double *lla) {
    UavPathPlanner p;
    const std::vector<WGS84Point> r = p.enuToWGS84_Batch(enu_rows(n, enu), WGS84Point{origin[0], origin[1], origin[2]});
    for (long long i = 0; i < n; ++i) {
        lla[3 * i] = r[static_cast<size_t>(i)].lon; lla[3 * i + 1] = r[static_cast<size_t>(i)].lat; lla[3 * i + 2] = r[static_cast<size_t>(i)].alt;
    }
}

// optimizeSegmentAltitudeENU (cpp:1329-1364) for B trajectories (CSR row_offset), OpenMP over trajectories.
//   params = {lambda_smooth, lambda_follow, max_climb_rate, uav_R, safe_distance}  (config_.altitude_optimization)
//   grid (may be NULL) = cost map [height][width] floats, top-left origin (ElevationCostMap, elevation_cost_map.hpp:49-55)
//   z1_out (may be NULL) = heights after optimizeHeights alone; ok_out[b] = the function's return value
// Returns the threads used.
int planner_ref_altitude_batch(const double *params, int B, const long long *row_offset, double *rows_inout, const float *grid,
                               int width, int height, double resolution, double origin_x, double origin_y, double *z1_out,
                               int *ok_out, int nthreads) {
    static Quiet quiet;
    int used = 1;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    used = nthreads;
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
#endif
    for (int b = 0; b < B; ++b) {
        UavPathPlanner p;
        p.config_.altitude_optimization.lambda_smooth = params[0];
        p.config_.altitude_optimization.lambda_follow = params[1];
        p.config_.altitude_optimization.max_climb_rate = params[2];
        p.config_.altitude_optimization.uav_R = params[3];
        p.config_.altitude_optimization.safe_distance = params[4];
        p.elev_cost_map_ = std::make_unique<ElevationCostMap>();
        if (grid) p.elev_cost_map_->createCostMap(width, height, resolution, origin_x, origin_y, grid);
        const long long r0 = row_offset[b], n = row_offset[b + 1] - r0;
        std::vector<ENUPoint> seg = enu_rows(n, rows_inout + 3 * r0);
        if (z1_out && n > 0) {  // the first pass alone, as optimizeSegmentAltitudeENU calls it (cpp:1336-1344)
            std::vector<Eigen::Vector3d> w;
            for (const auto &q : seg) w.emplace_back(q.east, q.north, q.up);
            std::vector<double> z;
            if (p.optimizeHeights(w, p.makeAltitudeParams(), z))
                for (long long i = 0; i < n; ++i) z1_out[r0 + i] = z[static_cast<size_t>(i)];
        }
        const bool ok = p.optimizeSegmentAltitudeENU(seg);
        if (ok_out) ok_out[b] = ok ? 1 : 0;
        for (long long i = 0; i < n; ++i) rows_inout[3 * (r0 + i) + 2] = seg[static_cast<size_t>(i)].up;
    }
    return used;
}

// generateFollowerTrajectories (cpp:3931-4074) + the formation generators (cpp:4076-4398) for one leader trajectory.
//   out[f][t][3] = {lon, lat, alt} of follower f at leader sample t; returns the number of followers written.
//   A negative config / input value means "not provided" exactly as in InputData (hpp:85-89).
int planner_ref_followers(const double *origin, long long n, const double *leader_enu, int formation_model, int n_followers,
                          const double *starts_wgs84, double cfg_formation_distance, double cfg_position_misalignment,
                          int cfg_max_row, double cfg_uav_R, double in_formation_distance, double in_position_misalignment,
                          double in_uav_R, int in_max_row, double *out) {
    static Quiet quiet;
    UavPathPlanner p;
    p.origin_ = WGS84Point{origin[0], origin[1], origin[2]};
    p.config_.path_planning.formation_distance = cfg_formation_distance;
    p.config_.path_planning.position_misalignment = cfg_position_misalignment;
    p.config_.path_planning.uav_formation_max_row = cfg_max_row;
    p.config_.altitude_optimization.uav_R = cfg_uav_R;
    InputData in;
    in.formation_using = 1;
    in.formation_model = formation_model;
    in.formation_distance = in_formation_distance;
    in.position_misalignment = in_position_misalignment;
    in.uav_R = in_uav_R;
    in.uav_formation_max_row = in_max_row;
    for (int f = 0; f < n_followers; ++f) {
        in.uavs_id.push_back(100 + f);
        in.uav_start_point_wgs84.emplace_back(starts_wgs84[3 * f], starts_wgs84[3 * f + 1], starts_wgs84[3 * f + 2]);
    }
    const std::vector<ENUPoint> traj = enu_rows(n, leader_enu);
    const json planes = p.generateFollowerTrajectories(in, traj, std::vector<WGS84Point>());
    for (size_t f = 0; f < planes.size(); ++f)
        for (long long t = 0; t < n; ++t) {
            const json &pt = planes[f][static_cast<size_t>(t) + 1];  // [0] is the uav id
            for (int a = 0; a < 3; ++a) out[(f * static_cast<size_t>(n) + static_cast<size_t>(t)) * 3 + a] = pt[static_cast<size_t>(a)].get<double>();
        }
    return static_cast<int>(planes.size());
}

}  // extern "C"
