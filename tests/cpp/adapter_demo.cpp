// tests/cpp/adapter_demo.cpp -- exercises include/minimum_snap_gpu.hpp the way UavPathPlanner::Minisnap_3D does
// (uavPathPlanning.cpp:4440-4474): an N x 3 route in, the reference's two methods called, results printed as text for
// the Python test to compare with the golden vectors.  Built against the oracle's Eigen shim in tests only (real
// Eigen is not installed in this image); the header itself is Eigen-API clean.
#include <cstdio>
#include <cstdlib>
#include <exception>

#include "bezier_gpu.hpp"
#include "geo_transform_gpu.hpp"
#include "minimum_snap_gpu.hpp"

// uavPathPlanning.hpp:145-156
struct WGS84Point {
    double lon, lat, alt;
};
struct ENUPoint {
    double east, north, up;
};

int main(int argc, char **argv) {
    // readme.md:14-20 -- ENU waypoints of the uav31_0 leader route
    const double enu[7][3] = {
        {-0.000000000046327, -0.000000000452815, 1669.000000000820137},
        {-22008.910310499257321, 32.799545377501204, 1636.091338242949178},
        {-22009.474804264991690, -2966.281837991115026, 1635.398165184439677},
        {-15007.552345050633448, -2983.825260306681230, 1655.674289593189314},
        {-1003.853909577760191, -2999.001544960936371, 1673.214552272680066},
        {-1003.446472092303907, 0.068179987007966, 1673.921199759593492},
        {-1003.432888336147585, 100.027485618222272, 1673.920415851918733}};
    Eigen::MatrixXd route(7, 3);
    for (int i = 0; i < 7; ++i)
        for (int a = 0; a < 3; ++a) route(i, a) = enu[i][a];
    MinimumSnapConfig cfg;  // minimum_snap_config.yaml:5-27
    cfg.order = 2;
    cfg.vel_zero_weight = 0.01;
    cfg.path_weight = 1e-7;
    cfg.V_avg = 200.0;
    cfg.min_time_s = 1.0;
    cfg.sample_distance = 300.0;
    const double leader_speed = argc > 1 ? std::atof(argv[1]) : 30.0;
    try {
        TrajectoryGeneratorTool generator_;
        Eigen::MatrixXd sampled = generator_.GenerateTrajectoryMatrix(route, cfg, 300.0, leader_speed);
        std::printf("samples %ld\n", static_cast<long>(sampled.rows()));
        for (long i = 0; i < static_cast<long>(sampled.rows()); ++i)
            std::printf("%.17g %.17g %.17g\n", sampled(i, 0), sampled(i, 1), sampled(i, 2));
        // a bare SolveQPClosedForm with explicit segment times ("waypoints plus segment times in")
        Eigen::VectorXd Time(6);
        const double T[6] = {110.0, 15.0, 35.0, 70.0, 15.0, 1.0};
        for (int i = 0; i < 6; ++i) Time(i) = T[i];
        double md = -1.0;
        Eigen::MatrixXd coeff = generator_.SolveQPClosedForm(2, route, Eigen::MatrixXd::Zero(2, 3),
                                                             Eigen::MatrixXd::Zero(2, 3), Time, 1e-7, 0.01, &md);
        std::printf("coeff %ld %ld max_dev %.17g\n", static_cast<long>(coeff.rows()), static_cast<long>(coeff.cols()), md);
        for (long i = 0; i < static_cast<long>(coeff.rows()); ++i) {
            for (long j = 0; j < static_cast<long>(coeff.cols()); ++j) std::printf("%.17g ", coeff(i, j));
            std::printf("\n");
        }
        // too-short input: empty matrix, as minimum_snap.cpp:54-57
        Eigen::MatrixXd one(1, 3);
        std::printf("short %ld\n", static_cast<long>(generator_.GenerateTrajectoryMatrix(one, cfg).rows()));
        // the planner's transforms on its own structs (readme.md:11): WGS84 -> ENU -> WGS84 about the leader's start
        const std::vector<WGS84Point> wgs = {{109.56059880227296, 40.86719901015758, 1669.0},
                                             {109.2995997466117, 40.86719901015758, 1674.0},
                                             {109.299698988346, 40.84019989401251, 1674.0},
                                             {109.38269994693026, 40.84019989401251, 1674.0},
                                             {109.54869918188973, 40.84019989401251, 1674.0},
                                             {109.54869918188973, 40.86719901015758, 1674.0},
                                             {109.54869918188973, 40.868098891288774, 1674.0}};
        const WGS84Point origin_{109.56059880227296, 40.86719901015758, 0.0};
        const std::vector<ENUPoint> e = msnap_geo::wgs84ToENU_Batch<ENUPoint>(generator_.handle(), wgs, origin_);
        const std::vector<WGS84Point> back = msnap_geo::enuToWGS84_Batch<WGS84Point>(generator_.handle(), e, origin_);
        std::printf("geo %zu\n", e.size());
        for (size_t i = 0; i < e.size(); ++i)
            std::printf("%.17g %.17g %.17g %.17g %.17g %.17g\n", e[i].east, e[i].north, e[i].up, back[i].lon, back[i].lat,
                        back[i].alt);
        // the alternative generator, as UavPathPlanner::Bezier_3D drives it (uavPathPlanning.cpp:4489-4499)
        math_util::Bezier bezier;
        math_util::BezierConfig bconfig;
        bezier.SetConfig(bconfig);
        Eigen::MatrixXd bz = bezier.GenerateTrajectoryMatrix(route, "", 300.0, leader_speed);
        std::printf("bezier %ld\n", static_cast<long>(bz.rows()));
        for (long i = 0; i < static_cast<long>(bz.rows()); ++i) std::printf("%.17g %.17g %.17g\n", bz(i, 0), bz(i, 1), bz(i, 2));
        std::printf("bezier_short %ld %ld\n", static_cast<long>(bezier.GenerateTrajectoryMatrix(one, "").rows()),
                    static_cast<long>(bezier.GenerateTrajectoryMatrix(one, "").cols()));
    } catch (const std::exception &e) {
        std::printf("exception %s\n", e.what());
        return 3;
    }
    return 0;
}
