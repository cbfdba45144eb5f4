// include/minimum_snap_gpu.hpp -- drop-in C++ replacement for the reference's math_util/minimum_snap.hpp.
//
// Same type names, same member signatures, same argument meaning and the same failure behaviour as
//   /root/reference/math_util/minimum_snap.hpp:9-33   struct MinimumSnapConfig
//   /root/reference/math_util/minimum_snap.hpp:36-63  class  TrajectoryGeneratorTool
// so that UavPathPlanner (uavPathPlanning.hpp:294 `TrajectoryGeneratorTool generator_;`, uavPathPlanning.cpp:4423 and
// 4461 `generator_.GenerateTrajectoryMatrix(route, config_.minimum_snap, distance_, v_avg_override)`) compiles and
// runs unchanged when this header is included instead of the reference's.  All arithmetic happens in the CUDA library
// behind include/msnap.h; this header only converts Eigen values to the C ABI's row-major buffers.
//
// It compiles against real Eigen 3 and against the oracle's mini shim (tests only).  There is no CPU fallback: if the
// library cannot create a handle (no sm_100 GPU), the first call throws std::runtime_error.
#ifndef MINIMUM_SNAP_GPU_HPP_
#define MINIMUM_SNAP_GPU_HPP_

#include <Eigen/Dense>

#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "msnap.h"

// minimum_snap.hpp:9-33 -- field for field
struct MinimumSnapConfig {
    int order = 3;
    double path_weight = 0.0;
    double vel_zero_weight = 0.0;
    double V_avg = 5.0;
    double min_time_s = 0.1;
    double sample_distance = 1.0;
    Eigen::Vector3d start_vel = Eigen::Vector3d::Zero();
    Eigen::Vector3d end_vel = Eigen::Vector3d::Zero();
    Eigen::Vector3d start_acc = Eigen::Vector3d::Zero();
    Eigen::Vector3d end_acc = Eigen::Vector3d::Zero();
};

class TrajectoryGeneratorTool {
public:
    explicit TrajectoryGeneratorTool(int device = 0) : device_(device) {}
    // the reference object is a plain value type (uavPathPlanning.hpp:294); copies get their own lazily created handle
    TrajectoryGeneratorTool(const TrajectoryGeneratorTool &o) : device_(o.device_) {}
    TrajectoryGeneratorTool &operator=(const TrajectoryGeneratorTool &o) {
        if (this != &o) {
            release();
            device_ = o.device_;
        }
        return *this;
    }
    ~TrajectoryGeneratorTool() { release(); }

    // minimum_snap.hpp:45-53.  Path (n x 3), Vel/Acc (2 x 3: row 0 start, row 1 end), Time (n-1) -> (n-1) x 3*2*order,
    // each row | x: c_{2o-1}..c_0 | y .. | z .. | (minimum_snap.cpp:220-225).
    Eigen::MatrixXd SolveQPClosedForm(int order, const Eigen::MatrixXd &Path, const Eigen::MatrixXd &Vel,
                                      const Eigen::MatrixXd &Acc, const Eigen::VectorXd &Time, double path_weight = 0.0,
                                      double vel_zero_weight = 0.0, double *max_deviation = nullptr) {
        const int ns = static_cast<int>(Time.size());
        if (ns < 1 || Path.rows() != ns + 1 || Path.cols() < 3 || Vel.rows() < 2 || Acc.rows() < 2) {
            std::cerr << "TrajectoryGeneratorTool::SolveQPClosedForm: inconsistent argument shapes" << std::endl;
            return Eigen::MatrixXd();
        }
        ensure();
        std::vector<double> wp = row_major_xyz(Path), vel(6), acc(6), t(static_cast<size_t>(ns));
        for (int r = 0; r < 2; ++r)
            for (int a = 0; a < 3; ++a) {
                vel[3 * r + a] = Vel(r, a);
                acc[3 * r + a] = Acc(r, a);
            }
        for (int i = 0; i < ns; ++i) t[static_cast<size_t>(i)] = Time(i);
        const int w = 3 * 2 * order;
        std::vector<double> coeff(static_cast<size_t>(ns) * (order >= 1 ? w : 0));
        double md = 0.0;
        const int rc = msnap_solve_qp_batch_host(h_, order, path_weight, vel_zero_weight, 1, ns, nullptr, wp.data(),
                                                 vel.data(), acc.data(), t.data(), coeff.data(), &md, nullptr, nullptr);
        if (rc != MSNAP_OK) return fail("SolveQPClosedForm", rc);
        Eigen::MatrixXd out(ns, w);
        for (int s = 0; s < ns; ++s)
            for (int j = 0; j < w; ++j) out(s, j) = coeff[static_cast<size_t>(s) * w + j];
        if (max_deviation) *max_deviation = md;
        return out;
    }

    // minimum_snap.hpp:60-61.  Path (n x 3) -> sampled trajectory (S x 3).  Fewer than 2 rows or 3 columns: message on
    // stderr and an empty matrix, exactly like minimum_snap.cpp:54-57; callers treat empty as failure
    // (uavPathPlanning.cpp:1850-1855).
    Eigen::MatrixXd GenerateTrajectoryMatrix(const Eigen::MatrixXd &Path, const MinimumSnapConfig &cfg,
                                             double sample_distance_override = -1.0, double v_avg_override = -1.0) {
        if (Path.rows() < 2 || Path.cols() < 3) {
            std::cerr << "TrajectoryGeneratorTool::GenerateTrajectoryMatrix: Path must be (N>=2 x 3)" << std::endl;
            return Eigen::MatrixXd();
        }
        ensure();
        const msnap_config c = to_c(cfg);
        const int n = static_cast<int>(Path.rows());
        std::vector<double> wp = row_major_xyz(Path);
        long long bound = 0;
        int rc = msnap_sample_bound_host(h_, &c, v_avg_override, 1, n - 1, nullptr, wp.data(), &bound);
        if (rc != MSNAP_OK) return fail("GenerateTrajectoryMatrix", rc);
        std::vector<double> samples(static_cast<size_t>(bound > 0 ? bound : 1) * 3);
        long long count = 0;
        rc = msnap_generate_one_host(h_, &c, sample_distance_override, v_avg_override, n, wp.data(), bound,
                                     samples.data(), &count);
        if (rc != MSNAP_OK) return fail("GenerateTrajectoryMatrix", rc);
        Eigen::MatrixXd out(count, 3);
        for (long long i = 0; i < count; ++i)
            for (int a = 0; a < 3; ++a) out(i, a) = samples[static_cast<size_t>(i) * 3 + a];
        return out;
    }

    // ---- batched addition (not in the reference): B uniform trajectories, row-major buffers, see include/msnap.h ----
    msnap_handle handle() {
        ensure();
        return h_;
    }
    static msnap_config to_c(const MinimumSnapConfig &cfg) {
        msnap_config c;
        c.order = cfg.order;
        c.path_weight = cfg.path_weight;
        c.vel_zero_weight = cfg.vel_zero_weight;
        c.V_avg = cfg.V_avg;
        c.min_time_s = cfg.min_time_s;
        c.sample_distance = cfg.sample_distance;
        for (int a = 0; a < 3; ++a) {
            c.start_vel[a] = cfg.start_vel(a);
            c.end_vel[a] = cfg.end_vel(a);
            c.start_acc[a] = cfg.start_acc(a);
            c.end_acc[a] = cfg.end_acc(a);
        }
        return c;
    }

private:
    void ensure() {
        if (h_) return;
        const int rc = msnap_create(device_, &h_);
        if (rc != MSNAP_OK)
            throw std::runtime_error(std::string("TrajectoryGeneratorTool: msnap_create failed: ") +
                                     msnap_status_string(rc) + " (this build has no CPU fallback)");
    }
    void release() {
        if (h_) msnap_destroy(h_);
        h_ = nullptr;
    }
    Eigen::MatrixXd fail(const char *where, int rc) {
        std::cerr << "TrajectoryGeneratorTool::" << where << ": " << msnap_status_string(rc) << " "
                  << msnap_last_error(h_) << std::endl;
        return Eigen::MatrixXd();  // the reference's failure value
    }
    static std::vector<double> row_major_xyz(const Eigen::MatrixXd &Path) {
        const long n = static_cast<long>(Path.rows());
        std::vector<double> wp(static_cast<size_t>(n) * 3);
        for (long i = 0; i < n; ++i)
            for (int a = 0; a < 3; ++a) wp[static_cast<size_t>(i) * 3 + a] = Path(i, a);
        return wp;
    }
    msnap_handle h_ = nullptr;
    int device_ = 0;
};

#endif  // MINIMUM_SNAP_GPU_HPP_
