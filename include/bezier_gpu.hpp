// include/bezier_gpu.hpp -- drop-in C++ replacement for the reference's math_util/bezier.hpp (the trajectory-generating
// part: what UavPathPlanner::Bezier_3D uses, /root/reference/uavPathPlanning.cpp:4477-4505).
//
// Same namespace, type names and member signatures as
//   /root/reference/math_util/bezier.hpp:91-96    struct math_util::BezierConfig
//   /root/reference/math_util/bezier.hpp:98-120   class  math_util::Bezier  (SetConfig, GenerateTrajectoryMatrix)
// so that `math_util::Bezier bezier; bezier.SetConfig(config); bezier.GenerateTrajectoryMatrix(route, "", distance_,
// V_avg_override)` (cpp:4489-4499) compiles and runs unchanged with this header included instead of the reference's.  The
// arithmetic happens in the CUDA library behind include/msnap.h (msnap_bezier_generate_batch_host); this header only
// converts Eigen values to the C ABI's row-major buffers.  The single-curve members Init / GeneratePath / GetResult /
// GetResultPath of the reference class are internals of GenerateTrajectoryMatrix there (bezier.cpp:163-166) and are not
// part of this boundary.  No CPU fallback: without an sm_100 GPU the first call throws std::runtime_error.
#ifndef BEZIER_GPU_HPP_
#define BEZIER_GPU_HPP_

#include <Eigen/Dense>

#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "msnap.h"

namespace math_util {

struct BezierConfig {  // bezier.hpp:91-96
    double min_radius;
    BezierConfig() { min_radius = 1.0; }
};

class Bezier {
public:
    explicit Bezier(int device = 0) : device_(device) {}
    Bezier(const Bezier &o) : config_(o.config_), device_(o.device_) {}
    Bezier &operator=(const Bezier &o) {
        if (this != &o) {
            release();
            config_ = o.config_;
            device_ = o.device_;
        }
        return *this;
    }
    ~Bezier() { release(); }

    void SetConfig(const BezierConfig &config) { config_ = config; }  // bezier.cpp:12-16

    // bezier.hpp:112, bezier.cpp:127-189.  Path (N x 3) -> sampled points (M x 3); fewer than 2 rows -> a 0 x 3 matrix
    // (bezier.cpp:129-131).  yaml_path and v_avg_override are unused, as in the reference.
    Eigen::MatrixXd GenerateTrajectoryMatrix(const Eigen::MatrixXd &Path, const std::string &yaml_path,
                                             double sample_distance_override = -1.0, double v_avg_override = -1.0) {
        (void)yaml_path;
        (void)v_avg_override;
        if (Path.rows() < 2) return Eigen::MatrixXd(0, 3);
        ensure();
        const int n = static_cast<int>(Path.rows());
        std::vector<double> wp(static_cast<size_t>(n) * 3);
        for (int i = 0; i < n; ++i)
            for (int a = 0; a < 3; ++a) wp[static_cast<size_t>(i) * 3 + a] = Path(i, a);
        long long off[2] = {0, 0};
        // sizing call: the exact row count, nothing written
        int rc = msnap_bezier_generate_batch_host(h_, sample_distance_override, config_.min_radius, 1, n - 1, nullptr, wp.data(),
                                                  0, off, nullptr, nullptr);
        if (rc != MSNAP_OK && rc != MSNAP_ERR_CAPACITY) return fail(rc);
        const long long rows = off[1];
        std::vector<double> s(static_cast<size_t>(rows > 0 ? rows : 1) * 3);
        rc = msnap_bezier_generate_batch_host(h_, sample_distance_override, config_.min_radius, 1, n - 1, nullptr, wp.data(), rows,
                                              off, s.data(), nullptr);
        if (rc != MSNAP_OK) return fail(rc);
        Eigen::MatrixXd out(rows, 3);
        for (long long i = 0; i < rows; ++i)
            for (int a = 0; a < 3; ++a) out(i, a) = s[static_cast<size_t>(i) * 3 + a];
        return out;
    }

    msnap_handle handle() {
        ensure();
        return h_;
    }

private:
    void ensure() {
        if (h_) return;
        const int rc = msnap_create(device_, &h_);
        if (rc != MSNAP_OK)
            throw std::runtime_error(std::string("math_util::Bezier: msnap_create failed: ") + msnap_status_string(rc) +
                                     " (this build has no CPU fallback)");
    }
    void release() {
        if (h_) msnap_destroy(h_);
        h_ = nullptr;
    }
    Eigen::MatrixXd fail(int rc) {
        std::cerr << "math_util::Bezier::GenerateTrajectoryMatrix: " << msnap_status_string(rc) << " " << msnap_last_error(h_)
                  << std::endl;
        return Eigen::MatrixXd(0, 3);
    }
    BezierConfig config_;
    msnap_handle h_ = nullptr;
    int device_ = 0;
};

}  // namespace math_util

#endif  // BEZIER_GPU_HPP_
