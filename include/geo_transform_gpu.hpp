// include/geo_transform_gpu.hpp -- drop-in bodies for UavPathPlanner's batched coordinate transforms
//   std::vector<ENUPoint>   UavPathPlanner::wgs84ToENU_Batch(const std::vector<WGS84Point>&, const WGS84Point&)
//   std::vector<WGS84Point> UavPathPlanner::enuToWGS84_Batch(const std::vector<ENUPoint>&,  const WGS84Point&)
// (/root/reference/uavPathPlanning.hpp:266-267, uavPathPlanning.cpp:1085-1108) and their one-point forms
// (hpp:264-265, cpp:1047-1083), over the C ABI in include/msnap.h.
//
// The templates take the reference's own point structs -- WGS84Point {lon, lat, alt} (hpp:145-149) and ENUPoint
// {east, north, up} (hpp:152-156): three doubles each, so a std::vector of them IS the C ABI's row-major [n][3]
// buffer and nothing is converted or copied on the host.  INTEGRATION.md section 6 shows the two member functions
// rewritten with these calls.  No CPU fallback: failures throw std::runtime_error with the library's message.
#ifndef GEO_TRANSFORM_GPU_HPP_
#define GEO_TRANSFORM_GPU_HPP_

#include <stdexcept>
#include <string>
#include <vector>

#include "msnap.h"

namespace msnap_geo {

template <class P>
constexpr void check_point_layout() {
    static_assert(sizeof(P) == 3 * sizeof(double) && alignof(P) == alignof(double),
                  "point struct must be three packed doubles (WGS84Point / ENUPoint of uavPathPlanning.hpp)");
}

inline void check(msnap_handle h, int rc, const char *where) {
    if (rc != MSNAP_OK)
        throw std::runtime_error(std::string(where) + ": " + msnap_status_string(rc) + " " + msnap_last_error(h));
}

// cpp:1085-1095
template <class ENU, class WGS>
std::vector<ENU> wgs84ToENU_Batch(msnap_handle h, const std::vector<WGS> &targets, const WGS &reference) {
    check_point_layout<ENU>();
    check_point_layout<WGS>();
    std::vector<ENU> results(targets.size());
    check(h, msnap_wgs84_to_enu_host(h, reinterpret_cast<const double *>(&reference), static_cast<long long>(targets.size()),
                                     reinterpret_cast<const double *>(targets.data()),
                                     reinterpret_cast<double *>(results.data())),
          "wgs84ToENU_Batch");
    return results;
}

// cpp:1098-1108
template <class WGS, class ENU>
std::vector<WGS> enuToWGS84_Batch(msnap_handle h, const std::vector<ENU> &targets, const WGS &reference) {
    check_point_layout<ENU>();
    check_point_layout<WGS>();
    std::vector<WGS> results(targets.size());
    check(h, msnap_enu_to_wgs84_host(h, reinterpret_cast<const double *>(&reference), static_cast<long long>(targets.size()),
                                     reinterpret_cast<const double *>(targets.data()),
                                     reinterpret_cast<double *>(results.data())),
          "enuToWGS84_Batch");
    return results;
}

// cpp:1047-1063 / cpp:1066-1083 (a batch of one)
template <class ENU, class WGS>
ENU wgs84ToENU(msnap_handle h, const WGS &target, const WGS &reference) {
    return wgs84ToENU_Batch<ENU, WGS>(h, std::vector<WGS>(1, target), reference)[0];
}
template <class WGS, class ENU>
WGS enuToWGS84(msnap_handle h, const ENU &enu, const WGS &reference) {
    return enuToWGS84_Batch<WGS, ENU>(h, std::vector<ENU>(1, enu), reference)[0];
}

}  // namespace msnap_geo

#endif  // GEO_TRANSFORM_GPU_HPP_
