/* oracle/geo_port.c -- TEST INFRASTRUCTURE (CPU restatement; never on the product path).
 *
 * Plain-C restatement of the reference's WGS84 <-> ECEF <-> ENU transforms, statement by statement, in the
 * reference's operation order (compile with -ffp-contract=off: the reference's CMake sets no -march/-O flags):
 *   constants WGS84_A / WGS84_E2, calcN, deg2rad, rad2deg   /root/reference/uavPathPlanning.hpp:134-173
 *   wgs84ToECEF                                              /root/reference/uavPathPlanning.cpp:894-910
 *   ecefToWGS84 (Bowring start + <= 10 fixed-point steps)    /root/reference/uavPathPlanning.cpp:926-968
 *   computeENURotationMatrix / ...Inverse                    /root/reference/uavPathPlanning.cpp:971-1020
 *   ecefToENU / enuToECEF                                    /root/reference/uavPathPlanning.cpp:1023-1044
 *   wgs84ToENU / enuToWGS84                                  /root/reference/uavPathPlanning.cpp:1047-1083
 *   wgs84ToENU_Batch / enuToWGS84_Batch (plain loops)        /root/reference/uavPathPlanning.cpp:1085-1108
 * The member functions live in class UavPathPlanner, whose translation unit needs yaml-cpp, GDAL and the out-of-tree
 * json.hpp / elog.h, so the reference itself cannot be compiled here; this port is PINNED on the
 * reference's own recorded run in /root/reference/readme.md:11-28 (7 WGS84 waypoints of uav31_0, their ENU values and
 * the WGS84 values recovered from those, printed to 15 decimals) -- tests/test_geo_oracle.py.
 *
 * Point layouts follow the reference structs: WGS84Point {lon, lat, alt} (degrees, degrees, metres; hpp:145-149),
 * ENUPoint {east, north, up} (hpp:152-156).
 */
#include <math.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static const double WGS84_A = 6378137.0;          /* hpp:134 */
static const double WGS84_E2 = 0.006694379990141; /* hpp:135 */

static double calcN(double lat_rad) { /* hpp:139-142 */
    double sin_lat = sin(lat_rad);
    return WGS84_A / sqrt(1.0 - WGS84_E2 * sin_lat * sin_lat);
}
static double deg2rad(double deg) { return deg * M_PI / 180.0; } /* hpp:166-168 */
static double rad2deg(double rad) { return rad * 180.0 / M_PI; } /* hpp:171-173 */

static void wgs84ToECEF(const double lla[3], double ecef[3]) { /* cpp:894-910 */
    double lat_rad = deg2rad(lla[1]);
    double lon_rad = deg2rad(lla[0]);
    double N = calcN(lat_rad);
    double cos_lat = cos(lat_rad), sin_lat = sin(lat_rad);
    double cos_lon = cos(lon_rad), sin_lon = sin(lon_rad);
    ecef[0] = (N + lla[2]) * cos_lat * cos_lon;
    ecef[1] = (N + lla[2]) * cos_lat * sin_lon;
    ecef[2] = (N * (1 - WGS84_E2) + lla[2]) * sin_lat;
}

/* returns the number of fixed-point steps executed (1..10), for the tests */
static int ecefToWGS84(const double ecef[3], double lla[3]) { /* cpp:926-968 */
    double p = sqrt(ecef[0] * ecef[0] + ecef[1] * ecef[1]);
    double theta = atan2(ecef[2] * WGS84_A, p * WGS84_A * (1 - WGS84_E2));
    double lat_rad = atan2(ecef[2] + WGS84_E2 * WGS84_A * (1 - WGS84_E2) * pow(sin(theta), 3) / (1 - WGS84_E2),
                           p - WGS84_E2 * WGS84_A * pow(cos(theta), 3));
    const int max_iterations = 10;
    const double tolerance = 1e-12;
    int steps = 0;
    for (int i = 0; i < max_iterations; ++i) {
        double N = calcN(lat_rad);
        double alt = p / cos(lat_rad) - N;
        double lat_new = atan2(ecef[2], p * (1 - WGS84_E2 * N / (N + alt)));
        ++steps;
        if (fabs(lat_new - lat_rad) < tolerance) {
            lat_rad = lat_new;
            break;
        }
        lat_rad = lat_new;
    }
    double lon_rad = atan2(ecef[1], ecef[0]);
    double N = calcN(lat_rad);
    double alt;
    if (p < 1e-12)
        alt = fabs(ecef[2]) - WGS84_A * sqrt(1 - WGS84_E2);
    else
        alt = p / cos(lat_rad) - N;
    lla[1] = rad2deg(lat_rad);
    lla[0] = rad2deg(lon_rad);
    lla[2] = alt;
    return steps;
}

static void ecefToENU(const double d[3], double ref_lat_rad, double ref_lon_rad, double enu[3]) { /* cpp:971-994, 1023-1032 */
    double cos_lat = cos(ref_lat_rad), sin_lat = sin(ref_lat_rad);
    double cos_lon = cos(ref_lon_rad), sin_lon = sin(ref_lon_rad);
    double R[3][3] = {{-sin_lon, cos_lon, 0.0},
                      {-sin_lat * cos_lon, -sin_lat * sin_lon, cos_lat},
                      {cos_lat * cos_lon, cos_lat * sin_lon, sin_lat}};
    enu[0] = R[0][0] * d[0] + R[0][1] * d[1] + R[0][2] * d[2];
    enu[1] = R[1][0] * d[0] + R[1][1] * d[1] + R[1][2] * d[2];
    enu[2] = R[2][0] * d[0] + R[2][1] * d[1] + R[2][2] * d[2];
}

static void enuToECEF(const double enu[3], double ref_lat_rad, double ref_lon_rad, double d[3]) { /* cpp:997-1020, 1035-1044 */
    double cos_lat = cos(ref_lat_rad), sin_lat = sin(ref_lat_rad);
    double cos_lon = cos(ref_lon_rad), sin_lon = sin(ref_lon_rad);
    double R[3][3] = {{-sin_lon, -sin_lat * cos_lon, cos_lat * cos_lon},
                      {cos_lon, -sin_lat * sin_lon, cos_lat * sin_lon},
                      {0.0, cos_lat, sin_lat}};
    d[0] = R[0][0] * enu[0] + R[0][1] * enu[1] + R[0][2] * enu[2];
    d[1] = R[1][0] * enu[0] + R[1][1] * enu[1] + R[1][2] * enu[2];
    d[2] = R[2][0] * enu[0] + R[2][1] * enu[1] + R[2][2] * enu[2];
}

/* ---- exported: one point ---------------------------------------------------------------------------------- */
void geo_port_wgs84_to_ecef(const double lla[3], double ecef[3]) { wgs84ToECEF(lla, ecef); }
int geo_port_ecef_to_wgs84(const double ecef[3], double lla[3]) { return ecefToWGS84(ecef, lla); }

void geo_port_wgs84_to_enu(const double target[3], const double reference[3], double enu[3]) { /* cpp:1047-1063 */
    double ref_ecef[3], target_ecef[3], delta[3];
    wgs84ToECEF(reference, ref_ecef);
    wgs84ToECEF(target, target_ecef);
    delta[0] = target_ecef[0] - ref_ecef[0];
    delta[1] = target_ecef[1] - ref_ecef[1];
    delta[2] = target_ecef[2] - ref_ecef[2];
    ecefToENU(delta, deg2rad(reference[1]), deg2rad(reference[0]), enu);
}

int geo_port_enu_to_wgs84(const double enu[3], const double reference[3], double lla[3]) { /* cpp:1066-1083 */
    double ref_ecef[3], delta[3], target_ecef[3];
    wgs84ToECEF(reference, ref_ecef);
    enuToECEF(enu, deg2rad(reference[1]), deg2rad(reference[0]), delta);
    target_ecef[0] = ref_ecef[0] + delta[0];
    target_ecef[1] = ref_ecef[1] + delta[1];
    target_ecef[2] = ref_ecef[2] + delta[2];
    return ecefToWGS84(target_ecef, lla);
}

/* ---- exported: batches (cpp:1085-1108 are plain loops; `threads` > 1 is the OpenMP-over-points CPU baseline) --- */
void geo_port_wgs84_to_enu_batch(long long n, const double *targets, const double reference[3], double *enu_out,
                                 int threads) {
    (void)threads;
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
    for (long long i = 0; i < n; ++i) geo_port_wgs84_to_enu(targets + 3 * i, reference, enu_out + 3 * i);
}

void geo_port_enu_to_wgs84_batch(long long n, const double *enu, const double reference[3], double *lla_out,
                                 int *steps_out, int threads) {
    (void)threads;
#pragma omp parallel for schedule(static) num_threads(threads) if (threads > 1)
    for (long long i = 0; i < n; ++i) {
        int s = geo_port_enu_to_wgs84(enu + 3 * i, reference, lla_out + 3 * i);
        if (steps_out) steps_out[i] = s;
    }
}
