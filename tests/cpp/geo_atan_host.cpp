// Host check of the atan2 kernel used by the direction-form ENU -> WGS84 path (cs_pathplan_b200/csrc/msnap_geo.cuh,
// geo_atan2): same coefficients (msnap_geo_atan.h), same Horner/fma sequence, same quadrant fix-ups; the division is
// the host's (the device uses a Newton-refined reciprocal with a residual correction, <= 1 ulp).
// Prints the largest |geo_atan2 - atan2| in units of 2^-53 over a dense sweep of directions and magnitudes.
#include <cmath>
#include <cstdio>

#include "msnap_geo_atan.h"

static double geo_atan2_host(double y, double x) {
    static const double c[GEO_ATAN_N] = GEO_ATAN_COEFFS;
    const double ax = std::fabs(x), ay = std::fabs(y);
    const double mx = ax > ay ? ax : ay, mn = ax > ay ? ay : ax;
    const double t = mn / mx, u = t * t;
    double pl = c[GEO_ATAN_N - 1];
    for (int i = GEO_ATAN_N - 2; i >= 0; --i) pl = std::fma(pl, u, c[i]);
    double a = std::fma(t * u, pl, t);
    if (ay > ax) a = (1.57079632679489655800e+00 - a) + 6.12323399573676603587e-17;
    if (x < 0.0) a = (3.14159265358979311600e+00 - a) + 1.22464679914735320717e-16;
    if (mx == 0.0) a = 0.0;
    return std::copysign(a, y);
}

int main() {
    double worst = 0.0;
    const int n = 2000000;
    for (int i = 0; i < n; ++i) {
        const double ang = -3.14159265358979 + 6.28318530717958 * (i + 0.37) / n;
        const double r = std::pow(10.0, -3.0 + 12.0 * (int)(((long long)i * 7919) % 1000) / 1000.0);
        const double y = r * std::sin(ang), x = r * std::cos(ang);
        const long double ref = atan2l((long double)y, (long double)x);
        const double err = std::fabs((double)((long double)geo_atan2_host(y, x) - ref)) / 1.1102230246251565e-16;
        if (err > worst) worst = err;
    }
    std::printf("%.3f\n", worst);
    return 0;
}
