// Host check of the atan2 kernel used by the direction-form ENU -> WGS84 path (cs_pathplan_b200/csrc/msnap_geo.cuh,
// geo_atan2): same coefficients (msnap_geo_atan.h), same Horner/fma sequence, same quadrant fix-ups; the division is
// the host's (the device uses a Newton-refined reciprocal with a residual correction, <= 1 ulp).
// Prints the largest |geo_atan2 - atan2| in units of 2^-53 over a dense sweep of directions and magnitudes.
#include <cmath>
#include <cstdio>

#include "msnap_geo_atan.h"

static double geo_atan2_host(double y, double x) {
    const double ax = std::fabs(x), ay = std::fabs(y);
    const bool sw = ay > ax, xneg = x < 0.0;
    const double mx = sw ? ay : ax, mn = sw ? ax : ay;
    const double t = mn / mx, u = t * t;
    const double pl = GEO_ATAN_POLY(std::fma, u);
    double a = std::fma(t * u, pl, t);
    const double off_hi = sw ? 1.57079632679489655800e+00 : (xneg ? 3.14159265358979311600e+00 : 0.0);
    const double off_lo = sw ? 6.12323399573676603587e-17 : (xneg ? 1.22464679914735320717e-16 : 0.0);
    a = (off_hi + (sw != xneg ? -a : a)) + off_lo;
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
