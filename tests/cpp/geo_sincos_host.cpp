// Host check of geo_sincos (cs_pathplan_b200/csrc/msnap_geo.cuh): same reduction constants and coefficients
// (msnap_geo_sincos.h), same fma sequence.  Prints the largest |error| of sin and cos in units of 2^-53 over a dense
// sweep of [-7, 7] (latitudes and longitudes in radians lie in [-pi, pi]) and a coarse one up to 1e5.
#include <cmath>
#include <cstdio>

#include "msnap_geo_sincos.h"

static void geo_sincos_host(double x, double &s_out, double &c_out) {
    static const double S[GEO_SIN_N] = GEO_SIN_COEFFS, C[GEO_COS_N] = GEO_COS_COEFFS;
    const double kf = std::rint(x * 0.63661977236758134308);
    const int k = (int)kf;
    double r = std::fma(-kf, GEO_PIO2_1, x);
    r = std::fma(-kf, GEO_PIO2_2, r);
    r = std::fma(-kf, GEO_PIO2_3, r);
    const double u = r * r;
    double ps = S[GEO_SIN_N - 1], pc = C[GEO_COS_N - 1];
    for (int i = GEO_SIN_N - 2; i >= 0; --i) ps = std::fma(ps, u, S[i]);
    for (int i = GEO_COS_N - 2; i >= 0; --i) pc = std::fma(pc, u, C[i]);
    const double sn = std::fma(r * u, ps, r), cs = std::fma(u * u, pc, std::fma(-0.5, u, 1.0));
    double s = (k & 1) ? cs : sn, c = (k & 1) ? sn : cs;
    if (k & 2) s = -s;
    if ((k + 1) & 2) c = -c;
    s_out = s;
    c_out = c;
}

int main() {
    double worst = 0.0;
    const int n = 2000000;
    for (int i = 0; i < 2 * n; ++i) {
        const double x = i < n ? -7.0 + 14.0 * (i + 0.31) / n : -1.0e5 + 2.0e5 * (i - n + 0.77) / n;
        double s, c;
        geo_sincos_host(x, s, c);
        const double es = std::fabs((double)((long double)s - sinl((long double)x))) / 1.1102230246251565e-16;
        const double ec = std::fabs((double)((long double)c - cosl((long double)x))) / 1.1102230246251565e-16;
        if (es > worst) worst = es;
        if (ec > worst) worst = ec;
    }
    std::printf("%.3f\n", worst);
    return 0;
}
