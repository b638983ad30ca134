// Bessel functions of the first kind, integer orders 0..nmax, by Miller's backward recurrence.
//
// Stands in for scipy.special.jv at the reference's call sites fit.py:106,108,160,275-276
// (scipy is a third-party dependency of the reference; its AMOS/cephes code is not restated --
// this is the textbook recurrence  J_{k-1}(x) = (2k/x) J_k(x) - J_{k+1}(x)  started above the
// turning point and normalised with  J_0 + 2 sum_{k>=1} J_{2k} = 1 ).
// tests/test_host_cores.py pins it against the committed scipy values (tests/golden/bessel_jv.npz).
#pragma once
#include "dfk_common.cuh"

namespace dfk {

constexpr double kBesselNoRescaleAbove = 1.0e-2;
constexpr double kBesselForwardAbove = 200.0;  // beyond this |x| > every supported order: go upward from J0, J1

// Starting order for the downward recurrence.  The truncation error of Miller's scheme is ~J_M(x), so M
// sits above |x| by 12*(|x|/2)^(1/3) + 5 (Airy decay past the turning point reaches 1e-17 there) and
// above the highest kept order.  Calibrated against a long-double reference on |x| <= 200, orders <= 65:
// worst absolute error 3.6e-16.
DFK_HD int bessel_start_order(double ax, int nmax) {
    int m = static_cast<int>(ax + 12.0 * cbrt(0.5 * ax) + 5.0);
    if (m < nmax + 2) m = nmax + 2;
    return m + (m & 1);  // even, so the normalisation sum ends on an even order
}

// out[k*stride] = J_k(x) for k = 0..nmax.  Returns the number of recurrence steps taken.
DFK_HD int bessel_j_upto(double x, int nmax, double* out, int stride) {
    const double ax = fabs(x);
    if (!(ax < 1.0e300)) {  // NaN or inf argument: propagate NaN like jv does
        for (int k = 0; k <= nmax; ++k) out[k * stride] = x - x;
        return 0;
    }
    if (ax == 0.0) {
        out[0] = 1.0;
        for (int k = 1; k <= nmax; ++k) out[k * stride] = 0.0;
        return 0;
    }
    int steps;
    if (ax > kBesselForwardAbove) {
        // nmax <= 65 < |x|: the upward recurrence is stable here.
        double jm = ::j0(ax), jk = ::j1(ax);
        out[0] = jm;
        if (nmax >= 1) out[stride] = jk;
        const double tox = 2.0 / ax;
        for (int k = 1; k < nmax; ++k) {
            const double jn = static_cast<double>(k) * tox * jk - jm;
            jm = jk;
            jk = jn;
            out[(k + 1) * stride] = jn;
        }
        steps = nmax;
    } else {
        const int mstart = bessel_start_order(ax, nmax);
        const double tox = 2.0 / ax;
        double bp = 0.0;      // b_{k+1}
        double b = 1.0e-200;  // b_k at k = mstart
        double sum = 0.0;     // b_0 + 2 * sum of even orders
        if (ax >= kBesselNoRescaleAbove) {
            // The unnormalised sequence ends near 1e-200 / |J_mstart(x)|; with mstart <= 70 that stays below
            // 1e130 for every |x| >= 1e-2, so the loop needs no range check.  mstart is even: two orders per trip.
            double kd = static_cast<double>(mstart);
            double even = 0.0;
            for (int k = mstart; k >= 2; k -= 2) {
                if (k <= nmax) out[k * stride] = b;
                even += b;
                double bm = (kd * tox) * b - bp;  // b_{k-1}
                bp = b;
                b = bm;
                kd -= 1.0;
                if (k - 1 <= nmax) out[(k - 1) * stride] = b;
                bm = (kd * tox) * b - bp;  // b_{k-2}
                bp = b;
                b = bm;
                kd -= 1.0;
            }
            sum = 2.0 * even;
        } else {
            for (int k = mstart; k >= 1; --k) {
                // here b = b_k, bp = b_{k+1}
                if (k <= nmax) out[k * stride] = b;
                if ((k & 1) == 0) sum += 2.0 * b;
                const double bm = static_cast<double>(k) * tox * b - bp;  // b_{k-1}
                bp = b;
                b = bm;
                if (fabs(b) > 1.0e200) {  // keep the unnormalised sequence in range (tiny |x|)
                    b *= 1.0e-200;
                    bp *= 1.0e-200;
                    sum *= 1.0e-200;
                    for (int q = k; q <= nmax; ++q) out[q * stride] *= 1.0e-200;
                }
            }
        }
        out[0] = b;
        sum += b;
        const double scale = 1.0 / sum;
        for (int k = 0; k <= nmax; ++k) out[k * stride] *= scale;
        steps = mstart;
    }
    if (x < 0.0) {  // J_k(-x) = (-1)^k J_k(x)
        for (int k = 1; k <= nmax; k += 2) out[k * stride] = -out[k * stride];
    }
    return steps;
}

}  // namespace dfk
