// One predict+update step of the 5-state DFMI EKF (reference loop body fitters.py:274-302).
// State x = [a, m, phi, psi, dc], full 5x5 covariance (the reference's simple-form update does not
// keep P symmetric, so no symmetry is assumed), scalar measurement.
#pragma once
#include "dfk_common.cuh"

namespace dfk {

DFK_HD double mul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);  // never contracted into an FMA
#else
    volatile double r = a * b;
    return r;
#endif
}
DFK_HD double add_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}

struct EkfState {
    double x[5];
    double P[5][5];
};

struct EkfConsts {
    double w_m;     // 2*pi*f_mod            (fitters.py:262)
    double f_samp;  // t_k = k / f_samp      (fitters.py:263)
    double q[5];    // process noise diagonal
    double r;       // measurement variance
};

// Sample index k is absolute: unlike the NLS lock-in, the EKF phase never restarts (fitters.py:280).
// kd: the absolute sample index as a double (exact for every index below 2^53).
DFK_HD void ekf_step(EkfState& s, double z, double kd, const EkfConsts& c) {
#pragma unroll
    for (int i = 0; i < 5; ++i) s.P[i][i] += c.q[i];  // P = F P F^T + Q with F = I (fitters.py:276)

    const double a = s.x[0], m = s.x[1], phi = s.x[2], psi = s.x[3], dc = s.x[4];
    // the angle is formed exactly as numpy does, w_m * (k / f_samp) + psi, each operation rounded once:
    // at t ~ 100 s its rounding (1e-10 rad) is the largest noise term the filter sees from arithmetic.
    const double t = kd / c.f_samp;
    const double theta = add_rn(mul_rn(c.w_m, t), psi);
    double st, ct;
    sincos_hd(theta, &st, &ct);
    const double arg = add_rn(phi, mul_rn(m, ct));
    double sa, ca;
    sincos_hd(arg, &sa, &ca);
    const double pred = a * ca + dc;
    const double H[5] = {ca, -a * sa * ct, -a * sa, a * m * sa * st, 1.0};  // fitters.py:287-293
    const double innov = z - pred;

    double PHt[5];
    double S = c.r;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < 5; ++j) acc += s.P[i][j] * H[j];
        PHt[i] = acc;
    }
    double hph = 0.0;
#pragma unroll
    for (int i = 0; i < 5; ++i) hph += H[i] * PHt[i];
    S += hph;
    const double invS = 1.0 / S;  // np.linalg.inv of the 1x1 innovation covariance (fitters.py:298)
    double K[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) K[i] = PHt[i] * invS;
#pragma unroll
    for (int i = 0; i < 5; ++i) s.x[i] += K[i] * innov;

    double HP[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int l = 0; l < 5; ++l) acc += H[l] * s.P[l][j];
        HP[j] = acc;
    }
    // P = (I - K H) P = P - K (H P)  (fitters.py:302, simple form, not Joseph)
#pragma unroll
    for (int i = 0; i < 5; ++i) {
#pragma unroll
        for (int j = 0; j < 5; ++j) s.P[i][j] -= K[i] * HP[j];
    }
}

}  // namespace dfk
