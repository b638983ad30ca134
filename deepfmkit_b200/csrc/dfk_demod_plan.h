// Host-side planning for the lock-in demodulation kernels (plain C++, shared by the CUDA
// launcher and the CPU-side unit tests).
//
// The reference demodulates with angle fl(fl((k+1)*w0) * t), t = 0..R-1 (fit.py:55-64).  When the
// modulation period is a whole number P of samples, cos/sin of the ideal angle 2*pi*(k+1)*t/P are
// periodic in t, so a buffer of n = R/P periods can be folded to P partial sums while it streams
// from HBM (1 add per 8 bytes) and the N harmonics are then taken from the folded period.
// The reference's harmonic frequency w_k = fl((k+1)*w0) differs from the ideal one by
// delta_k = w_k - 2*pi*(k+1)/P (a few 1e-16 rad/sample); over a buffer that is a phase ramp of up
// to delta_k*R ~ 1e-13 rad.  The folded kernel reproduces it to first order from a second folded
// sum U_j = sum_c (j + c*P) x[j + c*P]:
//     Q_k = (sum_j S_j cos(th_kj) - delta_k sum_j U_j sin(th_kj)) / R
//     I_k = (sum_j S_j sin(th_kj) + delta_k sum_j U_j cos(th_kj)) / R
// which brings the result to the reference's own rounding floor (measured 2.5e-14 of max|IQ|).
// Leaving the term out costs at most ~ramp/2 of max|IQ| (ramp = max_k |delta_k| R; measured 3.5e-14 at
// ramp = 2.2e-13), so it is only switched on when ramp > kDriftRamp = 4e-13, i.e. before the error could
// reach 2e-13 -- a fifth of the 1e-12 parity gate.  (Many harmonics or long buffers get there; the
// BASELINE configs with N = 10 do not.)
#pragma once
#include <cmath>
#include <cstdint>

namespace dfk {

constexpr int kMaxHarmonics = 64;
constexpr int64_t kMaxFoldPeriod = 2048;  // 256 consumer threads x 4 column pairs
constexpr double kDriftRamp = 4e-13;

struct DemodPlan {
    bool folded;     // integer even period -> folded TMA kernel; else general kernel
    bool drift;      // apply the first-order frequency-offset term
    int64_t P;       // samples per modulation period (folded only)
    int64_t periods; // R / P
    double delta[kMaxHarmonics];  // w_k - 2*pi*(k+1)/P, evaluated in double-double
};

// exp(i*2*pi*j/P) rounded to double (host version; the device uses sincospi).
inline void unit_circle(int64_t j, int64_t P, double* c, double* s) {
    const long double ang = 6.283185307179586476925286766559005768L * static_cast<long double>(j) /
                            static_cast<long double>(P);
    *c = static_cast<double>(cosl(ang));
    *s = static_cast<double>(sinl(ang));
}

inline DemodPlan make_demod_plan(int64_t R, double w0, int N) {
    DemodPlan pl;
    pl.folded = false;
    pl.drift = false;
    pl.P = 0;
    pl.periods = 0;
    for (int k = 0; k < kMaxHarmonics; ++k) pl.delta[k] = 0.0;
    if (!(w0 > 0.0) || R <= 0 || N <= 0 || N > kMaxHarmonics) return pl;
    const double two_pi_hi = 6.283185307179586, two_pi_lo = 2.4492935982947064e-16;
    const double pf = std::nearbyint(two_pi_hi / w0);
    if (!(pf >= 2.0) || pf > static_cast<double>(kMaxFoldPeriod)) return pl;
    const int64_t P = static_cast<int64_t>(pf);
    if ((P & 1) || (R % P) != 0) return pl;
    if (std::fabs(w0 * pf - two_pi_hi) > 1e-12 * two_pi_hi) return pl;
    double worst = 0.0;
    for (int k = 0; k < N; ++k) {
        const double kf = static_cast<double>(k + 1);
        const double wk = kf * w0;  // what Python computes for (n + 1) * w0
        // wk * P and 2*pi*(k+1), each as an unevaluated sum hi + lo
        const double a_hi = wk * pf, a_lo = std::fma(wk, pf, -a_hi);
        const double b_hi = kf * two_pi_hi, b_lo = std::fma(kf, two_pi_hi, -b_hi) + kf * two_pi_lo;
        pl.delta[k] = ((a_hi - b_hi) + (a_lo - b_lo)) / pf;
        const double ramp = std::fabs(pl.delta[k]) * static_cast<double>(R);
        if (ramp > worst) worst = ramp;
    }
    pl.folded = true;
    pl.P = P;
    pl.periods = R / P;
    pl.drift = worst > kDriftRamp;
    return pl;
}

}  // namespace dfk
