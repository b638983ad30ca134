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
constexpr int64_t kFoldRegisterPeriod = 2048;  // demod_fold_kernel: 256 consumer threads x 4 column pairs
constexpr int64_t kMaxFoldPeriod = 1 << 22;    // demod_fold_long_kernel takes longer fold lengths in column chunks
constexpr double kDriftRamp = 4e-13;
constexpr int kMaxFoldMul = 16;

struct DemodPlan {
    bool folded;     // a whole even fold length exists -> folded TMA kernels; else the direct kernel
    bool drift;      // apply the first-order frequency-offset term
    int64_t P;       // fold length in samples: kmul modulation periods (folded only)
    int64_t periods; // R / P
    int kmul;        // modulation periods per fold length: harmonic k of f_mod is harmonic kmul*k of the fold
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
    pl.kmul = 1;
    for (int k = 0; k < kMaxHarmonics; ++k) pl.delta[k] = 0.0;
    if (!(w0 > 0.0) || R <= 0 || N <= 0 || N > kMaxHarmonics) return pl;
    const double two_pi_hi = 6.283185307179586, two_pi_lo = 2.4492935982947064e-16;
    // The fold length is the smallest whole EVEN number of samples that holds a whole number q of modulation
    // periods (q = 1 for the BASELINE configs; q = 2 for an odd period such as the reference's own 30 kHz / 400 Hz
    // record, P = 75 -> fold 150; q up to kMaxFoldMul covers rational periods like 162.5 samples).  Harmonic k of
    // the modulation is then harmonic q*k of the fold, which only changes which twiddles the kernels tabulate.
    double pf = 0.0;
    int q = 0;
    for (int cand = 1; cand <= kMaxFoldMul; ++cand) {
        const double c = std::nearbyint(static_cast<double>(cand) * two_pi_hi / w0);
        if (!(c >= 2.0)) continue;
        if (c > static_cast<double>(kMaxFoldPeriod)) break;
        if (std::fabs(w0 * c - static_cast<double>(cand) * two_pi_hi) > 1e-12 * cand * two_pi_hi) continue;
        const int64_t ci = static_cast<int64_t>(c);
        if ((ci & 1) || (R % ci) != 0) continue;
        pf = c;
        q = cand;
        break;
    }
    if (q == 0) return pl;
    const int64_t P = static_cast<int64_t>(pf);
    const double qd = static_cast<double>(q);
    double worst = 0.0;
    for (int k = 0; k < N; ++k) {
        const double kf = static_cast<double>(k + 1);
        const double wk = kf * w0;  // what Python computes for (n + 1) * w0
        // wk * P and 2*pi*q*(k+1), each as an unevaluated sum hi + lo
        const double kq = kf * qd;
        const double a_hi = wk * pf, a_lo = std::fma(wk, pf, -a_hi);
        const double b_hi = kq * two_pi_hi, b_lo = std::fma(kq, two_pi_hi, -b_hi) + kq * two_pi_lo;
        pl.delta[k] = ((a_hi - b_hi) + (a_lo - b_lo)) / pf;
        const double ramp = std::fabs(pl.delta[k]) * static_cast<double>(R);
        if (ramp > worst) worst = ramp;
    }
    pl.folded = true;
    pl.P = P;
    pl.periods = R / P;
    pl.kmul = q;
    pl.drift = worst > kDriftRamp;
    return pl;
}

}  // namespace dfk
