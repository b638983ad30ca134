// One predict+update step of the 5-state DFMI EKF (reference loop body fitters.py:274-302).
// State x = [a, m, phi, psi, dc], scalar measurement z_k = a cos(phi + m cos(w_m t_k + psi)) + dc.
//
// The step is a ~35-operation dependent chain (psi -> theta -> sincos -> arg -> sincos -> P H^T -> S -> 1/S -> x),
// and with a few thousand channels there is at most one warp per SM sub-partition, so throughput is set by the
// latency of that chain and by the number of fp64 instructions a warp issues per step (measured on B200,
// benchmarks/micro/fp64_micro.cu: 8.7 cycles per dependent DFMA, 2.13 issue cycles per warp-wide fp64 instruction
// whatever the number of active lanes, 224 cycles per library sincos, 72 per division).  Hence:
//   * sincos_cw: Cody-Waite reduction (exact for |x| < 1e6) + fdlibm's minimax kernels in Estrin form -- 9 dependent
//     operations instead of ~26; the library routine remains for larger arguments;
//   * the Jacobian row is H = [ca, sa g1, sa g2, sa g3, 1] with g = (-a ct, -a, a m st) known before sincos(arg)
//     returns, so P H^T = (P_i0 ca + P_i4) + sa (P_i1 g1 + P_i2 g2 + P_i3 g3) is two operations deep once sa, ca
//     arrive, the bracket being formed in the shadow of the second sincos;
//   * 1/S by rcp.approx + two Newton steps (<= 1 ulp); t_k = k / f_samp by a Markstein-corrected product (bit-equal
//     to the IEEE quotient numpy forms, fitters.py:263);
//   * P is carried as its upper triangle.  The reference's simple-form update P <- (I - K H) P keeps P symmetric up
//     to rounding only; carrying the triangle differs from it by that rounding noise, which the filter contracts
//     (measured deviation of the states from the reference loop: tests/test_host_cores.py).
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

// sin and cos of x to <= 1 ulp-level absolute error (1.7e-16) for |x| < 1e6 (kSincosFastLimit); no branches.
// n = rint(x 2/pi) by the 1.5 2^52 trick; r = x - n pi/2 with pi/2 = c1 + c2, c1 holding 33 bits so that n c1 is
// exact for n < 2^20; sin r and cos r by the fdlibm minimax polynomials on |r| <= pi/4, evaluated pairwise.
constexpr double kSincosFastLimit = 1.0e6;

// The constants live in the constant bank on the device: an fp64 instruction takes a constant-bank operand directly,
// whereas a 64-bit literal costs two moves per use (ptxas re-materialised them every step: ~40 of 200 instructions).
#if defined(__CUDACC__)
#define DFK_SC_STORAGE __device__ __constant__
#else
#define DFK_SC_STORAGE static const
#endif
DFK_SC_STORAGE double kSc[16] = {
    6755399441055744.0,            // 0: 1.5 * 2^52
    6.36619772367581382433e-01,    // 1: 2 / pi
    1.57079632673412561417e+00,    // 2: pi/2, leading 33 bits
    6.07710050650619224932e-11,    // 3: pi/2 - kSc[2]
    -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,   // 4..9: S1..S6
    2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10,
    4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,    // 10..15: C1..C6
    -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11};

DFK_HD void sincos_cw_fast(double x, double* s, double* c) {
    const double kMagic = kSc[0];
    const double t = fma(x, kSc[1], kMagic);
    const double n = t - kMagic;
#if defined(__CUDA_ARCH__)
    const int q = __double2loint(t);
#else
    long long bits;
    __builtin_memcpy(&bits, &t, 8);
    const int q = static_cast<int>(bits & 0xffffffffll);
#endif
    double r = fma(-n, kSc[2], x);
    r = fma(-n, kSc[3], r);
    const double z = r * r;
    const double z2 = z * z, r3 = r * z;
    const double s01 = fma(z, kSc[5], kSc[4]);
    const double s23 = fma(z, kSc[7], kSc[6]);
    const double s45 = fma(z, kSc[9], kSc[8]);
    const double c01 = fma(z, kSc[11], kSc[10]);
    const double c23 = fma(z, kSc[13], kSc[12]);
    const double c45 = fma(z, kSc[15], kSc[14]);
    const double z4 = z2 * z2;
    const double half = fma(-0.5, z, 1.0);
    const double sp = fma(s45, z4, fma(s23, z2, s01));
    const double cp = fma(c45, z4, fma(c23, z2, c01));
    const double sr = fma(r3, sp, r);
    const double cr = fma(z2, cp, half);
    const double a = (q & 1) ? cr : sr;
    const double b = (q & 1) ? sr : cr;
    *s = (q & 2) ? -a : a;
    *c = ((q + 1) & 2) ? -b : b;
}

// The same with the library's Payne-Hanek path behind it for large, infinite or NaN arguments.
DFK_HD void sincos_cw(double x, double* s, double* c) {
    if (!(fabs(x) < kSincosFastLimit)) {
        sincos_hd(x, s, c);
        return;
    }
    sincos_cw_fast(x, s, c);
}

// 1 / s to <= 1 ulp for normal s with 1e-290 < |s| < 1e290 (kRecipLo, kRecipHi); no branches.
constexpr double kRecipLo = 1.0e-290, kRecipHi = 1.0e290;

DFK_HD double recip_fast(double s) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(s));
    double e = fma(-s, r, 1.0);
    r = fma(r, e, r);
    e = fma(-s, r, 1.0);
    return fma(r, e, r);
#else
    return 1.0 / s;
#endif
}

// The same, also returning the value after the first Newton step (relative error < 2^-45).
DFK_HD void recip_fast2(double s, double* after_one, double* converged) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(s));
    double e = fma(-s, r, 1.0);
    r = fma(r, e, r);
    *after_one = r;
    e = fma(-s, r, 1.0);
    *converged = fma(r, e, r);
#else
    *after_one = *converged = 1.0 / s;
#endif
}

DFK_HD double recip(double s) {
    const double as = fabs(s);
    if (!(as > kRecipLo && as < kRecipHi)) return 1.0 / s;
    return recip_fast(s);
}

struct EkfState {
    double x[5];
    // upper triangle of the covariance, row by row: 00 01 02 03 04 | 11 12 13 14 | 22 23 24 | 33 34 | 44
    double P[15];
    // covariance update of the last step, not yet applied: P -= kp hp^T.  Deferring it to the next call lets its 15
    // FMAs issue in the latency shadow of that step's first sincos instead of delaying the start of its chain
    // (the next angle needs only psi).  ekf_flush applies it; a fresh state holds zeros.
    double kp[5], hp[5];
};

DFK_HD int tri(int i, int j) {  // index of P_ij, i <= j
    return i * 5 - (i * (i - 1)) / 2 + (j - i);
}

struct EkfConsts {
    double w_m;       // 2*pi*f_mod            (fitters.py:262)
    double f_samp;    // t_k = k / f_samp      (fitters.py:263)
    double inv_fs;    // fl(1 / f_samp)
    double q[5];      // process noise diagonal
    double r;         // measurement variance
};

// fl(kd / f_samp): a Markstein correction of the product with the rounded reciprocal.
DFK_HD double sample_time(double kd, const EkfConsts& c) {
    const double q0 = kd * c.inv_fs;
    const double rem = fma(-q0, c.f_samp, kd);
    return fma(rem, c.inv_fs, q0);
}

// w_m * t_k exactly as numpy forms it (fitters.py:263,280), each operation rounded once: at t ~ 100 s the rounding of
// the carrier angle (1e-10 rad) is the largest arithmetic noise the filter sees.  Sample index k is absolute: unlike
// the NLS lock-in, the EKF phase never restarts.  kd: the index as a double (exact below 2^53).
DFK_HD double carrier_angle(double kd, const EkfConsts& c) { return mul_rn(c.w_m, sample_time(kd, c)); }

// P -= kp hp^T (fitters.py:302: P = (I - K H) P = P - K (H P), with H P = (P H^T)^T for the symmetric P).
DFK_HD void ekf_apply_pending(EkfState& s) {
    double* P = s.P;
    const double* k = s.kp;
    const double* h = s.hp;
    P[0] = fma(-k[0], h[0], P[0]); P[1] = fma(-k[0], h[1], P[1]); P[2] = fma(-k[0], h[2], P[2]);
    P[3] = fma(-k[0], h[3], P[3]); P[4] = fma(-k[0], h[4], P[4]);
    P[5] = fma(-k[1], h[1], P[5]); P[6] = fma(-k[1], h[2], P[6]); P[7] = fma(-k[1], h[3], P[7]);
    P[8] = fma(-k[1], h[4], P[8]);
    P[9] = fma(-k[2], h[2], P[9]); P[10] = fma(-k[2], h[3], P[10]); P[11] = fma(-k[2], h[4], P[11]);
    P[12] = fma(-k[3], h[3], P[12]); P[13] = fma(-k[3], h[4], P[13]);
    P[14] = fma(-k[4], h[4], P[14]);
}

DFK_HD void ekf_flush(EkfState& s) {
    ekf_apply_pending(s);
#pragma unroll
    for (int i = 0; i < 5; ++i) s.kp[i] = s.hp[i] = 0.0;
}

// One step on sample z whose carrier angle w_m t_k is A.
// FAST: the branch-free routines above; the return value tells whether every argument stayed inside their range
// (the caller replays the samples with FAST = false otherwise).  FAST = false: the checked routines, always true.
template <bool FAST>
DFK_HD bool ekf_step(EkfState& s, double z, double A, const EkfConsts& c) {
    double* P = s.P;
    const double a = s.x[0], m = s.x[1], phi = s.x[2], psi = s.x[3], dc = s.x[4];
    const double theta = add_rn(A, psi);  // fitters.py:280
    double st, ct;
    if (FAST) sincos_cw_fast(theta, &st, &ct); else sincos_cw(theta, &st, &ct);
    // (in the shadow of that sincos) the covariance update of the previous step, then
    // P = F P F^T + Q with F = I (fitters.py:276)
    ekf_apply_pending(s);
    P[0] += c.q[0]; P[5] += c.q[1]; P[9] += c.q[2]; P[12] += c.q[3]; P[14] += c.q[4];
    const double arg = fma(m, ct, phi);  // fitters.py:281
    // Jacobian row (fitters.py:287-293) without its common factor sin(arg): H = [ca, sa g1, sa g2, sa g3, 1]
    const double g1 = -(a * ct), g2 = -a, g3 = (a * m) * st;
    // rows of P (symmetric): G_i = P_i1 g1 + P_i2 g2 + P_i3 g3, formed while sincos(arg) is in flight
    const double G0 = fma(P[3], g3, fma(P[2], g2, P[1] * g1));
    const double G1 = fma(P[7], g3, fma(P[6], g2, P[5] * g1));
    const double G2 = fma(P[10], g3, fma(P[9], g2, P[6] * g1));
    const double G3 = fma(P[12], g3, fma(P[10], g2, P[7] * g1));
    const double G4 = fma(P[13], g3, fma(P[11], g2, P[8] * g1));
    const double gPg = fma(g3, G3, fma(g2, G2, g1 * G1));
    const double twoG0 = G0 + G0, twoG4 = G4 + G4, twoP04 = P[4] + P[4], rP44 = c.r + P[14];
    double sa, ca;
    if (FAST) sincos_cw_fast(arg, &sa, &ca); else sincos_cw(arg, &sa, &ca);
    const double innov = z - fma(a, ca, dc);  // fitters.py:283,296
    // P H^T
    const double h0 = fma(sa, G0, fma(P[0], ca, P[4]));
    const double h1 = fma(sa, G1, fma(P[1], ca, P[8]));
    const double h2 = fma(sa, G2, fma(P[2], ca, P[11]));
    const double h3 = fma(sa, G3, fma(P[3], ca, P[13]));
    const double h4 = fma(sa, G4, fma(P[4], ca, P[14]));
    // S = H P H^T + R (fitters.py:297) as a quadratic form in (ca, sa) whose coefficients were ready before sa, ca:
    //   S = (R + P44) + ca (ca P00 + 2 P04) + sa (sa gPg + 2 (ca G0 + G4)),   gPg = g^T P g
    // -- three operations deep instead of six through P H^T.
    const double S = fma(sa, fma(sa, gPg, fma(ca, twoG0, twoG4)), fma(ca, fma(ca, P[0], twoP04), rP44));
    // np.linalg.inv of the 1x1 innovation covariance (fitters.py:298).  The state update takes the reciprocal after
    // one Newton step (relative error < 2^-45: it scales an update of ~1e-4 of the state, i.e. acts like a change of
    // R in its 14th digit); the covariance takes the fully converged one.
    double inv1, invS;
    if (FAST) {
        recip_fast2(S, &inv1, &invS);
    } else {
        inv1 = invS = recip(S);
    }
    const double aS = fabs(S);
    const bool in_range = !FAST || (fabs(theta) < kSincosFastLimit && fabs(arg) < kSincosFastLimit && aS > kRecipLo &&
                                    aS < kRecipHi);
    const double gain = innov * inv1;
    s.x[0] = fma(h0, gain, a);
    s.x[1] = fma(h1, gain, m);
    s.x[2] = fma(h2, gain, phi);
    s.x[3] = fma(h3, gain, psi);
    s.x[4] = fma(h4, gain, dc);
    // K = P H^T / S; the update P -= K (H P) waits for the next call
    s.kp[0] = h0 * invS; s.kp[1] = h1 * invS; s.kp[2] = h2 * invS; s.kp[3] = h3 * invS; s.kp[4] = h4 * invS;
    s.hp[0] = h0; s.hp[1] = h1; s.hp[2] = h2; s.hp[3] = h3; s.hp[4] = h4;
    return in_range;
}

}  // namespace dfk
