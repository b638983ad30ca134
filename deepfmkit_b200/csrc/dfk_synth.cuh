// 'snr'-mode synthetic DFMI records generated in HBM (physics.py:475-530 restated for the device):
//   y[t] = A (1 + C cos(phi0 + m cos(2 pi f_mod t / f_samp + psi0))) + sigma * N(0, 1).
// The noise comes from a counter-based generator (Philox4x32-10 + Box-Muller), so a record is a pure
// function of (seed, channel, t): any slab can be generated on any GPU.  It is statistically, not
// bitwise, equivalent to the reference's MT19937 stream; parity tests use reference-generated input.
#pragma once
#include "dfk_common.cuh"

namespace dfk {

DFK_D void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct SynthParams {
    double* x;
    long long T, C;
    long long t0;    // absolute index of the first sample generated (even); the stream is a function of absolute t
    long long ld_c;  // output samples between consecutive channels
    long long P;  // samples per modulation period when it is a whole number, else 0
    double f_ratio;  // f_mod / f_samp
    double m, amp, vis, phi0, dphi, psi0;
    double sigma_scale;  // 10^(-snr_db / 20)
    unsigned long long seed;
};

// Standard deviation of the noiseless signal's AC part over whole modulation periods:
// var = (A C)^2 [ (1 + cos(2 phi) J0(2m)) / 2 - cos(phi)^2 J0(m)^2 ]  (Jacobi-Anger, k = 0 terms).
DFK_D double clean_ac_rms(double amp, double vis, double phi, double m) {
    const double j0m = ::j0(m), j02m = ::j0(2.0 * m);
    const double cp = cos(phi);
    const double var = 0.5 * (1.0 + cos(2.0 * phi) * j02m) - cp * cp * j0m * j0m;
    return amp * vis * sqrt(fmax(var, 0.0));
}

__global__ void __launch_bounds__(256) synth_snr_kernel(const SynthParams p) {
    const long long pairs_per_ch = (p.T + 1) / 2;
    const long long total = pairs_per_ch * p.C;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256ll) {
        const long long c = i / pairs_per_ch;
        const long long tl = (i - c * pairs_per_ch) * 2;  // index inside the slab
        const long long t0 = p.t0 + tl;                  // absolute sample index
        const double phi = p.phi0 + static_cast<double>(c) * p.dphi;
        const double sigma = clean_ac_rms(p.amp, p.vis, phi, p.m) * p.sigma_scale;
        uint32_t r[4];
        const unsigned long long key = p.seed + static_cast<unsigned long long>(c);
        philox4x32_10(static_cast<uint32_t>(t0), static_cast<uint32_t>(t0 >> 32), 0x5eedu, 0u,
                      static_cast<uint32_t>(key), static_cast<uint32_t>(key >> 32), r);
        // two 53-bit uniforms in (0, 1]
        const double u1 = (static_cast<double>((static_cast<unsigned long long>(r[0]) << 21) ^ (r[1] >> 11)) + 1.0) *
                          (1.0 / 9007199254740992.0);
        const double u2 = static_cast<double>((static_cast<unsigned long long>(r[2]) << 21) ^ (r[3] >> 11)) *
                          (1.0 / 9007199254740992.0);
        const double rad = sqrt(-2.0 * log(u1));
        double sn, cn;
        sincospi(2.0 * u2, &sn, &cn);
        const double g[2] = {rad * cn, rad * sn};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const long long t = t0 + e;
            if (tl + e >= p.T) break;
            double frac;
            if (p.P > 0) {
                frac = static_cast<double>(t % p.P) / static_cast<double>(p.P);
            } else {
                const double ph = static_cast<double>(t) * p.f_ratio;
                frac = ph - floor(ph);
            }
            const double th = cospi(2.0 * frac + p.psi0 * (1.0 / kPi));
            const double clean = p.amp * (1.0 + p.vis * cos(phi + p.m * th));
            p.x[c * p.ld_c + tl + e] = fma(sigma, g[e], clean);
        }
    }
}

}  // namespace dfk
