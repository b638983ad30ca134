// 'snr'-mode synthetic DFMI records generated in HBM (physics.py:475-530 restated for the device):
//   y[t] = A (1 + C cos(phi0 + m cos(2 pi f_mod t / f_samp + psi0))) + sigma * N(0, 1).
// The noise comes from a counter-based generator (Philox4x32-10 + Box-Muller), so a record is a pure
// function of (seed, channel, t): any slab, starting at any sample, can be generated on any GPU.  It is statistically, not
// bitwise, equivalent to the reference's MT19937 stream; parity tests use reference-generated input.
#pragma once
#include "dfk_common.cuh"
#include "dfk_rng.cuh"

namespace dfk {

struct SynthParams {
    double* x;
    long long T, C;
    long long t0;    // absolute index of the first sample generated; the stream is a function of absolute t
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

// Sample t of channel c:  clean_c[t] + sigma_c * z(seed + c, t).
//   * The noise uses one Philox4x32-10 block per four consecutive samples (counter = absolute t / 4, key = seed + c):
//     four 32-bit uniforms -> two Box-Muller pairs evaluated in fp32 (the noise is 20..60 dB below the signal, so its
//     own 2^-24 granularity is 7 digits below anything a fit can see; |z| <= 6.66).  fp64 log / sqrt / sincospi per
//     pair is what made the first generator compute-bound at 0.48 TB/s.
//   * The clean signal is periodic with P = f_samp / f_mod samples whenever that is a whole number: a block tabulates
//     one period of its channel in shared memory (cospi + cos in fp64, P evaluations) and reads it back by t mod P.
//     table == 0 (incommensurate or very long period, or records so short that a table per channel would cost more
//     than it saves) evaluates the two cosines per sample as before.
// Blocks: blockIdx.y = channel (or channel group when all channels share phi), blockIdx.x strides the quads.
constexpr int kSynthThreads = 256;
constexpr int kSynthMaxTable = 4096;  // doubles

DFK_D double synth_clean(const SynthParams& p, double phi, long long t) {
    double frac;
    if (p.P > 0) {
        frac = static_cast<double>(t % p.P) / static_cast<double>(p.P);
    } else {
        const double ph = static_cast<double>(t) * p.f_ratio;
        frac = ph - floor(ph);
    }
    const double th = cospi(2.0 * frac + p.psi0 * (1.0 / kPi));
    return p.amp * (1.0 + p.vis * cos(phi + p.m * th));
}

// The four samples t = 4q .. 4q+3 of channel c.  FIRST_PERIOD: the caller guarantees 4q + 3 < P (a record of one
// period), which spares the 64-bit remainder that finds the quad's place in the tabulated period; so does a caller
// that tracks the place itself and hands it over as jm0 = 4q mod P (>= 0).
template <bool FIRST_PERIOD = false>
DFK_D void synth_quad_values(const SynthParams& p, int table, const double* clean_tab, double phi, double sigma, long long c,
                             long long q, double y[4], int jm0 = -1) {
    const unsigned long long key = p.seed + static_cast<unsigned long long>(c);
    uint32_t r[4];
    philox4x32_10(static_cast<uint32_t>(q), static_cast<uint32_t>(q >> 32), 0x5eedu, 0u, static_cast<uint32_t>(key),
                  static_cast<uint32_t>(key >> 32), r);
    float z[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        box_muller(r[2 * h], r[2 * h + 1], z[2 * h], z[2 * h + 1]);
    }
    const long long t4 = q << 2;
    int jm = table ? (FIRST_PERIOD ? static_cast<int>(t4) : (jm0 >= 0 ? jm0 : static_cast<int>(t4 % p.P))) : 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        double clean;
        if (table) {
            clean = clean_tab[jm];
            if (!FIRST_PERIOD && ++jm == p.P) jm = 0;
            if (FIRST_PERIOD) ++jm;
        } else {
            clean = synth_clean(p, phi, t4 + e);
        }
        y[e] = fma(sigma, static_cast<double>(z[e]), clean);
    }
}

DFK_D void synth_quad(const SynthParams& p, int table, const double* clean_tab, double phi, double sigma, long long c,
                      long long q, int jm0 = -1) {
    double y[4];
    synth_quad_values(p, table, clean_tab, phi, sigma, c, q, y, jm0);
    const long long t4 = q << 2;
    double* dst = p.x + c * p.ld_c - p.t0 + t4;  // indexed by absolute t
    if (t4 >= p.t0 && t4 + 4 <= p.t0 + p.T && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
        reinterpret_cast<double2*>(dst)[0] = make_double2(y[0], y[1]);
        reinterpret_cast<double2*>(dst)[1] = make_double2(y[2], y[3]);
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (t4 + e >= p.t0 && t4 + e < p.t0 + p.T) dst[e] = y[e];
    }
}

__global__ void __launch_bounds__(kSynthThreads) synth_snr_kernel(const SynthParams p, int table, long long ch_per_block) {
    extern __shared__ double clean_tab[];
    __shared__ double sigma_sh;
    const long long q_first = p.t0 >> 2;  // first and one-past-last Philox block of the slab
    const long long q_end = (p.t0 + p.T + 3) >> 2;
    const long long nq = q_end - q_first;
    const long long c_lo = blockIdx.y * ch_per_block;
    const long long c_hi = (c_lo + ch_per_block < p.C) ? c_lo + ch_per_block : p.C;
    const long long stride = static_cast<long long>(gridDim.x) * kSynthThreads;
    const long long first = blockIdx.x * static_cast<long long>(kSynthThreads) + threadIdx.x;
    auto prepare = [&](double phi) {  // clean period and noise scale of a channel phase, once per block
        __syncthreads();
        if (threadIdx.x == 0) sigma_sh = clean_ac_rms(p.amp, p.vis, phi, p.m) * p.sigma_scale;
        if (table)
            for (int j = threadIdx.x; j < p.P; j += kSynthThreads) clean_tab[j] = synth_clean(p, phi, j);
        __syncthreads();
        return sigma_sh;
    };
    // The quads of one channel, strided over the grid's x dimension.  The place in the tabulated period advances by a
    // fixed step per trip, so the 64-bit remainder is taken once per thread and channel, not once per quad (software
    // 64-bit division is ~60 instructions next to the ~150 of a quad: the generator is bound by integer issue).
    auto channel = [&](long long c, double phi, double sigma) {
        long long q = q_first + first;
        int jm = table ? static_cast<int>((q << 2) % p.P) : -1;
        const int step = table ? static_cast<int>((stride << 2) % p.P) : 0;
        const int P = static_cast<int>(p.P);
        for (; q < q_end; q += stride) {
            synth_quad(p, table, clean_tab, phi, sigma, c, q, jm);
            if (table) {
                jm += step;
                if (jm >= P) jm -= P;
            }
        }
    };
    if (p.dphi == 0.0 || c_hi - c_lo == 1) {
        const double phi = p.phi0 + static_cast<double>(c_lo) * p.dphi;
        const double sigma = prepare(phi);
        if (c_hi - c_lo == 1) {
            channel(c_lo, phi, sigma);
            return;
        }
        // one phase for several short records: their (channel, quad) pairs are one flat index space, so that records
        // of one period (the Monte-Carlo shape) still fill every thread; 32-bit index arithmetic whenever it fits
        const long long total = (c_hi - c_lo) * nq;
        if (total < 0x7fffffffll && q_end < 0x1fffffffll) {
            const unsigned nq32 = static_cast<unsigned>(nq), P32 = static_cast<unsigned>(table ? p.P : 1);
            for (unsigned i = static_cast<unsigned>(first); i < static_cast<unsigned>(total); i += static_cast<unsigned>(stride)) {
                const unsigned dc = i / nq32;
                const unsigned q = static_cast<unsigned>(q_first) + (i - dc * nq32);
                synth_quad(p, table, clean_tab, phi, sigma, c_lo + dc, q, table ? static_cast<int>((q << 2) % P32) : -1);
            }
        } else {
            for (long long i = first; i < total; i += stride) {
                const long long dc = i / nq;
                synth_quad(p, table, clean_tab, phi, sigma, c_lo + dc, q_first + (i - dc * nq));
            }
        }
        return;
    }
    for (long long c = c_lo; c < c_hi; ++c) {
        const double phi = p.phi0 + static_cast<double>(c) * p.dphi;
        const double sigma = prepare(phi);
        channel(c, phi, sigma);
    }
}

}  // namespace dfk
