// Monte-Carlo sweep with the realisations made where they are used (BASELINE config 5; workers.py:132-189 simulates
// and fits one single-period record per trial).  sweep_period_kernel is demod_period_kernel with the TMA producer
// replaced by the 'snr' generator: a warp draws the eight records of its group straight into its shared-memory
// stage -- the same Philox blocks, the same tabulated clean period, hence bit for bit the samples
// dfk_synth_snr_slab_dev writes for the same (seed, channel) -- and demodulates them from there.  No record ever
// touches HBM: the launch reads nothing and writes 8 (2N + 1) bytes per realisation.
#pragma once
#include "dfk_demod.cuh"
#include "dfk_synth.cuh"

namespace dfk {

struct SweepParams {
    SynthParams synth;   // x, T, C, ld_c, t0 unused; P = samples per record; seed + channel keys the noise
    double phi;          // interferometric phase of every realisation
    double* qi;
    double* dc;
    long long nbuf;      // realisations (channels)
    long long c0;        // channel index of realisation 0 (the seed offset of this launch)
    int N;
};

struct SweepSmem {
    PeriodSmem period;   // T, rows (stage / barrier slots unused)
    size_t off_clean, off_stage, off_x, off_bar, total;
};

inline __host__ __device__ SweepSmem sweep_smem_layout(int P, int N) {
    SweepSmem S;
    S.period = period_smem_layout(P, N, 0, kFoldConsumerWarps);
    size_t o = 0;
    S.period.off_t = o;
    o += static_cast<size_t>(S.period.quarter + 1) * S.period.nrows * 8;
    S.period.off_rows = o;
    o += static_cast<size_t>(S.period.nrows) * 2 * sizeof(int);
    o = (o + 15) & ~static_cast<size_t>(15);
    S.off_clean = o;
    o += static_cast<size_t>(P) * 8;
    o = (o + 15) & ~static_cast<size_t>(15);
    S.off_stage = o;
    o += static_cast<size_t>(kFoldConsumerWarps) * kPeriodNbw * P * 8;
    S.off_x = o;
    o += kFoldConsumerWarps * S.period.x_per_warp * 8;
    o = (o + 7) & ~static_cast<size_t>(7);
    S.off_bar = o;
    o += static_cast<size_t>(2 * kFoldConsumerWarps) * 8;
    S.total = o;
    return S;
}

// 16 warps: warp w < 8 demodulates the groups that generator warp w + 8 draws for it.  A pair shares one stage of
// eight records: the generator refills it while its consumer is in the product phase, which reads only the
// transposed combinations.  Two mbarriers per pair (full / free), one arrival each.
constexpr int kSweepThreads = 2 * kFoldConsumerWarps * 32;

__global__ void __launch_bounds__(kSweepThreads, 1) sweep_period_kernel(const SweepParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double sigma_sh;
    const int P = static_cast<int>(p.synth.P), N = p.N;
    const SweepSmem S = sweep_smem_layout(P, N);
    const PeriodSmem& L = S.period;
    double* T = reinterpret_cast<double*>(smem_raw + L.off_t);
    int* row_type = reinterpret_cast<int*>(smem_raw + L.off_rows);
    int* row_out = row_type + L.nrows;
    double* clean = reinterpret_cast<double*>(smem_raw + S.off_clean);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + S.off_bar);
    uint64_t* empty = full + kFoldConsumerWarps;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int pair = warp & (kFoldConsumerWarps - 1);
    const bool generator = warp >= kFoldConsumerWarps;
    constexpr int NBW = kPeriodNbw;
    if (tid == 0) {
        sigma_sh = clean_ac_rms(p.synth.amp, p.synth.vis, p.phi, p.synth.m) * p.synth.sigma_scale;
        for (int w = 0; w < kFoldConsumerWarps; ++w) {
            mbar_init(&full[w], 1);
            mbar_init(&empty[w], 1);
        }
        mbar_fence_init();
    }
    for (int j = tid; j < P; j += kSweepThreads) clean[j] = synth_clean(p.synth, p.phi, j);
    period_build_tables(P, N, L.nrows, L.quarter, T, row_type, row_out, tid, kSweepThreads);
    const double sigma = sigma_sh;
    double* stage = reinterpret_cast<double*>(smem_raw + S.off_stage) + static_cast<size_t>(pair) * NBW * P;
    double* X = reinterpret_cast<double*>(smem_raw + S.off_x) + pair * L.x_per_warp;
    const int quads = P >> 2;  // P % 4 == 0
    // floor(i / quads) = umulhi(i, ceil(2^32 / quads)) for i < 2^16 (here i < 8 * 64)
    const unsigned quads_magic = static_cast<unsigned>((0x100000000ull + quads - 1) / quads);
    const long long ngroups = (p.nbuf + NBW - 1) / NBW;
    uint32_t phase = 0;
    for (long long g = static_cast<long long>(blockIdx.x) * kFoldConsumerWarps + pair; g < ngroups;
         g += static_cast<long long>(gridDim.x) * kFoldConsumerWarps, phase ^= 1u) {
        const long long b0 = g * NBW;
        const int nb = static_cast<int>(min(static_cast<long long>(NBW), p.nbuf - b0));
        if (generator) {
            mbar_wait(&empty[pair], phase ^ 1u);  // the first wait passes: nothing to free yet
            for (int i = lane; i < nb * quads; i += 32) {
                const int s = static_cast<int>(__umulhi(static_cast<unsigned>(i), quads_magic)), q = i - s * quads;  // i / quads
                double y[4];
                synth_quad_values<true>(p.synth, 1, clean, p.phi, sigma, p.c0 + b0 + s, q, y);
                double2* dst = reinterpret_cast<double2*>(stage + s * P + 4 * q);
                dst[0] = make_double2(y[0], y[1]);
                dst[1] = make_double2(y[2], y[3]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[pair]);
        } else {
            mbar_wait(&full[pair], phase);
            period_combos(stage, X, P, L.xrow, lane);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[pair]);
            period_product(X, T, row_type, row_out, P, N, L.nrows, L.xrow, nb, b0, p.qi, p.dc, lane);
            __syncwarp();
        }
    }
}

}  // namespace dfk
