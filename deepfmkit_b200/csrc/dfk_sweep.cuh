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
    size_t off_clean, off_stage, off_x, off_ctr, total;
};

// wc consumer warps (one transposed-combination scratch each), ns stages of eight records
inline __host__ __device__ SweepSmem sweep_smem_layout(int P, int N, int wc, int ns) {
    SweepSmem S;
    S.period = period_smem_layout(P, N, 0, wc);
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
    o += static_cast<size_t>(ns) * kPeriodNbw * P * 8;
    S.off_x = o;
    o += static_cast<size_t>(wc) * S.period.x_per_warp * 8;
    o = (o + 7) & ~static_cast<size_t>(7);
    S.off_ctr = o;
    o += static_cast<size_t>(2 * ns) * 8;
    S.total = o;
    return S;
}

// 16 warps: WG generator warps draw groups of eight records into a CTA-wide ring of NS stages, WC consumer warps
// demodulate them.  Group i of the CTA lives in stage i % NS, is drawn by generator i % WG and demodulated by consumer
// i % WC; a stage is released right after the combinations are formed (the product reads only the warp's transposed
// scratch), so it refills during the product.  Successive occupants of a stage belong to different warps on both
// sides, so the hand-over uses two monotonic counters per stage (groups written, groups released) rather than phase
// parities.  The generator's integer work is the longer half: 10 generator warps to 6 consumers measured best.
template <int WC, int WG, int NS>
__global__ void __launch_bounds__((WC + WG) * 32, 1) sweep_period_kernel(const SweepParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double sigma_sh;
    const int P = static_cast<int>(p.synth.P), N = p.N;
    const SweepSmem S = sweep_smem_layout(P, N, WC, NS);
    const PeriodSmem& L = S.period;
    double* T = reinterpret_cast<double*>(smem_raw + L.off_t);
    int* row_type = reinterpret_cast<int*>(smem_raw + L.off_rows);
    int* row_out = row_type + L.nrows;
    double* clean = reinterpret_cast<double*>(smem_raw + S.off_clean);
    volatile unsigned long long* written = reinterpret_cast<volatile unsigned long long*>(smem_raw + S.off_ctr);
    volatile unsigned long long* released = written + NS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool generator = warp >= WC;
    constexpr int NBW = kPeriodNbw;
    constexpr int kThreads = (WC + WG) * 32;
    if (tid == 0) sigma_sh = clean_ac_rms(p.synth.amp, p.synth.vis, p.phi, p.synth.m) * p.synth.sigma_scale;
    if (tid < 2 * NS) written[tid] = 0ull;  // (written and released are contiguous)
    for (int j = tid; j < P; j += kThreads) clean[j] = synth_clean(p.synth, p.phi, j);
    period_build_tables(P, N, L.nrows, L.quarter, T, row_type, row_out, tid, kThreads);
    const double sigma = sigma_sh;
    double* stage_base = reinterpret_cast<double*>(smem_raw + S.off_stage);
    const int quads = P >> 2;  // P % 4 == 0
    // floor(i / quads) = umulhi(i, ceil(2^32 / quads)) for i < 2^16 (here i < 8 * 64)
    const unsigned quads_magic = static_cast<unsigned>((0x100000000ull + quads - 1) / quads);
    const long long ngroups = (p.nbuf + NBW - 1) / NBW;
    const long long my_groups = ngroups > blockIdx.x ? (ngroups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (generator) {
        for (long long i = warp - WC; i < my_groups; i += WG) {
            const int st = static_cast<int>(i % NS);
            const unsigned long long k = static_cast<unsigned long long>(i / NS);  // occupants of the stage before this one
            const long long b0 = (blockIdx.x + i * gridDim.x) * NBW;
            const int nb = static_cast<int>(min(static_cast<long long>(NBW), p.nbuf - b0));
            double* stage = stage_base + static_cast<size_t>(st) * NBW * P;
            while (released[st] < k) __nanosleep(32);
            __threadfence_block();
            for (int q0 = lane; q0 < nb * quads; q0 += 32) {
                const int s = static_cast<int>(__umulhi(static_cast<unsigned>(q0), quads_magic)), q = q0 - s * quads;  // q0 / quads
                double y[4];
                synth_quad_values<true>(p.synth, 1, clean, p.phi, sigma, p.c0 + b0 + s, q, y);
                double2* dst = reinterpret_cast<double2*>(stage + s * P + 4 * q);
                dst[0] = make_double2(y[0], y[1]);
                dst[1] = make_double2(y[2], y[3]);
            }
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                written[st] = k + 1;
            }
        }
    } else {
        double* X = reinterpret_cast<double*>(smem_raw + S.off_x) + warp * L.x_per_warp;
        for (long long i = warp; i < my_groups; i += WC) {
            const int st = static_cast<int>(i % NS);
            const unsigned long long k = static_cast<unsigned long long>(i / NS);
            const long long b0 = (blockIdx.x + i * gridDim.x) * NBW;
            const int nb = static_cast<int>(min(static_cast<long long>(NBW), p.nbuf - b0));
            const double* stage = stage_base + static_cast<size_t>(st) * NBW * P;
            while (written[st] <= k) __nanosleep(32);
            __threadfence_block();
            period_combos(stage, X, P, L.xrow, lane);
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                released[st] = k + 1;
            }
            period_product(X, T, row_type, row_out, P, N, L.nrows, L.xrow, nb, b0, p.qi, p.dc, lane);
            __syncwarp();
        }
    }
}

}  // namespace dfk
