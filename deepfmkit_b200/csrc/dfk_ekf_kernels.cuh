// EKF kernels: per-channel record statistics (initial dc and default measurement variance,
// fitters.py:253,256) and the tracking loop itself, one thread per channel (fitters.py:274-308).
#pragma once
#include "dfk_ekf_core.cuh"

namespace dfk {

constexpr int kStatsThreads = 256;
constexpr int kEkfThreads = 32;
constexpr int kEkfPrefetch = 8;

// stats[c] = {mean(z_c), var(z_c)} with numpy's two-pass definition of var (ddof = 0).
__global__ void __launch_bounds__(kStatsThreads) channel_stats_kernel(const double* __restrict__ z, long long T,
                                                                      long long C, long long ld_t, long long ld_c,
                                                                      double* __restrict__ stats) {
    __shared__ double red[kStatsThreads / 32];
    __shared__ double mean_sh;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (long long c = blockIdx.x; c < C; c += gridDim.x) {
        const double* zc = z + c * ld_c;
        double acc = 0.0;
        for (long long t = tid; t < T; t += kStatsThreads) acc += zc[t * ld_t];
        acc = warp_sum(acc);
        __syncthreads();
        if (lane == 0) red[warp] = acc;
        __syncthreads();
        if (tid == 0) {
            double v = 0.0;
            for (int w = 0; w < kStatsThreads / 32; ++w) v += red[w];
            mean_sh = v / static_cast<double>(T);
        }
        __syncthreads();
        const double mean = mean_sh;
        acc = 0.0;
        for (long long t = tid; t < T; t += kStatsThreads) {
            const double d = zc[t * ld_t] - mean;
            acc = fma(d, d, acc);
        }
        acc = warp_sum(acc);
        __syncthreads();
        if (lane == 0) red[warp] = acc;
        __syncthreads();
        if (tid == 0) {
            double v = 0.0;
            for (int w = 0; w < kStatsThreads / 32; ++w) v += red[w];
            stats[2 * c] = mean;
            stats[2 * c + 1] = v / static_cast<double>(T);
        }
    }
}

constexpr int kEkfStateStride = 32;  // doubles per channel in a carried state: x[5], P[25], r, spare

struct EkfLaunch {
    long long k0;   // absolute index of the first sample of this call (a multiple of R)
    double* state;  // carried filter state, C x kEkfStateStride (nullptr: single-call mode)
    double init[4];
    double p0[5];
    double q[5];
    double r_val;  // NaN -> per-channel variance from stats
    double w_m, f_samp;
};

__global__ void __launch_bounds__(kEkfThreads) ekf_kernel(const double* __restrict__ z, long long T, long long C,
                                                          long long ld_t, long long ld_c, long long R, EkfLaunch a,
                                                          const double* __restrict__ stats, double* __restrict__ rows) {
    const long long c = blockIdx.x * static_cast<long long>(kEkfThreads) + threadIdx.x;
    if (c >= C) return;
    EkfState s;
    EkfConsts k;
    k.w_m = a.w_m;
    k.f_samp = a.f_samp;
#pragma unroll
    for (int i = 0; i < 5; ++i) k.q[i] = a.q[i];
    double* carried = a.state ? a.state + c * kEkfStateStride : nullptr;
    if (carried && a.k0 > 0) {  // continuation slab: resume the filter where the previous call left it
#pragma unroll
        for (int i = 0; i < 5; ++i) s.x[i] = carried[i];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
#pragma unroll
            for (int j = 0; j < 5; ++j) s.P[i][j] = carried[5 + 5 * i + j];
        }
        k.r = carried[30];
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) s.x[i] = a.init[i];
        s.x[4] = stats[2 * c];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
#pragma unroll
            for (int j = 0; j < 5; ++j) s.P[i][j] = (i == j) ? a.p0[i] : 0.0;
        }
        k.r = (a.r_val == a.r_val) ? a.r_val : stats[2 * c + 1];
    }

    const double* zc = z + c * ld_c;
    const long long nbuf = T / R;
    double* out = rows + c * nbuf * 8;
    // Each thread stages its next kEkfPrefetch samples in registers (independent loads in flight)
    // while it steps through the current ones from its private shared-memory column.
    __shared__ double stage[kEkfPrefetch][kEkfThreads];
    double nxt[kEkfPrefetch];
#pragma unroll
    for (int i = 0; i < kEkfPrefetch; ++i) nxt[i] = (i < T) ? __ldg(zc + i * ld_t) : 0.0;
    long long until_snap = R;  // samples left before the next state snapshot
    long long snaps = 0;
    double kd = static_cast<double>(a.k0);  // absolute sample index, exact in a double
    for (long long t0 = 0; t0 < T; t0 += kEkfPrefetch) {
#pragma unroll
        for (int i = 0; i < kEkfPrefetch; ++i) stage[i][threadIdx.x] = nxt[i];
#pragma unroll
        for (int i = 0; i < kEkfPrefetch; ++i) {
            const long long tn = t0 + kEkfPrefetch + i;
            nxt[i] = (tn < T) ? __ldg(zc + tn * ld_t) : 0.0;
        }
        const int lim = (T - t0 < kEkfPrefetch) ? static_cast<int>(T - t0) : kEkfPrefetch;
#pragma unroll 1
        for (int i = 0; i < lim; ++i) {
            ekf_step(s, stage[i][threadIdx.x], kd, k);
            kd += 1.0;
            if (--until_snap == 0) {
                if (snaps < nbuf) {
                    double* row = out + snaps * 8;
                    row[0] = s.x[0]; row[1] = s.x[1]; row[2] = s.x[2]; row[3] = s.x[3]; row[4] = s.x[4];
                    row[5] = 0.0; row[6] = 1.0; row[7] = 0.0;  // ssq = 0, fitok = 1 (fitters.py:313-318)
                }
                ++snaps;
                until_snap = R;
            }
        }
    }
    if (carried) {
#pragma unroll
        for (int i = 0; i < 5; ++i) carried[i] = s.x[i];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
#pragma unroll
            for (int j = 0; j < 5; ++j) carried[5 + 5 * i + j] = s.P[i][j];
        }
        carried[30] = k.r;
    }
}

}  // namespace dfk
