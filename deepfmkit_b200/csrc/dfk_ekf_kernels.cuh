// EKF kernels: per-channel record statistics (initial dc and default measurement variance,
// fitters.py:253,256) and the tracking loop itself, one thread per channel (fitters.py:274-308).
#pragma once
#include "dfk_ekf_core.cuh"

namespace dfk {

constexpr int kStatsThreads = 256;
constexpr int kEkfTile = 32;    // samples per staged tile and thread
constexpr int kEkfStages = 3;   // tiles in flight per thread (cp.async groups)

// slab[c] = {mean(z_c), var(z_c)} over the T samples given, numpy's two-pass definition of var (ddof = 0).
// With acc != nullptr the slab's moments are merged into the running {n, mean, M2} of the channel (Chan et al.),
// which is how a record streamed in slabs gets its whole-record mean and variance.
__global__ void __launch_bounds__(kStatsThreads) channel_stats_kernel(const double* __restrict__ z, long long T,
                                                                      long long C, long long ld_t, long long ld_c,
                                                                      double* __restrict__ stats,
                                                                      double* __restrict__ acc) {
    __shared__ double red[kStatsThreads / 32];
    __shared__ double mean_sh;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (long long c = blockIdx.x; c < C; c += gridDim.x) {
        const double* zc = z + c * ld_c;
        double sum = 0.0;
        for (long long t = tid; t < T; t += kStatsThreads) sum += zc[t * ld_t];
        sum = warp_sum(sum);
        __syncthreads();
        if (lane == 0) red[warp] = sum;
        __syncthreads();
        if (tid == 0) {
            double v = 0.0;
            for (int w = 0; w < kStatsThreads / 32; ++w) v += red[w];
            mean_sh = v / static_cast<double>(T);
        }
        __syncthreads();
        const double mean = mean_sh;
        sum = 0.0;
        for (long long t = tid; t < T; t += kStatsThreads) {
            const double d = zc[t * ld_t] - mean;
            sum = fma(d, d, sum);
        }
        sum = warp_sum(sum);
        __syncthreads();
        if (lane == 0) red[warp] = sum;
        __syncthreads();
        if (tid == 0) {
            double m2 = 0.0;
            for (int w = 0; w < kStatsThreads / 32; ++w) m2 += red[w];
            if (stats) {
                stats[2 * c] = mean;
                stats[2 * c + 1] = m2 / static_cast<double>(T);
            }
            if (acc) {
                double* a = acc + 3 * c;
                const double na = a[0], nb = static_cast<double>(T), n = na + nb;
                const double delta = mean - a[1];
                a[1] += delta * (nb / n);
                a[2] += m2 + delta * delta * (na * nb / n);
                a[0] = n;
            }
        }
    }
}

// stats[c] = {mean, M2 / n} from the merged moments.
__global__ void stats_finish_kernel(const double* __restrict__ acc, long long C, double* __restrict__ stats) {
    const long long c = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (c >= C) return;
    stats[2 * c] = acc[3 * c + 1];
    stats[2 * c + 1] = acc[3 * c + 2] / acc[3 * c];
}

constexpr int kEkfStateStride = 32;  // doubles per channel in a carried state: x[5], P[25], r, spare

struct EkfLaunch {
    long long k0;   // absolute index of the first sample of this call (a multiple of R)
    double* state;  // carried filter state, C x kEkfStateStride (nullptr: single-call mode)
    double init[4];
    double init_dc;  // NaN -> per-channel mean from stats
    double p0[5];
    double q[5];
    double r_val;  // NaN -> per-channel variance from stats
    double w_m, f_samp;
    int cpw;  // channels per warp (power of two <= 32): lanes cpw..31 of every one-warp block idle
};

DFK_D void cp_async8(double* smem_dst, const double* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr(smem_dst)), "l"(gmem_src) : "memory");
}
DFK_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
DFK_D void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// One-warp blocks, one thread per channel.  The filter is a serial chain per channel, so a warp's pace does not
// depend on how many of its lanes are active; with C channels and 4 sub-partitions per SM the launcher picks the
// fewest channels per warp that still gives every warp a sub-partition of its own (cfg 4: 4096 channels -> 8 per
// warp, 512 warps on 592 sub-partitions).  Samples arrive through each thread's private shared-memory column by
// cp.async, kEkfStages tiles ahead: no registers, no barrier (a thread only ever reads what it copied itself),
// and either record layout ([C][T] or [T][C]) costs one 8-byte copy per sample.
__global__ void __launch_bounds__(32) ekf_kernel(const double* __restrict__ z, long long T, long long C,
                                                 long long ld_t, long long ld_c, long long R, EkfLaunch a,
                                                 const double* __restrict__ stats, double* __restrict__ rows) {
    __shared__ double stage[kEkfStages][kEkfTile][32];
    const int lane = threadIdx.x;
    const long long c = blockIdx.x * static_cast<long long>(a.cpw) + lane;
    if (lane >= a.cpw || c >= C) return;
    EkfState s;
    EkfConsts k;
    k.w_m = a.w_m;
    k.f_samp = a.f_samp;
    k.inv_fs = 1.0 / a.f_samp;
#pragma unroll
    for (int i = 0; i < 5; ++i) k.q[i] = a.q[i];
    double* carried = a.state ? a.state + c * kEkfStateStride : nullptr;
    if (carried && a.k0 > 0) {  // continuation slab: resume the filter where the previous call left it
#pragma unroll
        for (int i = 0; i < 5; ++i) s.x[i] = carried[i];
#pragma unroll
        for (int i = 0; i < 15; ++i) s.P[i] = carried[5 + i];
        k.r = carried[30];
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) s.x[i] = a.init[i];
        s.x[4] = (a.init_dc == a.init_dc) ? a.init_dc : stats[2 * c];
#pragma unroll
        for (int i = 0; i < 15; ++i) s.P[i] = 0.0;
        s.P[0] = a.p0[0]; s.P[5] = a.p0[1]; s.P[9] = a.p0[2]; s.P[12] = a.p0[3]; s.P[14] = a.p0[4];
        k.r = (a.r_val == a.r_val) ? a.r_val : stats[2 * c + 1];
    }

    const double* zc = z + c * ld_c;
    const long long nbuf = T / R;
    double* out = rows + c * nbuf * 8;
    const long long ntiles = (T + kEkfTile - 1) / kEkfTile;
    auto issue = [&](long long tile) {
        if (tile < ntiles) {
            double* dst = &stage[tile % kEkfStages][0][lane];
            const long long t0 = tile * kEkfTile;
            const int lim = (T - t0 < kEkfTile) ? static_cast<int>(T - t0) : kEkfTile;
            const double* src = zc + t0 * ld_t;
#pragma unroll 8
            for (int i = 0; i < lim; ++i) cp_async8(dst + i * 32, src + i * ld_t);
        }
        cp_async_commit();  // an empty group keeps the wait arithmetic uniform at the tail
    };
#pragma unroll
    for (int p = 0; p < kEkfStages - 1; ++p) issue(p);

    long long until_snap = R;  // samples left before the next state snapshot
    long long snaps = 0;
    double kd = static_cast<double>(a.k0);  // absolute sample index, exact in a double
    for (long long tile = 0; tile < ntiles; ++tile) {
        issue(tile + kEkfStages - 1);
        cp_async_wait<kEkfStages - 1>();
        const double* src = &stage[tile % kEkfStages][0][lane];
        const long long t0 = tile * kEkfTile;
        const int lim = (T - t0 < kEkfTile) ? static_cast<int>(T - t0) : kEkfTile;
#pragma unroll 1
        for (int i = 0; i < lim; ++i) {
            ekf_step(s, src[i * 32], kd, k);
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
        for (int i = 0; i < 15; ++i) carried[5 + i] = s.P[i];
        carried[30] = k.r;
    }
}

}  // namespace dfk
