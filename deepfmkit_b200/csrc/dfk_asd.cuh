// 'asd'-mode physics of the reference on the device (physics.py:423-440, 557-611, 615-722), batched over Monte-Carlo
// trials: the exact model the Experiment runner simulates once per trial (experiments.py:15-88).
//   g(t)        = waveform(w_mod t + psi) / max|waveform|                      (unitless frequency modulation)
//   phi_mod(t)  = (2 pi / fs) cumsum((df + n_df(t)) g(t))                      (physics.py:675)
//   tau_dl(t)   = (A_arm sin(w_arm t + psi_arm) + phi lambda / 2 pi) / c       (dynamic main channel)
//   Phi(t)      = w0 ((tau_m + tau_dl) - tau_r) + interp(t - tau_m - tau_dl) - interp(t - tau_r)
//   y(t)        = (A + n_amp(t)) (1 + C cos Phi(t))
// with n_df, n_amp white, sigma = asd sqrt(fs / 2) (physics.py:591-597).  The coloured sources (laser frequency and
// arm-length noise) come from the third-party pyplnoise in the reference and are not generated here; like the
// reference's engine (generate(..., external_noise=...), physics.py:430-434) the kernel takes pre-computed noise series
// for all four sources instead -- laser frequency n_f(t) (w0 -> 2 pi (f0 + n_f)), amplitude, df and, on a dynamic
// channel, arm length (added to the path) -- which then replace the internal draws.
// One CTA per trial: block-wide prefix sum of the frequency waveform into an L2-resident scratch row, then the
// delayed-time interpolation np.interp performs, sample by sample.
#pragma once
#include "dfk_common.cuh"
#include "dfk_rng.cuh"

namespace dfk {

constexpr int kAsdThreads = 256;
constexpr int kAsdPerThread = 4;
constexpr int kAsdTerms = 6;
constexpr double kLightSpeed = 299792458.0;  // speed of light in m/s, the exact SI value the reference uses

struct AsdTrial {  // one Monte-Carlo trial, 26 doubles
    double amp, vis, df, w_mod, psi;
    double omega0;               // 2 pi c / lambda
    double tau_r, tau_m;         // arm delays
    double path0;                // phi lambda / (2 pi)
    double arm_amp, arm_w, arm_psi;
    double sigma_amp, sigma_df;  // white-noise standard deviations per sample
    double seed;                 // noise key (trial number), < 2^53
    double nterms;               // > 0: g = sum_k a_k cos(h_k theta + p_k);  0: g from the waveform table
    double table;                // waveform table row when nterms == 0
    double term[kAsdTerms][3];   // h_k, a_k, p_k   (only the first nterms)
    double dynamic;              // != 0: the channel's arm moves (takes the external arm-length noise)
    double noise_row;            // >= 0: row of the external noise series this trial reads; < 0: internal white noise
};
static_assert(sizeof(AsdTrial) == (19 + 3 * kAsdTerms) * sizeof(double), "AsdTrial layout");

struct AsdParams {
    const AsdTrial* trials;
    long long ntrials;
    long long N;            // samples per trial
    double fs;
    double scale;           // 2 pi / (1 / dt) with dt = t[1] - t[0], as physics.py:672-675 forms it
    const double* tables;   // [ntables][N] unnormalised waveform rows (arbitrary host-evaluated waveforms)
    double* scratch;        // [ntrials][N] running sum of the frequency waveform
    double* y;              // [ntrials][ld]
    long long ld;
    double* truth;          // optional [ntrials][ld]: ground-truth phase w0 ((tau_m + tau_dl) - tau_r)
    // external noise series [rows][N], each optional: laser frequency, amplitude, df, arm length (physics.py:430-434)
    const double* ext_fn;
    const double* ext_amp;
    const double* ext_df;
    const double* ext_arm;
};

DFK_D double asd_waveform(const AsdTrial& tr, const AsdParams& P, long long i) {
    const int nt = static_cast<int>(tr.nterms);
    if (nt == 0) return P.tables[static_cast<long long>(tr.table) * P.N + i];
    const double theta = tr.w_mod * (static_cast<double>(i) / P.fs) + tr.psi;  // omega_mod * time_axis + psi
    double g = 0.0;
    for (int k = 0; k < nt; ++k) g += tr.term[k][1] * cos(tr.term[k][0] * theta + tr.term[k][2]);
    return g;
}

// np.interp(x, t, f) on the uniform grid t_j = j / fs with the end values held outside it
DFK_D double asd_interp(const double* __restrict__ f, long long N, double fs, double scale, double x) {
    if (!(x > 0.0)) return scale * f[0];
    double u = x * fs;
    long long j = static_cast<long long>(u);
    if (j >= N - 1) return scale * f[N - 1];
    // guard the rounding of x * fs against the grid the reference searches (t_j = j / fs)
    double tj = static_cast<double>(j) / fs;
    if (tj > x) {
        --j;
        tj = static_cast<double>(j) / fs;
    }
    const double tj1 = static_cast<double>(j + 1) / fs;
    const double f0 = scale * f[j], f1 = scale * f[j + 1];
    const double slope = (f1 - f0) / (tj1 - tj);
    return slope * (x - tj) + f0;
}

__global__ void __launch_bounds__(kAsdThreads) synth_asd_kernel(const AsdParams P) {
    __shared__ double red[kAsdThreads / 32];
    __shared__ double carry_sh;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long j = blockIdx.x; j < P.ntrials; j += gridDim.x) {
        AsdTrial tr = P.trials[j];
        const unsigned long long key = static_cast<unsigned long long>(tr.seed);
        const bool ext = tr.noise_row >= 0.0;
        const long long erow = ext ? static_cast<long long>(tr.noise_row) * P.N : 0;
        const double* e_fn = ext && P.ext_fn ? P.ext_fn + erow : nullptr;
        const double* e_amp = ext && P.ext_amp ? P.ext_amp + erow : nullptr;
        const double* e_df = ext && P.ext_df ? P.ext_df + erow : nullptr;
        const double* e_arm = ext && P.ext_arm && tr.dynamic != 0.0 ? P.ext_arm + erow : nullptr;
        if (ext) tr.sigma_amp = tr.sigma_df = 0.0;  // external series replace the internal draws, missing ones are zero
        double* phi = P.scratch + j * P.N;
        // ---- normalisation of the waveform: max |g| over the trial (physics.py:668-669) ----
        double gmax = 0.0;
        for (long long i = threadIdx.x; i < P.N; i += kAsdThreads) gmax = fmax(gmax, fabs(asd_waveform(tr, P, i)));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
        __syncthreads();
        if (lane == 0) red[warp] = gmax;
        __syncthreads();
        gmax = 0.0;
#pragma unroll
        for (int w = 0; w < kAsdThreads / 32; ++w) gmax = fmax(gmax, red[w]);
        const double ginv = gmax != 0.0 ? 1.0 / gmax : 0.0;
        // ---- running sum of (df + n_df) g: tiles of 1024 samples, four consecutive samples per thread ----
        if (threadIdx.x == 0) carry_sh = 0.0;
        __syncthreads();
        for (long long base = 0; base < P.N; base += kAsdThreads * kAsdPerThread) {
            const long long i0 = base + static_cast<long long>(threadIdx.x) * kAsdPerThread;
            double v[kAsdPerThread];
            float z[4];
            if (tr.sigma_df != 0.0) normal4(key, static_cast<unsigned long long>(i0 >> 2), 0xdf00u, z);
            double run = 0.0;
#pragma unroll
            for (int e = 0; e < kAsdPerThread; ++e) {
                const long long i = i0 + e;
                double inc = 0.0;
                if (i < P.N) {
                    const double g = asd_waveform(tr, P, i) * ginv;
                    const double dfn = e_df ? tr.df + e_df[i]
                                            : (tr.sigma_df != 0.0 ? tr.df + tr.sigma_df * static_cast<double>(z[e]) : tr.df);
                    inc = dfn * g;
                }
                run += inc;
                v[e] = run;
            }
            // exclusive scan of the thread totals over the CTA
            double inc = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            if (lane == 31) red[warp] = inc;
            __syncthreads();
            double before = carry_sh + inc - run;
            for (int w = 0; w < warp; ++w) before += red[w];
#pragma unroll
            for (int e = 0; e < kAsdPerThread; ++e)
                if (i0 + e < P.N) phi[i0 + e] = before + v[e];
            __syncthreads();
            if (threadIdx.x == kAsdThreads - 1) carry_sh = before + run;
            __syncthreads();
        }
        // ---- delayed-time interpolation and the interferometer output ----
        const double scale = P.scale;
        for (long long i0 = static_cast<long long>(threadIdx.x) * 4; i0 < P.N; i0 += kAsdThreads * 4) {
            float z[4];
            if (tr.sigma_amp != 0.0) normal4(key, static_cast<unsigned long long>(i0 >> 2), 0xa300u, z);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const long long i = i0 + e;
                if (i >= P.N) break;
                const double t = static_cast<double>(i) / P.fs;
                double path;
                if (e_arm)  // physics.py:686-688: modulation + noise + static offset, in this order
                    path = (tr.arm_amp != 0.0 ? tr.arm_amp * sin(tr.arm_w * t + tr.arm_psi) : 0.0) + e_arm[i] + tr.path0;
                else
                    path = tr.arm_amp != 0.0 ? tr.arm_amp * sin(tr.arm_w * t + tr.arm_psi) + 0.0 + tr.path0 : tr.path0;
                const double tau_dl = path / kLightSpeed;
                const double pm_meas = asd_interp(phi, P.N, P.fs, scale, t - (tr.tau_m + tau_dl));
                const double pm_ref = asd_interp(phi, P.N, P.fs, scale, t - tr.tau_r);
                const double lag = (tr.tau_m + tau_dl) - tr.tau_r;
                const double geom = tr.omega0 * lag;
                // physics.py:703-705: the carrier term with the noisy optical frequency f0 + n_f(t)
                const double carrier = e_fn ? (2.0 * kPi * (tr.omega0 / (2.0 * kPi) + e_fn[i])) * lag : geom;
                const double phase = carrier + (pm_meas - pm_ref);
                const double a = e_amp ? tr.amp + e_amp[i]
                                       : (tr.sigma_amp != 0.0 ? tr.amp + tr.sigma_amp * static_cast<double>(z[e]) : tr.amp);
                P.y[j * P.ld + i] = a * (1.0 + tr.vis * cos(phase));
                if (P.truth) P.truth[j * P.ld + i] = geom;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Reduction of a batch of trial results to the statistics Experiment.run reports per grid point
// (experiments.py:432-446): nanmean, nanstd (population), nanmin, nanmax and the "worst" trial, the value farthest
// from the mean (or from center[point][col] when given).  values[point][trial][col] with a column stride; one CTA per
// (point, column).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kStatThreads = 256;

struct TrialStats {
    double mean, std, min, max, worst, count;
};

__global__ void __launch_bounds__(kStatThreads) trial_stats_kernel(const double* __restrict__ values, long long npoints,
                                                                  long long ntrials, int ncols, long long col_stride,
                                                                  const double* __restrict__ center,
                                                                  TrialStats* __restrict__ out) {
    __shared__ double sh[4][kStatThreads / 32];
    __shared__ double mean_sh;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long point = blockIdx.x / ncols;
    const int col = static_cast<int>(blockIdx.x % ncols);
    const double* v = values + point * ntrials * col_stride + col;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    double s = 0.0, cnt = 0.0, mn = __longlong_as_double(0x7ff0000000000000ll), mx = -mn;
    for (long long t = threadIdx.x; t < ntrials; t += kStatThreads) {
        const double x = v[t * col_stride];
        if (x == x) {
            s += x;
            cnt += 1.0;
            mn = fmin(mn, x);
            mx = fmax(mx, x);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) {
        sh[0][warp] = s;
        sh[1][warp] = cnt;
        sh[2][warp] = mn;
        sh[3][warp] = mx;
    }
    __syncthreads();
    s = cnt = 0.0;
    for (int w = 0; w < kStatThreads / 32; ++w) {
        s += sh[0][w];
        cnt += sh[1][w];
        mn = fmin(mn, sh[2][w]);
        mx = fmax(mx, sh[3][w]);
    }
    const double mean = cnt > 0.0 ? s / cnt : nan;
    if (threadIdx.x == 0) mean_sh = mean;
    __syncthreads();
    // second pass: centred sum of squares, and the trial farthest from the mean -- or from a given centre, e.g. the
    // true value -- (first one on ties, as argmax)
    const double ref = center ? center[blockIdx.x] : mean_sh;
    double q = 0.0, far = -1.0;
    long long far_t = 0x7fffffffffffffffll;
    for (long long t = threadIdx.x; t < ntrials; t += kStatThreads) {
        const double x = v[t * col_stride];
        if (x == x) {
            const double d = x - mean_sh;
            q = fma(d, d, q);
            const double ad = fabs(x - ref);
            if (ad > far) {
                far = ad;
                far_t = t;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        q += __shfl_xor_sync(0xffffffffu, q, o);
        const double of = __shfl_xor_sync(0xffffffffu, far, o);
        const long long ot = __shfl_xor_sync(0xffffffffu, far_t, o);
        if (of > far || (of == far && ot < far_t)) {
            far = of;
            far_t = ot;
        }
    }
    __shared__ long long far_sh[kStatThreads / 32];
    __syncthreads();
    if (lane == 0) {
        sh[0][warp] = q;
        sh[1][warp] = far;
        far_sh[warp] = far_t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        q = 0.0;
        far = -1.0;
        far_t = 0x7fffffffffffffffll;
        for (int w = 0; w < kStatThreads / 32; ++w) {
            q += sh[0][w];
            if (sh[1][w] > far || (sh[1][w] == far && far_sh[w] < far_t)) {
                far = sh[1][w];
                far_t = far_sh[w];
            }
        }
        TrialStats r;
        r.mean = mean;
        r.std = cnt > 0.0 ? sqrt(q / cnt) : nan;
        r.min = cnt > 0.0 ? mn : nan;
        r.max = cnt > 0.0 ? mx : nan;
        r.worst = cnt > 0.0 ? v[far_t * col_stride] : nan;
        r.count = cnt;
        out[blockIdx.x] = r;
    }
}

}  // namespace dfk
