// Monte-Carlo experiment entries of the C ABI (include/dfk_b200.h, "batched experiments"): the 'asd'-mode signal
// generator of the reference's physics engine, batched over trials, and the per-grid-point reduction of trial results.
#include "dfk_host.h"

#include <algorithm>
#include <cmath>

#include "dfk_asd.cuh"

extern "C" {

int dfk_synth_asd_dev(dfk_ctx* ctx, const double* trials_dev, int64_t ntrials, int64_t N, double f_samp,
                      const double* tables_dev, int64_t ntables, double* y_dev, int64_t ld, double* truth_dev) {
    return dfk_synth_asd_noise_dev(ctx, trials_dev, ntrials, N, f_samp, tables_dev, ntables, nullptr, 0, y_dev, ld, truth_dev);
}

int dfk_synth_asd_noise_dev(dfk_ctx* ctx, const double* trials_dev, int64_t ntrials, int64_t N, double f_samp,
                            const double* tables_dev, int64_t ntables, const double* const* noise_dev, int64_t noise_rows,
                            double* y_dev, int64_t ld, double* truth_dev) {
    DFK_ENTER(ctx);
    if (noise_rows < 0 || (noise_rows > 0 && !noise_dev)) return fail(DFK_ERR_ARG, "bad external noise series");
    if (ntrials < 0 || N < 0 || ld < N) return fail(DFK_ERR_ARG, "bad geometry: ntrials=%lld N=%lld ld=%lld", (long long)ntrials, (long long)N, (long long)ld);
    if (!(f_samp > 0.0)) return fail(DFK_ERR_ARG, "f_samp must be positive");
    if (ntrials == 0 || N == 0) return DFK_OK;
    if (!trials_dev || !y_dev) return fail(DFK_ERR_ARG, "null pointer");
    if (ntables < 0 || (ntables > 0 && !tables_dev)) return fail(DFK_ERR_ARG, "bad waveform tables");
    static_assert(sizeof(dfk::AsdTrial) == DFK_ASD_TRIAL_DOUBLES * sizeof(double), "ABI: trial record size");
    int rc = ensure(ctx, ctx->post[0], sizeof(double) * static_cast<size_t>(ntrials) * static_cast<size_t>(N));
    if (rc) return rc;
    dfk::AsdParams P;
    P.trials = reinterpret_cast<const dfk::AsdTrial*>(trials_dev);
    P.ntrials = ntrials;
    P.N = N;
    P.fs = f_samp;
    const double dt = N > 1 ? 1.0 / f_samp - 0.0 / f_samp : 1.0 / f_samp;  // t[1] - t[0] with t = arange(N) / f_samp
    P.scale = 2.0 * dfk::kPi / (1.0 / dt);
    P.tables = tables_dev;
    P.scratch = static_cast<double*>(ctx->post[0].ptr);
    P.y = y_dev;
    P.ld = ld;
    P.truth = truth_dev;
    P.ext_fn = noise_rows ? noise_dev[DFK_NOISE_LASER_FREQUENCY] : nullptr;
    P.ext_amp = noise_rows ? noise_dev[DFK_NOISE_AMPLITUDE] : nullptr;
    P.ext_df = noise_rows ? noise_dev[DFK_NOISE_DF] : nullptr;
    P.ext_arm = noise_rows ? noise_dev[DFK_NOISE_ARMLENGTH] : nullptr;
    const int grid = static_cast<int>(std::min<int64_t>(ntrials, static_cast<int64_t>(ctx->sm_count) * 32));
    dfk::synth_asd_kernel<<<grid, dfk::kAsdThreads, 0, ctx->stream()>>>(P);
    ctx->launches++;
    DFK_CUDA(cudaGetLastError());
    return DFK_OK;
}

int dfk_trial_stats_dev(dfk_ctx* ctx, const double* values_dev, int64_t npoints, int64_t ntrials, int32_t ncols,
                        int64_t col_stride, const double* center_dev, double* out_dev) {
    DFK_ENTER(ctx);
    if (npoints < 0 || ntrials < 1 || ncols < 1 || col_stride < ncols)
        return fail(DFK_ERR_ARG, "bad geometry: npoints=%lld ntrials=%lld ncols=%d stride=%lld", (long long)npoints,
                    (long long)ntrials, ncols, (long long)col_stride);
    if (npoints == 0) return DFK_OK;
    if (!values_dev || !out_dev) return fail(DFK_ERR_ARG, "null pointer");
    if (ncols > dfk::kStatMaxCols) return fail(DFK_ERR_ARG, "at most %d result columns", dfk::kStatMaxCols);
    static_assert(sizeof(dfk::TrialStats) == DFK_TRIAL_STATS_DOUBLES * sizeof(double), "ABI: statistics record size");
    // slices per grid point: enough CTAs to fill the GPU a few times over, at least a CTA's worth of trials each
    int64_t nslices = (static_cast<int64_t>(ctx->sm_count) * 4 + npoints - 1) / npoints;
    nslices = std::max<int64_t>(1, std::min<int64_t>({nslices, static_cast<int64_t>(dfk::kStatMaxSlices),
                                                      (ntrials + dfk::kStatThreads - 1) / dfk::kStatThreads}));
    const size_t nparts = static_cast<size_t>(npoints) * nslices * ncols;
    int rc = ensure(ctx, ctx->stats_part, nparts * (sizeof(dfk::StatPart1) + sizeof(dfk::StatPart2)));
    if (rc) return rc;
    auto* part1 = static_cast<dfk::StatPart1*>(ctx->stats_part.ptr);
    auto* part2 = reinterpret_cast<dfk::StatPart2*>(part1 + nparts);
    if (npoints * nslices > 0x7fffffffll) return fail(DFK_ERR_ARG, "too many grid points for one launch");
    const unsigned grid = static_cast<unsigned>(npoints * nslices);
    cudaStream_t st = ctx->stream();
    dfk::trial_stats_pass1<<<grid, dfk::kStatThreads, 0, st>>>(values_dev, ntrials, ncols, col_stride,
                                                               static_cast<int>(nslices), part1);
    dfk::trial_stats_pass2<<<grid, dfk::kStatThreads, 0, st>>>(values_dev, ntrials, ncols, col_stride,
                                                               static_cast<int>(nslices), center_dev, part1, part2);
    const int64_t nout = npoints * ncols;
    dfk::trial_stats_finish<<<static_cast<unsigned>((nout + 127) / 128), 128, 0, st>>>(
        values_dev, npoints, ntrials, ncols, col_stride, static_cast<int>(nslices), part1, part2,
        reinterpret_cast<dfk::TrialStats*>(out_dev));
    ctx->launches += 3;
    DFK_CUDA(cudaGetLastError());
    return DFK_OK;
}

}  // extern "C"
