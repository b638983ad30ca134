// Host build of the __host__ __device__ numerical cores, for CPU-side unit tests only.
// (The shipped library runs these cores on the GPU; this file lets `pytest -m "not gpu"`
// check the same source against the oracle where no GPU exists.)
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../deepfmkit_b200/csrc/dfk_bessel.cuh"
#include "../../deepfmkit_b200/csrc/dfk_ekf_core.cuh"
#include "../../deepfmkit_b200/csrc/dfk_lm_core.cuh"
#include "../../deepfmkit_b200/csrc/dfk_demod_plan.h"

using namespace dfk;

extern "C" {

int hh_bessel(double x, int nmax, double* out) { return bessel_j_upto(x, nmax, out, 1); }

static LmOpts make_opts(const double* o) {
    LmOpts r;
    r.max_steps = static_cast<int>(o[0]);
    r.conv_improve = o[1];
    r.conv_param = o[2];
    r.fitok_threshold = o[3];
    r.grid_min = o[4];
    r.grid_max = o[5];
    r.grid_step = o[6];
    r.bessel_thr = o[7];
    r.sincos_thr = o[8];
    return r;
}

void hh_eval_state(int N, const double* qi, const double* p, double* out /*1+16+4*/) {
    std::vector<double> bes(N + 2);
    bessel_j_upto(p[1], N + 1, bes.data(), 1);
    NormalEq ne;
    eval_state<1>(N, qi, 1, bes.data(), 1, p, ne);
    out[0] = ne.ssq;
    // the psi row/column is structurally zero off the diagonal (the reference holds rounding noise there)
    const double full[16] = {ne.a00, ne.a01, ne.a02, 0.0, ne.a01, ne.a11, ne.a12, 0.0,
                             ne.a02, ne.a12, ne.a22, 0.0, 0.0, 0.0, 0.0, ne.a33};
    std::memcpy(out + 1, full, sizeof(full));
    out[17] = ne.g0; out[18] = ne.g1; out[19] = ne.g2; out[20] = ne.g3;
}

double hh_eval_ssq(int N, const double* qi, const double* p) {
    std::vector<double> bes(N + 2);
    bessel_j_upto(p[1], N + 1, bes.data(), 1);
    return eval_ssq<1>(N, qi, 1, bes.data(), 1, p);
}

int hh_solve(const double* jtj16, const double* g, double lam, double* dp) {
    NormalEq ne;
    ne.ssq = 0;
    ne.a00 = jtj16[0]; ne.a01 = jtj16[1]; ne.a02 = jtj16[2];
    ne.a11 = jtj16[5]; ne.a12 = jtj16[6];
    ne.a22 = jtj16[10]; ne.a33 = jtj16[15];
    ne.g0 = g[0]; ne.g1 = g[1]; ne.g2 = g[2]; ne.g3 = g[3];
    return damped_solve(ne, lam, dp) ? 1 : 0;
}

void hh_grid_seed(int N, const double* qi, const double* opts, double* seed) {
    std::vector<double> bes(N + 2);
    LmCounts cnt = {};
    grid_seed<1>(N, qi, 1, bes.data(), 1, make_opts(opts), seed, cnt);
}

// Full fit; out = [status, amp, m, phi, psi, ssq, steps, n_state, n_ssq, n_solve, n_grid]
void hh_fit(int N, const double* qi, const double* p0, const double* opts, double* out) {
    std::vector<double> bes(N + 2);
    LmCounts cnt = {};
    double p[4] = {p0[0], p0[1], p0[2], p0[3]};
    double ssq;
    int steps;
    const int status = fit_full<1>(N, qi, 1, bes.data(), 1, make_opts(opts), p, ssq, steps, cnt);
    out[0] = status;
    out[1] = p[0]; out[2] = p[1]; out[3] = p[2]; out[4] = p[3];
    out[5] = ssq;
    out[6] = steps;
    out[7] = static_cast<double>(cnt.n_state);
    out[8] = static_cast<double>(cnt.n_ssq);
    out[9] = static_cast<double>(cnt.n_solve);
    out[10] = static_cast<double>(cnt.n_grid);
}

// EKF over one channel; rows[nbuf][5].
void hh_ekf(const double* z, int64_t T, int64_t R, double f_samp, double f_mod, const double* x0,
            const double* p0_diag, const double* q_diag, double r_val, double* rows) {
    EkfState s;
    for (int i = 0; i < 5; ++i) s.x[i] = x0[i];
    for (int i = 0; i < 15; ++i) s.P[i] = 0.0;
    for (int i = 0; i < 5; ++i) s.P[tri(i, i)] = p0_diag[i];
    for (int i = 0; i < 5; ++i) s.kp[i] = s.hp[i] = 0.0;
    EkfConsts c;
    c.w_m = 2 * kPi * f_mod;
    c.f_samp = f_samp;
    c.inv_fs = 1.0 / f_samp;
    for (int i = 0; i < 5; ++i) c.q[i] = q_diag[i];
    c.r = r_val;
    const int64_t nbuf = T / R;
    for (int64_t k = 0; k < T; ++k) {
        ekf_step<false>(s, z[k], carrier_angle(static_cast<double>(k), c), c);
        if ((k + 1) % R == 0) {
            const int64_t idx = (k + 1) / R - 1;
            if (idx < nbuf) std::memcpy(rows + idx * 5, s.x, 5 * sizeof(double));
        }
    }
}

// sincos_cw on n arguments: out[2i] = sin, out[2i+1] = cos.
void hh_sincos_cw(const double* x, int64_t n, double* out) {
    for (int64_t i = 0; i < n; ++i) sincos_cw(x[i], out + 2 * i, out + 2 * i + 1);
}

// Number of sample indices k in [k0, k1) for which sample_time(k) differs from the IEEE quotient k / f_samp.
int64_t hh_sample_time_mismatches(double f_samp, int64_t k0, int64_t k1) {
    EkfConsts c;
    c.f_samp = f_samp;
    c.inv_fs = 1.0 / f_samp;
    int64_t bad = 0;
    for (int64_t k = k0; k < k1; ++k) {
        volatile double ref = static_cast<double>(k) / f_samp;
        if (sample_time(static_cast<double>(k), c) != ref) ++bad;
    }
    return bad;
}

// Demod plan (host logic shared with the CUDA launcher) and a scalar emulation of the folded
// kernel's arithmetic: fold over periods, rotation recurrence from the unit-circle table, drift term.
int hh_demod_plan_mul(int64_t R, double w0, int N) { return make_demod_plan(R, w0, N).kmul; }

int hh_demod_plan(int64_t R, double w0, int N, int64_t* P, int* use_drift, double* delta /*N*/) {
    DemodPlan pl = make_demod_plan(R, w0, N);
    *P = pl.P;
    *use_drift = pl.drift ? 1 : 0;
    for (int k = 0; k < N; ++k) delta[k] = pl.delta[k];
    return pl.folded ? 1 : 0;
}

void hh_demod_fold_emulate(const double* x, int64_t R, int N, double w0, int force_drift, double* qi, double* dc) {
    DemodPlan pl = make_demod_plan(R, w0, N);
    const int64_t P = pl.P, n = R / P;
    const bool drift = force_drift < 0 ? pl.drift : (force_drift != 0);
    std::vector<double> S(P, 0.0), U(P, 0.0), wc(P), ws(P);
    for (int64_t j = 0; j < P; ++j) {
        double s = 0, t = 0;
        for (int64_t c = 0; c < n; ++c) {
            s += x[c * P + j];
            t = std::fma(static_cast<double>(c), x[c * P + j], t);
        }
        S[j] = s;
        U[j] = static_cast<double>(j) * s + static_cast<double>(P) * t;
        unit_circle(j, P, &wc[j], &ws[j]);
    }
    double sum = 0;
    for (int64_t j = 0; j < P; ++j) sum += S[j];
    *dc = sum / static_cast<double>(R);
    for (int k = 0; k < N; ++k) qi[k] = qi[N + k] = 0.0;
    const int KB = 8;
    for (int k0 = 0; k0 < N; k0 += KB) {
        std::vector<double> cs(KB, 0.0), ss(KB, 0.0), cu(KB, 0.0), su(KB, 0.0);
        for (int64_t j = 0; j < P; ++j) {
            const int64_t r0 = (static_cast<int64_t>(k0 + 1) * pl.kmul * j) % P;  // block start from the table
            const int64_t rs = (static_cast<int64_t>(pl.kmul) * j) % P;           // one harmonic further
            double c = wc[r0], s = ws[r0];
            for (int kk = 0; kk < KB; ++kk) {
                cs[kk] += S[j] * c; ss[kk] += S[j] * s;
                cu[kk] += U[j] * c; su[kk] += U[j] * s;
                const double cn = c * wc[rs] - s * ws[rs];
                s = s * wc[rs] + c * ws[rs];
                c = cn;
            }
        }
        for (int kk = 0; kk < KB && k0 + kk < N; ++kk) {
            const int k = k0 + kk;
            const double d = drift ? pl.delta[k] : 0.0;
            qi[k] = (cs[kk] - d * su[kk]) / static_cast<double>(R);
            qi[N + k] = (ss[kk] + d * cu[kk]) / static_cast<double>(R);
        }
    }
}

}  // extern "C"
