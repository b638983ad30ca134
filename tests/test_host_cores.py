"""CPU-side checks of the __host__ __device__ numerical cores (same source the CUDA kernels compile),
built with g++ into a scratch shared library and compared with the oracle / golden fixtures."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import dfmi_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
c_double_p = ctypes.POINTER(ctypes.c_double)


def _ptr(a):
    return a.ctypes.data_as(c_double_p)


@pytest.fixture(scope="session")
def hh():
    out = os.path.join(tempfile.mkdtemp(prefix="dfk_hh_"), "libhh.so")
    src = os.path.join(ROOT, "tests", "host", "host_harness.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-x", "c++", "-fPIC", "-shared", "-ffp-contract=off", "-mfma",
                           "-o", out, src])
    lib = ctypes.CDLL(out)
    lib.hh_eval_ssq.restype = ctypes.c_double
    return lib


OPTS = np.array([100, 1e-9, 1e-9, 1e-3, 5.0, 30.0, 0.5, 0.05, 0.1], dtype=float)


def hh_fit(hh, nh, qi, p0, opts=OPTS):
    out = np.zeros(11)
    hh.hh_fit(ctypes.c_int(nh), _ptr(np.ascontiguousarray(qi)), _ptr(np.asarray(p0, dtype=float)), _ptr(opts), _ptr(out))
    return out


def test_bessel_vs_scipy(hh, golden):
    g = golden("bessel_jv")
    worst = worst_fwd = 0.0
    for x, ref in zip(g["xs"], g["jv"]):
        out = np.zeros(66)
        hh.hh_bessel(ctypes.c_double(x), ctypes.c_int(65), _ptr(out))
        err = np.max(np.abs(out - ref))
        if abs(x) <= 200.0:
            worst = max(worst, err)
        else:
            worst_fwd = max(worst_fwd, err)
    # scipy's own error against a long-double recurrence is 1.5e-15 on this grid
    assert worst < 2.5e-15, worst
    assert worst_fwd < 1e-14, worst_fwd  # |x| > 200: upward recurrence from j0/j1


def test_bessel_small_orders_and_tiny_args(hh):
    from scipy.special import jv
    for x in (1e-12, 1e-6, 1e-3, 0.07, 199.9, 200.1, 1234.5, -1234.5):
        for nmax in (1, 11, 65):
            out = np.zeros(nmax + 1)
            hh.hh_bessel(ctypes.c_double(x), ctypes.c_int(nmax), _ptr(out))
            assert np.max(np.abs(out - jv(np.arange(nmax + 1), x))) < (3e-15 if abs(x) <= 200 else 2e-14), (x, nmax)
    out = np.zeros(5)
    hh.hh_bessel(ctypes.c_double(float("nan")), ctypes.c_int(4), _ptr(out))
    assert np.all(np.isnan(out))


def test_model_state_and_solve(hh, golden):
    g = golden("solver_units")
    nh = int(g["nh"])
    for c in range(len(g["params"])):
        out = np.zeros(21)
        hh.hh_eval_state(ctypes.c_int(nh), _ptr(g["data"][c].copy()), _ptr(g["params"][c].copy()), _ptr(out))
        scale = np.abs(g["jtj"][c]).max()
        assert abs(out[0] - g["ssq"][c]) <= 1e-13 * g["ssq"][c]
        assert np.max(np.abs(out[1:17] - g["jtj"][c])) <= 1e-13 * scale
        assert np.max(np.abs(out[17:21] - g["grad"][c])) <= 1e-13 * np.abs(g["grad"][c]).max()
        s = hh.hh_eval_ssq(ctypes.c_int(nh), _ptr(g["data"][c].copy()), _ptr(g["params"][c].copy()))
        assert abs(s - g["ssq_only"][c]) <= 1e-13 * g["ssq_only"][c]
        for li, lam in enumerate(orc.LAMBDA_LADDER):
            dp = np.zeros(4)
            ok = hh.hh_solve(_ptr(g["jtj"][c].copy()), _ptr(g["grad"][c].copy()), ctypes.c_double(lam), _ptr(dp))
            assert ok == 1
            ref = g["steps"][c, li]
            assert np.max(np.abs(dp - ref)) <= 1e-9 * np.abs(ref).max()


def test_solve_singular_gives_zero_step(hh):
    jtj = np.zeros(16)
    jtj[5] = jtj[10] = jtj[15] = 1.0  # amplitude row/column exactly zero (a == 0)
    dp = np.ones(4)
    ok = hh.hh_solve(_ptr(jtj), _ptr(np.ones(4)), ctypes.c_double(0.1), _ptr(dp))
    assert ok == 0 and np.all(dp == 0)
    assert np.all(orc.damped_step(0.1, jtj.reshape(4, 4), np.ones(4)) == 0)


def test_grid_seed(hh, golden):
    g = golden("solver_units")
    nh = int(g["nh"])
    for key_qi, key_seed in (("data", "seeds"), ("clean_qi", "clean_seed")):
        for c in range(len(g["params"])):
            seed = np.zeros(4)
            hh.hh_grid_seed(ctypes.c_int(nh), _ptr(g[key_qi][c].copy()), _ptr(OPTS), _ptr(seed))
            ref = g[key_seed][c]
            assert seed[1] == ref[1] and seed[3] == ref[3]
            assert abs(seed[0] - ref[0]) <= 1e-11 * max(1.0, abs(ref[0]))
            assert abs(seed[2] - ref[2]) <= 1e-12


def test_full_fit_on_clean_vectors(hh, golden):
    g = golden("solver_units")
    nh = int(g["nh"])
    n_checked = 0
    for c in range(len(g["params"])):
        out = hh_fit(hh, nh, g["clean_qi"][c], [1.6, 6.0, 0.0, 0.0])
        assert out[0] == g["clean_status"][c]
        if g["clean_status"][c] < 2:
            ref = g["clean_p"][c]
            assert abs(out[1] - ref[0]) <= 1e-8 * abs(ref[0])
            assert abs(out[2] - ref[1]) <= 1e-8 * abs(ref[1])
            assert abs(out[3] - ref[2]) <= 1e-8 and abs(out[4] - ref[3]) <= 1e-8
            n_checked += 1
    assert n_checked >= 5


def test_full_fit_flags_on_noise_vectors(hh, golden):
    g = golden("solver_units")
    nh = int(g["nh"])
    for c in range(len(g["params"])):
        out = hh_fit(hh, nh, g["data"][c], [1.6, 6.0, 0.0, 0.0])
        assert out[0] == g["fit_status"][c] == 2  # pure noise never fits


@pytest.mark.parametrize("name", ["cfg1_quickstart", "cfg2_1mhz", "cfg3_channel", "deep_mod_n62",
                                  "fallback_m16", "low_snr"])
def test_fit_rows_from_golden_qi(hh, golden, name):
    """GPU schedule (every buffer seeded from buffer 0) through the host-compiled LM core."""
    g = golden(name)
    meta = g["meta"]
    nh = int(meta[9])
    ref = g["rows_seq"]
    first = hh_fit(hh, nh, g["qi"][0], [meta[11], meta[12], 0.0, meta[13]])
    seed = first[1:5].copy()
    for b in range(len(ref)):
        out = first if b == 0 else hh_fit(hh, nh, g["qi"][b], seed)
        assert out[0] == ref[b, 6]
        assert abs(out[1] - ref[b, 0]) <= 1e-8 * abs(ref[b, 0])
        assert abs(out[2] - ref[b, 1]) <= 1e-8 * abs(ref[b, 1])
        assert abs(out[3] - ref[b, 2]) <= 1e-8
        assert abs(out[4] - ref[b, 3]) <= 1e-8
        assert abs(out[5] - ref[b, 5]) <= 1e-9 * max(ref[b, 5], 1e-12)


def test_pathological_flags(hh, golden):
    g = golden("pathological_m3")
    meta = g["meta"]
    ref = g["rows_seq"]
    out = hh_fit(hh, int(meta[9]), g["qi"][0], [1.6, 6.0, 0.0, 0.0])
    assert out[0] == ref[0, 6]


def test_cfg5_rows(hh, golden):
    g = golden("cfg5_crlb")
    for a, m in enumerate(g["ms"]):
        for t in range(int(g["trials"])):
            out = hh_fit(hh, 15, g["qi"][a, t], [1.6, float(m), 0.0, 0.0])
            ref = g["rows"][a, t]
            assert out[0] == ref[6]
            assert abs(out[1] - ref[0]) <= 1e-8 * abs(ref[0])
            assert abs(out[2] - ref[1]) <= 1e-8 * abs(ref[1])
            assert abs(out[3] - ref[2]) <= 1e-8 and abs(out[4] - ref[3]) <= 1e-8


@pytest.mark.parametrize("name", ["ekf_default", "ekf_offset"])
def test_ekf_core(hh, golden, name):
    from tests.test_oracle_golden import ekf_kwargs
    g = golden(name)
    kw = ekf_kwargs(g)
    z = g["x"]
    x0 = np.array([kw.get("init_a", 1.6), kw.get("init_m", 6.0), kw.get("init_phi", 0.0), kw.get("init_psi", 0.0),
                   np.mean(z)])
    p0 = np.asarray(kw.get("p0_diag", orc.EKF_P0_DIAG), dtype=float)
    q = np.asarray(kw.get("q_diag", orc.EKF_Q_DIAG), dtype=float)
    r = float(kw.get("r_val", np.var(z)))
    nbuf = len(z) // 4000
    rows = np.zeros((nbuf, 5))
    hh.hh_ekf(_ptr(z.copy()), ctypes.c_int64(len(z)), ctypes.c_int64(4000), ctypes.c_double(200e3),
              ctypes.c_double(1000.0), _ptr(x0), _ptr(p0), _ptr(q), ctypes.c_double(r), _ptr(rows))
    ref = g["rows"][:, :5]
    assert np.max(np.abs(rows - ref)) < 1e-13, np.max(np.abs(rows - ref))


def _run_ekf(hh, z, R=4000):
    x0 = np.array([1.6, 6.0, 0.0, 0.0, np.mean(z)])
    rows = np.zeros((len(z) // R, 5))
    hh.hh_ekf(_ptr(z.copy()), ctypes.c_int64(len(z)), ctypes.c_int64(R), ctypes.c_double(200e3), ctypes.c_double(1000.0),
              _ptr(x0), _ptr(np.asarray(orc.EKF_P0_DIAG, dtype=float)), _ptr(np.asarray(orc.EKF_Q_DIAG, dtype=float)),
              ctypes.c_double(float(np.var(z))), _ptr(rows))
    return rows


def test_ekf_core_200k_and_4e6_steps(hh, golden):
    """The kernel's step function (symmetric covariance, factored P H^T, own sincos) against EKFFitter.fit over the
    full-size fixtures: 8 channels x 200 000 steps and one channel x 4e6 steps (t up to 20 s).  The deviation is
    rounding noise that the filter contracts: it stays below 4e-13 at 4e6 steps."""
    g = golden("full_ekf_8ch")
    worst = 0.0
    for i, c in enumerate(g["channels"]):
        z = orc.snr_signal(6.0, 200e3, 1000.0, 1.0, 40.0, seed=int(c), phi0=2 * np.pi * c / 4096)
        worst = max(worst, np.max(np.abs(_run_ekf(hh, z) - g["rows"][i][:, :5])))
    assert worst < 1e-13, worst
    g = golden("full_ekf_deep")
    z = orc.snr_signal(6.0, 200e3, 1000.0, 20.0, 40.0, seed=777, phi0=2 * np.pi * 777 / 4096)
    dev = np.abs(_run_ekf(hh, z) - g["rows"][:, :5])
    assert dev[:250].max() < 5e-13 and dev.max() < 2e-12, (dev[:250].max(), dev.max())


def test_sincos_cw_accuracy(hh):
    rng = np.random.RandomState(5)
    xs = np.concatenate([rng.uniform(-40, 40, 100000), rng.uniform(-1e6, 1e6, 100000), np.linspace(-1, 1, 1001),
                         6.2e5 + rng.uniform(0, 10, 50000), [0.0, 1e-300, -1e-9, 1e6, -2e6, 3e9, np.inf, np.nan]])
    out = np.zeros(2 * len(xs))
    hh.hh_sincos_cw(_ptr(xs), ctypes.c_int64(len(xs)), _ptr(out))
    fin = np.isfinite(xs)
    xl = xs[fin].astype(np.longdouble)
    assert float(np.max(np.abs(out[0::2][fin] - np.sin(xl)))) < 2.5e-16
    assert float(np.max(np.abs(out[1::2][fin] - np.cos(xl)))) < 2.5e-16
    assert np.all(np.isnan(out[0::2][~fin])) and np.all(np.isnan(out[1::2][~fin]))


@pytest.mark.parametrize("f_samp", [200e3, 1e6, 48e3, 162.5e3, 30e3, 44.1e3, 1e6 / 3, 123456.789])
def test_sample_time_equals_ieee_quotient(hh, f_samp):
    """t_k = k / f_samp by the Markstein-corrected product is bit-equal to the division numpy performs."""
    hh.hh_sample_time_mismatches.restype = ctypes.c_int64
    assert hh.hh_sample_time_mismatches(ctypes.c_double(f_samp), ctypes.c_int64(0), ctypes.c_int64(3_000_000)) == 0
    hi = 2 ** 40
    assert hh.hh_sample_time_mismatches(ctypes.c_double(f_samp), ctypes.c_int64(hi), ctypes.c_int64(hi + 500_000)) == 0


@pytest.mark.parametrize("name,plan_drift,tol_plan", [("cfg1_quickstart", 0, 3e-13), ("cfg2_1mhz", 0, 3e-13),
                                                      ("deep_mod_n62", 1, 2e-13)])
def test_fold_demod_arithmetic(hh, golden, name, plan_drift, tol_plan):
    """Folding + rotation recurrence (+ drift term) reproduce the reference lock-in well within 1e-12 of max|IQ|:
    at the plan's own choice of the drift term, and at the reference's rounding floor with the term forced on."""
    from tests.test_oracle_golden import _signal_from_meta
    g = golden(name)
    meta = g["meta"]
    x = g["x"] if "x" in g.files else _signal_from_meta(meta)
    f_samp, f_mod, n, nh = meta[1], meta[2], int(meta[8]), int(meta[9])
    R = int(f_samp / f_mod * n)
    w0 = orc.rad_per_sample(f_samp, f_mod)
    P = ctypes.c_int64()
    drift = ctypes.c_int()
    delta = np.zeros(nh)
    assert hh.hh_demod_plan(ctypes.c_int64(R), ctypes.c_double(w0), ctypes.c_int(nh), ctypes.byref(P),
                            ctypes.byref(drift), _ptr(delta)) == 1
    assert P.value == int(f_samp / f_mod) and drift.value == plan_drift
    for force, tol in ((-1, tol_plan), (1, 2e-13 if nh > 20 else 1e-13)):
        for b in range(min(3, len(g["qi"]))):
            qi = np.zeros(2 * nh)
            dc = ctypes.c_double()
            buf = np.ascontiguousarray(x[b * R:(b + 1) * R])
            hh.hh_demod_fold_emulate(_ptr(buf), ctypes.c_int64(R), ctypes.c_int(nh), ctypes.c_double(w0),
                                     ctypes.c_int(force), _ptr(qi), ctypes.byref(dc))
            ref = g["qi"][b]
            assert np.max(np.abs(qi - ref)) <= tol * np.abs(ref).max()
            assert abs(dc.value - g["rows_seq"][b, 4]) <= 1e-14 * abs(g["rows_seq"][b, 4])


def test_demod_plan_fold_lengths(hh):
    """Even period: one period per fold.  Odd or rational period: the smallest even whole number of samples that
    holds whole periods (the reference's own record, 30 kHz / 400 Hz = 75 samples, folds at 150).  Otherwise, or if
    the buffer is not a whole number of folds: no fold (direct kernel)."""
    P = ctypes.c_int64()
    drift = ctypes.c_int()
    delta = np.zeros(10)

    def plan(R, f_samp, f_mod):
        w0 = orc.rad_per_sample(f_samp, f_mod)
        ok = hh.hh_demod_plan(ctypes.c_int64(R), ctypes.c_double(w0), ctypes.c_int(10), ctypes.byref(P),
                              ctypes.byref(drift), _ptr(delta))
        return ok, P.value, hh.hh_demod_plan_mul(ctypes.c_int64(R), ctypes.c_double(w0), ctypes.c_int(10))

    assert plan(4000, 200e3, 1000.0) == (1, 200, 1)
    assert plan(1500, 30e3, 400.0) == (1, 150, 2)        # odd period
    assert plan(4020, 201e3, 1000.0) == (1, 402, 2)
    assert plan(3250, 162.5e3, 1000.0) == (1, 650, 4)    # 162.5 samples per period
    assert plan(3240, 200e3, 1234.5)[0] == 0             # incommensurate
    assert plan(4100, 200e3, 1000.0)[0] == 0             # buffer is not a whole number of periods
    assert plan(75, 30e3, 400.0)[0] == 0                 # one odd period per buffer: no even fold fits
    assert plan(80000, 4e6, 1000.0) == (1, 4000, 1)      # beyond the register kernel: the column-chunked fold
    assert plan(200000, 10e6, 1000.0) == (1, 10000, 1)   # 10 MHz / 1 kHz


@pytest.mark.parametrize("f_samp,f_mod,n,nh", [(30e3, 400.0, 20, 10), (162.5e3, 1000.0, 20, 8)])
def test_fold_emulation_odd_and_rational_periods(hh, f_samp, f_mod, n, nh):
    x = orc.snr_signal(6.0, f_samp, f_mod, 0.2, 40.0, seed=9)
    R, _, nbuf = orc.buffer_geometry(len(x), f_samp, f_mod, n)
    w0 = orc.rad_per_sample(f_samp, f_mod)
    for b in range(2):
        buf = np.ascontiguousarray(x[b * R:(b + 1) * R])
        qi = np.zeros(2 * nh)
        dc = ctypes.c_double()
        hh.hh_demod_fold_emulate(_ptr(buf), ctypes.c_int64(R), ctypes.c_int(nh), ctypes.c_double(w0),
                                 ctypes.c_int(-1), _ptr(qi), ctypes.byref(dc))
        ref = orc.lockin_means(buf, w0, nh)
        assert np.max(np.abs(qi - ref)) <= 3e-13 * np.abs(ref).max()
