"""Parity of the CUDA path (through the C ABI) against the oracle and the golden fixtures.

Gates (BASELINE.json north_star): demodulated I/Q within 1e-12 of the buffer's max |I/Q| (fp64);
m and amp within 1e-8 relative, phi and psi within 1e-8 rad on rows the reference fits (fitok 0/1);
identical fitok flags everywhere.  EKF states within 1e-9 absolute of the reference loop.
"""
import os

import numpy as np
import pytest

from oracle import dfmi_oracle as orc
from tests.test_oracle_golden import _signal_from_meta, ekf_kwargs

pytestmark = pytest.mark.gpu

IQ_TOL = 1e-12
PARAM_TOL = 1e-8


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.fixture(scope="module")
def ctx(torch_mod):
    from deepfmkit_b200 import _lib
    c = _lib.Context(0)
    yield c
    c.close()


def _case(golden, name):
    g = golden(name)
    meta = g["meta"]
    x = g["x"] if "x" in g.files else _signal_from_meta(meta)
    f_samp, f_mod, n, nh = meta[1], meta[2], int(meta[8]), int(meta[9])
    R = int(f_samp / f_mod * n)
    return g, x, f_samp, f_mod, n, nh, R, orc.rad_per_sample(f_samp, f_mod)


def gpu_demod(torch, ctx, x, R, nh, w0):
    nbuf = len(x) // R
    xd = torch.from_numpy(np.ascontiguousarray(x[: nbuf * R])).cuda()
    qi = torch.empty((nbuf, 2 * nh), dtype=torch.float64, device="cuda")
    dc = torch.empty(nbuf, dtype=torch.float64, device="cuda")
    ctx.demod(xd.data_ptr(), nbuf, R, nh, w0, qi.data_ptr(), dc.data_ptr())
    ctx.synchronize()
    return qi.cpu().numpy(), dc.cpu().numpy()


def assert_rows_match(rows, ref, tol=PARAM_TOL):
    assert np.array_equal(rows[:, 6], ref[:, 6]), (rows[:, 6], ref[:, 6])
    ok = ref[:, 6] < 2
    assert np.all(np.abs(rows[ok, 0] - ref[ok, 0]) <= tol * np.abs(ref[ok, 0]))
    assert np.all(np.abs(rows[ok, 1] - ref[ok, 1]) <= tol * np.abs(ref[ok, 1]))
    assert np.all(np.abs(rows[ok, 2] - ref[ok, 2]) <= tol)
    assert np.all(np.abs(rows[ok, 3] - ref[ok, 3]) <= tol)
    assert np.all(np.abs(rows[ok, 5] - ref[ok, 5]) <= 1e-8 * np.maximum(ref[ok, 5], 1e-12))
    assert np.all(np.abs(rows[:, 4] - ref[:, 4]) <= 1e-14 * np.abs(ref[:, 4]))  # dc


# ---------------------------------------------------------------------------------------------------
def test_bessel_device_vs_scipy(torch_mod, ctx, golden):
    g = golden("bessel_jv")
    xs = torch_mod.from_numpy(g["xs"]).cuda()
    out = torch_mod.empty((len(g["xs"]), 66), dtype=torch_mod.float64, device="cuda")
    ctx.bessel_dev(xs.data_ptr(), len(g["xs"]), 65, out.data_ptr())
    ctx.synchronize()
    err = np.abs(out.cpu().numpy() - g["jv"]).max(axis=1)
    small = np.abs(g["xs"]) <= 200
    # |x| > 200 starts the upward recurrence from CUDA's j0/j1 (documented abs error ~1e-12 there)
    assert err[small].max() < 2.5e-15 and err[~small].max() < 2e-12


@pytest.mark.parametrize("name", ["cfg1_quickstart", "cfg2_1mhz", "cfg3_channel", "deep_mod_n62", "fallback_m16",
                                  "pathological_m3", "low_snr"])
def test_demod_folded_vs_reference(torch_mod, ctx, golden, name):
    from deepfmkit_b200 import _lib
    g, x, f_samp, f_mod, n, nh, R, w0 = _case(golden, name)
    assert _lib.demod_path(R, w0) == 1
    qi, dc = gpu_demod(torch_mod, ctx, x, R, nh, w0)
    ref = g["qi"]
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert np.max(np.abs(qi - ref) / scale) <= IQ_TOL, np.max(np.abs(qi - ref) / scale)
    assert np.all(np.abs(dc - g["rows_seq"][:, 4]) <= 1e-14 * np.abs(g["rows_seq"][:, 4]))


@pytest.mark.parametrize("f_samp,f_mod,n,nh", [(200e3, 1234.5, 20, 10), (201e3, 1000.0, 20, 10), (48e3, 440.0, 7, 13),
                                               (200e3, 1000.0, 1, 15), (30e3, 400.0, 20, 10), (162.5e3, 1000.0, 20, 10)])
def test_demod_vs_oracle_general(torch_mod, ctx, f_samp, f_mod, n, nh):
    """Incommensurate periods go through the direct kernel; odd (201, 75: the reference's own 30 kHz / 400 Hz record)
    and rational (162.5) periods fold over two or four periods; n = 1 (cfg 5) takes the single-period kernel."""
    x = orc.snr_signal(6.0, f_samp, f_mod, 0.05, 40.0, seed=3)
    R, _, nbuf = orc.buffer_geometry(len(x), f_samp, f_mod, n)
    w0 = orc.rad_per_sample(f_samp, f_mod)
    qi, dc = gpu_demod(torch_mod, ctx, x, R, nh, w0)
    for b in list(range(min(nbuf, 6))) + [nbuf - 1]:
        buf = x[b * R:(b + 1) * R]
        ref = orc.lockin_means(buf, w0, nh)
        assert np.max(np.abs(qi[b] - ref)) <= IQ_TOL * np.abs(ref).max()
        assert abs(dc[b] - buf.mean()) <= 1e-14 * abs(buf.mean())


@pytest.mark.parametrize("name", ["cfg1_quickstart", "fallback_m16"])
def test_tile_and_fold_kernels_agree(torch_mod, ctx, golden, name, monkeypatch):
    """Short periods take the barrier-free tile kernel; the DFK_NO_TILE override (an explicit library call, not an
    environment variable) routes the same data through the CTA-per-buffer fold kernel.  Both must sit at the reference's rounding floor."""
    g, x, f_samp, f_mod, n, nh, R, w0 = _case(golden, name)
    x = np.tile(x[: (len(x) // R) * R], 40)  # enough buffers for every warp of several CTAs, ragged last group
    x = x[: (len(x) // R - 3) * R]
    qi_tile, dc_tile = gpu_demod(torch_mod, ctx, x, R, nh, w0)
    from deepfmkit_b200 import _lib
    with _lib.dev_overrides(DFK_NO_TILE=1):
        qi_fold, dc_fold = gpu_demod(torch_mod, ctx, x, R, nh, w0)
    nb = len(g["qi"])
    ref = np.tile(g["qi"], (40, 1))[: len(qi_tile)]
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert np.max(np.abs(qi_tile - ref) / scale) <= IQ_TOL
    assert np.max(np.abs(qi_fold - ref) / scale) <= IQ_TOL
    assert np.max(np.abs(qi_tile - qi_fold) / scale) <= 3e-13
    assert np.max(np.abs(dc_tile - dc_fold)) <= 1e-14


@pytest.mark.parametrize("P,n,nh", [(2048, 3, 10), (1536, 2, 7), (1000, 1, 64), (6, 50, 2), (2, 40, 1), (256, 5, 31),
                                    (258, 4, 16), (200, 3, 40), (2050, 2, 5), (75, 20, 10),
                                    (100, 1, 10), (256, 1, 31), (4, 1, 1), (8, 1, 3), (200, 1, 40), (202, 1, 10),
                                    (12, 1, 5), (200, 1, 16), (75, 1, 10), (333, 6, 12), (1001, 2, 9),
                                    (4096, 3, 10), (10000, 2, 10), (16384, 2, 20), (2501, 2, 7), (10000, 20, 64),
                                    (6144, 1, 3), (2052, 5, 64)])
def test_demod_period_and_harmonic_extremes(torch_mod, ctx, P, n, nh):
    """Largest period of the register fold kernel (2048), smallest (2), the tile/fold boundary (256/258), N = 1 and
    N = 64, a table too big for the tile kernel (N = 40 at P = 200), an odd period in an odd number of periods (75 x 1:
    no even fold length -> direct kernel), one-period buffers (n = 1) through the quarter-wave kernel (P % 4 == 0) or
    the tile kernel (P = 202), and fold lengths beyond 2048 through the column-chunked kernel: 2050 and 2052 (a
    two- and four-column last chunk), 4096, 6144, 10 000 (10 MHz / 1 kHz), 16 384, 5002 (an odd period doubled)."""
    from deepfmkit_b200 import _lib
    f_mod = 1000.0
    f_samp = f_mod * P
    R = P * n
    w0 = orc.rad_per_sample(f_samp, f_mod)
    folds = P % 2 == 0 or n % 2 == 0
    assert _lib.demod_path(R, w0) == (1 if folds else 0)
    rng = np.random.RandomState(P + n)
    nbuf = 37
    t = np.arange(nbuf * R)
    x = 1.0 + np.cos(0.7 + 3.0 * np.cos(2 * np.pi * t / P + 0.2)) + 0.01 * rng.randn(nbuf * R)
    qi, dc = gpu_demod(torch_mod, ctx, x, R, nh, w0)
    for b in (0, 1, 17, nbuf - 1):
        buf = x[b * R:(b + 1) * R]
        ref = orc.lockin_means(buf, w0, nh)
        assert np.max(np.abs(qi[b] - ref)) <= IQ_TOL * max(np.abs(ref).max(), 1e-3), (P, n, nh, b)
        assert abs(dc[b] - buf.mean()) <= 1e-14 * abs(buf.mean())


def test_empty_and_degenerate_calls(torch_mod, ctx):
    from deepfmkit_b200 import _lib, nls_fit_batch
    torch = torch_mod
    w0 = orc.rad_per_sample(200e3, 1000.0)
    ctx.demod(0, 0, 4000, 10, w0, 0, 0)  # nothing to do, null pointers allowed
    ctx.nls_fit_dev(0, 0, 4000, 10, w0, [1.6, 6.0, 0, 0], True, None, 0)
    assert ctx.nls_fit_host(np.zeros(100), 4000, 10, w0, [1.6, 6.0, 0, 0]).shape == (0, 8)
    assert nls_fit_batch(np.zeros((3, 100)), 200e3, 1000.0, 20).shape == (3, 0, 8)
    x = torch.zeros(8000, dtype=torch.float64, device="cuda")
    rows = torch.zeros((2, 8), dtype=torch.float64, device="cuda")
    for bad in (dict(N=0), dict(N=65), dict(w0=0.0), dict(w0=float("nan")), dict(R=0)):
        kw = dict(N=10, w0=w0, R=4000)
        kw.update(bad)
        with pytest.raises(RuntimeError):
            ctx.nls_fit_dev(x.data_ptr(), 2, kw["R"], kw["N"], kw["w0"], [1.6, 6.0, 0, 0], True, None, rows.data_ptr())
    # an all-zero and a non-finite record are data, not errors; the flags follow the reference: the zero record
    # "fits" with the amplitude driven to nothing (fitok 0), the NaN buffer stays unfitted (fitok 2, ssq NaN)
    ctx.nls_fit_dev(x.data_ptr(), 2, 4000, 10, w0, [1.6, 6.0, 0, 0], True, None, rows.data_ptr())
    ctx.synchronize()
    r = rows.cpu().numpy()
    ref0 = orc.nls_fit(np.zeros(8000), 200e3, 1000.0, 20, 10)
    assert np.array_equal(r[:, 6], ref0[:, 6]) and np.all(r[:, 0] < 1e-100) and np.all(r[:, 5] == 0)
    x[100] = float("nan")
    ctx.nls_fit_dev(x.data_ptr(), 2, 4000, 10, w0, [1.6, 6.0, 0, 0], True, None, rows.data_ptr())
    ctx.synchronize()
    r = rows.cpu().numpy()
    xn = np.zeros(8000)
    xn[100] = np.nan
    with np.errstate(all="ignore"):
        refn = orc.nls_fit(xn, 200e3, 1000.0, 20, 10)
    assert np.array_equal(r[:, 6], refn[:, 6]) and r[0, 6] == 2 and np.isnan(r[0, 5])
    assert np.array_equal(r[0, :4], refn[0, :4])


def test_demod_unaligned_pointer_takes_direct_path(torch_mod, ctx, golden):
    g, x, f_samp, f_mod, n, nh, R, w0 = _case(golden, "cfg1_quickstart")
    nbuf = len(x) // R
    xd = torch_mod.zeros(len(x) + 1, dtype=torch_mod.float64, device="cuda")
    xd[1:] = torch_mod.from_numpy(x).cuda()
    qi = torch_mod.empty((nbuf, 2 * nh), dtype=torch_mod.float64, device="cuda")
    dc = torch_mod.empty(nbuf, dtype=torch_mod.float64, device="cuda")
    ctx.demod(xd.data_ptr() + 8, nbuf, R, nh, w0, qi.data_ptr(), dc.data_ptr())
    ctx.synchronize()
    ref = g["qi"]
    assert np.max(np.abs(qi.cpu().numpy() - ref) / np.abs(ref).max(axis=1, keepdims=True)) <= IQ_TOL


@pytest.mark.parametrize("lanes", [0, 1, 4, 32])
@pytest.mark.parametrize("name", ["cfg1_quickstart", "cfg2_1mhz", "cfg3_channel", "deep_mod_n62", "fallback_m16",
                                  "pathological_m3", "low_snr"])
def test_nls_rows_vs_reference(torch_mod, ctx, golden, name, lanes):
    from deepfmkit_b200 import _lib
    g, x, f_samp, f_mod, n, nh, R, w0 = _case(golden, name)
    meta = g["meta"]
    opts = _lib.default_lm_opts()
    opts.lanes_per_fit = lanes
    rows = ctx.nls_fit_host(x, R, nh, w0, [meta[11], meta[12], 0.0, meta[13]], seeded=True, opts=opts)
    assert_rows_match(rows, g["rows_seq"])
    if len(g["rows_par"]):
        assert_rows_match(rows, g["rows_par"])


def test_lm_fit_on_reference_harmonic_vectors(torch_mod, ctx, golden):
    """dfk_lm_fit alone, fed the reference's own I/Q means: isolates the solver from the demodulation."""
    g = golden("solver_units")
    nh = int(g["nh"])
    for key, st_key, p_key in (("clean_qi", "clean_status", "clean_p"), ("data", "fit_status", "fit_p")):
        qi = torch_mod.from_numpy(g[key]).cuda()
        nfit = qi.shape[0]
        guess = torch_mod.tensor([1.6, 6.0, 0.0, 0.0], dtype=torch_mod.float64, device="cuda")
        rows = torch_mod.zeros((nfit, 8), dtype=torch_mod.float64, device="cuda")
        ctx.lm_fit(qi.data_ptr(), nfit, nh, guess.data_ptr(), 0, None, None, rows.data_ptr())
        ctx.synchronize()
        rows = rows.cpu().numpy()
        assert np.array_equal(rows[:, 6], g[st_key])
        ok = g[st_key] < 2
        ref = g[p_key]
        assert np.all(np.abs(rows[ok, 0] - ref[ok, 0]) <= PARAM_TOL * np.abs(ref[ok, 0]))
        assert np.all(np.abs(rows[ok, 1] - ref[ok, 1]) <= PARAM_TOL * np.abs(ref[ok, 1]))
        assert np.all(np.abs(rows[ok, 2:4] - ref[ok, 2:4]) <= PARAM_TOL)


def test_cfg5_batch_per_channel_guess(torch_mod, golden):
    """CRLB sweep recipe: one period per realisation, N = 15, init_m = m_true, all in one launch."""
    from deepfmkit_b200 import nls_fit_batch
    g = golden("cfg5_crlb")
    ms, trials = g["ms"], int(g["trials"])
    recs, init_m = [], []
    for m in ms:
        for t in range(trials):
            recs.append(orc.snr_signal(float(m), 200e3, 1000, 1 / 1000, 40.0, seed=t))
            init_m.append(float(m))
    rows = nls_fit_batch(np.stack(recs), 200e3, 1000.0, 1, ndata=15, init_m=np.array(init_m), seeded=False)
    ref = g["rows"].reshape(-1, 7)
    assert_rows_match(rows[:, 0, :], ref)


def test_multichannel_batch_matches_per_channel_calls(torch_mod, ctx):
    from deepfmkit_b200 import nls_fit_batch
    chans = [orc.snr_signal(6.0, 200e3, 1000, 0.2, 40.0, seed=c, phi0=2 * np.pi * c / 5) for c in range(5)]
    rows = nls_fit_batch(np.stack(chans), 200e3, 1000.0, 20)
    w0 = orc.rad_per_sample(200e3, 1000.0)
    for c, x in enumerate(chans):
        single = ctx.nls_fit_host(x, 4000, 10, w0, [1.6, 6.0, 0.0, 0.0], seeded=True)
        assert np.array_equal(single, rows[c])
    ref = orc.nls_fit(chans[3], 200e3, 1000.0, 20, 10, schedule="gpu")
    assert_rows_match(rows[3], ref)


def test_host_entry_streams_slabs(torch_mod, ctx):
    """A record longer than one 128 MiB slab: rows must not depend on where the slab boundaries fall."""
    from deepfmkit_b200 import _lib
    R, nh = 4000, 10
    w0 = orc.rad_per_sample(200e3, 1000.0)
    nbuf = 9000  # 288 MB
    xd = torch_mod.empty(nbuf * R, dtype=torch_mod.float64, device="cuda")
    ctx.synth_snr_dev(xd.data_ptr(), nbuf * R, 1, 200e3, 1000.0, 6.0, snr_db=40.0, seed=5)
    ctx.synchronize()
    x = xd.cpu().numpy()
    rows = ctx.nls_fit_host(x, R, nh, w0, [1.6, 6.0, 0.0, 0.0], seeded=True)
    rows_dev = torch_mod.empty((nbuf, 8), dtype=torch_mod.float64, device="cuda")
    ctx.nls_fit_dev(xd.data_ptr(), nbuf, R, nh, w0, [1.6, 6.0, 0.0, 0.0], True, None, rows_dev.data_ptr())
    ctx.synchronize()
    assert np.array_equal(rows, rows_dev.cpu().numpy())
    assert np.all(rows[:, 6] == 0)
    for b in (0, 1, 4194, 4195, 8999):  # around the slab boundary (128 MiB / 32 kB = 4194.3 buffers)
        ref = orc.nls_fit(x[b * R:(b + 1) * R], 200e3, 1000.0, 20, nh,
                          init_a=rows[0, 0] if b else 1.6, init_m=rows[0, 1] if b else 6.0,
                          init_psi=rows[0, 3] if b else 0.0)
        # the oracle cold-starts phi at 0 while the GPU seeds it from buffer 0: same minimum, 1e-8 gate
        assert_rows_match(rows[b:b + 1], ref)


@pytest.mark.parametrize("name", ["ekf_default", "ekf_offset"])
def test_ekf_vs_reference(torch_mod, ctx, golden, name):
    from deepfmkit_b200 import _lib
    g = golden(name)
    kw = ekf_kwargs(g)
    opts = _lib.default_ekf_opts()
    for i, key in enumerate(("init_a", "init_m", "init_phi", "init_psi")):
        if key in kw:
            opts.init[i] = kw[key]
    for i in range(5):
        if "p0_diag" in kw:
            opts.p0_diag[i] = kw["p0_diag"][i]
        if "q_diag" in kw:
            opts.q_diag[i] = kw["q_diag"][i]
    if "r_val" in kw:
        opts.r_val = kw["r_val"]
    rows = ctx.ekf_host(g["x"][None, :], 4000, 200e3, 1000.0, opts)[0]
    ref = g["rows"]
    assert np.max(np.abs(rows[:, :5] - ref[:, :5])) < 1e-9
    assert np.all(rows[:, 5] == 0) and np.all(rows[:, 6] == 1)


def test_ekf_batch_layouts_agree(torch_mod, golden):
    from deepfmkit_b200 import ekf_fit_batch
    g = golden("ekf_default")
    x = g["x"][:20000]
    z = np.stack([x, x[::-1].copy(), x * 1.01])
    a = ekf_fit_batch(z, 200e3, 1000.0, 20)
    b = ekf_fit_batch(np.ascontiguousarray(z.T), 200e3, 1000.0, 20, time_major=True)
    # the time-major record sums its mean / variance in a different order: equal to rounding, not to the bit
    assert np.max(np.abs(a - b)) <= 1e-12
    ref = orc.ekf_track(x, 200e3, 1000.0, 20)
    assert np.max(np.abs(a[0, :, :5] - ref[:, :5])) < 1e-11


def _ekf_case(args):
    x, kw = args
    return orc.ekf_track(x, 200e3, 1000.0, 20, **kw)


def test_ekf_random_settings_vs_oracle(torch_mod, ctx):
    """Eight channels with random true parameters, SNR, initial guesses, P0, Q and R through dfk_ekf_dev one call per
    setting (options are per call), each against the oracle loop over the same 12000 samples."""
    from multiprocessing import Pool
    from deepfmkit_b200 import _lib
    rng = np.random.RandomState(77)
    jobs, opts_list = [], []
    for c in range(8):
        x = orc.snr_signal(rng.uniform(3, 12), 200e3, 1000.0, 0.06, rng.choice([20.0, 40.0]), seed=200 + c,
                           phi0=rng.uniform(-3, 3), psi0=rng.uniform(-0.3, 0.3))
        kw = dict(init_a=rng.uniform(0.8, 2.0), init_m=rng.uniform(4, 10), init_phi=rng.uniform(-1, 1),
                  init_psi=rng.uniform(-0.2, 0.2), p0_diag=rng.uniform(0.1, 2.0, 5), q_diag=10.0 ** rng.uniform(-9, -5, 5),
                  r_val=None if c % 2 else float(10.0 ** rng.uniform(-5, -2)))
        jobs.append((x, kw))
    with Pool(8) as pool:
        refs = pool.map(_ekf_case, jobs)
    for (x, kw), ref in zip(jobs, refs):
        o = _lib.default_ekf_opts()
        o.init[0], o.init[1], o.init[2], o.init[3] = kw["init_a"], kw["init_m"], kw["init_phi"], kw["init_psi"]
        for i in range(5):
            o.p0_diag[i], o.q_diag[i] = kw["p0_diag"][i], kw["q_diag"][i]
        o.r_val = float("nan") if kw["r_val"] is None else kw["r_val"]
        rows = ctx.ekf_host(x[None, :], 4000, 200e3, 1000.0, o)[0]
        scale = np.maximum(np.abs(ref[:, :5]), 1.0)
        assert np.max(np.abs(rows[:, :5] - ref[:, :5]) / scale) < 1e-9, kw


def test_ekf_slabwise_equals_whole_record(torch_mod, ctx, golden):
    """dfk_ekf_stream_dev over three slabs with carried state == one pass over the record (same initial dc and R)."""
    from deepfmkit_b200 import _lib
    torch = torch_mod
    g = golden("ekf_default")
    x = g["x"][:36000]
    R, C = 4000, 3
    z = np.stack([x, x[::-1].copy(), 0.5 * x + 0.1])
    zd = torch.from_numpy(z).cuda()
    state = torch.zeros((C, 32), dtype=torch.float64, device="cuda")
    opts = _lib.default_ekf_opts()
    parts = []
    for k0, T in ((0, 12000), (12000, 8000), (20000, 16000)):
        slab = zd[:, k0:k0 + T].contiguous()
        rows = torch.empty((C, T // R, 8), dtype=torch.float64, device="cuda")
        ctx.ekf_stream_dev(slab.data_ptr(), T, C, 1, T, R, 200e3, 1000.0, opts, k0, state.data_ptr(), rows.data_ptr())
        ctx.synchronize()
        parts.append(rows.cpu().numpy())
    got = np.concatenate(parts, axis=1)
    for c in range(C):
        first = z[c, :12000]
        ref = orc.ekf_track(z[c], 200e3, 1000.0, 20, r_val=np.var(first), init_dc=np.mean(first))
        assert np.max(np.abs(got[c, :, :5] - ref[:, :5])) < 1e-9
    with pytest.raises(RuntimeError):
        ctx.ekf_stream_dev(zd.data_ptr(), 8000, C, 1, 36000, R, 200e3, 1000.0, opts, 4000, None, 0)


def test_synth_slabs_tile_the_record(torch_mod, ctx):
    torch = torch_mod
    T, C = 30000, 3
    whole = torch.empty((C, T), dtype=torch.float64, device="cuda")
    ctx.synth_snr_dev(whole.data_ptr(), T, C, 200e3, 1000.0, 6.0, dphi=0.4, seed=21)
    parts = torch.zeros((C, T), dtype=torch.float64, device="cuda")
    for t0, n in ((0, 10000), (10000, 4000), (14000, 16000)):
        ctx.synth_snr_slab_dev(parts.data_ptr() + t0 * 8, n, C, T, t0, 200e3, 1000.0, 6.0, dphi=0.4, seed=21)
    ctx.synchronize()
    assert torch.equal(whole, parts)


# ---- the reference-facing Python API --------------------------------------------------------------
def test_fitter_api_and_result_frame(torch_mod, golden):
    import pandas as pd
    from deepfmkit_b200 import DeepFitFramework, DeepRawObject, StandardNLSFitter
    g = golden("facade_quickstart")
    x = orc.snr_signal(6.0, 200e3, 1000, 1, 40.0, seed=0)
    raw = DeepRawObject(data=pd.DataFrame(x, columns=["ch0"]), f_samp=200e3, f_mod=1000, label="dynamic_channel")
    df = StandardNLSFitter({"n": 20}).fit(raw, parallel=True, n_cores=3)
    assert list(df.columns) == list(g["columns"][:7])
    assert [str(t) for t in df.dtypes] == list(g["dtypes"][:7])
    assert_rows_match(df.to_numpy(dtype=float), g["values"][:, :7])

    class _Laser:
        df = float(g["laser_df"])

    class _Sim:
        label = "dynamic_channel"
        laser = _Laser()
        fit_n = 20

    raw.sim = _Sim()
    dff = DeepFitFramework()
    dff.load_raw_object(raw)
    dff.sims["dynamic_channel"] = raw.sim
    fobj = dff.fit("dynamic_channel", parallel=False)
    out = dff.fits_df["dynamic_channel_nls"]
    assert list(out.columns) == list(g["columns"])
    assert [str(t) for t in out.dtypes] == list(g["dtypes"])
    assert np.allclose(out["tau"].to_numpy(), g["values"][:, 7], rtol=1e-8, atol=0)
    assert np.array_equal(fobj.time, g["time"])
    assert [fobj.n, fobj.R, fobj.fs, fobj.nbuf, fobj.ndata, fobj.init_a, fobj.init_m, fobj.f_samp, fobj.f_mod] == \
        list(g["scalars"])
    assert dff.fit("nope") is None and dff.fit("dynamic_channel", method="bogus") is None


def test_fitter_edge_cases(torch_mod):
    import pandas as pd
    from deepfmkit_b200 import DeepRawObject, EKFFitter, StandardNLSFitter
    with pytest.raises(ValueError):
        StandardNLSFitter({})
    short = DeepRawObject(data=np.ones(100), f_samp=200e3, f_mod=1000)
    assert StandardNLSFitter({"n": 20}).fit(short).empty  # nbuf == 0 -> empty frame (fitters.py:363)
    x = orc.snr_signal(6.0, 200e3, 1000, 0.0417, 40.0, seed=2)  # ragged tail: 2 buffers + 340 samples
    raw = DeepRawObject(data=x, f_samp=200e3, f_mod=1000)
    df = StandardNLSFitter({"n": 20, "ndata": 12}).fit(raw, init_m=6.2)
    assert len(df) == 2 and set(df["fitok"]) == {0}
    before = raw.data.to_numpy().copy()
    e = EKFFitter({"n": 20}).fit(raw, verbose=False)
    assert len(e) == 2 and list(e.columns) == ["amp", "m", "phi", "psi", "dc", "ssq", "fitok"]
    assert e["fitok"].dtype == np.int64 and set(e["fitok"]) == {1}
    assert np.array_equal(raw.data.to_numpy(), before)  # caller's record untouched


def test_tunables_are_read_at_call_time(torch_mod, golden):
    from deepfmkit_b200 import DeepRawObject, StandardNLSFitter
    from deepfmkit_b200 import fit as tun
    g, x, f_samp, f_mod, n, nh, R, w0 = _case(golden, "fallback_m16")
    raw = DeepRawObject(data=x, f_samp=f_samp, f_mod=f_mod)
    base = StandardNLSFitter({"n": n, "ndata": nh}).fit(raw)
    assert 1 in set(base["fitok"])
    old = tun.M_GRID_MAX
    try:
        tun.M_GRID_MAX = 10.0  # the grid no longer reaches m = 16: buffer 0 stays unfitted
        patched = StandardNLSFitter({"n": n, "ndata": nh}).fit(raw)
    finally:
        tun.M_GRID_MAX = old
    ref = orc.nls_fit(x, f_samp, f_mod, n, nh, schedule="gpu", tun=orc.with_tunables(m_grid_max=10.0))
    assert np.array_equal(patched["fitok"].to_numpy(), ref[:, 6].astype(np.int64))
    assert patched["fitok"].iloc[0] == 2


# ---- full-size properties (BASELINE configs at scale, no oracle possible) -------------------------------
def test_large_record_properties(torch_mod, ctx):
    """cfg 2 geometry (1 MHz, R = 20000) on a 1.6 GB device-generated record: linearity of the lock-in,
    agreement of the two demod kernels, and fits that recover the generating parameters."""
    torch = torch_mod
    R, nh, nbuf = 20000, 10, 10000
    w0 = orc.rad_per_sample(1e6, 1000.0)
    x = torch.empty(nbuf * R, dtype=torch.float64, device="cuda")
    ctx.synth_snr_dev(x.data_ptr(), nbuf * R, 1, 1e6, 1000.0, 6.0, snr_db=40.0, seed=11)
    qi = torch.empty((nbuf, 2 * nh), dtype=torch.float64, device="cuda")
    dc = torch.empty(nbuf, dtype=torch.float64, device="cuda")
    ctx.demod(x.data_ptr(), nbuf, R, nh, w0, qi.data_ptr(), dc.data_ptr())
    ctx.synchronize()
    # direct kernel on a misaligned copy of the first 64 buffers
    y = torch.empty(64 * R + 1, dtype=torch.float64, device="cuda")
    y[1:] = x[: 64 * R]
    qi2 = torch.empty((64, 2 * nh), dtype=torch.float64, device="cuda")
    dc2 = torch.empty(64, dtype=torch.float64, device="cuda")
    ctx.demod(y.data_ptr() + 8, 64, R, nh, w0, qi2.data_ptr(), dc2.data_ptr())
    ctx.synchronize()
    scale = qi[:64].abs().max(dim=1, keepdim=True).values
    assert float(((qi[:64] - qi2).abs() / scale).max()) <= IQ_TOL
    # linearity: demod(2x + 1) = 2 demod(x) (harmonics reject the constant), dc -> 2 dc + 1
    z = x[: 256 * R] * 2.0 + 1.0
    qi3 = torch.empty((256, 2 * nh), dtype=torch.float64, device="cuda")
    dc3 = torch.empty(256, dtype=torch.float64, device="cuda")
    ctx.demod(z.data_ptr(), 256, R, nh, w0, qi3.data_ptr(), dc3.data_ptr())
    ctx.synchronize()
    assert float((qi3 - 2 * qi[:256]).abs().max()) < 1e-13
    assert float((dc3 - (2 * dc[:256] + 1)).abs().max()) < 1e-13
    # whole readout
    rows = torch.empty((nbuf, 8), dtype=torch.float64, device="cuda")
    ctx.nls_fit_dev(x.data_ptr(), nbuf, R, nh, w0, [1.6, 6.0, 0.0, 0.0], True, None, rows.data_ptr())
    ctx.synchronize()
    r = rows.cpu().numpy()
    assert np.all(r[:, 6] == 0)
    assert abs(r[:, 1].mean() - 6.0) < 1e-4 and r[:, 1].std() < 2e-3
    assert abs(r[:, 0].mean() - 1.0) < 1e-4 and abs(r[:, 2].mean()) < 1e-3


def test_synth_statistics(torch_mod, ctx):
    torch = torch_mod
    T = 4_000_000
    x = torch.empty(2 * T, dtype=torch.float64, device="cuda")
    ctx.synth_snr_dev(x.data_ptr(), T, 2, 200e3, 1000.0, 6.0, phi0=0.3, dphi=0.5, snr_db=20.0, seed=9)
    ctx.synchronize()
    for c in range(2):
        clean = orc.snr_signal(6.0, 200e3, 1000.0, 0.05, 300.0, seed=0, phi0=0.3 + 0.5 * c)  # 300 dB = noise free
        xc = x[c * T:(c + 1) * T].cpu().numpy()
        noise = xc[: len(clean)] - clean
        full_noise = xc - np.tile(clean[:200], T // 200)
        sigma = np.sqrt(np.mean((clean - clean.mean()) ** 2) / 100.0)
        assert abs(full_noise.mean()) < 5 * sigma / np.sqrt(T)
        assert abs(full_noise.std() / sigma - 1) < 5e-3
        assert abs(np.corrcoef(full_noise[:-1], full_noise[1:])[0, 1]) < 5e-3
        assert np.abs(noise).max() < 7 * sigma
    a = x[:T].cpu().numpy()
    b = x[T:].cpu().numpy()
    assert abs(np.corrcoef(a - a.mean(), b - b.mean())[0, 1]) < 0.9  # different seeds per channel


@pytest.mark.parametrize("script,token,port", [("sharded_record.py", "SHARDED_OK", 29577),
                                               ("sharded_sweep.py", "SWEEP_OK", 29578),
                                               ("sharded_experiment.py", "EXPERIMENT_OK", 29579)])
def test_sharded_over_gpus(torch_mod, script, token, port):
    """One record cut into contiguous slabs / one Monte-Carlo sweep split by realisation / one Experiment split by grid
    point over all visible GPUs (torchrun, nccl) == the one-GPU result."""
    import subprocess
    import sys
    n = torch_mod.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    n = min(n, 4)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(root, "tests", "multi", script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert f"{token} world={n}" in out.stdout


def test_monte_carlo_sweep_reaches_crlb(torch_mod):
    """Config-5 workload through the batched driver: device-generated realisations, per-m statistics.  The NLS fit
    is efficient (notebook 1.1_CRLB-test): std(m_hat) sits on the reference's CRLB, the mean on the truth."""
    from deepfmkit_b200 import crlb_sigma_m, nls_sweep
    ms = [3.0, 6.0, 11.0, 17.0]
    n_trials = 20000
    out = nls_sweep(ms, n_trials, snr_db=40.0, ndata=15, seed=5, max_resident_bytes=48 << 20, return_rows=True)
    assert out["rows"].shape == (4, n_trials, 8)
    assert np.all(out["fitok"][:, 0] > 0.999)
    # helpers.py:16-45 assumes an AC power of 0.5; 'snr' mode scales the noise to the record's actual AC power
    # (physics.py:522-527), which depends on m -- put the bound on the same footing before comparing
    ac_power = np.array([np.var(orc.snr_signal(m, 200e3, 1000.0, 1e-3, 300.0)) for m in ms])
    ratio = out["m_std"] / (out["crlb_sigma_m"] * np.sqrt(ac_power / 0.5))
    assert np.all((ratio > 0.96) & (ratio < 1.05)), ratio
    assert np.all(np.abs(out["m_mean"] - np.array(ms)) < 5 * out["crlb_sigma_m"] / np.sqrt(n_trials) + 2e-5)
    assert np.allclose(out["rows"][:, :, 1].std(axis=1), out["m_std"], rtol=1e-6)
    assert np.allclose(out["crlb_sigma_m"], [orc.crlb_sigma_m(m, 15, 40.0, 200) for m in ms], rtol=1e-12)
    # a handful of device-generated realisations re-fitted by the oracle: same rows within the gate
    import torch
    from deepfmkit_b200 import _lib
    ctx = _lib.get_context(0)
    x = torch.empty((3, 200), dtype=torch.float64, device="cuda")
    ctx.synth_snr_slab_dev(x.data_ptr(), 200, 3, 200, 0, 200e3, 1000.0, 6.0, snr_db=40.0, seed=5 + 1 * n_trials)
    ctx.synchronize()
    xs = x.cpu().numpy()
    for r in range(3):
        ref = orc.nls_fit(xs[r], 200e3, 1000.0, 1, 15, init_m=6.0)[0]
        got = out["rows"][1, r]
        assert got[6] == ref[6] and np.max(np.abs(got[:4] - ref[:4])) < 1e-8


def _bulk_case(args):
    m, snr, seed, init_m, phi0, psi0 = args
    x = orc.snr_signal(m, 200e3, 1000.0, 2e-3, snr, seed=seed, phi0=phi0, psi0=psi0)
    with np.errstate(all="ignore"):
        row = orc.nls_fit(x, 200e3, 1000.0, 2, 12, init_m=init_m)[0]
    return x, row


def test_bulk_parity_over_random_cold_starts(torch_mod):
    """1200 random single-buffer fits -- m in [2, 22], SNR 3..40 dB, random phases, half of them cold-started at
    m = 6 so that the grid fallback and the failing branch are exercised at scale -- GPU batch vs oracle, one by one:
    identical flags everywhere, parameters within the gate wherever the reference itself fits."""
    from multiprocessing import Pool
    from deepfmkit_b200 import nls_fit_batch
    rng = np.random.RandomState(2024)
    n = 1200
    ms = rng.uniform(2.0, 22.0, n)
    snrs = rng.choice([3.0, 10.0, 20.0, 40.0], n)
    init_m = np.where(rng.rand(n) < 0.5, 6.0, ms)
    phis, psis = rng.uniform(-3, 3, n), rng.uniform(-0.5, 0.5, n)
    jobs = [(ms[i], snrs[i], 1000 + i, init_m[i], phis[i], psis[i]) for i in range(n)]
    with Pool(min(16, os.cpu_count() or 1)) as pool:
        res = pool.map(_bulk_case, jobs, chunksize=25)
    x = np.stack([r[0] for r in res])
    ref = np.stack([r[1] for r in res])
    rows = nls_fit_batch(x, 200e3, 1000.0, 2, ndata=12, init_m=init_m, seeded=False)[:, 0, :]
    flags_equal = rows[:, 6] == ref[:, 6]
    assert {0.0, 1.0, 2.0} <= set(ref[:, 6]), "the sample must exercise all three flags"
    assert np.all(flags_equal), (np.flatnonzero(~flags_equal), rows[~flags_equal, :7], ref[~flags_equal])
    ok = ref[:, 6] < 2
    err = np.abs(rows[ok, :4] - ref[ok, :4])
    err[:, :2] /= np.abs(ref[ok, :2])
    assert err.max() <= PARAM_TOL, (err.max(), np.unravel_index(err.argmax(), err.shape))


def test_many_parked_fits_take_the_flat_retry_kernel(torch_mod):
    """>= 4096 parked fits switch the retry stage from a warp per fit to a thread per fit.  The same 9000 cold fits
    (init_m = 6, m_true in 9..21: nearly all need the grid fallback) in one call (thread per fit) and in chunks of
    1500 (warp per fit) must agree, and a sample of them must agree with the oracle."""
    from multiprocessing import Pool
    from deepfmkit_b200 import nls_fit_batch
    rng = np.random.RandomState(11)
    n = 9000
    ms = rng.uniform(9.0, 21.0, n)
    phis = rng.uniform(-3, 3, n)
    jobs = [(ms[i], 30.0, 5000 + i, 6.0, phis[i], 0.0) for i in range(n)]
    sample = list(range(0, n, 45))
    with Pool(min(16, os.cpu_count() or 1)) as pool:
        res = pool.map(_bulk_case, [jobs[i] for i in sample], chunksize=10)
    t = np.arange(400) / 200e3
    x = np.stack([orc.snr_signal(ms[i], 200e3, 1000.0, 2e-3, 30.0, seed=5000 + i, phi0=phis[i]) for i in range(n)])
    whole = nls_fit_batch(x, 200e3, 1000.0, 2, ndata=12, init_m=6.0, seeded=False)[:, 0, :]
    assert np.mean(whole[:, 6] >= 1) > 0.9, "the sample must park most fits"
    chunks = np.concatenate([nls_fit_batch(x[i:i + 1500], 200e3, 1000.0, 2, ndata=12, init_m=6.0, seeded=False)[:, 0, :]
                             for i in range(0, n, 1500)])
    assert np.array_equal(whole[:, 6], chunks[:, 6])
    ok = whole[:, 6] < 2
    assert np.max(np.abs(whole[ok, :4] - chunks[ok, :4])) < PARAM_TOL
    ref = np.stack([r[1] for r in res])
    got = whole[sample]
    assert np.array_equal(got[:, 6], ref[:, 6])
    okr = ref[:, 6] < 2
    err = np.abs(got[okr, :4] - ref[okr, :4])
    err[:, :2] /= np.abs(ref[okr, :2])
    assert err.max() <= PARAM_TOL


@pytest.mark.parametrize("R,nh,nbuf,c0", [(200, 15, 1003, 0), (200, 10, 64, 4096), (100, 20, 517, 7), (400, 10, 300, 1), (80, 62, 33, 0)])
def test_fused_sweep_equals_generate_then_demodulate(torch_mod, ctx, R, nh, nbuf, c0):
    """dfk_sweep_demod_dev makes the realisations inside the demodulation kernel (one period per record) or through
    scratch (R = 400 here: two periods): either way bit for bit what dfk_synth_snr_slab_dev + dfk_demod give."""
    f_samp, f_mod, m, seed = 200e3, 1000.0 if R != 100 and R != 80 else (2000.0 if R == 100 else 2500.0), 7.25, 12345
    w0 = 2.0 * np.pi * f_mod / f_samp
    x = torch_mod.empty((nbuf, R), dtype=torch_mod.float64, device="cuda")
    ctx.synth_snr_slab_dev(x.data_ptr(), R, nbuf, R, 0, f_samp, f_mod, m, phi0=0.3, psi0=0.1, snr_db=35.0, seed=seed + c0)
    qi0 = torch_mod.empty((nbuf, 2 * nh), dtype=torch_mod.float64, device="cuda")
    dc0 = torch_mod.empty(nbuf, dtype=torch_mod.float64, device="cuda")
    ctx.demod(x.data_ptr(), nbuf, R, nh, w0, qi0.data_ptr(), dc0.data_ptr())
    qi1 = torch_mod.full((nbuf, 2 * nh), float("nan"), dtype=torch_mod.float64, device="cuda")
    dc1 = torch_mod.full((nbuf,), float("nan"), dtype=torch_mod.float64, device="cuda")
    ctx.sweep_demod_dev(nbuf, c0, R, nh, f_samp, f_mod, m, qi1.data_ptr(), dc1.data_ptr(), phi0=0.3, psi0=0.1, snr_db=35.0,
                        seed=seed)
    ctx.synchronize()
    assert torch_mod.equal(qi0, qi1) and torch_mod.equal(dc0, dc1)


def test_time_major_batch_equals_channel_major(torch_mod):
    """Interleaved [T, C] records (the layout of acquisition hardware) folded in place == the channel-major readout."""
    from deepfmkit_b200 import nls_fit_batch
    xs = np.stack([orc.snr_signal(6.0 + 0.5 * c, 200e3, 1000.0, 0.1, 40.0, seed=c, phi0=0.2 * c) for c in range(5)])
    a = nls_fit_batch(xs, 200e3, 1000.0, 20)
    b = nls_fit_batch(np.ascontiguousarray(xs.T), 200e3, 1000.0, 20, time_major=True)
    c = nls_fit_batch(torch_mod.from_numpy(np.ascontiguousarray(xs.T)).cuda(), 200e3, 1000.0, 20, time_major=True, return_tensor=True)
    assert np.array_equal(b, c.cpu().numpy())
    assert np.array_equal(a[:, :, 6], b[:, :, 6]) and np.max(np.abs(a[:, :, :4] - b[:, :, :4])) < 1e-9
    assert np.max(np.abs(a[:, :, 4] - b[:, :, 4])) < 1e-14
    # a geometry that cannot fold (75 samples per period, one period per buffer): the transposing fallback, same rows
    ys = np.stack([orc.snr_signal(6.0, 30e3, 400.0, 0.05, 40.0, seed=c) for c in range(3)])
    a = nls_fit_batch(ys, 30e3, 400.0, 1)
    b = nls_fit_batch(np.ascontiguousarray(ys.T), 30e3, 400.0, 1, time_major=True)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("P,n,nh,C,nbuf", [(200, 20, 10, 256, 7), (200, 20, 10, 1, 9), (1000, 20, 10, 2, 5), (200, 3, 62, 5, 11),
                                           (128, 1, 15, 33, 4), (2000, 40, 40, 3, 3), (76, 2, 7, 27, 6)])
def test_time_major_demod_matches_reference_lockin(torch_mod, ctx, P, n, nh, C, nbuf):
    """dfk_demod_tm_dev on x[t, c] against the oracle's lock-in of every channel: the drift term on long / many-harmonic
    buffers, one channel, odd channel counts, super-periods that are no multiple of the 2048-column chunk."""
    f_mod = 1000.0
    f_samp = f_mod * P
    R = P * n
    w0 = orc.rad_per_sample(f_samp, f_mod)
    rng = np.random.RandomState(P + C)
    t = np.arange(nbuf * R)
    x = np.stack([1.0 + np.cos(0.3 * c + (3.0 + 0.01 * c) * np.cos(2 * np.pi * t / P + 0.2)) + 0.01 * rng.randn(nbuf * R)
                  for c in range(C)], axis=1)  # [T, C]
    xd = torch_mod.from_numpy(np.ascontiguousarray(x)).cuda()
    qi = torch_mod.full((C * nbuf, 2 * nh), float("nan"), dtype=torch_mod.float64, device="cuda")
    dc = torch_mod.full((C * nbuf,), float("nan"), dtype=torch_mod.float64, device="cuda")
    ctx.demod_tm(xd.data_ptr(), nbuf, C, R, nh, w0, qi.data_ptr(), dc.data_ptr())
    ctx.synchronize()
    qi, dc = qi.cpu().numpy().reshape(C, nbuf, 2 * nh), dc.cpu().numpy().reshape(C, nbuf)
    for c in sorted({0, C // 2, C - 1}):
        for b in sorted({0, nbuf - 1}):
            buf = x[b * R:(b + 1) * R, c]
            ref = orc.lockin_means(buf, w0, nh)
            assert np.max(np.abs(qi[c, b] - ref)) <= IQ_TOL * max(np.abs(ref).max(), 1e-3), (c, b)
            assert abs(dc[c, b] - buf.mean()) <= 1e-14 * abs(buf.mean())


@pytest.mark.parametrize("R,nh,f_mod", [(3240, 10, 1234.5), (3241, 8, 1234.5), (5, 3, 1234.5), (31, 12, 977.7), (33, 16, 977.7),
                                        (4097, 40, 1234.5), (1000, 64, 977.7), (300_001, 10, 1234.5),
                                        (700_000, 20, 1001.3), (64, 1, 1234.5)])
def test_direct_lockin_shapes(torch_mod, ctx, R, nh, f_mod):
    """Records that cannot fold (incommensurate modulation period) through the table-driven lock-in: buffers shorter
    than a warp, ragged last steps, 8 / 12 / 16 harmonics per pass, several passes (N = 20, 40, 64), buffers longer
    than one table chunk (300 001 and 700 000 samples), against the reference's own formulation (oracle)."""
    from deepfmkit_b200 import _lib
    f_samp = 200e3
    w0 = orc.rad_per_sample(f_samp, f_mod)
    assert _lib.demod_path(R, w0) == 0
    nbuf = 19 if R < 100_000 else 3
    rng = np.random.RandomState(R + nh)
    t = np.arange(nbuf * R)
    x = 1.0 + np.cos(0.4 + 5.0 * np.cos(w0 * t + 0.3)) + 0.01 * rng.randn(nbuf * R)
    qi, dc = gpu_demod(torch_mod, ctx, x, R, nh, w0)
    worst = 0.0
    for b in range(nbuf):
        buf = x[b * R:(b + 1) * R]
        ref = orc.lockin_means(buf, w0, nh)
        worst = max(worst, np.max(np.abs(qi[b] - ref)) / max(np.abs(ref).max(), 1e-3))
        assert abs(dc[b] - buf.mean()) <= 1e-14 * abs(buf.mean())
    assert worst <= IQ_TOL, worst


@pytest.mark.parametrize("R,nh", [(517, 10), (4000, 7), (64, 9), (40001, 10)])
def test_direct_lockin_pairs(torch_mod, ctx, R, nh):
    """With enough pieces the table-driven lock-in takes two per warp: same numbers as one per warp (the per-lane
    order of operations is the same), an odd last one included, and the oracle's on a sample of the buffers.  40 001
    samples: three chunks per buffer, the last one ragged, so that pairs straddle buffers and lengths."""
    from deepfmkit_b200 import _lib
    f_samp, f_mod = 200e3, 1234.5
    w0 = orc.rad_per_sample(f_samp, f_mod)
    assert _lib.demod_path(R, w0) == 0
    nbuf = 2 * 12 * 160 + 1 if R < 10000 else 1281
    rng = np.random.RandomState(R)
    t = np.arange(nbuf * R)
    x = 1.0 + np.cos(0.4 + 5.0 * np.cos(w0 * t + 0.3)) + 0.01 * rng.randn(nbuf * R)
    qi, dc = gpu_demod(torch_mod, ctx, x, R, nh, w0)
    with _lib.dev_overrides(DFK_DIRECT_PAIR=0):
        qi1, dc1 = gpu_demod(torch_mod, ctx, x, R, nh, w0)
    assert np.array_equal(qi, qi1) and np.array_equal(dc, dc1)
    for b in list(range(0, nbuf, 397)) + [nbuf - 2, nbuf - 1]:
        buf = x[b * R:(b + 1) * R]
        ref = orc.lockin_means(buf, w0, nh)
        assert np.max(np.abs(qi[b] - ref)) <= IQ_TOL * np.abs(ref).max(), b
        assert abs(dc[b] - buf.mean()) <= 1e-14 * abs(buf.mean())
