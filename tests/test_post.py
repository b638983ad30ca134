"""Post-fit step: block means and the log-frequency spectral estimate.

CPU part: the oracle against the reference-minted downsample fixture, the C ABI's frequency plan against the oracle's
bit for bit, physical sanity of the oracle estimate.  GPU part (-m gpu): the CUDA kernels through the C ABI against the
oracle.  Gates: block means within 4 ulp of the mean's magnitude (the summation order differs from numpy's pairwise
one); spectra within 1e-10 relative (window, twiddle and sums are fp64 throughout; measured ~1e-13).
"""
import numpy as np
import pytest

from oracle import post_oracle as po

SPEC_TOL = 1e-10


def _series(seed, n, fs):
    rng = np.random.RandomState(seed)
    t = np.arange(n) / fs
    return 0.7 + 1e-3 * np.cumsum(rng.randn(n)) + 0.05 * rng.randn(n) + 0.2 * np.sin(2 * np.pi * 0.123 * fs * t / 10)


def _golden_input(seed, n):
    rng = np.random.RandomState(seed)
    return 1.0 + 0.3 * rng.randn(n) + np.sin(np.arange(n) * 1e-3)


# ---------------------------------------------------------------------------------------------------------- CPU
def test_oracle_downsample_matches_reference_fixture(golden):
    g = golden("post_downsample")
    for seed, n, R in g["cases"]:
        y = po.vectorized_downsample(_golden_input(int(seed), int(n)), int(R))
        assert np.array_equal(y, g[f"y{seed}"])
    assert po.vectorized_downsample(np.ones(10), 0).size == 0
    assert po.vectorized_downsample(np.ones(10), 2.0).size == 0


@pytest.mark.parametrize("N,fs,J,K,olap,bmin,lmin", [(20000, 50.0, 200, 100, None, 1, 0), (180000, 50.0, 500, 100, None, 1, 0),
                                                      (5000, 10.0, 500, 100, 0.5, 1, 0), (12345, 1.0, 50, 10, 0.3, 2.5, 64),
                                                      (300, 2.0, 500, 100, None, 1, 0), (2, 1.0, 10, 10, None, 1, 0)])
def test_plan_matches_oracle_bitwise(N, fs, J, K, olap, bmin, lmin):
    from deepfmkit_b200 import _lib
    o = _lib.default_lpsd_opts()
    o.jdes, o.kdes, o.bmin, o.lmin = J, K, bmin, lmin
    if olap is not None:
        o.olap = olap
    p = _lib.lpsd_plan(N, fs, o)
    f, r, m, L, Kk = po.ltf_plan(N, fs, po.default_overlap("kaiser") if olap is None else olap, bmin, lmin, J, K)
    assert np.array_equal(f, p["f"]) and np.array_equal(r, p["r"]) and np.array_equal(m, p["m"])
    assert np.array_equal(L, p["L"]) and np.array_equal(Kk, p["K"])
    # every segment lies inside the record and the last one ends at its end
    for l, k in zip(L, Kk):
        s = po.segment_starts(N, int(l), int(k))
        assert s[0] == 0 and s[-1] + l <= N
        if k > 1:
            assert s[-1] + l == N


def test_plan_rejects_bad_arguments():
    from deepfmkit_b200 import _lib
    o = _lib.default_lpsd_opts()
    with pytest.raises(RuntimeError):
        _lib.lpsd_plan(1, 1.0, o)
    o.order = 3
    with pytest.raises(RuntimeError):
        _lib.lpsd_plan(100, 1.0, o)
    o = _lib.default_lpsd_opts()
    o.olap = 1.0
    with pytest.raises(RuntimeError):
        _lib.lpsd_plan(100, 1.0, o)


def test_oracle_spectrum_is_calibrated():
    """White noise of variance s^2 has the one-sided density 2 s^2 / fs; a sine of amplitude A has power A^2 / 2."""
    rng = np.random.RandomState(3)
    N, fs = 30000, 20.0
    x = 0.1 * rng.randn(N) + 2.0
    f, ps, psd, enbw, K = po.lpsd(x, fs, Jdes=150)
    hi = f > 0.5
    assert abs(np.mean(psd[hi]) / (2 * 0.01 / fs) - 1) < 0.05
    assert np.allclose(ps, psd * enbw, rtol=1e-12)
    f0 = f[len(f) // 2]  # a sine exactly on a planned frequency: no scalloping
    y = 0.5 * np.sin(2 * np.pi * f0 * np.arange(N) / fs + 0.3)
    _, ps2, _, _, _ = po.lpsd(y, fs, Jdes=150)
    assert abs(ps2[len(f) // 2] / 0.125 - 1) < 1e-6
    assert abs(po.kaiser_alpha(200) - 8.0858879) < 1e-9 and 0.7 < po.default_overlap("kaiser") < 0.8


# ---------------------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def torch_mod():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.mark.gpu
def test_downsample_device_and_host_paths(torch_mod, golden):
    from deepfmkit_b200 import _lib, vectorized_downsample
    g = golden("post_downsample")
    for seed, n, R in g["cases"]:
        x = _golden_input(int(seed), int(n))
        ref = g[f"y{seed}"]
        got = vectorized_downsample(x, int(R))
        assert got.shape == ref.shape
        if ref.size:
            assert np.max(np.abs(got - ref)) <= 4 * np.finfo(float).eps * np.max(np.abs(ref))
        if n // R:
            gd = vectorized_downsample(torch_mod.from_numpy(x).cuda(), int(R))
            assert gd.is_cuda and np.array_equal(gd.cpu().numpy(), got)
            # misaligned start (odd offset) takes the scalar path
            if n - 1 >= R:
                g1 = vectorized_downsample(torch_mod.from_numpy(x).cuda()[1:], int(R)).cpu().numpy()
                r1 = po.vectorized_downsample(x[1:], int(R))
                assert np.max(np.abs(g1 - r1)) <= 4 * np.finfo(float).eps * np.max(np.abs(r1))
    assert vectorized_downsample(np.ones(5), 0).size == 0 and vectorized_downsample(np.ones(5), 9).size == 0
    # streamed host path: small slabs force several double-buffered copies
    ctx = _lib.get_context(0)
    x = _golden_input(11, 4000 * 301 + 17)
    ctx.set_host_slab_bytes(4000 * 8 * 7)
    try:
        got = ctx.downsample_host(x, 4000)
    finally:
        ctx.set_host_slab_bytes(0)
    ref = po.vectorized_downsample(x, 4000)
    assert np.max(np.abs(got - ref)) <= 4 * np.finfo(float).eps * np.max(np.abs(ref))


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(), dict(order=-1), dict(order=1), dict(order=2), dict(win="hann", olap=0.5),
                                dict(Jdes=60, Kdes=20, bmin=2.0, Lmin=32), dict(psll=120, olap=0.6)])
def test_lpsd_matches_oracle(torch_mod, kw):
    from deepfmkit_b200 import lpsd
    N, fs = 6000, 50.0
    x = _series(5, N, fs)
    okw = dict(Jdes=120)
    okw.update(kw)
    f, ps, psd, enbw, K = po.lpsd(x, fs, **okw)
    gf, gps, gpsd, genbw, gK, plan = lpsd(x, fs, **okw)
    assert np.array_equal(gf, f) and np.array_equal(gK, K)
    assert np.max(np.abs(gps / ps - 1)) < SPEC_TOL
    assert np.max(np.abs(gpsd / psd - 1)) < SPEC_TOL
    assert np.max(np.abs(genbw / enbw - 1)) < 1e-12
    assert len(plan["L"]) == len(f)


@pytest.mark.gpu
def test_lpsd_long_segments_strided_and_batched(torch_mod):
    """CTA-per-group path (L >= 2048), a column of a row table read in place (stride 8), several series per launch."""
    from deepfmkit_b200 import _lib, lpsd
    N, fs = 40000, 50.0
    xs = np.stack([_series(s, N, fs) for s in (1, 2, 3)])
    ref = [po.lpsd(x, fs, Jdes=40, Kdes=8) for x in xs]
    out = lpsd(torch_mod.from_numpy(xs).cuda(), fs, Jdes=40, Kdes=8, return_type="dict")
    assert out["psd"].shape == (3, len(ref[0][0])) and int(np.max(_lib.lpsd_plan(N, fs, _opts(40, 8))["L"])) >= 2048
    for c in range(3):
        assert np.max(np.abs(out["psd"][c] / ref[c][2] - 1)) < SPEC_TOL
    rows = torch_mod.zeros((N, 8), dtype=torch_mod.float64, device="cuda")
    rows[:, 2] = torch_mod.from_numpy(xs[0]).cuda()
    ctx = _lib.get_context(0)
    got = ctx.lpsd_dev(rows.data_ptr() + 2 * 8, N, 8, 1, 0, fs, _opts(40, 8))
    assert np.max(np.abs(got["psd"][0] / ref[0][2] - 1)) < SPEC_TOL


def _opts(J, K):
    from deepfmkit_b200 import _lib
    o = _lib.default_lpsd_opts()
    o.jdes, o.kdes = J, K
    return o


@pytest.mark.gpu
def test_facade_calc_lpsd(torch_mod):
    """fit -> calc_lpsd through the facade, the step every science notebook takes after the fit (core.py:590-609)."""
    from deepfmkit_b200 import DeepFitFramework, DeepRawObject
    from oracle import dfmi_oracle as orc
    x = orc.snr_signal(6.0, 200e3, 1000.0, 2.0, 40.0, seed=4)
    dff = DeepFitFramework()
    dff.load_raw_object(DeepRawObject(data=x, f_samp=200e3, f_mod=1000.0, label="r"))
    fit = dff.fit("r", n=4)
    assert fit.f is None and fit.olap == "default" and fit.Jdes == 500 and fit.psll == 200
    dff.calc_lpsd()
    f, ps, psd, enbw, K = po.lpsd(fit.phi, fit.fs)
    assert np.array_equal(fit.f, f) and np.max(np.abs(fit.Sxx / psd - 1)) < SPEC_TOL
    dff.calc_lpsd(labels=["nope"])  # logs, does not raise
