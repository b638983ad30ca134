"""Host-side logic of the callers around the path (no GPU): Monte-Carlo partial statistics, the CRLB restatement,
option mapping of the spectral estimate, the .npy header reader of the binary ingest."""
import numpy as np
import pytest

from oracle import dfmi_oracle as orc


def test_sweep_partials_combine_independently_of_the_split():
    from deepfmkit_b200.montecarlo import _combine_partials
    rng = np.random.RandomState(0)
    ms = np.array([3.0, 6.0, 11.0])
    truth = np.stack([np.ones(3), ms, np.zeros(3), np.zeros(3)], 1)
    vals = truth[:, None, :] + 1e-3 * rng.randn(3, 1000, 4)
    flags = rng.choice([0, 1, 2], size=(3, 1000), p=[0.9, 0.08, 0.02])
    ssq = rng.rand(3, 1000) * 1e-5

    def part(lo, hi):
        v, f = vals[:, lo:hi], flags[:, lo:hi]
        return {"n": hi - lo, "truth": truth, "sum": v.sum(1), "sumsq": ((v - truth[:, None, :]) ** 2).sum(1), "min": v.min(1),
                "max": v.max(1), "worst": np.abs(v - truth[:, None, :]).max(1),
                "ok": np.stack([(f == s).sum(1) for s in (0, 1, 2)], 1).astype(float), "ssq": ssq[:, lo:hi].sum(1)}

    whole = _combine_partials(ms, [part(0, 1000)], 15, 40.0, 200)
    split = _combine_partials(ms, [part(0, 333), part(333, 900), part(900, 1000)], 15, 40.0, 200)
    for key in whole:
        assert np.allclose(whole[key], split[key], rtol=1e-12, atol=1e-15), key
    assert np.allclose(whole["m_mean"], vals[:, :, 1].mean(1)) and np.allclose(whole["m_std"], vals[:, :, 1].std(1))
    assert np.allclose(whole["fitok"].sum(1), 1.0) and whole["n_trials"] == 1000
    assert np.allclose(whole["m_worst"], np.abs(vals[:, :, 1] - ms[:, None]).max(1))


@pytest.mark.parametrize("m,nh,snr,R", [(6.0, 10, 40.0, 4000), (2.0, 15, 40.0, 200), (20.0, 15, 20.0, 200), (11.5, 30, 60.0, 1000)])
def test_crlb_restatement_matches_oracle(m, nh, snr, R):
    from deepfmkit_b200 import crlb_sigma_m
    assert abs(crlb_sigma_m(m, nh, snr, R) / orc.crlb_sigma_m(m, nh, snr, R) - 1) < 1e-9


def test_lpsd_option_mapping():
    from deepfmkit_b200.spectra import lpsd_opts
    o = lpsd_opts()
    assert (o.olap, o.bmin, o.lmin, o.jdes, o.kdes, o.order, o.window, o.psll) == (-1.0, 1.0, 0, 500, 100, 0, 0, 200.0)
    o = lpsd_opts(olap=0.5, bmin=2, Lmin=64, Jdes=50, Kdes=10, order=2, win=np.hanning, psll=120)
    assert (o.olap, o.bmin, o.lmin, o.jdes, o.kdes, o.order, o.window, o.psll) == (0.5, 2.0, 64, 50, 10, 2, 1, 120.0)
    assert lpsd_opts(win="kaiser").window == 0 and lpsd_opts(win="hann").window == 1 and lpsd_opts(win=np.kaiser).window == 0
    with pytest.raises(ValueError):
        lpsd_opts(win=np.blackman)


def test_npy_header_reader(tmp_path):
    from deepfmkit_b200.io import _npy_header
    a = (np.arange(24).reshape(6, 4) - 7).astype(np.int16)
    p = str(tmp_path / "a.npy")
    np.save(p, a)
    name, shape, fortran, off = _npy_header(p)
    assert (name, shape, fortran) == ("int16", (6, 4), False)
    assert np.array_equal(np.fromfile(p, dtype=np.int16, offset=off).reshape(shape), a)
    np.save(p, np.asfortranarray(a.astype(np.float32)))
    name, shape, fortran, off = _npy_header(p)
    assert (name, shape, fortran) == ("float32", (6, 4), True)
    np.save(p, a.astype(">i2"))
    with pytest.raises(ValueError):
        _npy_header(p)
    np.save(p, a.astype(np.uint8))
    with pytest.raises(ValueError):
        _npy_header(p)
    with open(p, "wb") as f:
        f.write(b"not numpy")
    with pytest.raises(ValueError):
        _npy_header(p)


def test_raw_object_is_lazy_about_its_frame():
    from deepfmkit_b200 import DeepRawObject
    r = DeepRawObject(data=np.arange(5.0), f_samp=10.0, f_mod=1.0, label="x")
    assert len(r) == 5 and r.device_data is None and r.data.columns.tolist() == ["ch0"]
    r.data = r.data * 2
    assert r.data["ch0"].tolist() == [0.0, 2.0, 4.0, 6.0, 8.0]
    empty = DeepRawObject()
    assert len(empty) == 0 and empty.data is None
