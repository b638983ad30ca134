"""Raw-data ingest: DFMSWPM text records and the binary fast path.

CPU part: the oracle's restatement of pandas' float converter and of parse_header / load_raw against the fixture the
unmodified reference (and pandas itself) minted -- bit for bit -- and the C ABI's host-only header parser.
GPU part (-m gpu): the device parser and the widening kernels through the C ABI, bit for bit against the fixture
and the oracle.
"""
import os

import numpy as np
import pytest

from oracle import ingest_oracle as io_orc


def _bits(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64)).view(np.int64)


def _write(tmp_path, name, data: bytes):
    p = os.path.join(str(tmp_path), name)
    with open(p, "wb") as f:
        f.write(data)
    return p


# ---------------------------------------------------------------------------------------------------------- CPU
def test_oracle_float_converter_matches_pandas_bitwise(golden):
    g = golden("ingest_text")
    got = np.array([io_orc.precise_xstrtod(str(s)) for s in g["tokens"]])
    assert np.array_equal(_bits(got), _bits(g["token_values"]))
    # it is NOT the correctly rounded value: that is the point of restating it
    exact = np.array([float(str(s)) for s in g["tokens"]])
    assert 0.1 < np.mean(got != exact) < 0.4


def test_oracle_load_raw_matches_reference_fixture(golden, tmp_path):
    g = golden("ingest_text")
    for name in g["names"]:
        path = _write(tmp_path, f"{name}.txt", g[f"{name}__bytes"].tobytes())
        hdr, vals = io_orc.load_raw(path)
        ref_h = g[f"{name}__header"]
        assert [hdr["channels"], hdr["t0"], hdr["f_samp"], hdr["f_mod"]] == list(ref_h)
        ref = g[f"{name}__values"].astype(np.float64)
        assert vals.shape == ref.shape and np.array_equal(_bits(vals), _bits(ref)), name


def test_cabi_header_parser(golden, tmp_path):
    from deepfmkit_b200 import parse_header
    g = golden("ingest_text")
    for name in g["names"]:
        raw = g[f"{name}__bytes"].tobytes()
        path = _write(tmp_path, f"{name}.txt", raw)
        hdr = parse_header(path)
        ref = io_orc.parse_header(path)
        assert hdr == ref
        assert [hdr["channels"], hdr["t0"], hdr["f_samp"], hdr["f_mod"]] == list(g[f"{name}__header"])
        assert raw[:hdr["data_offset"]].count(b"\n") == 13
    # the reference keeps only "0-9." of a header line: an exponent loses its 'e', a sign is dropped
    lines = ["% raw_data", "% msg", "% Number of channels: 2", "% Start time: -5", "% Sampling frequency: 2e5",
             "% Modulation frequency: 1000.0 Hz", "%", "%", "%", "%", "%", "%", "ch0 ch1 "]
    path = _write(tmp_path, "quirk.txt", ("\n".join(lines) + "\n1 2 \n").encode())
    hdr = parse_header(path)
    assert hdr["channels"] == 2 and hdr["t0"] == 5 and hdr["f_samp"] == 25.0 and hdr["f_mod"] == 1000.0
    with pytest.raises(RuntimeError):
        parse_header(_write(tmp_path, "short.txt", b"% a\n% b\n"))
    with pytest.raises(RuntimeError):
        parse_header(os.path.join(str(tmp_path), "missing.txt"))
    bad = list(lines)
    bad[2] = "% Number of channels: two"
    with pytest.raises(RuntimeError):  # int('') raises in the reference
        parse_header(_write(tmp_path, "bad.txt", ("\n".join(bad) + "\n").encode()))


def test_oracle_widen():
    a = np.arange(-6, 6, dtype=np.int16).reshape(4, 3)
    w = io_orc.widen(a, 4, 3, True, 0.5, 1.0)
    assert w.shape == (3, 4) and np.array_equal(w[1], 0.5 * np.array([-5, -2, 1, 4]) + 1.0)
    assert np.array_equal(io_orc.widen(a.T.copy(), 4, 3, False), a.T.astype(float))


# ---------------------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def torch_mod():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.mark.gpu
def test_device_parser_matches_reference_fixture(torch_mod, golden, tmp_path):
    from deepfmkit_b200 import load_raw_device
    g = golden("ingest_text")
    for name in g["names"]:
        path = _write(tmp_path, f"{name}.txt", g[f"{name}__bytes"].tobytes())
        data, hdr = load_raw_device(path)
        ref = g[f"{name}__values"].astype(np.float64)
        assert hdr["nbad"] == 0 and tuple(data.shape) == ref.shape, name
        assert np.array_equal(_bits(data.cpu().numpy()), _bits(ref)), name


@pytest.mark.gpu
def test_device_float_converter_matches_pandas_bitwise(torch_mod, golden):
    from deepfmkit_b200 import _lib
    g = golden("ingest_text")
    toks = [str(s) for s in g["tokens"]]
    text = "".join(t + " \n" for t in toks).encode()
    ctx = _lib.get_context(0)
    nrows = ctx.text_load_host(text)
    assert nrows == len(toks)
    out = torch_mod.empty(nrows, dtype=torch_mod.float64, device="cuda")
    assert ctx.text_parse_dev(1, out.data_ptr(), nrows) == 0
    assert np.array_equal(_bits(out.cpu().numpy()), _bits(g["token_values"]))
    ctx.text_release()


@pytest.mark.gpu
def test_device_parser_edge_cases(torch_mod):
    from deepfmkit_b200 import _lib
    ctx = _lib.get_context(0)

    def run(text, ncols, usecols=None):
        n = ctx.text_load_host(text)
        out = torch_mod.full((ncols, max(n, 1)), -7.0, dtype=torch_mod.float64, device="cuda")
        nbad = ctx.text_parse_dev(ncols, out.data_ptr(), max(n, 1), usecols=usecols) if n else 0
        return n, out.cpu().numpy()[:, :n], nbad

    # blank lines, CRLF, tabs, leading blanks, no final newline
    text = b"1.5 2.5\r\n\r\n   \n\t3.5\t4.5 \n\n5.5 6.5"
    n, v, nbad = run(text, 2)
    ref, rbad = io_orc.parse_text(text, 2)
    assert n == 3 and nbad == rbad == 0 and np.array_equal(v, ref)
    # short rows and non-numeric fields become NaN and are counted; extra fields are ignored
    text = b"1 2 3\n4\nx 5 6 7\n8 9e 10\n"
    n, v, nbad = run(text, 3)
    ref, rbad = io_orc.parse_text(text, 3)
    assert n == 4 and nbad == rbad == 4 and np.array_equal(np.isnan(v), np.isnan(ref))
    assert np.array_equal(v[~np.isnan(v)], ref[~np.isnan(ref)])
    # usecols picks file columns
    text = b"".join(b"%d %d %d %d \n" % (i, 10 * i, 100 * i, 1000 * i) for i in range(1, 300))
    n, v, nbad = run(text, 2, usecols=[1, 3])
    assert n == 299 and nbad == 0 and np.array_equal(v[0], 10.0 * np.arange(1, 300)) and np.array_equal(v[1], 1000.0 * np.arange(1, 300))
    # empty and blank-only inputs
    assert run(b"", 1)[0] == 0 and run(b"\n\n  \n", 1)[0] == 0
    with pytest.raises(RuntimeError):
        ctx.text_load_host(b"1 2\n")
        ctx.text_parse_dev(2, 0, 1, usecols=[1, 0])
    ctx.text_release()


@pytest.mark.gpu
def test_device_parser_large_file_spans_many_chunks(torch_mod, tmp_path):
    """~6 MB of text: a hundred 64 kB counting chunks, rows straddling chunk and 16-byte boundaries everywhere."""
    from deepfmkit_b200 import load_raw_device
    rng = np.random.RandomState(3)
    data = rng.randn(120000, 2) * [1.0, 1e-3]
    lines = ["% raw_data", "% m", "% Number of channels: 2", "% Start time: 1", "% Sampling frequency: 200000.0",
             "% Modulation frequency: 1000.0", "%", "%", "%", "%", "%", "%", "ch0 ch1 "]
    body = "".join(f"{float(a)!r} {float(b)!r} \n" for a, b in data)
    path = _write(tmp_path, "big.txt", ("\n".join(lines) + "\n" + body).encode())
    got, hdr = load_raw_device(path)
    assert tuple(got.shape) == (2, 120000) and hdr["nbad"] == 0
    got = got.cpu().numpy()
    # every value against the oracle's restatement of pandas' converter, bit for bit.  (It is not the correctly
    # rounded value: leading zeros of "0.000436..." count towards the 17 digits pandas keeps, so small numbers in
    # fixed notation come back with ~1e-12 relative error -- from pandas and therefore from here.)
    ref = np.array([[io_orc.precise_xstrtod(repr(float(v))) for v in data[:, c]] for c in range(2)])
    assert np.array_equal(_bits(got), _bits(ref))
    assert np.max(np.abs(got - data.T) / np.abs(data.T)) < 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["int16", "int32", "float32", "float64"])
@pytest.mark.parametrize("time_major", [True, False])
def test_binary_ingest(torch_mod, tmp_path, dtype, time_major):
    from deepfmkit_b200 import _lib, load_binary
    rng = np.random.RandomState(11)
    T, C = 70001, 3
    shape = (T, C) if time_major else (C, T)
    a = (rng.randn(*shape) * 2000).astype(dtype)
    scale, offset = 2.5 / 32768.0, 0.25
    ref = io_orc.widen(a, T, C, time_major, scale, offset)
    raws = load_binary(a, 200e3, 1000.0, time_major=time_major, scale=scale, offset=offset)
    got = np.stack([r.device_data.cpu().numpy() for r in raws])
    assert got.shape == (C, T) and np.array_equal(got, ref)
    # the same from a raw file and from a .npy file, streamed in small slabs
    ctx = _lib.get_context(0)
    ctx.set_host_slab_bytes(64 * 1024)
    try:
        p = os.path.join(str(tmp_path), "rec.bin")
        with open(p, "wb") as f:
            f.write(b"HDR!" * 4)
            f.write(a.tobytes())
        raws = load_binary(p, 200e3, 1000.0, channels=C, dtype=dtype, time_major=time_major, scale=scale, offset=offset,
                           byte_offset=16)
        assert np.array_equal(np.stack([r.device_data.cpu().numpy() for r in raws]), ref)
        pn = os.path.join(str(tmp_path), "rec.npy")
        np.save(pn, a)
        raws = load_binary(pn, 200e3, 1000.0, time_major=time_major, scale=scale, offset=offset)
        assert np.array_equal(np.stack([r.device_data.cpu().numpy() for r in raws]), ref)
    finally:
        ctx.set_host_slab_bytes(0)
    assert raws[1].data.columns.tolist() == ["ch1"] and np.array_equal(raws[1].data.values.flatten(), ref[1])


@pytest.mark.gpu
@pytest.mark.parametrize("C", [2, 16, 17, 40])
def test_widen_channel_counts(torch_mod, C):
    """Interleaved records of few channels (thread per time step) and of many (shared-memory tile transpose)."""
    from deepfmkit_b200 import load_binary
    rng = np.random.RandomState(C)
    T = 10007
    a = (rng.randn(T, C) * 3000).astype(np.int16)
    raws = load_binary(a, 200e3, 1000.0, time_major=True, scale=1e-3, offset=-0.5)
    got = np.stack([r.device_data.cpu().numpy() for r in raws])
    assert np.array_equal(got, io_orc.widen(a, T, C, True, 1e-3, -0.5))


@pytest.mark.gpu
def test_facade_load_raw_then_fit(torch_mod, tmp_path):
    """load_raw -> fit on the record where it lies == the fit of the same samples handed over as a pandas frame."""
    from deepfmkit_b200 import DeepFitFramework, DeepRawObject
    from oracle import dfmi_oracle as orc
    x0 = orc.snr_signal(6.0, 200e3, 1000.0, 0.2, 40.0, seed=2)
    x1 = orc.snr_signal(7.5, 200e3, 1000.0, 0.2, 40.0, seed=3, phi0=0.7)
    lines = ["% raw_data", "% m", "% Number of channels: 2", "% Start time: 20210818171519", "% Sampling frequency: 200000.0",
             "% Modulation frequency: 1000.0", "%", "%", "%", "%", "%", "%", "ch0 ch1 "]
    body = "".join(f"{float(a)!r} {float(b)!r} \n" for a, b in zip(x0, x1))
    path = _write(tmp_path, "two.txt", ("\n".join(lines) + "\n" + body).encode())
    dff = DeepFitFramework()
    dff.load_raw(path, labels=["a", "b"])
    assert dff.channr == 2 and dff.t0 == 20210818171519 and dff.f_samp == 200e3 and dff.f_mod == 1000.0
    assert set(dff.raws) == {"a", "b"} and dff.raws["b"].device_data.is_cuda and dff.raws["b"].raw_file == path
    fit_b = dff.fit("b", n=20)
    _, vals = io_orc.load_raw(path)
    dff.load_raw_object(DeepRawObject(data=vals[1], f_samp=200e3, f_mod=1000.0, label="host"))
    fit_h = dff.fit("host", n=20)
    for col in ("amp", "m", "phi", "psi", "dc", "ssq"):
        assert np.array_equal(getattr(fit_b, col), getattr(fit_h, col)), col
    assert np.array_equal(dff.raws["b"].data["ch1"].to_numpy(), vals[1])  # the lazy frame is the parsed record
    e_dev = dff.fit("a", method="ekf", n=20)
    dff.load_raw_object(DeepRawObject(data=vals[0], f_samp=200e3, f_mod=1000.0, label="host0"))
    e_host = dff.fit("host0", method="ekf", n=20)
    assert np.allclose(e_dev.m, e_host.m, rtol=0, atol=1e-12)
