#!/usr/bin/env python
"""Mint the raw-ingest fixture by executing the UNMODIFIED reference and pandas (build container only).

    python tests/golden/make_golden_ingest.py

ingest_text.npz holds, for several DFMSWPM raw_data files written here (their bytes are stored, they are small):
the header fields ``DeepFitFramework.parse_header`` reads, and every channel as ``DeepFitFramework.load_raw`` returns
it (core.py:129-174, 259-286 -> pandas.read_csv).  Plus ``tokens`` / ``token_values``: number strings in many formats
and what pandas' C parser makes of them (the converter is third-party code, pinned here by its output).
"""
import io
import logging
import os
import sys
import tempfile

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

core = mg.core
logging.disable(logging.CRITICAL)


def header(C, t0="20210818171519", f_samp="200000.0", f_mod="1000.0", names=True):
    lines = ["% raw_data", "% Message goes here", f"% Number of channels: {C}", f"% Start time: {t0}",
             f"% Sampling frequency: {f_samp}", f"% Modulation frequency: {f_mod}", "% n: 0", "% Downsampling factor: 0",
             "% Fit data rate: 0", "% Initial amplitude: 0", "% Initial modulation depth: 0", "%",
             " ".join(f"ch{c}" for c in range(C)) + " "]
    return "\n".join(lines) + "\n"


def body(data, fmt=repr, trailing=True, eol="\n"):
    return "".join(" ".join(fmt(float(v)) for v in row) + (" " if trailing else "") + eol for row in data)


def files():
    rng = np.random.RandomState(7)
    out = {}
    out["one_channel"] = header(1) + body(1.0 + 0.8 * rng.randn(700, 1))
    out["three_channels"] = header(3) + body(rng.randn(300, 3) * [1.0, 1e-6, 1e5])
    out["no_trailing_blank"] = header(2) + body(rng.randn(64, 2), trailing=False)
    out["fixed_and_sci"] = header(2, f_samp="1000000.0", f_mod="1000.0") + \
        body(rng.randn(100, 2), fmt=lambda v: "%.6f" % v) + body(rng.randn(100, 2) * 1e-9, fmt=lambda v: "%.12e" % v)
    out["integers_no_final_newline"] = (header(1, t0="7", f_samp="30000.0", f_mod="400.0") +
                                        body(np.round(rng.randn(50, 1) * 1000), fmt=lambda v: "%d" % v)).rstrip("\n")
    return out


TOKENS = None


def tokens():
    rng = np.random.RandomState(1)
    vals = np.concatenate([rng.randn(4000), rng.randn(1000) * 1e-7, rng.randn(1000) * 1e9, rng.rand(1000) * 3.3,
                           np.array([0.0, 1.0, -1.0, 1e-300, 1e300, 123456789012345678.0, 0.1, 1 / 3, 5e-324,
                                     2.2250738585072014e-308, 1.7976931348623157e308])])
    strs = [repr(float(v)) for v in vals] + ["%.6f" % v for v in vals[:800]] + ["%.10e" % v for v in vals[:800]] + \
           ["%d" % int(v * 1000) for v in vals[:400]] + \
           ["1.", ".5", "-.5e3", "+7", "1e5", "1E-5", "12345678901234567890123",
            "0.000000000000000000001234567890123456789", "00012.5000", "1e-320", "1e-400", "1e400", "-0.0", "0",
            "9007199254740993", "0.30000000000000004", "123456789.123456789123456789"]
    return strs


if __name__ == "__main__":
    out = {}
    names = []
    with tempfile.TemporaryDirectory() as tmp:
        for name, text in files().items():
            path = os.path.join(tmp, name + ".txt")
            with open(path, "w", newline="") as f:
                f.write(text)
            dff = core.DeepFitFramework()
            dff.load_raw(path)
            chans = [dff.raws[f"{path}_ch{c}"].data.values.flatten() for c in range(dff.channr)]
            out[f"{name}__bytes"] = np.frombuffer(text.encode(), dtype=np.uint8)
            out[f"{name}__values"] = np.stack(chans)
            out[f"{name}__header"] = np.array([dff.channr, dff.t0, dff.f_samp, dff.f_mod], dtype=np.float64)
            names.append(name)
            print(name, out[f"{name}__values"].shape, out[f"{name}__header"])
    strs = tokens()
    df = pd.read_csv(io.StringIO("".join(s + " \n" for s in strs)), sep=" ", usecols=[0], names=["ch0"])
    assert df["ch0"].dtype == np.float64 and len(df) == len(strs)
    out["tokens"] = np.array(strs)
    out["token_values"] = df["ch0"].to_numpy()
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "ingest_text.npz"), **out)
    print("pandas", pd.__version__, len(strs), "tokens;", os.path.getsize(os.path.join(HERE, "ingest_text.npz")) // 1024, "KiB")
