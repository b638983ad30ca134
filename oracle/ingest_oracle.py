"""CPU oracle of the raw-data ingest step.  TEST INFRASTRUCTURE ONLY: imported by tests/, smoke() and
the cpu_baseline legs of bench.py -- never by the product path.

Restates ``DeepFitFramework.parse_header`` (core.py:129-174) and ``load_raw`` (core.py:259-286).  ``load_raw`` is
``pd.read_csv(raw_file, sep=' ', skiprows=13, usecols=[c], names=['ch<c>'])`` per channel, so the arithmetic lives in
a third-party dependency absent from /root/reference: pandas (3.0.2 in this image; the reference pins nothing).  Its
C parser converts fields with ``precise_xstrtod`` (pandas/_libs/src/parser/tokenizer.c, the default since pandas 1.2,
``float_precision=None`` == 'high'): at most 17 significant digits accumulated as ``number = number * 10 + digit`` in
fp64, then one multiplication or division by a tabulated power of ten.  It is not correctly rounded (24 % of
``repr(float)`` strings come back 1 ulp off), so ``float(token)`` would be the wrong oracle; the routine is restated
here and pinned bit for bit against pandas and against the reference's own ``load_raw`` by tests/golden/ingest_text.npz
(minted by tests/golden/make_golden_ingest.py).
"""
from __future__ import annotations

import numpy as np

_POW10 = [float("1e%d" % i) for i in range(309)]
_BLANK = " \t\r"


def parse_header(path):
    """core.py:129-174 for file_select='raw': of lines 2..5 keep the characters '1234567890.', read int, int, float,
    float.  Also returns the byte offset of the first data row (after the 13 lines read_csv skips)."""
    with open(path, "rb") as f:
        raw = f.read(1 << 16)
    lines = raw.split(b"\n")
    values = ["".join(ch for ch in lines[v].decode("latin1") if ch in "1234567890.") for v in range(2, 6)]
    off = 0
    for _ in range(13):
        nxt = raw.find(b"\n", off)
        if nxt < 0:
            off = len(raw)
            break
        off = nxt + 1
    return {"channels": int(values[0]), "t0": int(values[1]), "f_samp": float(values[2]), "f_mod": float(values[3]),
            "data_offset": off}


def precise_xstrtod(s: str) -> float:
    """pandas' default float converter (tokenizer.c precise_xstrtod), decimal '.', exponent 'e'/'E'.
    Returns NaN when the token is not a number in full."""
    p, n = 0, len(s)
    neg = False
    if p < n and s[p] in "+-":
        neg = s[p] == "-"
        p += 1
    number, exponent, num_digits, num_decimals = 0.0, 0, 0, 0
    while p < n and "0" <= s[p] <= "9":
        if num_digits < 17:
            number = number * 10.0 + (ord(s[p]) - 48)
            num_digits += 1
        else:
            exponent += 1
        p += 1
    if p < n and s[p] == ".":
        p += 1
        while num_digits < 17 and p < n and "0" <= s[p] <= "9":
            number = number * 10.0 + (ord(s[p]) - 48)
            p += 1
            num_digits += 1
            num_decimals += 1
        if num_digits >= 17:
            while p < n and "0" <= s[p] <= "9":
                p += 1
        exponent -= num_decimals
    if num_digits == 0:
        return float("nan")
    if neg:
        number = -number
    if p < n and s[p] in "eE":
        q = p + 1
        eneg = False
        if q < n and s[q] in "+-":
            eneg = s[q] == "-"
            q += 1
        e, nd = 0, 0
        while q < n and "0" <= s[q] <= "9":
            e = e * 10 + (ord(s[q]) - 48)
            nd += 1
            q += 1
        if nd:
            exponent += -e if eneg else e
            p = q
    if p != n:
        return float("nan")
    if exponent > 308:
        return float("-inf") if neg else float("inf")
    if exponent > 0:
        return number * _POW10[exponent]
    if exponent < -308:
        if exponent < -616:
            return -0.0 if neg else 0.0
        return number / _POW10[-308 - exponent] / _POW10[308]
    return number / _POW10[-exponent]


def parse_text(text: bytes, ncols: int, usecols=None):
    """Rows of a blank-separated text region -> (values[ncols, nrows], nbad).  Lines holding only blanks are skipped;
    missing or non-numeric fields are NaN (counted in nbad)."""
    usecols = list(range(ncols)) if usecols is None else list(usecols)
    rows = []
    nbad = 0
    for line in text.split(b"\n"):
        fields = line.decode("latin1").replace("\t", " ").replace("\r", " ").split()
        if not fields:
            continue
        row = []
        for c in usecols:
            v = precise_xstrtod(fields[c]) if c < len(fields) else float("nan")
            if v != v:
                nbad += 1
            row.append(v)
        rows.append(row)
    out = np.array(rows, dtype=np.float64).reshape(len(rows), ncols).T.copy()
    return out, nbad


def load_raw(path):
    """core.py:259-286 restated: header + every channel.  Returns (header dict, values[channels, nrows])."""
    hdr = parse_header(path)
    with open(path, "rb") as f:
        f.seek(hdr["data_offset"])
        text = f.read()
    vals, _ = parse_text(text, hdr["channels"])
    return hdr, vals


def widen(src, T, C, time_major, scale=1.0, offset=0.0):
    """Binary record -> fp64 channel-major [C, T]:  scale * sample + offset (one fused multiply-add, as the device)."""
    a = np.asarray(src).reshape((T, C) if time_major else (C, T)).astype(np.float64)
    a = a.T if time_major else a
    # fma(scale, x, offset): exact product in extended precision where it matters (|x| < 2^31, scale arbitrary)
    prod = a.astype(np.longdouble) * np.longdouble(scale) + np.longdouble(offset)
    return np.ascontiguousarray(prod.astype(np.float64))
