"""CPU oracle of the post-fit step.  TEST INFRASTRUCTURE ONLY: imported by tests/, smoke() and the
cpu_baseline legs of bench.py -- never by the product path.

``vectorized_downsample`` restates dsp.py:3-56 (block means of the ground-truth phase at the fit rate).

``lpsd`` restates what ``DeepFitFramework.calc_lpsd`` / ``DeepFitObject.calc_lpsd`` (core.py:590-609, data.py:239-244)
ask of the third-party package ``spectools`` (``from spectools.lpsd import lpsd``; absent from /root/reference, from
this image and from the wheelhouse; the reference pins no version -- it has no requirements file).  **Parity unpinned**:
the routine is restated from the published algorithm the call signature belongs to (``olap, bmin, Lmin, Jdes, Kdes,
order, win, psll``): M. Troebs and G. Heinzel, "Improved spectrum estimation from digitized time series on a
logarithmic frequency axis", Measurement 39 (2006) 120-129, with the LTPDA scheduler (``ltf_plan``: frequencies,
segment lengths, averages from Ndata, fs, olap, bmin, Lmin, Jdes, Kdes) and the Kaiser-window relations of G. Heinzel
et al., "Spectrum and spectral density estimation by the DFT, including a comprehensive list of window functions"
(2002): alpha(PSLL), recommended overlap ROV(alpha).  Anchors in the reference: the call sites above, the defaults of
``DeepFitObject`` (data.py:150-158: olap "default", bmin 1, Lmin 0, Jdes 500, Kdes 100, order 0, win np.kaiser,
psll 200) and the use of the third return value as a one-sided PSD in rad^2/Hz (data.py:262: sqrt(Sxx) is plotted as
"Phase ASD rad/sqrt(Hz)").
"""
from __future__ import annotations

import numpy as np


def vectorized_downsample(signal, R):
    """dsp.py:3-56: trim to a multiple of R, reshape (-1, R), mean along axis 1; empty array when nothing fits."""
    if not isinstance(R, int) or R <= 0:
        return np.array([])
    signal = np.asarray(signal)
    trimmed = (len(signal) // R) * R
    if trimmed == 0:
        return np.array([])
    return signal[:trimmed].reshape(-1, R).mean(axis=1)


def kaiser_alpha(psll):
    """Kaiser alpha for a peak side-lobe level in dB (Heinzel 2002, the fit used by lpsd.c / LTPDA specwin)."""
    x = psll / 100.0
    return ((0.0889732 * x - 0.493285) * x + 4.71469) * x - 0.0821377


def kaiser_rov(alpha):
    """Recommended overlap (fraction) of a Kaiser window (same source)."""
    x = alpha
    return (100.0 - 1.0 / (((4.42204e-05 * x - 0.000925946) * x + 0.00912223) * x + 0.0061076)) / 100.0


def window(name, L, psll=200.0):
    """Periodic ("DFT-even") window of length L: np.kaiser(L + 1, pi * alpha)[:-1] or the Hann window."""
    n = np.arange(L)
    if name == "kaiser":
        return np.kaiser(L + 1, np.pi * kaiser_alpha(psll))[:-1]
    if name == "hann":
        return 0.5 * (1.0 - np.cos(2.0 * np.pi * n / L))
    raise ValueError(name)


def default_overlap(name, psll=200.0):
    return kaiser_rov(kaiser_alpha(psll)) if name == "kaiser" else 0.5


def ltf_plan(ndata, fs, olap, bmin, lmin, jdes, kdes):
    """Frequencies f, resolutions r, (fractional) bins m, segment lengths L and averages K (LTPDA ltf_plan)."""
    xov = 1.0 - olap
    fmin = fs / ndata * bmin
    fmax = fs / 2.0
    fresmin = fs / ndata
    freslim = fresmin * (1.0 + xov * (kdes - 1))
    logfact = (ndata / 2.0) ** (1.0 / jdes) - 1.0
    f, r, b, L, K = [], [], [], [], []
    fi = fmin
    while fi < fmax:
        fres = fi * logfact
        if fres <= freslim:
            fres = np.sqrt(fres * freslim)
        if fres < fresmin:
            fres = fresmin
        fbin = fi / fres
        if fbin < bmin:
            fbin = bmin
            fres = fi / fbin
        dftlen = int(np.floor(fs / fres + 0.5))
        if dftlen > ndata:
            dftlen = ndata
        if dftlen < lmin:
            dftlen = lmin
        nseg = int(np.floor((ndata - dftlen) / (xov * dftlen) + 1.0 + 0.5))
        if nseg == 1:
            dftlen = ndata
        fres = fs / dftlen
        fbin = fi / fres
        f.append(fi)
        r.append(fres)
        b.append(fbin)
        L.append(dftlen)
        K.append(nseg)
        fi = fi + fres
    return (np.array(f), np.array(r), np.array(b), np.array(L, dtype=np.int64), np.array(K, dtype=np.int64))


def segment_starts(ndata, L, K):
    """Start sample of each of the K segments: evenly spread so that the last one ends at the record's end (lpsd.c)."""
    shift = 1.0 if K == 1 else (ndata - L) / (K - 1)
    if shift < 1.0:
        shift = 1.0
    return np.floor(np.arange(K) * shift + 0.5).astype(np.int64)


def detrend(seg, order):
    if order < 0:
        return seg
    if order == 0:
        return seg - seg.mean()
    n = np.arange(len(seg), dtype=np.float64)
    u = n - 0.5 * (len(seg) - 1)
    return seg - np.polyval(np.polyfit(u, seg, order), u)


def lpsd(x, fs, olap="default", bmin=1, Lmin=0, Jdes=500, Kdes=100, order=0, win="kaiser", psll=200.0):
    """Returns (f, ps, psd, enbw, navs): power spectrum, one-sided power spectral density, equivalent noise bandwidth
    and the number of averages per frequency."""
    x = np.asarray(x, dtype=np.float64)
    if olap == "default" or olap is None:
        olap = default_overlap(win, psll)
    f, r, m, L, K = ltf_plan(len(x), fs, olap, bmin, Lmin, Jdes, Kdes)
    ps = np.zeros(len(f))
    psd = np.zeros(len(f))
    enbw = np.zeros(len(f))
    cache = {}
    for j in range(len(f)):
        l = int(L[j])
        if l not in cache:
            cache[l] = window(win, l, psll)
        w = cache[l]
        c = w * np.exp(2j * np.pi * m[j] / l * np.arange(l))
        total = 0.0
        for s in segment_starts(len(x), l, int(K[j])):
            a = np.dot(c, detrend(x[s:s + l], order))
            total += a.real * a.real + a.imag * a.imag
        avg = total / K[j]
        s1, s2 = w.sum(), (w * w).sum()
        ps[j] = 2.0 * avg / (s1 * s1)
        psd[j] = 2.0 * avg / (fs * s2)
        enbw[j] = fs * s2 / (s1 * s1)
    return f, ps, psd, enbw, K
