"""Batched Monte-Carlo sweep of single-buffer NLS fits (BASELINE config 5).

The reference runs this study one trial per process: ``run_efficiency_trial`` (workers.py:132-189) simulates
one buffer in 'snr' mode with ``trial_num`` as the noise seed and fits it with ``init_m = m_true`` and
``parallel=False``; ``Experiment.run`` (experiments.py:288-458) maps that over a grid x trials with a
``multiprocessing.Pool`` and reduces each grid point to mean / std / min / max / worst-case.  Notebook
``1.1_CRLB-test`` (cells 3-6) compares Var(m_hat) with ``calculate_crlb_for_m`` (helpers.py:16-45).

Here the whole grid is one pass: every realisation is generated in HBM by the counter-based 'snr' generator
(``dfk_synth_snr_dev`` -- statistically, not bitwise, the reference's MT19937 noise), demodulated and fitted by
``dfk_nls_fit_batch_dev`` with a per-realisation cold start, and reduced on the device.  Waves bound the
resident record (default 8 GB).
"""
from __future__ import annotations

import math

import numpy as np

from . import _lib
from . import fit as fit_tunables


def bessel_j(nmax: int, x: float) -> np.ndarray:
    """J_0..J_nmax(x) by Miller's backward recurrence (host copy of csrc/dfk_bessel.cuh, |x| <= 200)."""
    ax = abs(float(x))
    out = np.zeros(nmax + 1)
    if ax == 0.0:
        out[0] = 1.0
        return out
    mstart = max(int(ax + 12.0 * (0.5 * ax) ** (1.0 / 3.0) + 5.0), nmax + 2)
    mstart += mstart & 1
    tox = 2.0 / ax
    bp, b, total = 0.0, 1.0e-200, 0.0
    for k in range(mstart, 0, -1):
        if k <= nmax:
            out[k] = b
        if k % 2 == 0:
            total += 2.0 * b
        bp, b = b, k * tox * b - bp
        if abs(b) > 1e200:
            b *= 1e-200
            bp *= 1e-200
            total *= 1e-200
            out *= 1e-200
    out[0] = b
    out /= total + b
    if x < 0:
        out[1::2] *= -1.0
    return out


def crlb_sigma_m(m_true: float, ndata: int, snr_db: float, buffer_size: int) -> float:
    """Cramer-Rao bound on sigma(m) as the reference defines it (helpers.py:16-45): Fisher matrix = J^T J of the
    harmonic model at [a, m, phi, psi] = [1, m_true, 0, 0], I/Q noise variance 0.5 / SNR / (2 R)."""
    sigma_iq_sq = 0.5 / 10 ** (snr_db / 10.0) / (2 * buffer_size)
    j = np.arange(1, ndata + 1)
    bes = bessel_j(ndata + 1, m_true)
    B, dB = bes[1:ndata + 1], 0.5 * (bes[0:ndata] - bes[2:ndata + 2])
    q, dq = np.cos(j * np.pi / 2.0), np.cos(j * np.pi / 2.0 + np.pi / 2.0)
    # at psi = 0 the sine half of the model vanishes except in the psi column
    jac = np.zeros((2 * ndata, 4))
    jac[:ndata, 0] = q * B
    jac[:ndata, 1] = q * dB
    jac[:ndata, 2] = dq * B
    jac[ndata:, 3] = -(q * B) * j
    try:
        cov = np.linalg.inv(jac.T @ jac)
    except np.linalg.LinAlgError:
        return float("nan")
    return float(math.sqrt(cov[1, 1] * sigma_iq_sq))


def nls_sweep(m_values, n_trials, snr_db=40.0, f_samp=200e3, f_mod=1000.0, n=1, ndata=15, init_a=1.6, init_m=None,
              amp=1.0, visibility=1.0, phi0=0.0, psi0=0.0, seed=0, device=0, max_resident_bytes=8 << 30,
              tunables_from=None, return_rows=False, _seed_stride=None, _raw_stats=False):
    """Fit ``n_trials`` independent single-buffer realisations at every m in ``m_values``.

    init_m: None -> each realisation starts from its true m (workers.py:167-173); a float -> the same cold start
    everywhere (the grid fallback then does the work, as in the reference).
    Returns a dict of per-m arrays: ``m_mean, m_std, m_min, m_max, m_worst`` (largest |m_hat - m|), the same for
    ``amp, phi, psi``, ``ssq_mean``, ``fitok`` fractions ``[len(m), 3]``, ``crlb_sigma_m``, and with return_rows the
    raw ``[len(m), n_trials, 8]`` table.
    """
    import torch
    ms = np.atleast_1d(np.asarray(m_values, dtype=np.float64))
    ctx = _lib.get_context(device)
    dev = torch.device("cuda", device)
    R = int(f_samp / f_mod * n)
    w0 = 2.0 * np.pi * f_mod / f_samp
    opts = fit_tunables.current_lm_opts(tunables_from)
    per_wave = max(1, min(int(n_trials), int(max_resident_bytes // (len(ms) * R * 8))))
    seed_stride = int(n_trials) if _seed_stride is None else int(_seed_stride)
    stats = {k: [] for k in ("sum", "sumsq", "min", "max", "worst", "ok", "ssq")}
    all_rows = [] if return_rows else None
    with torch.cuda.device(dev):
        x = torch.empty((len(ms), per_wave, R), dtype=torch.float64, device=dev)
        rows = torch.empty((len(ms), per_wave, _lib.ROW_STRIDE), dtype=torch.float64, device=dev)
        guess = torch.zeros((len(ms), per_wave, 4), dtype=torch.float64, device=dev)
        guess[:, :, 0] = init_a
        guess[:, :, 1] = torch.from_numpy(ms if init_m is None else np.full_like(ms, float(init_m))).to(dev)[:, None]
        truth = torch.from_numpy(np.stack([np.full_like(ms, amp), ms, np.full_like(ms, phi0), np.full_like(ms, psi0)], 1)).to(dev)
        ctx.use_torch_stream()
        try:
            done = 0
            while done < n_trials:
                nw = min(per_wave, n_trials - done)
                for i, m in enumerate(ms):  # one generator launch per grid point: realisation = "channel" = seed
                    ctx.synth_snr_slab_dev(x[i].data_ptr(), R, nw, R, 0, f_samp, f_mod, float(m), amp, visibility, phi0,
                                           0.0, psi0, snr_db, seed + done + i * seed_stride)
                xv, rv, gv = x[:, :nw], rows[:, :nw], guess[:, :nw]
                if nw != per_wave:
                    xv, rv, gv = xv.contiguous(), rv.contiguous(), gv.contiguous()
                ctx.nls_fit_batch_dev(xv.data_ptr(), len(ms) * nw, 1, R, R, int(ndata), w0, None, gv.data_ptr(), 4, False,
                                      opts, rv.data_ptr())
                p = rv[:, :, :4]
                stats["sum"].append(p.sum(1))
                stats["sumsq"].append(((p - truth[:, None, :]) ** 2).sum(1))
                stats["min"].append(p.min(1).values)
                stats["max"].append(p.max(1).values)
                stats["worst"].append((p - truth[:, None, :]).abs().max(1).values)
                stats["ok"].append(torch.stack([(rv[:, :, 6] == s).sum(1) for s in (0, 1, 2)], 1))
                stats["ssq"].append(rv[:, :, 5].sum(1))
                if return_rows:
                    all_rows.append(rv.cpu().numpy().copy())
                done += nw
            torch.cuda.current_stream(dev).synchronize()
        finally:
            ctx.use_default_stream()
    part = {"n": int(n_trials), "truth": truth.cpu().numpy(),
            "sum": torch.stack(stats["sum"]).sum(0).cpu().numpy(), "sumsq": torch.stack(stats["sumsq"]).sum(0).cpu().numpy(),
            "min": torch.stack(stats["min"]).min(0).values.cpu().numpy(),
            "max": torch.stack(stats["max"]).max(0).values.cpu().numpy(),
            "worst": torch.stack(stats["worst"]).max(0).values.cpu().numpy(),
            "ok": torch.stack(stats["ok"]).sum(0).cpu().numpy().astype(np.float64),
            "ssq": torch.stack(stats["ssq"]).sum(0).cpu().numpy()}
    if _raw_stats:
        return part
    out = _combine_partials(ms, [part], ndata, snr_db, R)
    if return_rows:
        out["rows"] = np.concatenate(all_rows, axis=1)
    return out


def nls_sweep_sharded(m_values, n_trials, group=None, device=None, seed=0, **kwargs):
    """``nls_sweep`` with the realisations split over the ranks of a ``torch.distributed`` group (one process per GPU).

    Realisations are independent, so the split is by realisation index: rank r draws trials
    ``slab_bounds(n_trials, world, r)`` with its own disjoint seed range and nothing crosses the GPUs but the per-m
    sufficient statistics at the end (a few hundred bytes, gathered as Python objects).  Every rank returns the
    combined result; ``return_rows`` is not supported here.
    """
    import torch
    import torch.distributed as dist
    from .sharding import slab_bounds

    if kwargs.pop("return_rows", False):
        raise ValueError("return_rows is a single-GPU option")
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = slab_bounds(int(n_trials), world, rank)
    if device is None:
        device = torch.cuda.current_device()
    ms = np.atleast_1d(np.asarray(m_values, dtype=np.float64))
    part = None
    if hi > lo:
        # the seed of realisation t of grid point i must not depend on the split: nls_sweep uses
        # seed + done + i * n_trials with `done` counting from 0, so shift by lo and keep the per-m stride global
        part = _sweep_partial(ms, hi - lo, seed + lo, int(n_trials), device, **kwargs)
    parts = [None] * world
    dist.all_gather_object(parts, part, group=group)
    return _combine_partials(ms, [p for p in parts if p is not None], kwargs.get("ndata", 15), kwargs.get("snr_db", 40.0),
                             int(kwargs.get("f_samp", 200e3) / kwargs.get("f_mod", 1000.0) * kwargs.get("n", 1)))


def _sweep_partial(ms, n_local, seed0, seed_stride, device, **kwargs):
    """Sufficient statistics of n_local realisations per m (numpy, host)."""
    out = nls_sweep(ms, n_local, seed=seed0, device=device, _seed_stride=seed_stride, _raw_stats=True, **kwargs)
    return out


def _combine_partials(ms, parts, ndata, snr_db, R):
    nt = float(sum(p["n"] for p in parts))
    total = sum(p["sum"] for p in parts)
    sq = sum(p["sumsq"] for p in parts)
    mean = total / nt
    tr = parts[0]["truth"]
    var = np.maximum(sq / nt - (mean - tr) ** 2, 0.0)
    out = {"m_values": ms, "n_trials": int(nt), "fitok": sum(p["ok"] for p in parts) / nt,
           "ssq_mean": sum(p["ssq"] for p in parts) / nt,
           "crlb_sigma_m": np.array([crlb_sigma_m(m, ndata, snr_db, R) for m in ms])}
    mn = np.min([p["min"] for p in parts], axis=0)
    mx = np.max([p["max"] for p in parts], axis=0)
    worst = np.max([p["worst"] for p in parts], axis=0)
    for c, name in enumerate(("amp", "m", "phi", "psi")):
        out[f"{name}_mean"], out[f"{name}_std"] = mean[:, c], np.sqrt(var[:, c])
        out[f"{name}_min"], out[f"{name}_max"], out[f"{name}_worst"] = mn[:, c], mx[:, c], worst[:, c]
    return out
