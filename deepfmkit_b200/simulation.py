"""Device-side signal generation behind the reference's ``SignalGenerator.generate`` (physics.py:380-530, 615-722).

``simulate_asd_batch`` produces every trial of a batch in one launch of the 'asd'-mode generator
(csrc/dfk_asd.cuh); ``simulate`` is the one-channel call the facade's ``simulate`` uses.  Internally drawn noise is white only
(``amp_n``, ``df_n``): the coloured sources of the reference (``f_n``, ``arml_mod_n``) come from the third-party
``pyplnoise`` generator, which is not available, and raise here rather than being silently dropped.  Any noise --
coloured laser-frequency and arm-length noise included -- can be handed in as pre-computed series instead
(``external_noise``, the reference engine's own input of that name, physics.py:380-434).
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .physics import SPEED_OF_LIGHT
from .waveforms import harmonic_terms

_MAX_TERMS = 6


class WaveformTables:
    """Host-evaluated waveform rows for callables the device cannot evaluate, de-duplicated by (function, kwargs, psi)."""

    def __init__(self, n_samples, f_samp):
        self.n_samples = int(n_samples)
        self.f_samp = float(f_samp)
        self.rows = []
        self._index = {}

    def row_for(self, laser):
        kwargs = laser.waveform_kwargs or {}
        try:
            key = (id(laser.waveform_func), repr(sorted(kwargs.items())), float(laser.psi), float(laser.f_mod))
        except Exception:
            key = None
        if key is not None and key in self._index:
            return self._index[key]
        t = np.arange(self.n_samples) / self.f_samp
        phase_axis = 2 * np.pi * laser.f_mod * t + laser.psi  # physics.py:663
        g = np.asarray(laser.waveform_func(phase_axis, **kwargs), dtype=np.float64).reshape(-1)
        if g.shape[0] != self.n_samples:
            raise ValueError("waveform function must return one value per sample")
        self.rows.append(g)
        if key is not None:
            self._index[key] = len(self.rows) - 1
        return len(self.rows) - 1

    def array(self):
        return np.stack(self.rows) if self.rows else None


def pack_asd_trial(laser, ifo, f_samp, trial_num, tables: WaveformTables, dynamic=True, noise_row=-1):
    """One DFK_ASD_TRIAL_DOUBLES record (include/dfk_b200.h) from the reference's configuration objects.
    noise_row >= 0: the trial reads that row of the external noise series instead of drawing noise."""
    if noise_row < 0 and (getattr(laser, "f_n", 0.0) != 0.0 or getattr(ifo, "arml_mod_n", 0.0) != 0.0):
        raise NotImplementedError("coloured noise sources (laser.f_n, ifo.arml_mod_n) need pyplnoise, which the reference "
                                  "takes from a third-party package; only the white sources amp_n and df_n are generated")
    rec = np.zeros(_lib.ASD_TRIAL_DOUBLES)
    rec[0] = laser.amp
    rec[1] = laser.visibility
    rec[2] = laser.df
    rec[3] = 2 * np.pi * laser.f_mod
    rec[4] = laser.psi
    rec[5] = 2 * np.pi * ((SPEED_OF_LIGHT / laser.wavelength) + 0.0)  # physics.py:703-704
    rec[6] = ifo.ref_arml / SPEED_OF_LIGHT
    rec[7] = ifo.meas_arml / SPEED_OF_LIGHT
    rec[8] = ifo.phi * laser.wavelength / (2 * np.pi)
    rec[9] = ifo.arml_mod_amp if dynamic else 0.0
    rec[10] = 2 * np.pi * ifo.arml_mod_f
    rec[11] = ifo.arml_mod_psi
    rec[12] = laser.amp_n * np.sqrt(f_samp / 2.0)  # physics.py:591-593
    rec[13] = laser.df_n * np.sqrt(f_samp / 2.0)
    rec[14] = float(int(trial_num))
    terms = harmonic_terms(laser.waveform_func, laser.waveform_kwargs)
    if terms is not None and len(terms) <= _MAX_TERMS:
        rec[15] = len(terms)
        for k, (h, a, p) in enumerate(terms):
            rec[17 + 3 * k: 20 + 3 * k] = (h, a, p)
    else:
        rec[15] = 0
        rec[16] = tables.row_for(laser)
    rec[35] = 1.0 if dynamic else 0.0
    rec[36] = float(noise_row)
    return rec


def _noise_rows(external_noise, n_samples, dev):
    """The four optional series of an ``external_noise`` dictionary as device tensors ``[rows, n_samples]`` (a 1-D
    series is one row; scalars and missing keys mean "no noise from this source", physics.py:432-434)."""
    import torch
    out, rows = [], 0
    for key in _lib.Context.NOISE_KEYS:
        v = external_noise.get(key, 0.0) if external_noise else 0.0
        if v is None or np.isscalar(v):
            if v:
                raise ValueError(f"external noise '{key}' must be a series of one value per sample (or 0)")
            out.append(None)
            continue
        t = v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64))
        t = t.to(device=dev, dtype=torch.float64).reshape(-1, n_samples).contiguous()
        if rows and t.shape[0] != rows:
            raise ValueError("external noise series must have the same number of rows")
        rows = t.shape[0]
        out.append(t)
    return out, rows


def simulate_asd_batch(trials: np.ndarray, n_samples: int, f_samp: float, tables: WaveformTables = None, device=0,
                       with_truth=False, out=None, external_noise=None):
    """All trials in one launch: returns a CUDA tensor ``[n_trials, n_samples]`` (and the ground-truth phase).
    external_noise: ``{'laser_frequency' | 'amplitude' | 'df' | 'armlength': series [rows, n_samples]}``; trials name
    their row in record field 36 (``pack_asd_trial(noise_row=...)``)."""
    import torch
    dev = torch.device("cuda", device)
    ctx = _lib.get_context(device)
    trials = np.ascontiguousarray(trials, dtype=np.float64).reshape(-1, _lib.ASD_TRIAL_DOUBLES)
    J = trials.shape[0]
    with torch.cuda.device(dev):
        td = torch.from_numpy(trials).to(dev)
        tab = tables.array() if tables is not None else None
        tabd = torch.from_numpy(tab).to(dev) if tab is not None else None
        y = out if out is not None else torch.empty((J, n_samples), dtype=torch.float64, device=dev)
        truth = torch.empty((J, n_samples), dtype=torch.float64, device=dev) if with_truth else None
        noise, noise_rows = _noise_rows(external_noise, n_samples, dev)
        if (trials[:, 36] >= noise_rows).any():
            raise ValueError("a trial names an external noise row that was not given")
        ctx.use_torch_stream()
        try:
            kw = dict(tables_ptr=tabd.data_ptr() if tabd is not None else None, ntables=0 if tab is None else tab.shape[0],
                      truth_ptr=truth.data_ptr() if truth is not None else None)
            if noise_rows:
                ctx.synth_asd_noise_dev(td.data_ptr(), J, n_samples, f_samp, y.data_ptr(), y.stride(0) if J else n_samples,
                                        [t.data_ptr() if t is not None else None for t in noise], noise_rows, **kw)
            else:
                ctx.synth_asd_dev(td.data_ptr(), J, n_samples, f_samp, y.data_ptr(), y.stride(0) if J else n_samples, **kw)
        finally:
            ctx.use_default_stream()
        torch.cuda.current_stream(dev).synchronize()  # td / tabd may be freed once we return
    return (y, truth) if with_truth else y


def simulate(sim, n_seconds, mode="asd", snr_db=None, trial_num=0, device=0, witness=None, external_noise=None):
    """One channel: ``SignalGenerator.generate(main_config, n_seconds, mode, trial_num, snr_db=...)['main']``
    (physics.py:380-421) as a ``DeepRawObject`` whose samples live on the device.

    witness: a second ``DFMIObject`` ('asd' mode only, physics.py:458-471).  Returns ``(main, witness)`` then: the
    witness sees the same noise realisation -- same trial number, noise levels taken from the main channel's laser,
    as the reference draws one set of noise arrays from ``main_config`` -- through its own static interferometer.
    Both records come out of one launch.

    external_noise ('asd' mode): pre-computed noise series keyed 'laser_frequency', 'amplitude', 'df', 'armlength'
    (numpy arrays or CUDA tensors, one value per sample) used INSTEAD of internally drawn noise, missing keys meaning
    none (physics.py:430-434) -- the way coloured noise enters.  The witness shares them, except the arm-length one."""
    import torch
    from .core import DeepRawObject
    n = int(n_seconds * sim.f_samp)
    sim.N = n
    if mode == "asd":
        tables = WaveformTables(n, sim.f_samp)
        row = 0 if external_noise else -1
        recs = [pack_asd_trial(sim.laser, sim.ifo, sim.f_samp, trial_num, tables, dynamic=True, noise_row=row)]
        if witness is not None:
            w = pack_asd_trial(witness.laser, witness.ifo, sim.f_samp, trial_num, tables, dynamic=False, noise_row=row)
            w[12:14] = recs[0][12:14]
            recs.append(w)
        y, truth = simulate_asd_batch(np.stack(recs), n, sim.f_samp, tables, device=device, with_truth=True,
                                      external_noise=external_noise)
        raw = DeepRawObject(device_data=y[0], f_samp=sim.f_samp, f_mod=sim.laser.f_mod, label=sim.label, sim=sim)
        raw.phi_sim = truth[0]
        if witness is None:
            return raw
        raw_w = DeepRawObject(device_data=y[1], f_samp=witness.f_samp, f_mod=witness.laser.f_mod, label=witness.label,
                              sim=witness)
        raw_w.phi_sim = truth[1]
        return raw, raw_w
    if mode == "snr":
        if snr_db is None:
            raise ValueError("SNR mode requires a value for 'snr_db'.")
        if witness is not None:  # physics.py:414-415 hands the witness to the 'asd' engine only
            raise ValueError("witness channels are generated in 'asd' mode only")
        ctx = _lib.get_context(device)
        dev = torch.device("cuda", device)
        y = torch.empty(n, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            ctx.use_torch_stream()
            try:
                ctx.synth_snr_dev(y.data_ptr(), n, 1, sim.f_samp, sim.laser.f_mod, sim.m, amp=sim.laser.amp,
                                  visibility=sim.laser.visibility, phi0=sim.ifo.phi, psi0=sim.laser.psi, snr_db=snr_db,
                                  seed=int(trial_num))
            finally:
                ctx.use_default_stream()
        return DeepRawObject(device_data=y, f_samp=sim.f_samp, f_mod=sim.laser.f_mod, label=sim.label, sim=sim, t0=0)
    raise ValueError(f"Unknown simulation mode: '{mode}'")
