"""The façade call of the reference, ``DeepFitFramework.fit`` (core.py:424-517), over the GPU fitters.

Only the fitting entry point and the two containers it touches are provided: simulation, file I/O,
plotting and the experimental W-DFMI fitters stay with the reference (INTEGRATION.md shows how the
reference's own ``DeepFitFramework`` is pointed at these fitters instead).  The call signature, the
default ``fit_label``, the ``tau`` column, the error behaviour (log + ``None``) and the scalar fields
of the returned fit object are those of the reference.
"""
from __future__ import annotations

import logging

import numpy as np

from .fitters import EKFFitter, StandardNLSFitter


def _reference_fitters():
    """The experimental W-DFMI / HW-DFMI fitters of the reference (bound to a general-purpose optimiser, outside this package), under
    the method names core.py:452-459 gives them -- available when the reference itself is importable as ``DeepFMKit``."""
    try:
        from DeepFMKit import fitters as ref
    except Exception:
        return {}
    names = {"wdfmi_ortho": "WDFMI_OrthogonalFitter", "wdfmi_nls": "WDFMI_NLSFitter", "wdfmi_seq": "WDFMI_SequentialFitter",
             "hwdfmi": "HWDFMI_Fitter"}
    return {method: getattr(ref, cls) for method, cls in names.items() if hasattr(ref, cls)}


class DeepRawObject:
    """One channel of raw data: what the fitters read (data.py:16-118 holds the full container).

    ``data`` is the one-column pandas frame of the reference.  A record that already lives on the GPU (io.load_raw,
    io.load_binary, the device generator) is carried as ``device_data`` -- a 1-D float64 CUDA tensor the fitters read
    in place -- and the frame is built from it only when ``data`` is asked for."""

    def __init__(self, data=None, f_samp=None, f_mod=None, label=None, sim=None, t0=0, device_data=None, column="ch0"):
        import pandas as pd
        if data is not None and not hasattr(data, "columns"):
            data = pd.DataFrame(np.asarray(data, dtype=np.float64).reshape(-1), columns=[column])
        self._data = data
        self.device_data = device_data
        self._column = column
        self.f_samp = f_samp
        self.f_mod = f_mod
        self.label = label
        self.sim = sim
        self.t0 = t0
        self.phi_sim = None
        self.raw_file = None

    @property
    def data(self):
        if self._data is None and self.device_data is not None:
            import pandas as pd
            self._data = pd.DataFrame(self.device_data.detach().cpu().numpy().reshape(-1), columns=[self._column])
        return self._data

    @data.setter
    def data(self, value):
        self._data = value
        self.device_data = None  # the frame is the truth once somebody assigns it

    def __len__(self):
        if self.device_data is not None:
            return int(self.device_data.shape[0])
        return 0 if self._data is None else int(self._data.shape[0])


class DeepFitObject:
    """Result arrays + fit geometry, the fields core.py:372-386 fills."""

    def __init__(self):
        self.label = None
        self.n = self.R = self.fs = self.nbuf = None
        self.ndata = self.init_a = self.init_m = 0
        self.t0 = 0
        self.f_samp = self.f_mod = None
        self.ssq = self.amp = self.m = self.tau = self.phi = self.psi = self.dc = self.time = None
        # spectral-estimate settings and results (data.py:148-158)
        self.f = None
        self.Sxx = None
        self.olap = "default"
        self.bmin = 1
        self.Lmin = 0
        self.Jdes = 500
        self.Kdes = 100
        self.order = 0
        self.win = np.kaiser
        self.psll = 200

    def calc_lpsd(self):
        """LPSD of the fitted interferometric phase (data.py:239-244), on the device."""
        from .spectra import lpsd
        self.f, _, self.Sxx, _, _, _ = lpsd(self.phi, self.fs, self.olap, self.bmin, self.Lmin, self.Jdes, self.Kdes,
                                            self.order, self.win, self.psll, return_type="legacy")


class DeepFitFramework:
    def __init__(self):
        self.raw_file = None
        self.sims = {}
        self.raws = {}
        self.fits = {}
        self.fits_df = {}

    def parse_header(self, file_select="raw"):
        """Header of ``self.raw_file`` (core.py:129-174, raw files only: fit files stay with the reference)."""
        if file_select != "raw" or getattr(self, "raw_file", None) is None:
            logging.error("No files specified !!")
            return
        from .io import parse_header
        hdr = parse_header(self.raw_file)
        self.channr, self.t0, self.f_samp, self.f_mod = hdr["channels"], hdr["t0"], hdr["f_samp"], hdr["f_mod"]
        logging.info("Number of channels: {}".format(self.channr))
        logging.info("Starting time: {}".format(self.t0))
        logging.info("Sampling frequency: {}".format(self.f_samp))
        logging.info("Modulation frequency: {}".format(self.f_mod))

    def load_raw(self, raw_file=None, labels=None):
        """Load a raw_data file (core.py:259-286): parsed on the GPU, one ``DeepRawObject`` per channel whose samples
        stay on the device for the fitters."""
        from .io import load_raw
        if raw_file is not None:
            self.raw_file = raw_file
        if getattr(self, "raw_file", None) is None:
            logging.error("No raw file specified !!")
            return
        self.parse_header(file_select="raw")
        for raw in load_raw(self.raw_file, labels=labels):
            self.raws[raw.label] = raw

    def load_sim(self, sim):
        self.sims[sim.label] = sim

    def simulate(self, main_label, n_seconds, mode="asd", witness_label=None, snr_db=None, trial_num=0, verbose=False):
        """Simulate the channel ``self.sims[main_label]`` on the GPU (core.py:176-243 -> SignalGenerator.generate) and
        store the record in ``self.raws``.  Witness channels belong to the W-DFMI fitters and are not generated."""
        import time
        from .simulation import simulate
        t0 = time.time()
        if main_label not in self.sims:
            logging.error(f"Main simulation label '{main_label}' not found!")
            return
        if witness_label:
            logging.error("Witness channels are outside this package (W-DFMI stays with the reference).")
            return
        if mode == "snr" and snr_db is None:
            logging.error("SNR mode requires a value for 'snr_db'.")
            logging.error("Simulation failed to generate data.")
            return
        if mode not in ("asd", "snr"):
            logging.error(f"Unknown simulation mode: '{mode}'")
            logging.error("Simulation failed to generate data.")
            return
        main_config = self.sims[main_label]
        raw = simulate(main_config, n_seconds, mode=mode, snr_db=snr_db, trial_num=trial_num)
        self.raws[raw.label] = raw
        main_config.simtime = time.time() - t0

    def load_raw_object(self, raw: DeepRawObject, label=None):
        label = label or raw.label
        raw.label = label
        self.raws[label] = raw
        return raw

    def calc_lpsd(self, labels=None):
        """LPSD of phi for the fits under ``labels``, or for every fit (core.py:590-609)."""
        if labels is not None:
            for label in labels:
                try:
                    self.fits[label].calc_lpsd()
                except KeyError:
                    logging.warning("Specified label is invalid!")
        else:
            for fit in self.fits.values():
                fit.calc_lpsd()

    def fit_init(self, label, n):
        """(R, fs, nbuf) of a fitting run (core.py:390-422)."""
        raw = self.raws[label]
        R = int(raw.f_samp / raw.f_mod * n)
        fs = raw.f_samp / R
        nbuf = int(len(raw) / R)
        if nbuf == 0:
            logging.error("Check buffer size !! Calculated nbuf is zero.")
        return R, fs, nbuf

    def fit(self, main_label, method="nls", fit_label=None, **kwargs):
        fitter_map = {"nls": StandardNLSFitter, "ekf": EKFFitter}
        fitter_map.update(_reference_fitters())  # the W-DFMI family stays with the reference (core.py:452-459)
        if method not in fitter_map:
            logging.error(f"Unknown fit method: '{method}'. Available: {list(fitter_map.keys())}")
            return
        FitterClass = fitter_map[method]
        if main_label not in self.raws:
            logging.error(f"Invalid raw data label: '{main_label}' !!")
            return
        main_raw = self.raws[main_label]
        if fit_label is None:
            fit_label = f"{main_label}_{method}"
        n_cycles = kwargs.get("n")
        if n_cycles is None:
            sim = getattr(main_raw, "sim", None)
            sim_obj = self.sims.get(sim.label if sim else main_label)
            n_cycles = sim_obj.fit_n if sim_obj else 20
        R, fs, nbuf = self.fit_init(main_label, n_cycles)

        phi_sim = getattr(main_raw, "phi_sim", None)
        if phi_sim is not None and len(phi_sim) > 0:  # core.py:480-481: ground truth at the fit rate
            from .spectra import vectorized_downsample
            main_raw.phi_sim_downsamp = vectorized_downsample(phi_sim, R)

        fit_config = {"n": n_cycles}
        fitter_args = {"main_raw": main_raw}
        if "wdfmi" in method:  # core.py:492-496 (reference fitters only)
            witness_label = kwargs.get("witness_label")
            if not witness_label or witness_label not in self.raws:
                logging.error(f"W-DFMI method '{method}' requires a valid 'witness_label'.")
                return
            fitter_args["witness_raw"] = self.raws[witness_label]
        fitter = FitterClass(fit_config)
        results_df = fitter.fit(**fitter_args, **kwargs)
        if results_df is None or results_df.empty:
            logging.error(f"{FitterClass.__name__} returned no results.")
            return None
        sim = getattr(main_raw, "sim", None)
        results_df["tau"] = results_df["m"] / (2 * np.pi * sim.laser.df) if sim else 0.0  # core.py:506-507
        self.fits_df[fit_label] = results_df

        fit = DeepFitObject()
        fit.n, fit.R, fit.fs, fit.nbuf = n_cycles, R, fs, nbuf
        fit.ndata, fit.init_a, fit.init_m = fit_config.get("ndata", 0), 0, 0  # Q5: the reference stores zeros
        fit.t0 = getattr(main_raw, "t0", 0)
        fit.f_samp, fit.f_mod = main_raw.f_samp, main_raw.f_mod
        for col in ("ssq", "amp", "m", "tau", "phi", "psi", "dc"):
            setattr(fit, col, np.asarray(results_df[col], dtype=np.float64) if col != "tau" or sim
                    else np.zeros(len(results_df)))
        fit.time = np.arange(0, fit.ssq.shape[0] / fit.fs, 1.0 / fit.fs)
        fit.label = fit_label
        self.fits[fit_label] = fit
        return fit
