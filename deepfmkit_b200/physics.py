"""Configuration objects of the reference's physics engine (physics.py:20-260), attribute for attribute, so that
experiment factories written for the reference run unchanged.  Only the containers live here: the simulation itself
is the device generator (csrc/dfk_asd.cuh for 'asd' mode, csrc/dfk_synth.cuh for 'snr' mode)."""
from __future__ import annotations

import logging

import numpy as np

SPEED_OF_LIGHT = 299792458.0  # speed of light in m/s, the exact SI value the reference uses


class LaserConfig:
    """Physical properties of the laser source (physics.py:20-60)."""

    def __init__(self, label="laser_source", psi=None):
        self.label = label
        self.wavelength = 1.064e-6
        self.amp = 1.0
        self.visibility = 1.0
        self.f_mod = 1000
        self.df = 3e9
        self.psi = psi if psi else 0.0
        self.waveform_func = lambda t_phase: np.cos(t_phase)
        self.waveform_kwargs = {}
        self.f_n = 0.0
        self.df_n = 0.0
        self.amp_n = 0.0


class InterferometerConfig:
    """Optical path of one interferometer (physics.py:211-236)."""

    def __init__(self, label="interferometer_path"):
        self.label = label
        self.phi = 0.0
        self.ref_arml = 0.1
        self.meas_arml = 0.3
        self.arml_mod_f = 5.0
        self.arml_mod_amp = 0.0
        self.arml_mod_psi = 0.0
        self.arml_mod_n = 0.0


class DFMIObject:
    """A simulation channel: one laser feeding one interferometer (physics.py:238-300)."""

    def __init__(self, label, laser_config, ifo_config, f_samp=200000):
        self.label = label
        self.laser = laser_config
        self.ifo = ifo_config
        self.f_samp = float(f_samp)
        self.N = 0
        self.simtime = None
        self.fit_n = 20
        self.f_fit = float(self.laser.f_mod / self.fit_n)

    @property
    def m(self):
        delta_l = self.ifo.meas_arml - self.ifo.ref_arml
        if delta_l == 0:
            return 0.0
        return 2 * np.pi * self.laser.df * delta_l / SPEED_OF_LIGHT

    def info(self):
        """Log a summary of the channel (physics.py:297-325)."""
        opd = self.ifo.meas_arml - self.ifo.ref_arml
        logging.info(
            f"\nDFMI channel '{self.label}'\n"
            f"  laser '{self.laser.label}': wavelength {self.laser.wavelength * 1e6:.3f} um, f_mod {self.laser.f_mod} Hz, "
            f"df {self.laser.df / 1e9:.3f} GHz, amplitude {self.laser.amp:.2f}, visibility {self.laser.visibility:.2f}\n"
            f"  interferometer '{self.ifo.label}': arms {self.ifo.ref_arml:.4f} / {self.ifo.meas_arml:.4f} m, "
            f"OPD {opd * 100:.2f} cm, motion {self.ifo.arml_mod_amp * 1e9:.2f} nm\n"
            f"  m {self.m:.4f} rad, f_samp {self.f_samp / 1e3:.1f} kHz, fit_n {self.fit_n}, output rate {self.f_fit} Hz, "
            f"simtime {self.simtime if self.simtime else 'N/A'}, N {self.N if self.N > 0 else 'N/A'}\n")


class SignalGenerator:
    """The reference's physics engine by name (physics.py:362-421): ``generate`` returns the records keyed 'main' and
    'witness', produced by the device generators (simulation.py); ``external_noise`` series replace the internally
    drawn noise as in the reference (and are ignored in 'snr' mode, physics.py:418)."""

    def generate(self, main_config, n_seconds, mode="asd", trial_num=0, witness_config=None, snr_db=None,
                 external_noise=None):
        from .simulation import simulate
        if mode == "asd":
            out = simulate(main_config, n_seconds, mode="asd", trial_num=trial_num, witness=witness_config,
                           external_noise=external_noise)
            return {"main": out[0], "witness": out[1]} if witness_config is not None else {"main": out}
        if mode == "snr":
            if snr_db is None:
                logging.error("SNR mode requires a value for 'snr_db'.")
                return {}
            return {"main": simulate(main_config, n_seconds, mode="snr", snr_db=snr_db, trial_num=trial_num)}
        logging.error(f"Unknown simulation mode: '{mode}'")
        return {}
