"""Configuration objects of the reference's physics engine (physics.py:20-260), attribute for attribute, so that
experiment factories written for the reference run unchanged.  Only the containers live here: the simulation itself
is the device generator (csrc/dfk_asd.cuh for 'asd' mode, csrc/dfk_synth.cuh for 'snr' mode)."""
from __future__ import annotations

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
