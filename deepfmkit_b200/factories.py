"""Experiment configuration factories (factories.py:9-225 of the reference): the abstract base the reference asks
users to subclass, the standard DFMI and W-DFMI factories and the amplitude-offset example.  Objects of the reference's own classes are accepted wherever these
are (they are read by attribute)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Callable, Set

import numpy as np

from . import physics


class ExperimentFactory(ABC):
    @abstractmethod
    def __call__(self, params: dict) -> dict:
        """Physics configurations of one trial: ``{'laser_config': ..., 'main_ifo_config': ...}``."""

    @abstractmethod
    def _get_expected_params_keys(self) -> Set[str]:
        """Names of the parameters ``__call__`` reads (used to validate axes and static parameters)."""


class StandardDFMIExperimentFactory(ExperimentFactory):
    """factories.py:49-113: m_main sets the laser's modulation amplitude for a given optical path difference."""

    def __init__(self, waveform_function: Callable, opd_main: float = 0.1):
        if not callable(waveform_function):
            raise TypeError("waveform_function must be a callable.")
        self.waveform_func_to_use = waveform_function
        self.opd_main = opd_main

    def _get_expected_params_keys(self) -> Set[str]:
        return {"m_main", "psi", "phi", "distortion_amp", "distortion_phase", "waveform_kwargs"}

    def __call__(self, params: dict) -> dict:
        m_main = params["m_main"]
        waveform_kwargs = {"distortion_amp": params.get("distortion_amp", 0.0),
                           "distortion_phase": params.get("distortion_phase", 0.0)}
        laser_config = physics.LaserConfig()
        laser_config.psi = params.get("psi", 0)
        main_ifo_config = physics.InterferometerConfig(label="main_ifo")
        main_ifo_config.ref_arml = 0.1
        main_ifo_config.meas_arml = main_ifo_config.ref_arml + self.opd_main
        main_ifo_config.phi = params.get("phi", 0)
        laser_config.waveform_func = self.waveform_func_to_use
        laser_config.waveform_kwargs = waveform_kwargs
        laser_config.df = (m_main * physics.SPEED_OF_LIGHT) / (2 * np.pi * self.opd_main)
        return {"laser_config": laser_config, "main_ifo_config": main_ifo_config}


class StandardWDFMIExperimentFactory(ExperimentFactory):
    """factories.py:113-180: as the DFMI factory, plus a static witness interferometer on the same laser whose path
    difference gives ``m_witness`` and whose phase offset puts it at mid-fringe."""

    def __init__(self, waveform_function: Callable, opd_main: float = 0.2):
        if not callable(waveform_function):
            raise TypeError("waveform_function must be a callable.")
        self.waveform_func_to_use = waveform_function
        self.opd_main = opd_main

    def _get_expected_params_keys(self) -> Set[str]:
        return {"m_main", "m_witness", "psi", "phi", "distortion_amp", "distortion_phase", "waveform_kwargs"}

    def __call__(self, params: dict) -> dict:
        out = StandardDFMIExperimentFactory(self.waveform_func_to_use, self.opd_main)(params)
        laser_config = out["laser_config"]
        m_witness = params.get("m_witness", 0.0)
        witness_ifo_config = physics.InterferometerConfig(label="witness_ifo")
        if laser_config.df > 0 and m_witness > 0:
            opd_witness = (m_witness * physics.SPEED_OF_LIGHT) / (2 * np.pi * laser_config.df)
            witness_ifo_config.ref_arml = 0.01
            witness_ifo_config.meas_arml = witness_ifo_config.ref_arml + opd_witness
            f0 = physics.SPEED_OF_LIGHT / laser_config.wavelength
            witness_ifo_config.phi = (np.pi / 2.0) - (2 * np.pi * f0 * opd_witness) / physics.SPEED_OF_LIGHT
        out["witness_ifo_config"] = witness_ifo_config
        return out


class VairableAmplitudeOffset(ExperimentFactory):
    """factories.py:188-225 (the reference's spelling): signal amplitude = nominal + offset at a given m_main."""

    def __init__(self, opd_main: float = 0.1):
        self.opd_main = opd_main

    def _get_expected_params_keys(self) -> Set[str]:
        return {"m_main", "nominal_amplitude", "amplitude_offset", "waveform_kwargs"}

    def __call__(self, params: dict) -> dict:
        laser_config = physics.LaserConfig(label="ExperimentLaser")
        laser_config.amp = params["nominal_amplitude"] + params["amplitude_offset"]
        if self.opd_main == 0:
            raise ValueError("opd_main cannot be zero in the factory.")
        laser_config.df = (params["m_main"] * physics.SPEED_OF_LIGHT) / (2 * np.pi * self.opd_main)
        main_ifo_config = physics.InterferometerConfig(label="main_ifo")
        main_ifo_config.ref_arml = 0.1
        main_ifo_config.meas_arml = main_ifo_config.ref_arml + self.opd_main
        return {"laser_config": laser_config, "main_ifo_config": main_ifo_config}
