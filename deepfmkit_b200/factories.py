"""Experiment configuration factories (factories.py:9-113 of the reference): the abstract base the reference asks
users to subclass, and the standard DFMI factory.  Objects of the reference's own classes are accepted wherever these
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
