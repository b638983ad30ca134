"""The trial workers of the reference (workers.py:8-189) over the GPU façade: configure, simulate, fit.

In the reference these are the functions a ``multiprocessing.Pool`` maps over a parameter grid, one process per
trial.  Here each call runs its one trial on the device; a whole study belongs in ``Experiment.run`` or ``nls_sweep``,
which do every trial in one batched pass."""
from __future__ import annotations

from typing import Optional

import numpy as np

from .core import DeepFitFramework
from .physics import SPEED_OF_LIGHT, DFMIObject


def calculate_ambiguity_boundary_point(params):
    """|coarse phase error| = 2 pi (delta_l / c) (f0 / delta_f) at one grid point (workers.py:8-42)."""
    delta_f, delta_l, f0 = params["delta_f"], params["delta_l"], params["f0"]
    grid_i, grid_j = params["grid_i"], params["grid_j"]
    if delta_f == 0:
        return (grid_i, grid_j, float("inf"))
    return (grid_i, grid_j, np.abs(-2 * np.pi * (delta_l / SPEED_OF_LIGHT) * (f0 / delta_f)))


def run_single_trial(laser_config, main_ifo_config, fitter_method: str, fitter_kwargs: Optional[dict] = None,
                     witness_ifo_config=None, n_seconds: Optional[float] = None, trial_num: int = 0):
    """One configure-simulate-fit trial (workers.py:44-130): a fresh façade, channel ``main_trial`` (and
    ``witness_trial`` on the same laser), an 'asd'-mode record of ``n_seconds`` (one buffer by default), the fit.
    Returns the ``DeepFitObject`` or None."""
    if fitter_kwargs is None:
        fitter_kwargs = {}
    dff = DeepFitFramework()
    main_label = "main_trial"
    main_channel = DFMIObject(main_label, laser_config, main_ifo_config)
    dff.sims[main_label] = main_channel
    witness_label = None
    if witness_ifo_config:
        witness_label = "witness_trial"
        dff.sims[witness_label] = DFMIObject(witness_label, laser_config, witness_ifo_config)
    if n_seconds is None:
        n_seconds = fitter_kwargs.get("n", main_channel.fit_n) / laser_config.f_mod
    dff.simulate(main_label, n_seconds=n_seconds, witness_label=witness_label, trial_num=trial_num)
    if "wdfmi" in fitter_method:
        fitter_kwargs["witness_label"] = witness_label
    fitter_kwargs["verbose"] = False
    return dff.fit(main_label, method=fitter_method, **fitter_kwargs)


def run_efficiency_trial(params: dict) -> float:
    """Fitted m of one single-buffer NLS trial started at the true m (workers.py:132-189); NaN if the fit failed."""
    laser_config = params["laser_config"]
    n_seconds = params["n_seconds"]
    fitter_kwargs = {"n": int(laser_config.f_mod * n_seconds), "ndata": params["ndata"], "init_m": params["m_true"],
                     "parallel": False}
    fit_obj = run_single_trial(laser_config=laser_config, main_ifo_config=params["ifo_config"], fitter_method="nls",
                               fitter_kwargs=fitter_kwargs, n_seconds=n_seconds, trial_num=params["trial_num"])
    if fit_obj and fit_obj.m.size > 0:
        return fit_obj.m[0]
    return np.nan
