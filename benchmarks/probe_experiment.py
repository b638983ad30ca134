#!/usr/bin/env python
"""Experiment runner at a realistic size (development tool): a 2-axis sweep x trials with 'asd' white noise."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepfmkit_b200 import Experiment, factories, physics, waveforms  # noqa: E402


class NoisyFactory(factories.StandardDFMIExperimentFactory):
    def __call__(self, params):
        cfg = super().__call__(params)
        cfg["laser_config"].amp_n = 1e-6
        cfg["laser_config"].df_n = 1e3
        return cfg


def main():
    n_m, n_d, n_trials = (int(a) for a in (sys.argv[1:4] or (40, 10, 100)))
    exp = Experiment("probe")
    exp.set_config_factory(NoisyFactory(waveforms.second_harmonic_distortion, opd_main=0.2))
    exp.add_axis("m_main", np.linspace(3, 20, n_m))
    exp.add_axis("distortion_amp", np.linspace(0, 0.05, n_d))
    exp.n_trials = n_trials
    exp.add_analysis("nls", "nls", fitter_kwargs={"ndata": 15})
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = exp.run()
        dt = time.perf_counter() - t0
        trials = n_m * n_d * n_trials
        print(json.dumps({"trials": trials, "samples_per_trial": 2000, "wall_s": dt, "trials_per_s": trials / dt,
                          "m_std_first_point": float(res["nls"]["m"]["std"][0, 0])}), flush=True)


if __name__ == "__main__":
    main()
