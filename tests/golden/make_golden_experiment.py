#!/usr/bin/env python
"""Mint the Monte-Carlo experiment fixture by executing the UNMODIFIED reference (build container only).

    python tests/golden/make_golden_experiment.py

experiment_ref.npz holds
  * asd_*: records of the reference's 'asd'-mode physics engine (SignalGenerator.generate, physics.py:423-440,
    615-722) for a handful of noiseless channel configurations -- default cosine, second-harmonic distortion,
    dfm_wave (a waveform only a table can carry), arm-length modulation, non-zero psi / phi -- with the parameters that
    made them; the ground-truth phase beside them;
  * exp_*: the result dictionary of the reference's own ``Experiment.run`` (experiments.py:288-458, process pool) for a
    2-axis x 3-trial study with one stochastic variable, plus ``get_params_for_point`` of two grid points.
The standard factory leaves every noise ASD at zero, so these are deterministic given numpy's global seed.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

core = mg.core
from DeepFMKit import experiments as rexp  # noqa: E402
from DeepFMKit import factories as rfac  # noqa: E402
from DeepFMKit import waveforms as rwave  # noqa: E402

ASD_CASES = [
    # name, m, f_samp, n_seconds, psi, phi, waveform, kwargs, arm amp, arm f, arm psi
    ("cos", 6.0, 200e3, 0.01, 0.0, 0.0, None, {}, 0.0, 5.0, 0.0),
    ("cos_psi_phi", 7.3, 200e3, 0.02, 0.4, 1.1, None, {}, 0.0, 5.0, 0.0),
    ("shd", 5.0, 200e3, 0.01, 0.1, 0.3, "second_harmonic_distortion", {"distortion_amp": 0.08, "distortion_phase": 0.7}, 0.0, 5.0, 0.0),
    ("dfm_wave", 4.0, 200e3, 0.01, 0.0, 0.2, "dfm_wave", {"m": 1.3, "phi": 0.4}, 0.0, 5.0, 0.0),
    ("arm_mod", 6.0, 200e3, 0.05, 0.0, 0.0, None, {}, 2e-7, 40.0, 0.5),
    ("one_mhz", 9.0, 1e6, 0.004, 0.2, 0.0, None, {}, 1e-8, 100.0, 0.0),
]


def asd_record(m, f_samp, n_seconds, psi, phi, wf, kw, aamp, af, apsi):
    laser = core.LaserConfig(psi=psi)
    ifo = core.InterferometerConfig()
    ifo.phi = phi
    ifo.arml_mod_amp, ifo.arml_mod_f, ifo.arml_mod_psi = aamp, af, apsi
    core.set_laser_df_for_effect(laser, ifo, m)
    if wf is not None:
        laser.waveform_func = getattr(rwave, wf)
        laser.waveform_kwargs = dict(kw)
    sim = core.DFMIObject("ch", laser, ifo, f_samp=f_samp)
    raw = core.SignalGenerator().generate(sim, n_seconds, mode="asd", trial_num=0)["main"]
    return raw.data.values.flatten(), np.asarray(raw.phi_sim, dtype=float), laser.df


def phi_generator():
    return np.random.uniform(-1.0, 1.0)


def run_experiment():
    exp = rexp.Experiment(description="golden")
    exp.set_config_factory(rfac.StandardDFMIExperimentFactory(waveform_function=rwave.second_harmonic_distortion, opd_main=0.2))
    exp.add_axis("m_main", np.array([4.0, 6.5, 9.0]))
    exp.add_axis("distortion_amp", np.array([0.0, 0.03]))
    exp.set_static({"distortion_phase": 0.5, "psi": 0.2})
    exp.add_stochastic_variable("phi", phi_generator)
    exp.n_trials = 3
    exp.n_fit_buffers_per_trial = 10
    exp.f_samp = 200000
    exp.add_analysis("nls15", "nls", fitter_kwargs={"ndata": 15, "init_m": 6.0})
    exp.add_analysis("ekf", "ekf", result_cols=["m", "phi"])
    np.random.seed(123)
    res = exp.run(n_cores=4)
    np.random.seed(5)
    p0 = exp.get_params_for_point((0, 1))
    p1 = exp.get_params_for_point((2, 0))
    return res, p0, p1


if __name__ == "__main__":
    out = {}
    names = []
    for name, m, f_samp, n_seconds, psi, phi, wf, kw, aamp, af, apsi in ASD_CASES:
        y, truth, df = asd_record(m, f_samp, n_seconds, psi, phi, wf, kw, aamp, af, apsi)
        out[f"asd_{name}__y"] = y
        out[f"asd_{name}__truth"] = truth
        out[f"asd_{name}__par"] = np.array([m, f_samp, n_seconds, psi, phi, aamp, af, apsi, df])
        out[f"asd_{name}__wf"] = np.array([wf or "", repr(sorted(kw.items()))])
        names.append(name)
        print(name, y.shape, y[:3])
    out["asd_names"] = np.array(names)
    res, p0, p1 = run_experiment()
    for an in ("nls15", "ekf"):
        for col, d in res[an].items():
            for stat, v in d.items():
                out[f"exp_{an}__{col}__{stat}"] = np.asarray(v)
        print(an, sorted(res[an].keys()))
    out["exp_point_0_1"] = np.array([p0["m_main"], p0["distortion_amp"], p0["distortion_phase"], p0["psi"], p0["phi"]])
    out["exp_point_2_0"] = np.array([p1["m_main"], p1["distortion_amp"], p1["distortion_phase"], p1["psi"], p1["phi"]])
    np.savez_compressed(os.path.join(HERE, "experiment_ref.npz"), **out)
    print(os.path.getsize(os.path.join(HERE, "experiment_ref.npz")) // 1024, "KiB")
