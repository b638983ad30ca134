#!/usr/bin/env python
"""Mint the fixture for the façade's file formats and witness channels by executing the UNMODIFIED reference
(build container only):

    python tests/golden/make_golden_facade_io.py

facade_io.npz holds
  * fitfile_*: the text the reference's own ``DeepFitObject.to_txt`` (data.py:180-213) writes for (a) the first rows
    of the reference's test record ``test/fit_data.txt`` as its ``load_fit`` (core.py:288-332) read them, and (b) a
    fit of a simulated record; beside each text, the arrays and header fields ``load_fit`` returns for it;
  * help_*: ``helpers.py`` (CRLB, SNR -> noise density, Jacobian, m precision), the sawtooth / square waveforms, the
    W-DFMI and amplitude-offset factories, ``workers.py`` (ambiguity point, ``run_efficiency_trial`` and
    ``run_single_trial`` on noise-free 'asd' records);
  * wit_*: ``create_witness_channel`` (core.py:519-588) for three ways of asking, and a noise-free main + witness
    pair of records from ``simulate(..., witness_label=...)`` in 'asd' mode with the ground-truth phases.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

core = mg.core
COLS = ("ssq", "amp", "m", "phi", "psi", "dc")


def loaded(path, label):
    dff = core.DeepFitFramework()
    dff.load_fit(path, labels=[label])
    fit = dff.fits[label]
    hdr = np.array([dff.channr, dff.t0, dff.f_samp, dff.f_mod, dff.n, dff.R, dff.fs], dtype=float)
    return fit, hdr


def fit_file_cases(out, tmp):
    # (a) the reference's own test record, first 40 buffers, re-written by the reference's writer
    fit, _ = loaded(os.path.join(mg.REF, "test", "fit_data.txt"), "t")
    for c in COLS:
        setattr(fit, c, getattr(fit, c)[:40])
    pa = os.path.join(tmp, "a.txt")
    fit.to_txt(pa)
    # (b) a fit of a simulated record
    raw = mg.make_raw(m=7.0, f_samp=200e3, n_seconds=0.5, snr_db=30.0, trial=3, phi=0.4)
    dff = core.DeepFitFramework()
    dff.raws["r"] = raw
    raw.t0 = 20240131120000
    fobj = dff.fit("r", method="nls", n=20, parallel=False)
    fobj.t0 = raw.t0
    pb = os.path.join(tmp, "b.txt")
    fobj.to_txt(pb)
    for name, path in (("a", pa), ("b", pb)):
        out[f"fitfile_{name}_text"] = np.frombuffer(open(path, "rb").read(), dtype=np.uint8)
        f2, hdr = loaded(path, "x")
        out[f"fitfile_{name}_hdr"] = hdr
        out[f"fitfile_{name}_cols"] = np.stack([getattr(f2, c) for c in COLS])
        out[f"fitfile_{name}_time"] = f2.time
    out["fitfile_b_scalars"] = np.array([fobj.n, fobj.R, fobj.fs, fobj.init_a, fobj.init_m, fobj.t0, fobj.f_samp, fobj.f_mod],
                                        dtype=float)


def external_noise(n):
    """Deterministic noise series (the tests regenerate them from the same seed)."""
    rng = np.random.RandomState(11)
    return {"laser_frequency": np.cumsum(rng.randn(n)) * 2e3, "amplitude": rng.randn(n) * 1e-3,
            "df": rng.randn(n) * 3e4, "armlength": np.cumsum(rng.randn(n)) * 1e-10}


def witness_cases(out):
    dff = core.DeepFitFramework()
    laser = core.LaserConfig(psi=0.3)
    ifo = core.InterferometerConfig()
    ifo.phi = 0.7
    ifo.arml_mod_amp, ifo.arml_mod_f = 1e-7, 30.0
    core.set_laser_df_for_effect(laser, ifo, 6.5)
    main = core.DFMIObject("main", laser, ifo, f_samp=200e3)
    main.fit_n = 10
    dff.sims["main"] = main
    rows = []
    for label, kw in (("w_default", {}), ("w_m", {"m_witness": 0.07}), ("w_dl", {"delta_l_witness": 2.5e-3})):
        w = dff.create_witness_channel("main", label, **kw)
        rows.append([w.ifo.ref_arml, w.ifo.meas_arml, w.ifo.phi, w.m, w.fit_n, w.f_samp, w.ifo.arml_mod_amp])
    out["wit_configs"] = np.array(rows)
    out["wit_laser"] = np.array([laser.df, laser.wavelength, laser.psi, laser.f_mod])
    out["wit_main_ifo"] = np.array([ifo.ref_arml, ifo.meas_arml, ifo.phi, ifo.arml_mod_amp, ifo.arml_mod_f, ifo.arml_mod_psi])
    dff.simulate("main", 0.02, mode="asd", witness_label="w_m", trial_num=0)
    for key, label in (("main", "main"), ("wit", "w_m")):
        raw = dff.raws[label]
        out[f"wit_{key}_data"] = raw.data.values.flatten()
        out[f"wit_{key}_phi_sim"] = np.asarray(raw.phi_sim, dtype=float)
    # the same pair with pre-computed noise in all four sources (SignalGenerator.generate(external_noise=...)):
    # random-walk laser-frequency and arm-length noise, white amplitude and df noise
    chans = core.SignalGenerator().generate(main, 0.02, mode="asd", trial_num=0, witness_config=dff.sims["w_m"],
                                            external_noise=external_noise(4000))
    for key in ("main", "witness"):
        out[f"ext_{key}_data"] = chans[key].data.values.flatten()
        out[f"ext_{key}_phi_sim"] = np.asarray(chans[key].phi_sim, dtype=float)
    only_f = {"laser_frequency": external_noise(4000)["laser_frequency"]}
    out["ext_onlyf_data"] = core.SignalGenerator().generate(main, 0.02, mode="asd", external_noise=only_f)["main"].data.values.flatten()


def helper_cases(out):
    from DeepFMKit import factories as rfac
    from DeepFMKit import helpers as rh
    from DeepFMKit import waveforms as rw
    from DeepFMKit import workers as rwork
    crlb_in = [(6.0, 10, 40.0, 4000), (2.0, 15, 40.0, 200), (20.0, 15, 30.0, 200), (11.3, 30, 20.0, 1000)]
    out["help_crlb_in"] = np.array(crlb_in)
    out["help_crlb"] = np.array([rh.calculate_crlb_for_m(m, int(n), s, int(R)) for m, n, s, R in crlb_in])
    out["help_asd"] = np.array([rh.snr_to_asd(40.0, 200e3), rh.snr_to_asd(17.5, 1e6)])
    out["help_jac"] = np.stack([rh.calculate_jacobian(10, np.array(p)) for p in ([1.3, 6.2, 0.4, -0.2], [0.0, 3.0, 1.0, 0.5])])
    out["help_mprec"] = rh.calculate_m_precision(np.array([2.0, 5.5, 9.0, 14.0]), 12, 35.0)
    laser, ifo = core.LaserConfig(), core.InterferometerConfig()
    ifo.ref_arml, ifo.meas_arml = 0.25, 0.1
    rh.set_laser_df_for_effect(laser, ifo, 7.7)
    out["help_df"] = np.array([laser.df])
    t = np.linspace(-15.0, 40.0, 5001)
    out["help_wave_t"] = t
    out["help_wave_tri"] = np.stack([rw.triangle_wave(t), rw.triangle_wave(t, width=0.2), rw.triangle_wave(t, width=1.0)])
    out["help_wave_sq"] = np.stack([rw.square_wave(t), rw.square_wave(t, duty=0.3)])
    # factories
    fw = rfac.StandardWDFMIExperimentFactory(rw.second_harmonic_distortion, opd_main=0.25)
    cfg = fw({"m_main": 8.0, "m_witness": 0.09, "psi": 0.2, "phi": 0.6, "distortion_amp": 0.05, "distortion_phase": 0.3})
    l, mi, wi = cfg["laser_config"], cfg["main_ifo_config"], cfg["witness_ifo_config"]
    out["help_fac_w"] = np.array([l.df, l.psi, mi.ref_arml, mi.meas_arml, mi.phi, wi.ref_arml, wi.meas_arml, wi.phi,
                                  l.waveform_kwargs["distortion_amp"], l.waveform_kwargs["distortion_phase"]])
    cfg0 = fw({"m_main": 8.0})
    out["help_fac_w0"] = np.array([cfg0["witness_ifo_config"].ref_arml, cfg0["witness_ifo_config"].meas_arml,
                                   cfg0["witness_ifo_config"].phi])
    out["help_fac_w_keys"] = np.array(sorted(fw._get_expected_params_keys()))
    fa = rfac.VairableAmplitudeOffset(opd_main=0.15)
    cfga = fa({"m_main": 5.0, "nominal_amplitude": 1.2, "amplitude_offset": -0.15})
    out["help_fac_a"] = np.array([cfga["laser_config"].amp, cfga["laser_config"].df, cfga["main_ifo_config"].ref_arml,
                                  cfga["main_ifo_config"].meas_arml])
    out["help_fac_a_keys"] = np.array(sorted(fa._get_expected_params_keys()))
    # workers
    out["help_ambiguity"] = np.array([rwork.calculate_ambiguity_boundary_point(
        {"delta_f": 3e9, "delta_l": 1e-6, "f0": 2.8e14, "grid_i": 3, "grid_j": 4})[2]])
    eff_in = [(4.0, 15, 0.001, 0.3), (9.5, 15, 0.002, 1.2), (17.0, 20, 0.001, -0.8)]
    eff = []
    for m, nd, secs, phi in eff_in:
        laser, ifo = core.LaserConfig(), core.InterferometerConfig()
        ifo.phi = phi
        rh.set_laser_df_for_effect(laser, ifo, m)
        eff.append(rwork.run_efficiency_trial({"laser_config": laser, "ifo_config": ifo, "n_seconds": secs, "ndata": nd,
                                               "m_true": m, "trial_num": 2}))
    out["help_eff_in"] = np.array(eff_in)
    out["help_eff_m"] = np.array(eff)
    laser, ifo = core.LaserConfig(psi=0.1), core.InterferometerConfig()
    ifo.phi = 0.5
    rh.set_laser_df_for_effect(laser, ifo, 6.0)
    fobj = rwork.run_single_trial(laser, ifo, "nls", {"n": 5, "ndata": 12}, n_seconds=0.05, trial_num=1)
    out["help_single"] = np.stack([fobj.amp, fobj.m, fobj.phi, fobj.psi, fobj.dc, fobj.ssq])
    out["help_single_scalars"] = np.array([fobj.n, fobj.R, fobj.fs, fobj.nbuf, laser.df])


if __name__ == "__main__":
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        fit_file_cases(out, tmp)
    witness_cases(out)
    helper_cases(out)
    path = os.path.join(HERE, "facade_io.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})
