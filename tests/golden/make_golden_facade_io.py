#!/usr/bin/env python
"""Mint the fixture for the façade's file formats and witness channels by executing the UNMODIFIED reference
(build container only):

    python tests/golden/make_golden_facade_io.py

facade_io.npz holds
  * fitfile_*: the text the reference's own ``DeepFitObject.to_txt`` (data.py:180-213) writes for (a) the first rows
    of the reference's test record ``test/fit_data.txt`` as its ``load_fit`` (core.py:288-332) read them, and (b) a
    fit of a simulated record; beside each text, the arrays and header fields ``load_fit`` returns for it;
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


if __name__ == "__main__":
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        fit_file_cases(out, tmp)
    witness_cases(out)
    path = os.path.join(HERE, "facade_io.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})
