"""The façade's file formats and witness channels (the callers either side of the path).

Fixture tests/golden/facade_io.npz, minted by the unmodified reference: ``fit_data`` text written by the reference's
``DeepFitObject.to_txt`` with what its ``load_fit`` reads back; ``create_witness_channel`` results; a main + witness
pair of 'asd'-mode records.  CPU part: ``load_fit`` returns the reference's arrays bit for bit and ``to_txt`` writes
the reference's bytes.  GPU part (-m gpu): witness simulation against the reference's records (gate 2e-8, as for every
'asd' record: tests/test_experiment.py), and a raw record written by ``DeepRawObject.to_txt`` read back bit-identically
by the device text parser.
"""
import numpy as np
import pytest

COLS = ("ssq", "amp", "m", "phi", "psi", "dc")


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _write(tmp_path, g, name):
    path = tmp_path / f"{name}.txt"
    path.write_bytes(g[f"fitfile_{name}_text"].tobytes())
    return str(path)


@pytest.mark.parametrize("name", ["a", "b"])
def test_load_fit_reads_what_the_reference_reads(golden, tmp_path, name):
    import deepfmkit_b200 as dfk
    g = golden("facade_io")
    path = _write(tmp_path, g, name)
    dff = dfk.DeepFitFramework()
    dff.load_fit(path, labels=["x"])
    hdr = g[f"fitfile_{name}_hdr"]
    assert [dff.channr, dff.t0, dff.f_samp, dff.f_mod, dff.n, dff.R, dff.fs] == list(hdr)
    assert isinstance(dff.t0, int) and isinstance(dff.n, int) and isinstance(dff.R, int)
    fit = dff.fits["x"]
    cols = g[f"fitfile_{name}_cols"]
    for c, ref in zip(COLS, cols):
        assert np.array_equal(getattr(fit, c), ref), c  # bit for bit
    assert np.array_equal(fit.time, g[f"fitfile_{name}_time"])
    assert fit.nbuf == cols.shape[1] and fit.label == "x"
    assert (fit.n, fit.R, fit.fs, fit.ndata, fit.init_a, fit.init_m) == (dff.n, dff.R, dff.fs, 10, 1.6, 6.0)
    # default labels: the reference needs a raw file name there; without one the fit file's is used
    dff2 = dfk.DeepFitFramework(fit_file=path)
    assert list(dff2.fits) == [path + "_ch0"]
    fresh = dfk.core.DeepFitObject()
    fresh.fit_file = path
    fresh.parse_header()
    assert [fresh.t0, fresh.f_samp, fresh.f_mod, fresh.n, fresh.R, fresh.fs] == list(hdr[1:])


def test_to_txt_writes_the_reference_bytes(golden, tmp_path):
    import deepfmkit_b200 as dfk
    g = golden("facade_io")
    # (a) load, then write: the framework's defaults for init_a / init_m go into the header as in the reference
    path = _write(tmp_path, g, "a")
    dff = dfk.DeepFitFramework()
    dff.load_fit(path, labels=["t"])
    dff.to_txt(str(tmp_path) + "/out_")
    assert (tmp_path / "out_t.txt").read_bytes() == g["fitfile_a_text"].tobytes()
    dff.to_txt(str(tmp_path) + "/sel_", labels=["t"])
    assert (tmp_path / "sel_t.txt").read_bytes() == g["fitfile_a_text"].tobytes()
    # (b) a fit object as DeepFitFramework.fit leaves it (init_a = init_m = 0, integer t0)
    n, R, fs, init_a, init_m, t0, f_samp, f_mod = g["fitfile_b_scalars"]
    fit = dfk.core.DeepFitObject()
    fit.n, fit.R, fit.fs, fit.init_a, fit.init_m = int(n), int(R), float(fs), int(init_a), int(init_m)
    fit.t0, fit.f_samp, fit.f_mod = int(t0), float(f_samp), int(f_mod)
    for c, ref in zip(COLS, g["fitfile_b_cols"]):
        setattr(fit, c, ref)
    out = tmp_path / "b_out.txt"
    fit.to_txt(str(out))
    assert out.read_bytes() == g["fitfile_b_text"].tobytes()


def test_create_witness_channel_matches_reference(golden):
    import deepfmkit_b200 as dfk
    from deepfmkit_b200 import physics
    g = golden("facade_io")
    df, wavelength, psi, f_mod = g["wit_laser"]
    laser = physics.LaserConfig(psi=psi)
    laser.df, laser.wavelength, laser.f_mod = df, wavelength, f_mod
    ifo = physics.InterferometerConfig()
    ifo.ref_arml, ifo.meas_arml, ifo.phi, ifo.arml_mod_amp, ifo.arml_mod_f, ifo.arml_mod_psi = g["wit_main_ifo"]
    main = physics.DFMIObject("main", laser, ifo, f_samp=200e3)
    main.fit_n = 10
    dff = dfk.DeepFitFramework()
    dff.load_sim(main)
    for (label, kw), ref in zip((("w_default", {}), ("w_m", {"m_witness": 0.07}), ("w_dl", {"delta_l_witness": 2.5e-3})),
                                g["wit_configs"]):
        w = dff.create_witness_channel("main", label, **kw)
        got = [w.ifo.ref_arml, w.ifo.meas_arml, w.ifo.phi, w.m, w.fit_n, w.f_samp, w.ifo.arml_mod_amp]
        assert np.allclose(got, ref, rtol=1e-15, atol=0), (label, got, list(ref))
        assert w.laser is laser and dff.sims[label] is w
    with pytest.raises(KeyError):
        dff.create_witness_channel("nope", "w")
    with pytest.raises(ValueError):
        dff.create_witness_channel("main", "w", m_witness=0.1, delta_l_witness=1e-3)
    laser.df = 0
    with pytest.raises(ValueError):
        dff.create_witness_channel("main", "w")
    assert dff.new_sim("fresh") == "fresh" and dff.sims["fresh"].m > 0
    assert len(dff.new_sim()) == 15  # time-stamp label


def _main_and_witness(g):
    import deepfmkit_b200 as dfk
    from deepfmkit_b200 import physics
    df, wavelength, psi, f_mod = g["wit_laser"]
    laser = physics.LaserConfig(psi=psi)
    laser.df, laser.wavelength, laser.f_mod = df, wavelength, f_mod
    ifo = physics.InterferometerConfig()
    ifo.ref_arml, ifo.meas_arml, ifo.phi, ifo.arml_mod_amp, ifo.arml_mod_f, ifo.arml_mod_psi = g["wit_main_ifo"]
    dff = dfk.DeepFitFramework()
    dff.load_sim(physics.DFMIObject("main", laser, ifo, f_samp=200e3))
    dff.create_witness_channel("main", "w_m", m_witness=0.07)
    return dff


@pytest.mark.gpu
def test_witness_simulation_matches_reference(torch_mod, golden):
    g = golden("facade_io")
    dff = _main_and_witness(g)
    dff.simulate("main", 0.02, mode="asd", witness_label="w_m", trial_num=0)
    assert set(dff.raws) == {"main", "w_m"}
    for key, label in (("main", "main"), ("wit", "w_m")):
        raw = dff.raws[label]
        y = raw.data.values.flatten()
        ref = g[f"wit_{key}_data"]
        assert y.shape == ref.shape and np.max(np.abs(y - ref)) < 2e-8, (label, np.max(np.abs(y - ref)))
        truth = raw.phi_sim.cpu().numpy()
        assert np.max(np.abs(truth - g[f"wit_{key}_phi_sim"])) <= 4e-15 * max(1.0, np.max(np.abs(truth))), label
        assert raw.sim is dff.sims[label] and raw.f_mod == dff.sims[label].laser.f_mod
    # the witness shares the main channel's noise realisation: with white amplitude noise switched on, the two
    # records' deviations from their noise-free selves are the same multiplicative factor
    dff.sims["main"].laser.amp_n = 1e-4
    clean_main, clean_wit = dff.raws["main"].data.values.flatten(), dff.raws["w_m"].data.values.flatten()
    dff.simulate("main", 0.02, mode="asd", witness_label="w_m", trial_num=7)
    rm = dff.raws["main"].data.values.flatten() / clean_main
    rw = dff.raws["w_m"].data.values.flatten() / clean_wit
    assert np.std(rm) > 1e-3 and np.max(np.abs(rm - rw)) < 1e-9
    # the engine under its own name
    from deepfmkit_b200 import physics
    chans = physics.SignalGenerator().generate(dff.sims["main"], 0.02, mode="asd", trial_num=7, witness_config=dff.sims["w_m"])
    assert set(chans) == {"main", "witness"}
    assert np.array_equal(chans["witness"].data.values, dff.raws["w_m"].data.values)
    assert physics.SignalGenerator().generate(dff.sims["main"], 0.02, mode="snr") == {}
    assert physics.SignalGenerator().generate(dff.sims["main"], 0.02, mode="nope") == {}
    dff.sims["main"].info()
    # 'snr' mode generates the main channel only, as the reference's engine does
    dff.raws.clear()
    dff.simulate("main", 0.01, mode="snr", snr_db=30.0, witness_label="w_m")
    assert set(dff.raws) == {"main"}


@pytest.mark.gpu
def test_raw_to_txt_round_trips_through_the_device_parser(torch_mod, tmp_path):
    import deepfmkit_b200 as dfk
    rng = np.random.RandomState(4)
    x = np.concatenate([rng.randn(3000) * 10.0 ** rng.randint(-8, 8, 3000), [0.0, -0.0, 1e-300, 1.7976931348623157e308]])
    raw = dfk.DeepRawObject(data=x, f_samp=200000.0, f_mod=1000.0, label="r", t0=20240131120000)
    path = str(tmp_path / "raw.txt")
    raw.to_txt(path)
    dff = dfk.DeepFitFramework(raw_file=path, raw_labels=["back"])
    back = dff.raws["back"]
    assert (dff.channr, dff.t0, dff.f_samp, dff.f_mod) == (1, 20240131120000, 200000.0, 1000.0)
    got = back.data.values.flatten()
    # the device parser reproduces pandas' reader bit for bit (the oracle restates it); that reader is not correctly
    # rounded -- it can lose ~1e-12 on fixed-notation numbers with leading zeros -- hence the looser gate to the input
    from oracle import ingest_oracle as io_orc
    _, vals = io_orc.load_raw(path)
    assert got.shape == x.shape and np.array_equal(got, np.asarray(vals).reshape(-1))
    assert np.allclose(got, x, rtol=1e-11, atol=0)
    again = dfk.DeepRawObject()
    again.raw_file = path
    again.parse_header()
    assert (again.t0, again.f_samp, again.f_mod) == (20240131120000, 200000.0, 1000.0)


# ------------------------------------------------------------------------ helpers, waveforms, factories, workers
def test_helpers_match_reference(golden):
    from deepfmkit_b200 import helpers, physics
    g = golden("facade_io")
    got = np.array([helpers.calculate_crlb_for_m(m, int(n), s, int(R)) for m, n, s, R in g["help_crlb_in"]])
    assert np.allclose(got, g["help_crlb"], rtol=1e-9, atol=0)
    assert np.allclose([helpers.snr_to_asd(40.0, 200e3), helpers.snr_to_asd(17.5, 1e6)], g["help_asd"], rtol=1e-15)
    for p, ref in zip(([1.3, 6.2, 0.4, -0.2], [0.0, 3.0, 1.0, 0.5]), g["help_jac"]):
        jac = helpers.calculate_jacobian(10, np.array(p))
        assert jac.shape == ref.shape and np.max(np.abs(jac - ref)) < 1e-14
    got = helpers.calculate_m_precision(np.array([2.0, 5.5, 9.0, 14.0]), 12, 35.0)
    assert np.allclose(got, g["help_mprec"], rtol=1e-9, atol=0)
    laser, ifo = physics.LaserConfig(), physics.InterferometerConfig()
    ifo.ref_arml, ifo.meas_arml = 0.25, 0.1
    helpers.set_laser_df_for_effect(laser, ifo, 7.7)
    assert np.isclose(laser.df, g["help_df"][0], rtol=1e-15, atol=0)


def test_waveforms_and_factories_match_reference(golden):
    from deepfmkit_b200 import factories, waveforms
    g = golden("facade_io")
    t = g["help_wave_t"]
    tri = np.stack([waveforms.triangle_wave(t), waveforms.triangle_wave(t, width=0.2), waveforms.triangle_wave(t, width=1.0)])
    sq = np.stack([waveforms.square_wave(t), waveforms.square_wave(t, duty=0.3)])
    assert np.array_equal(tri, g["help_wave_tri"]) and np.array_equal(sq, g["help_wave_sq"])
    assert waveforms.harmonic_terms(waveforms.triangle_wave, {}) is None  # shipped as a table, never guessed at
    fw = factories.StandardWDFMIExperimentFactory(waveforms.second_harmonic_distortion, opd_main=0.25)
    cfg = fw({"m_main": 8.0, "m_witness": 0.09, "psi": 0.2, "phi": 0.6, "distortion_amp": 0.05, "distortion_phase": 0.3})
    l, mi, wi = cfg["laser_config"], cfg["main_ifo_config"], cfg["witness_ifo_config"]
    got = [l.df, l.psi, mi.ref_arml, mi.meas_arml, mi.phi, wi.ref_arml, wi.meas_arml, wi.phi,
           l.waveform_kwargs["distortion_amp"], l.waveform_kwargs["distortion_phase"]]
    assert np.allclose(got, g["help_fac_w"], rtol=1e-15, atol=0)
    w0 = fw({"m_main": 8.0})["witness_ifo_config"]
    assert np.allclose([w0.ref_arml, w0.meas_arml, w0.phi], g["help_fac_w0"], rtol=0, atol=0)
    assert sorted(fw._get_expected_params_keys()) == list(g["help_fac_w_keys"])
    fa = factories.VairableAmplitudeOffset(opd_main=0.15)
    cfga = fa({"m_main": 5.0, "nominal_amplitude": 1.2, "amplitude_offset": -0.15})
    got = [cfga["laser_config"].amp, cfga["laser_config"].df, cfga["main_ifo_config"].ref_arml, cfga["main_ifo_config"].meas_arml]
    assert np.allclose(got, g["help_fac_a"], rtol=1e-15, atol=0)
    assert sorted(fa._get_expected_params_keys()) == list(g["help_fac_a_keys"])
    with pytest.raises(TypeError):
        factories.StandardWDFMIExperimentFactory("not callable")
    with pytest.raises(ValueError):
        factories.VairableAmplitudeOffset(opd_main=0)({"m_main": 1.0, "nominal_amplitude": 1.0, "amplitude_offset": 0.0})


def test_ambiguity_point_matches_reference(golden):
    from deepfmkit_b200 import workers
    g = golden("facade_io")
    i, j, v = workers.calculate_ambiguity_boundary_point({"delta_f": 3e9, "delta_l": 1e-6, "f0": 2.8e14, "grid_i": 3, "grid_j": 4})
    assert (i, j) == (3, 4) and np.isclose(v, g["help_ambiguity"][0], rtol=1e-15, atol=0)
    assert workers.calculate_ambiguity_boundary_point({"delta_f": 0, "delta_l": 1e-6, "f0": 1.0, "grid_i": 0, "grid_j": 0})[2] == float("inf")


@pytest.mark.gpu
def test_trial_workers_match_reference(torch_mod, golden):
    from deepfmkit_b200 import helpers, physics, workers
    g = golden("facade_io")
    for (m, nd, secs, phi), ref in zip(g["help_eff_in"], g["help_eff_m"]):
        laser, ifo = physics.LaserConfig(), physics.InterferometerConfig()
        ifo.phi = phi
        helpers.set_laser_df_for_effect(laser, ifo, m)
        got = workers.run_efficiency_trial({"laser_config": laser, "ifo_config": ifo, "n_seconds": secs, "ndata": int(nd),
                                            "m_true": m, "trial_num": 2})
        assert abs(got - ref) < 1e-7, (m, got, ref)  # records agree to 2e-8 (running-sum order); m follows
    laser, ifo = physics.LaserConfig(psi=0.1), physics.InterferometerConfig()
    ifo.phi = 0.5
    helpers.set_laser_df_for_effect(laser, ifo, 6.0)
    fobj = workers.run_single_trial(laser, ifo, "nls", {"n": 5, "ndata": 12}, n_seconds=0.05, trial_num=1)
    n, R, fs, nbuf, df = g["help_single_scalars"]
    assert (fobj.n, fobj.R, fobj.fs, fobj.nbuf) == (int(n), int(R), fs, int(nbuf)) and laser.df == df
    got = np.stack([fobj.amp, fobj.m, fobj.phi, fobj.psi, fobj.dc, fobj.ssq])
    ref = g["help_single"]
    assert got.shape == ref.shape
    assert np.max(np.abs(got[:5] - ref[:5])) < 1e-7 and np.max(np.abs(got[5] - ref[5])) < 1e-9


@pytest.mark.gpu
def test_experiment_accepts_the_wdfmi_factory_for_gpu_analyses(torch_mod):
    """The W-DFMI factory adds a witness interferometer; 'nls' / 'ekf' analyses read the main channel only, so the
    study gives what the plain DFMI factory gives (experiments.py:46-76)."""
    import deepfmkit_b200 as dfk
    from deepfmkit_b200 import factories, waveforms
    out = []
    for fac in (factories.StandardWDFMIExperimentFactory, factories.StandardDFMIExperimentFactory):
        exp = dfk.Experiment("w")
        exp.set_config_factory(fac(waveforms.second_harmonic_distortion, opd_main=0.2))
        exp.add_axis("m_main", np.array([5.0, 7.0]))
        exp.set_static({"m_witness": 0.1} if fac is factories.StandardWDFMIExperimentFactory else {})
        exp.n_trials = 2
        exp.n_fit_buffers_per_trial = 5
        exp.add_analysis("nls", "nls", fitter_kwargs={"ndata": 12})
        out.append(exp.run())
    a, b = (o["nls"]["m"]["all_trials"] for o in out)
    assert np.array_equal(a, b) and np.all(np.abs(a - np.array([5.0, 7.0])[:, None]) < 1e-2)  # ('asd' physics: m to ~1e-3)
    exp.add_analysis("w", "wdfmi_ortho")
    with pytest.raises(NotImplementedError):
        exp.run()


# ------------------------------------------------------------------ pre-computed noise series (external_noise)
def _external_noise(n):
    rng = np.random.RandomState(11)  # as tests/golden/make_golden_facade_io.py
    return {"laser_frequency": np.cumsum(rng.randn(n)) * 2e3, "amplitude": rng.randn(n) * 1e-3,
            "df": rng.randn(n) * 3e4, "armlength": np.cumsum(rng.randn(n)) * 1e-10}


def test_oracle_external_noise_matches_reference(golden):
    """The oracle's 'asd' physics with all four noise sources injected == the reference engine's records."""
    from oracle import dfmi_oracle as orc
    g = golden("facade_io")
    df, wavelength, psi, f_mod = g["wit_laser"]
    ref_arml, meas_arml, phi, aamp, af, apsi = g["wit_main_ifo"]
    noise = _external_noise(4000)
    common = dict(m_target=0.0, f_samp=200e3, f_mod=f_mod, n_seconds=0.02, psi0=psi, wavelength=wavelength, df=df)
    y, truth = orc.asd_signal(phi0=phi, ref_arml=ref_arml, meas_arml=meas_arml, arml_mod_amp=aamp, arml_mod_f=af,
                              arml_mod_psi=apsi, external_noise=noise, **common)
    assert np.max(np.abs(y - g["ext_main_data"])) < 1e-12 and np.array_equal(truth, g["ext_main_phi_sim"])
    w_ref, w_meas, w_phi = g["wit_configs"][1][:3]
    yw, tw = orc.asd_signal(phi0=w_phi, ref_arml=w_ref, meas_arml=w_meas, dynamic=False, external_noise=noise, **common)
    assert np.max(np.abs(yw - g["ext_witness_data"])) < 1e-12 and np.array_equal(tw, g["ext_witness_phi_sim"])
    yf, _ = orc.asd_signal(phi0=phi, ref_arml=ref_arml, meas_arml=meas_arml, arml_mod_amp=aamp, arml_mod_f=af,
                           arml_mod_psi=apsi, external_noise={"laser_frequency": noise["laser_frequency"]}, **common)
    assert np.max(np.abs(yf - g["ext_onlyf_data"])) < 1e-12
    # the noise matters at the level the gate can see
    assert np.max(np.abs(g["ext_main_data"] - g["wit_main_data"])) > 1e-3


@pytest.mark.gpu
def test_device_external_noise_matches_reference(torch_mod, golden):
    from deepfmkit_b200 import physics
    g = golden("facade_io")
    dff = _main_and_witness(g)
    noise = _external_noise(4000)
    main, wit = dff.sims["main"], dff.sims["w_m"]
    main.laser.amp_n = 5e-3  # internal noise settings are ignored once series are handed in (physics.py:430-434)
    chans = physics.SignalGenerator().generate(main, 0.02, mode="asd", trial_num=0, witness_config=wit, external_noise=noise)
    for key in ("main", "witness"):
        y = chans[key].data.values.flatten()
        dev = np.max(np.abs(y - g[f"ext_{key}_data"]))
        assert dev < 2e-8, (key, dev)
        truth = chans[key].phi_sim.cpu().numpy()
        assert np.max(np.abs(truth - g[f"ext_{key}_phi_sim"])) <= 4e-15 * np.max(np.abs(truth)), key
    # a subset of the sources, CUDA tensors as input
    only_f = {"laser_frequency": torch_mod.from_numpy(noise["laser_frequency"]).cuda()}
    y = physics.SignalGenerator().generate(main, 0.02, mode="asd", external_noise=only_f)["main"].data.values.flatten()
    assert np.max(np.abs(y - g["ext_onlyf_data"])) < 2e-8
    # coloured-noise settings no longer stop a simulation whose noise is handed in ...
    main.laser.f_n = 10.0
    physics.SignalGenerator().generate(main, 0.02, mode="asd", external_noise=noise)
    # ... and still do when it would have to be generated
    with pytest.raises(NotImplementedError):
        physics.SignalGenerator().generate(main, 0.02, mode="asd")
    with pytest.raises(ValueError):
        physics.SignalGenerator().generate(main, 0.02, mode="asd", external_noise={"df": 3.0})
