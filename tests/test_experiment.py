"""Batched Monte-Carlo experiments: the 'asd'-mode generator on the device and the Experiment runner.

CPU part: the oracle's 'asd' physics against records minted by the reference's own SignalGenerator; the runner's job
enumeration, parameter packing and ``get_params_for_point`` against the reference's Experiment (fixture
tests/golden/experiment_ref.npz).  GPU part (-m gpu): device records against the reference's, the full
``Experiment.run`` against the reference's result dictionary, white-noise statistics, the trial-statistics kernel.
Gates: records within 2e-8 (the reference's own running sum of a ~1e6 rad phase carries ~1e-9 of rounding, and the
device sums in a different order); fitted parameters of converged fits within 1e-8; identical flags.
"""
import numpy as np
import pytest

from oracle import dfmi_oracle as orc

REC_TOL = 2e-8


def _laser_ifo(par, wf):
    """Configuration objects of this package for a fixture case."""
    from deepfmkit_b200 import physics, waveforms
    m, f_samp, n_seconds, psi, phi, aamp, af, apsi, df = par
    laser = physics.LaserConfig(psi=psi)
    ifo = physics.InterferometerConfig()
    ifo.phi = phi
    ifo.arml_mod_amp, ifo.arml_mod_f, ifo.arml_mod_psi = aamp, af, apsi
    laser.df = df
    name, kw = str(wf[0]), eval(str(wf[1]))  # noqa: S307 - our own fixture
    if name:
        laser.waveform_func = getattr(waveforms, name)
        laser.waveform_kwargs = dict(kw)
    return laser, ifo, f_samp, n_seconds


# ---------------------------------------------------------------------------------------------------------- CPU
def test_oracle_asd_matches_reference_records(golden):
    g = golden("experiment_ref")
    for name in g["asd_names"]:
        par = g[f"asd_{name}__par"]
        if str(g[f"asd_{name}__wf"][0]):
            continue  # the oracle restates the default waveform only
        m, f_samp, n_seconds, psi, phi, aamp, af, apsi, df = par
        y, truth = orc.asd_signal(m, f_samp, 1000.0, n_seconds, arml_mod_amp=aamp, arml_mod_f=af, arml_mod_psi=apsi, phi0=phi,
                                  psi0=psi)
        assert np.array_equal(y, g[f"asd_{name}__y"]), name
        assert np.array_equal(truth, g[f"asd_{name}__truth"]), name


def test_waveform_recognition_and_packing(golden):
    from deepfmkit_b200 import _lib, physics, waveforms
    from deepfmkit_b200.simulation import WaveformTables, pack_asd_trial
    assert waveforms.harmonic_terms(physics.LaserConfig().waveform_func, {}) == [(1.0, 1.0, 0.0)]
    assert waveforms.harmonic_terms(lambda t: np.sin(t), {}) is None
    assert waveforms.harmonic_terms(waveforms.second_harmonic_distortion, {"distortion_amp": 0.1, "distortion_phase": 0.3}) == \
        [(1.0, 1.0, 0.0), (2.0, 0.1, 0.3)]
    assert waveforms.harmonic_terms(waveforms.dfm_like_wave, {}) == [(1.0, 1.0, 0.0), (2.0, 0.1, 0.0), (3.0, 0.05, 0.0)]
    assert waveforms.harmonic_terms(waveforms.dfm_wave, {"m": 1.0}) is None
    g = golden("experiment_ref")
    laser, ifo, f_samp, n_seconds = _laser_ifo(g["asd_dfm_wave__par"], g["asd_dfm_wave__wf"])
    tables = WaveformTables(int(n_seconds * f_samp), f_samp)
    rec = pack_asd_trial(laser, ifo, f_samp, 7, tables)
    assert rec.shape == (_lib.ASD_TRIAL_DOUBLES,) and rec[15] == 0 and rec[16] == 0 and rec[14] == 7
    assert pack_asd_trial(laser, ifo, f_samp, 8, tables)[16] == 0 and len(tables.rows) == 1  # de-duplicated
    t = np.arange(int(n_seconds * f_samp)) / f_samp
    assert np.array_equal(tables.rows[0], np.cos(0.4 + 1.3 * np.cos(2 * np.pi * 1000 * t + laser.psi)))
    laser.f_n = 1.0
    with pytest.raises(NotImplementedError):
        pack_asd_trial(laser, ifo, f_samp, 0, tables)


def _phi_generator():
    return np.random.uniform(-1.0, 1.0)


def _make_experiment():
    from deepfmkit_b200 import Experiment, factories, waveforms
    exp = Experiment(description="golden")
    exp.set_config_factory(factories.StandardDFMIExperimentFactory(waveform_function=waveforms.second_harmonic_distortion,
                                                                   opd_main=0.2))
    exp.add_axis("m_main", np.array([4.0, 6.5, 9.0]))
    exp.add_axis("distortion_amp", np.array([0.0, 0.03]))
    exp.set_static({"distortion_phase": 0.5, "psi": 0.2})
    exp.add_stochastic_variable("phi", _phi_generator)
    exp.n_trials = 3
    exp.n_fit_buffers_per_trial = 10
    exp.f_samp = 200000
    exp.add_analysis("nls15", "nls", fitter_kwargs={"ndata": 15, "init_m": 6.0})
    exp.add_analysis("ekf", "ekf", result_cols=["m", "phi"])
    return exp


def test_experiment_configuration_mirrors_reference(golden):
    g = golden("experiment_ref")
    exp = _make_experiment()
    np.random.seed(5)
    state = np.random.get_state()[1].copy()
    p0, p1 = exp.get_params_for_point((0, 1)), exp.get_params_for_point((2, 0))
    assert np.array_equal(np.random.get_state()[1], state)  # the global random state is left alone
    assert [p0[k] for k in ("m_main", "distortion_amp", "distortion_phase", "psi", "phi")] == list(g["exp_point_0_1"])
    assert [p1[k] for k in ("m_main", "distortion_amp", "distortion_phase", "psi", "phi")] == list(g["exp_point_2_0"])
    with pytest.raises(ValueError):
        exp.add_axis("not_a_parameter", [1, 2])
    with pytest.raises(ValueError):
        exp.get_params_for_point((0,))
    with pytest.raises(TypeError):
        exp.set_config_factory(lambda p: p)
    # job list: product order over the axes, trial numbers counting up, one generator call per job in that order
    np.random.seed(123)
    jobs = list(exp._jobs())
    assert [j[2] for j in jobs] == list(range(18)) and jobs[0][0] == (0, 0) and jobs[3][0] == (0, 1) and jobs[-1][0] == (2, 1)
    np.random.seed(123)
    assert [j[3]["phi"] for j in jobs] == [np.random.uniform(-1.0, 1.0) for _ in range(18)]
    with pytest.raises(RuntimeError):
        exp.save_results("/tmp/never.pkl")


# ---------------------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def torch_mod():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.mark.gpu
def test_device_asd_records_match_reference(torch_mod, golden):
    from deepfmkit_b200 import _lib
    from deepfmkit_b200.simulation import WaveformTables, pack_asd_trial, simulate_asd_batch
    g = golden("experiment_ref")
    worst = 0.0
    for name in g["asd_names"]:
        laser, ifo, f_samp, n_seconds = _laser_ifo(g[f"asd_{name}__par"], g[f"asd_{name}__wf"])
        n = int(n_seconds * f_samp)
        tables = WaveformTables(n, f_samp)
        rec = np.stack([pack_asd_trial(laser, ifo, f_samp, t, tables) for t in range(3)])  # three identical trials
        y, truth = simulate_asd_batch(rec, n, f_samp, tables, with_truth=True)
        y, truth = y.cpu().numpy(), truth.cpu().numpy()
        ref = g[f"asd_{name}__y"]
        assert y.shape == (3, len(ref)) and np.array_equal(y[0], y[1]) and np.array_equal(y[0], y[2])
        dev = np.max(np.abs(y[0] - ref))
        worst = max(worst, dev)
        assert dev < REC_TOL, (name, dev)
        assert np.max(np.abs(truth[0] - g[f"asd_{name}__truth"])) <= 1e-15 * np.max(np.abs(truth[0])) * 4, name
    print("max record deviation", worst)


@pytest.mark.gpu
def test_experiment_run_matches_reference(torch_mod, golden, tmp_path):
    g = golden("experiment_ref")
    exp = _make_experiment()
    np.random.seed(123)
    res = exp.run(n_cores=4, filename=str(tmp_path / "res.pkl"))
    assert set(res) == {"axes", "nls15", "ekf"} and sorted(res["nls15"]) == sorted(["amp", "dc", "fitok", "m", "phi", "psi", "ssq", "tau"])
    assert sorted(res["ekf"]) == ["m", "phi"]
    flags = g["exp_nls15__fitok__all_trials"]
    assert np.array_equal(res["nls15"]["fitok"]["all_trials"], flags)
    ok = flags < 2
    for col in ("amp", "m", "phi", "psi", "dc", "tau"):
        got, ref = res["nls15"][col]["all_trials"], g[f"exp_nls15__{col}__all_trials"]
        assert got.shape == ref.shape == (3, 2, 3)
        scale = np.maximum(np.abs(ref[ok]), 1.0) if col in ("amp", "m", "dc") else 1.0
        if col == "tau":
            scale = np.abs(ref[ok])
        assert np.max(np.abs(got[ok] - ref[ok]) / scale) < 1e-8, col
        # fits the reference itself flags as failed (fitok 2) are path-dependent -- the grid fallback can land in a
        # mirror solution (phi + pi) -- so only their bulk is held to the reference: most agree to ~1e-9 anyway
        close = np.abs(got[~ok] - ref[~ok]) / np.maximum(np.abs(ref[~ok]), 1e-9 if col == "tau" else 1.0) < 1e-6
        assert close.mean() >= 0.8, (col, close)
    for stat in ("mean", "std", "min", "max", "worst"):
        got, ref = res["nls15"]["m"][stat], g[f"exp_nls15__m__{stat}"]
        assert got.shape == ref.shape == (3, 2)
        assert np.max(np.abs(got[:2, 0] - ref[:2, 0])) < 1e-8, stat  # the grid points whose fits converge
    for col in ("m", "phi"):
        assert np.max(np.abs(res["ekf"][col]["all_trials"] - g[f"exp_ekf__{col}__all_trials"])) < 1e-6, col
    from deepfmkit_b200 import Experiment
    again = Experiment(filename=str(tmp_path / "res.pkl"))
    assert np.array_equal(again.results["nls15"]["m"]["mean"], res["nls15"]["m"]["mean"])


@pytest.mark.gpu
def test_trial_stats_kernel_matches_numpy(torch_mod):
    from deepfmkit_b200 import _lib
    rng = np.random.RandomState(0)
    P, T, C = 7, 1000, 5
    v = rng.randn(P, T, C) * [1.0, 1e-3, 5.0, 1.0, 1.0] + [0.0, 6.0, 0.0, 1e3, 0.0]
    v[rng.rand(P, T, C) < 0.02] = np.nan
    v[3, :, 2] = np.nan  # a column with no finite trial
    vd = torch_mod.from_numpy(v).cuda()
    out = torch_mod.empty((P, C, 6), dtype=torch_mod.float64, device="cuda")
    ctx = _lib.get_context(0)
    ctx.trial_stats_dev(vd.data_ptr(), P, T, C, C, out.data_ptr())
    ctx.synchronize()
    s = out.cpu().numpy()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mean, std = np.nanmean(v, axis=1), np.nanstd(v, axis=1)
        mn, mx = np.nanmin(v, axis=1), np.nanmax(v, axis=1)
    good = np.isfinite(mean)
    assert np.allclose(s[..., 0][good], mean[good], rtol=1e-13, atol=1e-15) and np.all(np.isnan(s[..., 0][~good]))
    assert np.allclose(s[..., 1][good], std[good], rtol=1e-12)
    assert np.array_equal(s[..., 2][good], mn[good]) and np.array_equal(s[..., 3][good], mx[good])
    dev = np.abs(v - mean[:, None, :])
    for p in range(P):
        for c in range(C):
            if good[p, c]:
                assert s[p, c, 4] == v[p, np.nanargmax(dev[p, :, c]), c]
                assert s[p, c, 5] == np.sum(np.isfinite(v[p, :, c]))
    # "worst" measured from a given centre (the true value in nls_sweep) instead of the mean
    cen = rng.randn(P, C)
    ctx.trial_stats_dev(vd.data_ptr(), P, T, C, C, out.data_ptr(), center_ptr=torch_mod.from_numpy(cen).cuda().data_ptr())
    ctx.synchronize()
    s2 = out.cpu().numpy()
    assert np.array_equal(s2[..., :4], s[..., :4], equal_nan=True)
    for p in range(P):
        for c in range(C):
            if good[p, c]:
                assert s2[p, c, 4] == v[p, np.nanargmax(np.abs(v[p, :, c] - cen[p, c])), c]


@pytest.mark.gpu
def test_white_noise_sources_have_the_reference_statistics(torch_mod):
    """amp_n and df_n: sigma = asd * sqrt(fs / 2) per sample (physics.py:591-597).  Amplitude noise is read back from
    the record directly; modulation-depth noise through the scatter it gives the fitted m, compared with the oracle's
    MT19937 realisations of the same physics."""
    from deepfmkit_b200 import nls_fit_batch, physics
    from deepfmkit_b200.simulation import WaveformTables, pack_asd_trial, simulate_asd_batch
    f_samp, n = 200e3, 4000
    laser, ifo = physics.LaserConfig(), physics.InterferometerConfig()
    laser.df = orc.laser_df(6.0, ifo.ref_arml, ifo.meas_arml)
    tables = WaveformTables(n, f_samp)
    clean = simulate_asd_batch(pack_asd_trial(laser, ifo, f_samp, 0, tables)[None], n, f_samp, tables).cpu().numpy()[0]
    laser.amp_n = 1e-5
    recs = np.stack([pack_asd_trial(laser, ifo, f_samp, t, tables) for t in range(64)])
    y = simulate_asd_batch(recs, n, f_samp, tables).cpu().numpy()
    z = (y / clean - 1.0) / (laser.amp_n * np.sqrt(f_samp / 2))  # (A + n) / A - 1 with A = 1
    assert abs(z.std() - 1.0) < 0.01 and abs(z.mean()) < 0.01
    assert abs(np.corrcoef(z[0], z[1])[0, 1]) < 0.05 and abs(np.corrcoef(z[0, :-1], z[0, 1:])[0, 1]) < 0.05
    laser.amp_n, laser.df_n = 0.0, 2e4
    ntr = 400
    recs = np.stack([pack_asd_trial(laser, ifo, f_samp, 1000 + t, tables) for t in range(ntr)])
    yd = simulate_asd_batch(recs, n, f_samp, tables)
    m_dev = nls_fit_batch(yd, f_samp, 1000.0, 20, seeded=False)[:, 0, 1]
    m_ref = []
    for t in range(120):
        sig, _ = orc.asd_signal(6.0, f_samp, 1000.0, n / f_samp, trial=t, df_n=laser.df_n)
        m_ref.append(orc.nls_fit(sig, f_samp, 1000.0, 20, 10, schedule="seq")[0, 1])
    m_ref = np.array(m_ref)
    assert abs(m_dev.mean() - 6.0) < 5 * m_dev.std() / np.sqrt(ntr) + 1e-6
    assert 0.75 < m_dev.std() / m_ref.std() < 1.3, (m_dev.std(), m_ref.std())


@pytest.mark.gpu
def test_facade_simulate_then_fit(torch_mod):
    """new channel -> simulate ('asd' and 'snr') -> fit, all on the device (core.py:176-243, 424-517)."""
    from deepfmkit_b200 import DeepFitFramework, physics
    dff = DeepFitFramework()
    laser, ifo = physics.LaserConfig(), physics.InterferometerConfig()
    laser.df = orc.laser_df(6.0, ifo.ref_arml, ifo.meas_arml)
    dff.load_sim(physics.DFMIObject("main", laser, ifo, f_samp=200e3))
    dff.simulate("main", n_seconds=0.5, mode="asd")
    raw = dff.raws["main"]
    assert raw.device_data.is_cuda and len(raw) == 100000 and raw.phi_sim.shape[0] == 100000
    fit = dff.fit("main")  # n from sims['main'].fit_n = 20
    truth = raw.phi_sim_downsamp  # core.py:480-481: the ground-truth phase brought to the fit rate
    assert truth.shape[0] == 25 and np.allclose(truth.cpu().numpy(), raw.phi_sim.cpu().numpy().reshape(25, -1).mean(1), rtol=1e-14)
    assert fit.nbuf == 25 and np.all(np.abs(fit.m - 6.0) < 1e-3) and np.allclose(fit.tau, fit.m / (2 * np.pi * laser.df))
    y_ref, _ = orc.asd_signal(6.0, 200e3, 1000.0, 0.5)
    assert np.max(np.abs(raw.data["ch0"].to_numpy() - y_ref)) < REC_TOL
    ref = orc.nls_fit(y_ref, 200e3, 1000.0, 20, 10, schedule="gpu")  # the reference's fit of the reference's record
    assert np.max(np.abs(np.stack([fit.amp, fit.m, fit.phi, fit.psi], 1) - ref[:, :4])) < 1e-8
    dff.simulate("main", n_seconds=0.2, mode="snr", snr_db=40.0, trial_num=3)
    fit = dff.fit("main", fit_label="snr")
    assert fit.nbuf == 10 and np.all(np.abs(fit.m - 6.0) < 0.01)
    dff.simulate("main", n_seconds=0.2, mode="snr")  # logs, stores nothing new
    dff.simulate("nope", n_seconds=0.2)
