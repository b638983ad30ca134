#!/usr/bin/env python
"""Mint the FULL-SIZE parity fixtures of SURVEY 8(d) by executing the UNMODIFIED reference.

    python tests/golden/make_golden_full.py          (build container only; ~10 min on 8 cores)

Companion of make_golden.py (same import shim, same rule: only reference *outputs* are stored).  Inputs are
never stored: every record is a pure function of its seed and is regenerated in the tests by the oracle's
generators, which tests/test_oracle_golden.py pins bit-exactly against the reference's (x_head / x_tail / x_sum
are stored for that).  Fixtures:

  full_cfg1          cfg 1, all 500 buffers (10 s at 200 kHz), sequential + pool (n_cores = 8) schedules, I/Q means
  full_cfg2_first    cfg 2, first 20 s slab (1000 buffers of 20000 samples, trial_num = 0)
  full_cfg2_last     cfg 2, last 20 s slab (trial_num = 179)
  full_cfg3_c{c}     cfg 3, channels c = 0, 37, 128, 255 x first 10 s (phi0 = 2 pi c / 256, trial_num = c)
  full_cfg5          cfg 5, 1000 realisations x 19 integer m (n = 1, ndata = 15, init_m = m_true, parallel=False)
  full_ekf_8ch       cfg 4, 8 channels x 1 s (200 000 steps each) through EKFFitter.fit
  full_ekf_deep      one channel x 20 s (4e6 steps): past t = 16.8 s where |w_m t| > 105615 rad and CUDA's sincos
                     leaves its Cody-Waite path; rows at 5 s (1e6 steps) are rows[:250]
  drift_phi_*        'asd'-mode records whose interferometric phase walks (arm-length modulation), sequential and
                     pool (n_cores = 3 and 8) schedules: the warm-start chain is observable here
"""
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (imports the reference through the shim)

core, rfit, rfitters = mg.core, mg.rfit, mg.rfitters
PAR_CORES = 8


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB", flush=True)


def fingerprint(x):
    return dict(x_head=x[:16].copy(), x_tail=x[-16:].copy(), x_sum=np.array([x.sum(), np.abs(x).sum()]))


def nls_full(name, n=20, nh=10, **sig):
    t0 = time.time()
    raw = mg.make_raw(**sig)
    fitter = rfitters.StandardNLSFitter({"n": n, "ndata": nh})
    seq = mg.rows_of(fitter.fit(raw, parallel=False))
    par = mg.rows_of(fitter.fit(raw, parallel=True, n_cores=PAR_CORES))
    R = int(raw.f_samp / raw.f_mod * n)
    x = raw.data.values.flatten()
    save(name, rows_seq=seq, rows_par=par, qi=mg.qi_of(raw, R, nh, len(seq)),
         meta=np.array([sig.get("m", 6.0), raw.f_samp, raw.f_mod, sig.get("n_seconds", 1.0), sig.get("snr_db", 40.0),
                        sig.get("trial", 0), sig.get("phi", 0.0), sig.get("psi") or 0.0, n, nh, PAR_CORES, 1.6, 6.0, 0.0]),
         **fingerprint(x))
    print(f"  {name}: {len(seq)} buffers in {time.time() - t0:.0f} s", flush=True)


def cfg5_column(m):
    trials = 1000
    rows = np.zeros((trials, 7))
    fitter = rfitters.StandardNLSFitter({"n": 1, "ndata": 15})
    for t in range(trials):
        raw = mg.make_raw(m=float(m), f_samp=200e3, n_seconds=1 / 1000, snr_db=40.0, trial=t)
        rows[t] = mg.rows_of(fitter.fit(raw, parallel=False, init_m=float(m)))[0]
    return rows


def ekf_channel(args):
    c, n_channels, n_seconds = args
    raw = mg.make_raw(m=6.0, f_samp=200e3, n_seconds=n_seconds, snr_db=40.0, trial=c, phi=2 * np.pi * c / n_channels)
    df = rfitters.EKFFitter({"n": 20}).fit(raw, verbose=False)
    x = raw.data.values.flatten()
    return mg.rows_of(df), fingerprint(x)


def dyn_raw(arml_amp, arml_f, n_seconds, m=6.0, amp_n=1e-4, trial=0, f_samp=200e3):
    laser = core.LaserConfig()
    laser.f_mod = 1000
    laser.amp_n = amp_n
    ifo = core.InterferometerConfig()
    ifo.arml_mod_amp = arml_amp
    ifo.arml_mod_f = arml_f
    core.set_laser_df_for_effect(laser, ifo, m)
    sim = core.DFMIObject("ch", laser, ifo, f_samp=f_samp)
    return core.SignalGenerator().generate(sim, n_seconds, mode="asd", trial_num=trial)["main"]


def drift_case(name, arml_amp, arml_f, n_seconds, trial):
    raw = dyn_raw(arml_amp, arml_f, n_seconds, trial=trial)
    x = raw.data.values.flatten()
    fitter = rfitters.StandardNLSFitter({"n": 20, "ndata": 10})
    seq = mg.rows_of(fitter.fit(raw, parallel=False))
    par3 = mg.rows_of(fitter.fit(raw, parallel=True, n_cores=3))
    par8 = mg.rows_of(fitter.fit(raw, parallel=True, n_cores=8))
    save(name, rows_seq=seq, rows_par3=par3, rows_par8=par8, qi=mg.qi_of(raw, 4000, 10, len(seq)),
         phi_sim=np.asarray(raw.phi_sim)[::4000].copy(),
         meta=np.array([6.0, 200e3, 1000.0, n_seconds, trial, 1e-4, arml_amp, arml_f]), **fingerprint(x))


def main():
    only = set(sys.argv[1:])

    def want(name):
        return not only or name in only

    pool = ProcessPoolExecutor(max_workers=6)
    fut_ekf8 = fut_deep = fut_cfg5 = None
    ekf_channels = [0, 1, 511, 1024, 2047, 3000, 4000, 4095]
    if want("ekf"):
        fut_deep = pool.submit(ekf_channel, (777, 4096, 20.0))
        fut_ekf8 = [pool.submit(ekf_channel, (c, 4096, 1.0)) for c in ekf_channels]
    if want("cfg5"):
        fut_cfg5 = [pool.submit(cfg5_column, m) for m in range(2, 21)]

    if want("drift"):
        drift_case("drift_phi_1um", 1e-6, 1.0, 1.0, trial=21)
        drift_case("drift_phi_slow", 3e-7, 0.5, 2.0, trial=22)
    if want("cfg1"):
        nls_full("full_cfg1", m=6.0, f_samp=200e3, n_seconds=10.0, snr_db=40.0, trial=0)
    if want("cfg3"):
        for c in (0, 37, 128, 255):
            nls_full(f"full_cfg3_c{c}", m=6.0, f_samp=200e3, n_seconds=10.0, snr_db=40.0, trial=c,
                     phi=2 * np.pi * c / 256)
    if want("cfg2"):
        nls_full("full_cfg2_first", m=6.0, f_samp=1e6, n_seconds=20.0, snr_db=40.0, trial=0)
        nls_full("full_cfg2_last", m=6.0, f_samp=1e6, n_seconds=20.0, snr_db=40.0, trial=179)

    if fut_cfg5:
        rows = np.stack([f.result() for f in fut_cfg5])
        save("full_cfg5", rows=rows, ms=np.arange(2, 21, dtype=float), trials=np.array(1000))
    if fut_ekf8:
        res = [f.result() for f in fut_ekf8]
        save("full_ekf_8ch", rows=np.stack([r[0] for r in res]), channels=np.array(ekf_channels, dtype=float),
             n_channels=np.array(4096.0), x_head=np.stack([r[1]["x_head"] for r in res]),
             x_sum=np.stack([r[1]["x_sum"] for r in res]),
             meta=np.array([6.0, 200e3, 1000.0, 1.0, 40.0, 20]))
    if fut_deep:
        rows, fp = fut_deep.result()
        save("full_ekf_deep", rows=rows, channel=np.array(777.0), n_channels=np.array(4096.0),
             meta=np.array([6.0, 200e3, 1000.0, 20.0, 40.0, 20]), **fp)
    pool.shutdown()


if __name__ == "__main__":
    main()
