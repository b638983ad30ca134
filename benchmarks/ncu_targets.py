#!/usr/bin/env python
"""One launch of every kernel family at a representative size, for `ncu --set full -k regex:...` (development tool).

    python benchmarks/ncu_targets.py [fold ekf sweep period long post text asd widen]

Kernels that only read their big input run at full size (the cfg-2 record); generators write at most ~2 GB so that
ncu's save/restore between replay passes stays cheap.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepfmkit_b200 import _lib, lpsd, nls_sweep, vectorized_downsample  # noqa: E402
from deepfmkit_b200 import fit as tun  # noqa: E402


def main():
    what = sys.argv[1:] or ["fold", "ekf", "sweep", "period", "long", "post", "text", "asd", "widen", "tm"]
    ctx = _lib.get_context(0)
    opts = tun.current_lm_opts()
    dev = "cuda"
    if "fold" in what or "post" in what or "long" in what:
        T = 3_600_000_000 if "fold" in what else 1_000_000_000
        x = torch.empty(T, dtype=torch.float64, device=dev)
        for off in range(0, T, 200_000_000):  # generated in 1.6 GB slabs
            n = min(200_000_000, T - off)
            ctx.synth_snr_slab_dev(x.data_ptr() + off * 8, n, 1, n, off, 1e6, 1000.0, 6.0, seed=1)
        if "fold" in what:
            nbuf = T // 20000
            rows = torch.empty((nbuf, 8), dtype=torch.float64, device=dev)
            ctx.nls_fit_dev(x.data_ptr(), nbuf, 20000, 10, 2 * np.pi * 1e-3, [1.6, 6.0, 0, 0], 16, opts, rows.data_ptr())
            torch.cuda.synchronize()
        if "long" in what:
            P, n, N = 10000, 20, 10
            nbuf = min(T, 1_000_000_000) // (P * n)
            qi = torch.empty((nbuf, 2 * N), dtype=torch.float64, device=dev)
            dc = torch.empty(nbuf, dtype=torch.float64, device=dev)
            ctx.demod(x.data_ptr(), nbuf, P * n, N, 2 * np.pi / P, qi.data_ptr(), dc.data_ptr())
            torch.cuda.synchronize()
        if "post" in what:
            y = vectorized_downsample(x[:1_000_000_000], 20000)
            phi = torch.from_numpy(np.cumsum(np.random.RandomState(0).randn(180000)) * 1e-3).to(dev)
            lpsd(phi, 50.0)
            torch.cuda.synchronize()
            del y
        del x
    if "ekf" in what:
        C, T, R = 4096, 200_000, 4000
        z = torch.empty((C, T), dtype=torch.float64, device=dev)
        ctx.synth_snr_dev(z.data_ptr(), T, C, 200e3, 1000.0, 6.0, dphi=2 * np.pi / C, seed=3)
        rows = torch.empty((C, T // R, 8), dtype=torch.float64, device=dev)
        ctx.ekf_dev(z.data_ptr(), T, C, 1, T, R, 200e3, 1000.0, _lib.default_ekf_opts(), rows.data_ptr())
        torch.cuda.synchronize()
        del z
    if "sweep" in what:
        nls_sweep([6.0, 12.0], 1_000_000, ndata=15, seed=2)
        torch.cuda.synchronize()
    if "period" in what:
        C, R, N = 4_000_000, 200, 15
        x = torch.empty(C * R, dtype=torch.float64, device=dev)
        ctx.synth_snr_dev(x.data_ptr(), R, C, 200e3, 1000.0, 6.0, seed=5)
        qi = torch.empty((C, 2 * N), dtype=torch.float64, device=dev)
        dc = torch.empty(C, dtype=torch.float64, device=dev)
        ctx.demod(x.data_ptr(), C, R, N, 2 * np.pi / 200, qi.data_ptr(), dc.data_ptr())
        torch.cuda.synchronize()
        del x
    if "tm" in what:  # time-major: 256 channels x 20 s interleaved (8.2 GB)
        C, T, R, N = 256, 4_000_000, 4000, 10
        xc = torch.empty((C, T), dtype=torch.float64, device=dev)
        ctx.synth_snr_dev(xc.data_ptr(), T, C, 200e3, 1000.0, 6.0, dphi=2 * np.pi / C, seed=5)
        xt = xc.t().contiguous()
        del xc
        qi = torch.empty((C * (T // R), 2 * N), dtype=torch.float64, device=dev)
        dc = torch.empty(C * (T // R), dtype=torch.float64, device=dev)
        ctx.demod_tm(xt.data_ptr(), T // R, C, R, N, 2 * np.pi / 200, qi.data_ptr(), dc.data_ptr())
        torch.cuda.synchronize()
        del xt
    if "text" in what:
        vals = 1.0 + np.random.RandomState(0).randn(1_000_000, 2)
        text = "".join(f"{a!r} {b!r} \n" for a, b in vals.tolist()).encode() * 8
        nr = ctx.text_load_host(text)
        out = torch.empty((2, nr), dtype=torch.float64, device=dev)
        ctx.text_parse_dev(2, out.data_ptr(), nr)
        ctx.text_release()
    if "asd" in what:
        from deepfmkit_b200 import physics
        from deepfmkit_b200.simulation import WaveformTables, pack_asd_trial, simulate_asd_batch
        laser, ifo = physics.LaserConfig(), physics.InterferometerConfig()
        laser.amp_n, laser.df_n = 1e-5, 1e3
        tables = WaveformTables(2000, 200e3)
        rec = pack_asd_trial(laser, ifo, 200e3, 0, tables)
        recs = np.repeat(rec[None, :], 100_000, axis=0)
        recs[:, 14] = np.arange(100_000)
        simulate_asd_batch(recs, 2000, 200e3, tables)
    if "widen" in what:
        a = (np.random.RandomState(1).randn(100_000_000, 2) * 3000).astype(np.int16)
        out = torch.empty((2, 100_000_000), dtype=torch.float64, device=dev)
        ad = torch.from_numpy(a).to(dev)
        ctx.widen_dev(ad.data_ptr(), "int16", 100_000_000, 2, True, out.data_ptr(), 100_000_000, scale=1e-4)
        torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
