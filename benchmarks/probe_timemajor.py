#!/usr/bin/env python
"""Time-major (interleaved) multi-channel readout: native fold vs transpose-then-fit (development tool)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepfmkit_b200 import _lib  # noqa: E402
from deepfmkit_b200 import fit as tun  # noqa: E402


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    ctx = _lib.Context(0); ctx.use_torch_stream()
    opts = tun.current_lm_opts()
    for C, secs, f_samp in ((256, 100.0, 200e3), (8, 1600.0, 200e3), (2, 1000.0, 1e6)):
        R = int(f_samp / 1000.0 * 20); N = 10
        T = int(secs * f_samp) // R * R
        bpc = T // R
        w0 = 2 * np.pi * 1000.0 / f_samp
        xc = torch.empty((C, T), dtype=torch.float64, device="cuda")
        ctx.synth_snr_dev(xc.data_ptr(), T, C, f_samp, 1000.0, 6.0, dphi=2 * np.pi / C, seed=5)
        xt = xc.t().contiguous()  # [T, C]
        rows_a = torch.empty((C, bpc, 8), dtype=torch.float64, device="cuda")
        rows_b = torch.empty((C, bpc, 8), dtype=torch.float64, device="cuda")
        ta = timed(lambda: ctx.nls_fit_batch_dev(xc.data_ptr(), C, bpc, T, R, N, w0, [1.6, 6.0, 0, 0], None, 0, True, opts, rows_a.data_ptr()))
        tb = timed(lambda: ctx.nls_fit_batch_tm_dev(xt.data_ptr(), C, bpc, R, N, w0, [1.6, 6.0, 0, 0], None, 0, True, opts, rows_b.data_ptr()))
        tmp = torch.empty((C, T), dtype=torch.float64, device="cuda")
        tc = timed(lambda: ctx.widen_dev(xt.data_ptr(), "float64", T, C, True, tmp.data_ptr(), T))
        qi = torch.empty((C * bpc, 2 * N), dtype=torch.float64, device="cuda"); dc = torch.empty(C * bpc, dtype=torch.float64, device="cuda")
        td = timed(lambda: ctx.demod_tm(xt.data_ptr(), bpc, C, R, N, w0, qi.data_ptr(), dc.data_ptr()))
        dev = float((rows_a[:, :, :4] - rows_b[:, :, :4]).abs().max())
        gb = C * T * 8 / 1e9
        print(json.dumps({"channels": C, "GB": gb, "channel_major_ms": ta, "time_major_native_ms": tb, "transpose_ms": tc,
                          "time_major_demod_ms": td, "time_major_demod_GBps": gb / td * 1e3, "max_param_diff": dev,
                          "flags_equal": bool(torch.equal(rows_a[:, :, 6], rows_b[:, :, 6]))}), flush=True)
        del xc, xt, tmp, rows_a, rows_b, qi, dc
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
