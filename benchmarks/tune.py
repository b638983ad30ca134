#!/usr/bin/env python
"""Development sweep of kernel launch knobs (env overrides read by the library at launch time).

    python benchmarks/tune.py demod|lm [cfg2|cfg3|cfg5small]
"""
import itertools
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepfmkit_b200 import _lib  # noqa: E402

GEOM = {"cfg2": (1e6, 20, 10, 1, 2000.0), "cfg3": (200e3, 20, 10, 64, 100.0), "cfg5small": (200e3, 1, 15, 4_000_000, 1e-3)}


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    what = sys.argv[1]
    cfg = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
    f_samp, n, nd, C, secs = GEOM[cfg]
    R = int(f_samp / 1000.0 * n)
    T = int(secs * f_samp) // R * R
    nbuf = C * (T // R)
    w0 = 2 * np.pi * 1000.0 / f_samp
    ctx = _lib.Context(0)
    ctx.use_torch_stream()
    x = torch.empty(C * T, dtype=torch.float64, device="cuda")
    ctx.synth_snr_dev(x.data_ptr(), T, C, f_samp, 1000.0, 6.0, seed=1)
    qi = torch.empty((nbuf, 2 * nd), dtype=torch.float64, device="cuda")
    dc = torch.empty(nbuf, dtype=torch.float64, device="cuda")
    rows = torch.empty((nbuf, 8), dtype=torch.float64, device="cuda")
    alg = (8 * R + 8 * (2 * nd + 1)) * nbuf
    if what == "demod":
        ref = None
        from deepfmkit_b200 import _lib as L
        stages = [int(v) for v in os.environ.get("TUNE_STAGES", "16384,24576,32768,40960,49152,65536").split(",")]
        for drift, ctas, nst, stage in itertools.product(("",), (1, 2, 3), (2, 3, 4, 5, 6, 7), stages):
            L.load_library().dfk_dev_clear()
            for key, val in (("DFK_FOLD_CTAS", ctas), ("DFK_FOLD_NSTAGES", nst), ("DFK_FOLD_STAGE_BYTES", stage)):
                L.load_library().dfk_dev_set(key.encode(), int(val))
            if drift:
                L.load_library().dfk_dev_set(b"DFK_FOLD_DRIFT", int(drift))
            qi.zero_()
            try:
                t = timed(lambda: ctx.demod(x.data_ptr(), nbuf, R, nd, w0, qi.data_ptr(), dc.data_ptr()))
            except RuntimeError as e:
                print(json.dumps({"drift": drift, "ctas": ctas, "nst": nst, "stage": stage, "err": str(e)[:80]}))
                torch.cuda.synchronize()
                continue
            q = qi[:64].clone()
            if ref is None:
                ref = q
            dev = float((q - ref).abs().max())
            print(json.dumps({"cfg": cfg, "drift": drift or "plan", "ctas": ctas, "nst": nst, "stage": stage, "ms": round(t, 4),
                              "GBps": round(alg / t / 1e6, 1), "dev_vs_first": dev}), flush=True)
    else:
        ctx.demod(x.data_ptr(), nbuf, R, nd, w0, qi.data_ptr(), dc.data_ptr())
        guess = torch.tensor([1.0, 6.0, 0.0, 0.0], dtype=torch.float64, device="cuda")
        from deepfmkit_b200 import _lib as L
        for minb, lanes in itertools.product((4, 6, 8), (1, 2, 4, 8)):
            L.load_library().dfk_dev_set(b"DFK_LM_MINB", int(minb))
            opts = _lib.default_lm_opts()
            opts.lanes_per_fit = lanes
            t = timed(lambda: ctx.lm_fit(qi.data_ptr(), nbuf, nd, guess.data_ptr(), 0, dc.data_ptr(), opts, rows.data_ptr()))
            print(json.dumps({"cfg": cfg, "minb": minb, "lanes": lanes, "ms": round(t, 4), "fits_per_s": round(nbuf / t * 1e3),
                              "m_mean": float(rows[:, 1].mean())}), flush=True)


if __name__ == "__main__":
    main()
