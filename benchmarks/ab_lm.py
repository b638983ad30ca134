import os, sys, json
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from deepfmkit_b200 import _lib
ctx = _lib.Context(0); ctx.use_torch_stream()
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); best = min(best, a.elapsed_time(b))
    return best
for name, f_samp, n, nd, C, secs, cold in (("cfg2", 1e6, 20, 10, 1, 3600.0, False), ("cfg3", 200e3, 20, 10, 64, 100.0, False), ("cfg5", 200e3, 1, 15, 4_000_000, 1e-3, True)):
    R = int(f_samp / 1000 * n); T = int(secs * f_samp) // R * R; nbuf = C * (T // R); w0 = 2 * np.pi * 1000 / f_samp
    x = torch.empty(C * T, dtype=torch.float64, device="cuda"); ctx.synth_snr_dev(x.data_ptr(), T, C, f_samp, 1000.0, 6.0, seed=1)
    qi = torch.empty((nbuf, 2 * nd), dtype=torch.float64, device="cuda"); dc = torch.empty(nbuf, dtype=torch.float64, device="cuda")
    ctx.demod(x.data_ptr(), nbuf, R, nd, w0, qi.data_ptr(), dc.data_ptr())
    guess = torch.tensor([1.6 if cold else 1.0, 6.0, 0.0, 0.0], dtype=torch.float64, device="cuda")
    out = {}
    for flat in ("0", "2"):
        for blocks in ("4",):
            _lib.load_library().dfk_dev_set(b"DFK_LM_FLAT", int(flat)); _lib.load_library().dfk_dev_set(b"DFK_LM_FLAT_BLOCKS", int(blocks))
            rows = torch.zeros((nbuf, 8), dtype=torch.float64, device="cuda")
            opts = _lib.default_lm_opts(); opts.lanes_per_fit = 1
            ctx.lm_counters(reset=True)
            t = timed(lambda: ctx.lm_fit(qi.data_ptr(), nbuf, nd, guess.data_ptr(), 0, dc.data_ptr(), opts, rows.data_ptr()))
            cnt = ctx.lm_counters(reset=True)
            out[(flat, blocks)] = rows.clone()
            print(name, "flat", flat, "blocks/SM", blocks, round(t, 4), "ms", round(nbuf / t / 1e3), "Mfits/s", {k: round(v / (6 * nbuf), 2) for k, v in cnt.items()}, flush=True)
    a, b = out[("0", "4")], out[("2", "4")]
    print("  flags equal:", bool((a[:, 6] == b[:, 6]).all()), "max param diff:", float((a[:, :4] - b[:, :4]).abs().max()), "steps equal:", bool((a[:, 7] == b[:, 7]).all()))
    del x, qi
