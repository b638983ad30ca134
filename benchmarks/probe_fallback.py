#!/usr/bin/env python
"""Monte-Carlo sweep with a fixed cold start (init_m = 6 for every m in 2..20): most fits need the grid fallback."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepfmkit_b200 import _lib, nls_sweep

n_trials = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
ms = list(range(2, 21))
ctx = _lib.get_context(0)
for init_m in (None, 6.0):
    nls_sweep(ms, n_trials, init_m=init_m)  # warm-up at the same size (allocations)
    torch.cuda.synchronize()
    ctx.lm_counters(reset=True)
    t0 = time.perf_counter()
    out = nls_sweep(ms, n_trials, init_m=init_m)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    cnt = ctx.lm_counters(reset=True)
    nf = len(ms) * n_trials
    print(f"init_m={init_m}: {nf} fits in {dt * 1e3:.1f} ms = {nf / dt / 1e6:.2f} Mfits/s (generation included); grid fallbacks per fit "
          f"{cnt['n_grid'] / nf:.3f}; fitok fractions {out['fitok'].mean(axis=0).round(4)}")
