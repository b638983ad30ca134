#!/usr/bin/env python
"""Throughput of the lock-in at long fold lengths (development tool): 10 MHz / 1 kHz and friends."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepfmkit_b200 import _lib  # noqa: E402


def main():
    ctx = _lib.Context(0)
    ctx.use_torch_stream()
    lib = _lib.load_library()
    total = 2_000_000_000  # 16 GB record
    x = torch.empty(total, dtype=torch.float64, device="cuda")
    ctx.synth_snr_dev(x.data_ptr(), total, 1, 1e6, 1000.0, 6.0, seed=1)
    for P, n, N in ((1000, 20, 10), (2048, 20, 10), (4096, 20, 10), (10000, 20, 10), (16384, 20, 10), (10000, 20, 30),
                    (65536, 4, 10), (10000, 1, 10)):
        R = P * n
        nbuf = total // R
        w0 = 2 * np.pi / P
        qi = torch.empty((nbuf, 2 * N), dtype=torch.float64, device="cuda")
        dc = torch.empty(nbuf, dtype=torch.float64, device="cuda")
        for mode in ((0,) if P <= 2048 else (0, 1)):
            lib.dfk_dev_clear()
            if mode:
                lib.dfk_dev_set(b"DFK_NO_FOLD_LONG", 1)
            best = 1e30
            for _ in range(3 if not mode else 1):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ctx.demod(x.data_ptr(), nbuf, R, N, w0, qi.data_ptr(), dc.data_ptr())
                b.record(); b.synchronize()
                best = min(best, a.elapsed_time(b))
            print(json.dumps({"P": P, "n": n, "N": N, "buffers": nbuf, "kernel": "direct" if mode else "folded", "ms": best,
                              "GBps": nbuf * (R * 8 + 8 * (2 * N + 1)) / best / 1e6}), flush=True)
    lib.dfk_dev_clear()


if __name__ == "__main__":
    main()
