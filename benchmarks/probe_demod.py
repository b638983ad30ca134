#!/usr/bin/env python
"""Run the demodulation alone on a device-generated batch (profiling target).  usage: probe_demod.py R N nbuf"""
import sys
import os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepfmkit_b200 import _lib

R, nd, C = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
f_samp = 200e3
w0 = 2 * np.pi * 1000 / f_samp
ctx = _lib.Context(0)
ctx.use_torch_stream()
x = torch.empty(C * R, dtype=torch.float64, device="cuda")
ctx.synth_snr_dev(x.data_ptr(), R, C, f_samp, 1000.0, 6.0, seed=1)
qi = torch.empty((C, 2 * nd), dtype=torch.float64, device="cuda")
dc = torch.empty(C, dtype=torch.float64, device="cuda")
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ctx.demod(x.data_ptr(), C, R, nd, w0, qi.data_ptr(), dc.data_ptr())
    b.record()
    b.synchronize()
    print(round(a.elapsed_time(b), 3), "ms", round(C * (R * 8 + (2 * nd + 1) * 8) / a.elapsed_time(b) / 1e6), "GB/s")
