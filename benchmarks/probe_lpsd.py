"""LPSD of a fit-rate phase series on the device: wall time of lpsd() and of the kernels alone."""
import json, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from deepfmkit_b200 import lpsd
for n in (5000, 180000, 1_800_000):
    phi = torch.from_numpy(np.cumsum(np.random.RandomState(0).randn(n)) * 1e-3).cuda()
    lpsd(phi, 50.0); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        out = lpsd(phi, 50.0)
    torch.cuda.synchronize()
    print(json.dumps({"points": n, "ms": (time.perf_counter() - t0) / 3 * 1e3, "frequencies": len(out[0])}), flush=True)
