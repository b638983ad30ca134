import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from deepfmkit_b200 import _lib
ctx = _lib.Context(0)
R, nd = 20000, 10
w0 = 2 * np.pi * 1000 / 1e6
nbuf = 18000  # 2.88 GB
xd = torch.empty(nbuf * R, dtype=torch.float64, device="cuda")
ctx.use_torch_stream(); ctx.synth_snr_dev(xd.data_ptr(), nbuf * R, 1, 1e6, 1000.0, 6.0, seed=1); torch.cuda.synchronize(); ctx.use_own_stream()
pinned = torch.empty(nbuf * R, dtype=torch.float64, pin_memory=True); pinned.copy_(xd); torch.cuda.synchronize()
pageable = pinned.numpy().copy()
for name, arr in (("pinned", pinned.numpy()), ("pageable", pageable)):
    ctx.nls_fit_host(arr, R, nd, w0, [1.6, 6, 0, 0])
    t0 = time.perf_counter()
    for _ in range(3):
        rows = ctx.nls_fit_host(arr, R, nd, w0, [1.6, 6, 0, 0])
    dt = (time.perf_counter() - t0) / 3
    print(name, round(dt * 1e3, 1), "ms", round(arr.nbytes / dt / 1e9, 1), "GB/s", rows[:, 6].sum())
