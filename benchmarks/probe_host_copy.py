"""Staged (pageable) vs direct (pinned) host entry: GB/s of dfk_nls_fit_host over a cfg-2-shaped record, and the
geometry of the staged copy (stage size, stagers in flight, copy threads, streaming stores)."""
import itertools, json, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from deepfmkit_b200 import _lib

ctx = _lib.Context(0)
R, nd = 20000, 10
w0 = 2 * np.pi * 1000 / 1e6
nbuf = int(sys.argv[1]) if len(sys.argv) > 1 else 45000  # 7.2 GB
xd = torch.empty(nbuf * R, dtype=torch.float64, device="cuda")
ctx.use_torch_stream(); ctx.synth_snr_dev(xd.data_ptr(), nbuf * R, 1, 1e6, 1000.0, 6.0, seed=1); torch.cuda.synchronize(); ctx.use_own_stream()
pinned = torch.empty(nbuf * R, dtype=torch.float64, pin_memory=True); pinned.copy_(xd); torch.cuda.synchronize()
del xd
pageable = pinned.numpy().copy()


def run(arr, reps=3):
    ctx.nls_fit_host(arr, R, nd, w0, [1.6, 6, 0, 0])
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        rows = ctx.nls_fit_host(arr, R, nd, w0, [1.6, 6, 0, 0])
        best = min(best, time.perf_counter() - t0)
    return best, float(rows[:, 6].sum())


dt, ok = run(pinned.numpy())
print(json.dumps({"input": "pinned", "ms": round(dt * 1e3, 1), "GBps": round(pageable.nbytes / dt / 1e9, 1), "fitok": ok}), flush=True)
for kb, stagers, threads, nt in itertools.product((4096, 16384, 65536), (3, 6), (8, 16, 24), (0, 1)):
    with _lib.dev_overrides(DFK_STAGE_KB=kb, DFK_STAGERS=stagers, DFK_COPY_THREADS=threads, DFK_COPY_NT=nt):
        dt, ok = run(pageable, 2)
    print(json.dumps({"input": "pageable", "stage_kb": kb, "stagers": stagers, "threads": threads, "nt": nt,
                      "ms": round(dt * 1e3, 1), "GBps": round(pageable.nbytes / dt / 1e9, 1), "fitok": ok}), flush=True)
