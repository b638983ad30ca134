"""The lock-in for records that cannot fold (incommensurate modulation period): GB/s on cfg-2- and cfg-3-shaped
records whose f_mod is off the sample grid."""
import json, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from deepfmkit_b200 import _lib
ctx = _lib.Context(0); ctx.use_torch_stream()
for name, f_samp, f_mod, R, nbuf, N in (("cfg2-like", 1e6, 1000.3, 20000, 90_000, 10), ("cfg3-like", 200e3, 1000.3, 4000, 1_280_000 // 2, 10),
                                          ("N=20", 1e6, 1000.3, 20000, 45_000, 20), ("long", 1e6, 1000.3, 2_000_000, 450, 10)):
    w0 = 2 * np.pi * f_mod / f_samp
    assert _lib.demod_path(R, w0) == 0
    x = torch.randn(nbuf * R, dtype=torch.float64, device="cuda")
    qi = torch.empty((nbuf, 2 * N), dtype=torch.float64, device="cuda"); dc = torch.empty(nbuf, dtype=torch.float64, device="cuda")
    fn = lambda: ctx.demod(x.data_ptr(), nbuf, R, N, w0, qi.data_ptr(), dc.data_ptr())
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); [fn() for _ in range(3)]; b.record(); b.synchronize()
    ms = a.elapsed_time(b) / 3
    print(json.dumps({"what": name, "R": R, "N": N, "GB": x.numel() * 8 / 1e9, "ms": ms, "GBps": x.numel() * 8 / ms / 1e6}), flush=True)
    del x
