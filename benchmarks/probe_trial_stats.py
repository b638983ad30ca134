import sys; sys.path.insert(0, "/root/repo")
import torch, numpy as np
from deepfmkit_b200 import _lib
ctx = _lib.get_context(0)
for P, T, C, S in ((19, 1_000_000, 7, 8), (400, 100, 8, 8), (2000, 500, 8, 8)):
    v = torch.randn(P, T, S, dtype=torch.float64, device="cuda")
    out = torch.empty((P, C, 6), dtype=torch.float64, device="cuda")
    cen = torch.zeros((P, C), dtype=torch.float64, device="cuda")
    ctx.use_torch_stream()
    for _ in range(2): ctx.trial_stats_dev(v.data_ptr(), P, T, C, S, out.data_ptr(), center_ptr=cen.data_ptr())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): ctx.trial_stats_dev(v.data_ptr(), P, T, C, S, out.data_ptr(), center_ptr=cen.data_ptr())
    b.record(); b.synchronize()
    print(P, T, C, "ms", a.elapsed_time(b) / 5, "GB/s", v.numel() * 8 / (a.elapsed_time(b) / 5) / 1e6, float(out[0, 0, 1]))
    ctx.use_default_stream()
