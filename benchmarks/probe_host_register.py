"""How fast can an ordinary numpy array be page-locked in place (cudaHostRegister), against staging it?"""
import json, sys, time
import numpy as np, torch
torch.cuda.init()
rt = torch.cuda.cudart()
print(open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(), flush=True)
for gb in (1.0, 7.2, 28.8):
    n = int(gb * 1e9 / 8)
    x = np.empty(n); x[::512] = 1.0  # touched
    t0 = time.perf_counter(); rc = rt.cudaHostRegister(x.ctypes.data, x.nbytes, 0); t1 = time.perf_counter()
    d = torch.empty(min(n, 900_000_000), dtype=torch.float64, device="cuda")
    h = torch.from_numpy(x[: d.numel()])
    torch.cuda.synchronize(); t2 = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); t3 = time.perf_counter()
    t4 = time.perf_counter(); rc2 = rt.cudaHostUnregister(x.ctypes.data); t5 = time.perf_counter()
    print(json.dumps({"GB": gb, "register_s": round(t1 - t0, 3), "register_GBps": round(gb / (t1 - t0), 1), "rc": int(rc),
                      "copy_GBps": round(d.numel() * 8 / (t3 - t2) / 1e9, 1), "unregister_s": round(t5 - t4, 3), "rc2": int(rc2)}), flush=True)
    del x, d, h
