"""How fast can an ordinary numpy array be page-locked in place (cudaHostRegister), against staging it?  One call over
the whole record, then disjoint chunks registered concurrently by several threads (ctypes releases the GIL)."""
import ctypes, json, sys, threading, time
import numpy as np, torch
torch.cuda.init(); torch.zeros(1, device="cuda")
rt = ctypes.CDLL("/usr/local/cuda/lib64/libcudart.so")
rt.cudaHostRegister.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint]
rt.cudaHostUnregister.argtypes = [ctypes.c_void_p]
print(open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(), flush=True)
gb = float(sys.argv[1]) if len(sys.argv) > 1 else 14.4
n = int(gb * 1e9 / 8)
x = np.empty(n); x[::512] = 1.0  # touched
base = x.ctypes.data
t0 = time.perf_counter(); rc = rt.cudaHostRegister(base, x.nbytes, 0); t1 = time.perf_counter()
rc2 = rt.cudaHostUnregister(base); t2 = time.perf_counter()
print(json.dumps({"GB": gb, "threads": 1, "chunk_MB": "all", "register_GBps": round(gb / (t1 - t0), 1),
                  "unregister_GBps": round(gb / (t2 - t1), 1), "rc": [rc, rc2]}), flush=True)
for threads in (1, 2, 4, 8):
    for chunk_mb in (64, 256):
        chunk = chunk_mb << 20
        offs = list(range(0, x.nbytes - chunk + 1, chunk))
        errs = []
        def work(k, fn):
            for off in offs[k::threads]:
                r = fn(base + off, chunk, 0) if fn is rt.cudaHostRegister else fn(base + off)
                if r: errs.append(r)
        res = {}
        for name, fn in (("register", rt.cudaHostRegister), ("unregister", rt.cudaHostUnregister)):
            ts = [threading.Thread(target=work, args=(k, fn)) for k in range(threads)]
            t0 = time.perf_counter(); [t.start() for t in ts]; [t.join() for t in ts]; dt = time.perf_counter() - t0
            res[name + "_GBps"] = round(len(offs) * chunk / dt / 1e9, 1)
        print(json.dumps({"GB": gb, "threads": threads, "chunk_MB": chunk_mb, **res, "errors": len(errs)}), flush=True)
