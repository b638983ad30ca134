#!/usr/bin/env python
"""Phases of the device text parser on a ~1 GB DFMSWPM record (development tool)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepfmkit_b200 import _lib  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    rng = np.random.RandomState(0)
    vals = 1.0 + rng.randn(1_000_000, 2)
    block = "".join(f"{a!r} {b!r} \n" for a, b in vals.tolist()).encode()
    text = block * reps
    nrows_expect = 1_000_000 * reps
    ctx = _lib.get_context(0)
    pinned = torch.empty(len(text), dtype=torch.uint8, pin_memory=True)
    pinned.numpy()[:] = np.frombuffer(text, dtype=np.uint8)
    lib = _lib.load_library()
    import ctypes
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nr = ctypes.c_int64()
        rc = lib.dfk_text_load_host(ctx._h, ctypes.c_void_p(pinned.data_ptr()), len(text), ctypes.byref(nr))
        assert rc == 0 and nr.value == nrows_expect, (rc, nr.value)
        t1 = time.perf_counter()
        out = torch.empty((2, nr.value), dtype=torch.float64, device="cuda")
        nbad = ctx.text_parse_dev(2, out.data_ptr(), nr.value)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        assert nbad == 0
        print(json.dumps({"bytes": len(text), "rows": nr.value, "load_and_index_ms": (t1 - t0) * 1e3, "parse_ms": (t2 - t1) * 1e3,
                          "load_GBps": len(text) / (t1 - t0) / 1e9, "parse_GBps_text": len(text) / (t2 - t1) / 1e9,
                          "parse_samples_per_s": 2 * nr.value / (t2 - t1)}), flush=True)
    assert np.array_equal(out[:, :1000].cpu().numpy(), out[:, 1_000_000:1_001_000].cpu().numpy())
    ctx.text_release()


if __name__ == "__main__":
    main()
