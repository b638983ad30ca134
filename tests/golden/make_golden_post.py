#!/usr/bin/env python
"""Mint the post-fit fixtures by executing the UNMODIFIED reference (build container only).

    python tests/golden/make_golden_post.py

post_downsample.npz: outputs of the reference's own ``vectorized_downsample`` (dsp.py:3-56) on seeded inputs, ragged
tails and degenerate arguments included.  Inputs are regenerated from the stored seeds in the tests.
(The spectral estimate has no reference-side fixture: the reference delegates it to ``spectools``, which is not
installed here -- see oracle/post_oracle.py.)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402,F401  (puts the reference on sys.path as the package DeepFMKit)

from DeepFMKit import dsp  # noqa: E402

CASES = [(0, 4000 * 37, 4000), (1, 20000 * 5 + 123, 20000), (2, 1000, 7), (3, 199, 200), (4, 200, 200), (5, 4097 * 3, 4097),
         (6, 65536, 1)]


def case_input(seed, n):
    rng = np.random.RandomState(seed)
    return 1.0 + 0.3 * rng.randn(n) + np.sin(np.arange(n) * 1e-3)


if __name__ == "__main__":
    out = {"cases": np.array(CASES, dtype=np.int64)}
    for seed, n, R in CASES:
        out[f"y{seed}"] = dsp.vectorized_downsample(case_input(seed, n), R)
    np.savez_compressed(os.path.join(HERE, "post_downsample.npz"), **out)
    print({k: v.shape for k, v in out.items()})
