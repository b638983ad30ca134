"""CPU oracle for the DFMI readout hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain numpy/scipy restatement of the algorithm in the
reference's ``fit.py`` / ``fitters.py`` / ``physics.py`` ('snr' mode).  It is
the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
Nothing under ``deepfmkit_b200/`` imports it, and the product path raises if
its CUDA library is missing -- there is no CPU fallback.

Parity pinning: the reference ships no usable golden vectors for this path
(its only fixture, ``test/fit_data.txt``, lacks its input file).  The oracle is
therefore pinned against outputs of the *unmodified reference executed in the
build container* (``tests/golden/make_golden.py`` imports ``/root/reference``
and writes ``tests/golden/*.npz``); ``tests/test_oracle_golden.py`` checks the
oracle against those fixtures.
"""
