#!/usr/bin/env python
"""Place the UNMODIFIED reference (mdovale/DeepFMKit) under baseline/_ref/ as an importable package.

The contract's recipe, ``pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target
baseline/_ref /root/reference``, fails here: the checkout has neither setup.py nor pyproject.toml ("Directory
'/root/reference' is not installable").  The checkout *is* the package directory (flat modules plus an empty
__init__.py; its notebooks import it as ``DeepFMKit``), so what pip would have done is reproduced by hand: the
module files are copied verbatim into baseline/_ref/DeepFMKit/.  baseline/_ref/ is git-ignored (the reference's
sources never enter this repository's history) but not gpurun-ignored, so it travels to the GPU box, where
/root/reference does not exist; ``bench.py --impl reference`` and the CPU legs of bench.py import it from there.

    python baseline/install_ref.py            (build container; __graft_entry__.build() calls it too)
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref", "DeepFMKit")


def install(src=None, quiet=False):
    src = src or os.environ.get("DFK_REFERENCE", "/root/reference")
    if not os.path.isdir(src):
        return False
    os.makedirs(DEST, exist_ok=True)
    n = 0
    for name in sorted(os.listdir(src)):
        if name.endswith(".py"):
            shutil.copyfile(os.path.join(src, name), os.path.join(DEST, name))
            n += 1
    if not quiet:
        print(f"installed {n} reference modules into {DEST}")
    return n > 0


def import_reference():
    """Import the installed reference as ``DeepFMKit``; returns (core, fit, fitters) or raises ImportError.

    matplotlib and pyplnoise are not in this image and the readout path never touches them ('snr'-mode synthesis,
    StandardNLSFitter, EKFFitter, fit.py): they are stubbed before the import, nothing in the reference is patched."""
    from unittest.mock import MagicMock
    if not os.path.exists(os.path.join(DEST, "fitters.py")):
        raise ImportError(f"{DEST} is empty: run baseline/install_ref.py where /root/reference exists")
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.dates", "matplotlib.cm",
                 "pyplnoise"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = MagicMock()
    root = os.path.dirname(DEST)
    if root not in sys.path:
        sys.path.insert(0, root)
    import DeepFMKit.core as core
    import DeepFMKit.fit as fit
    import DeepFMKit.fitters as fitters
    return core, fit, fitters


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
