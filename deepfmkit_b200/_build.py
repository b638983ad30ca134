"""Build the sm_100a shared library in-tree (nvcc cross-compiles without a GPU).

Every ``csrc/*.cu`` is one translation unit, compiled in parallel and linked into ``libdfk_b200.so``.
"""
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libdfk_b200.so")
OBJ_DIR = os.path.join(PKG_DIR, "..", "build", "obj")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def sources():
    out = [os.path.join(PKG_DIR, "..", "include", "dfk_b200.h")]
    for name in sorted(os.listdir(CSRC)):
        out.append(os.path.join(CSRC, name))
    return out


def translation_units():
    return [os.path.join(CSRC, n) for n in sorted(os.listdir(CSRC)) if n.endswith(".cu")]


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in sources())


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libdfk_b200.so")
    return nvcc


def build_library(force=False, verbose=False):
    """Compile csrc/*.cu -> libdfk_b200.so. Returns the library path."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [s for s in sources() if not s.endswith(".cu")]
    newest_header = max(os.path.getmtime(h) for h in headers)

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), newest_header):
            return obj, ""
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed on {os.path.basename(src)}:\n" + proc.stdout + proc.stderr)
        return obj, proc.stderr

    units = translation_units()
    with ThreadPoolExecutor(max_workers=len(units)) as pool:
        results = list(pool.map(compile_one, units))
    objs = [o for o, _ in results]
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH + ".tmp"] + objs
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + proc.stdout + proc.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    if verbose:
        for _, log in results:
            print(log)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build_library(force="--incremental" not in sys.argv, verbose="-q" not in sys.argv))
