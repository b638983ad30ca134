"""Raw-data ingest on the GPU: the reference's ``parse_header`` / ``load_raw`` (core.py:129-174,
259-286) and a binary fast path.

The reference reads a DFMSWPM ``raw_data`` text file once per channel with ``pandas.read_csv``; here the file's data
region goes to the device once and every channel is parsed there (``dfk_text_load_file`` / ``dfk_text_parse_dev``),
bit-identical to what pandas returns.  The record stays on the device: the ``DeepRawObject`` of a channel carries a
CUDA view of its samples (``device_data``) that the fitters read in place, and builds the pandas frame the reference
exposes as ``.data`` only if somebody asks for it.  There is no CPU parser here.
"""
from __future__ import annotations

import ast
import logging
import os

import numpy as np

from . import _lib


def parse_header(raw_file) -> dict:
    """channels, t0, f_samp, f_mod (and the byte offset of the first data row) of a raw_data file."""
    return _lib.raw_parse_header(raw_file)


def load_raw_device(raw_file, device=0, usecols=None):
    """The whole file as one CUDA tensor ``[channels, nrows]`` (float64) plus its header.

    usecols: ascending file columns to keep (default: every channel the header announces)."""
    import torch
    hdr = parse_header(raw_file)
    ctx = _lib.get_context(device)
    dev = torch.device("cuda", device)
    cols = list(range(hdr["channels"])) if usecols is None else [int(c) for c in usecols]
    with torch.cuda.device(dev):
        ctx.use_torch_stream()
        try:
            _, nrows = ctx.text_load_file(raw_file, hdr["data_offset"])
            out = torch.empty((len(cols), nrows), dtype=torch.float64, device=dev)
            nbad = ctx.text_parse_dev(len(cols), out.data_ptr(), max(nrows, 1), usecols=cols) if nrows else 0
            ctx.text_release()
        finally:
            ctx.use_default_stream()
    if nbad:
        logging.warning(f"{raw_file}: {nbad} fields were missing or not numbers (stored as NaN)")
    hdr["nbad"] = nbad
    return out, hdr


def load_raw(raw_file, labels=None, device=0):
    """``DeepFitFramework.load_raw`` (core.py:259-286): one ``DeepRawObject`` per channel, labelled
    ``<raw_file>_ch<c>`` unless ``labels`` names them.  Returns them in channel order."""
    from .core import DeepRawObject
    data, hdr = load_raw_device(raw_file, device=device)
    C = hdr["channels"]
    if labels is None:
        labels = [f"{raw_file}_ch{c}" for c in range(C)]
    else:
        assert len(labels) == C
    raws = []
    for c in range(C):
        raw = DeepRawObject(device_data=data[c], column=f"ch{c}", f_samp=hdr["f_samp"], f_mod=hdr["f_mod"], label=labels[c],
                            t0=hdr["t0"])
        raw.raw_file = raw_file
        raws.append(raw)
    return raws


_NP_TO_RAW = {"int16": "int16", "int32": "int32", "float32": "float32", "float64": "float64"}


def _npy_header(path):
    """(dtype name, shape, fortran_order, data offset) of a .npy file (format 1.0 - 3.0)."""
    with open(path, "rb") as f:
        magic = f.read(8)
        if magic[:6] != b"\x93NUMPY":
            raise ValueError(f"{path} is not a .npy file")
        major = magic[6]
        hlen = int.from_bytes(f.read(2 if major == 1 else 4), "little")
        meta = ast.literal_eval(f.read(hlen).decode("latin1"))
        offset = f.tell()
    dt = np.dtype(meta["descr"])
    if dt.byteorder == ">" or dt.name not in _NP_TO_RAW:
        raise ValueError(f"{path}: unsupported sample type {meta['descr']} (little-endian int16/int32/float32/float64)")
    return dt.name, tuple(meta["shape"]), bool(meta["fortran_order"]), offset


def load_binary(source, f_samp, f_mod, channels=None, dtype=None, time_major=True, scale=1.0, offset=0.0, byte_offset=0,
                labels=None, device=0, t0=0):
    """Binary fast path (additive): a record of int16 / int32 / float32 / float64 samples -> DeepRawObjects whose
    samples are fp64 on the device, ``scale * sample + offset`` (ADC counts -> volts).

    source: a numpy array ``[T, C]`` (time_major) or ``[C, T]``, a ``.npy`` file, or a raw little-endian file (then
    ``dtype`` and ``channels`` are required; ``byte_offset`` skips a header)."""
    import torch
    from .core import DeepRawObject
    ctx = _lib.get_context(device)
    dev = torch.device("cuda", device)
    if isinstance(source, np.ndarray):
        a = np.ascontiguousarray(source)
        if a.ndim == 1:
            a = a[:, None] if time_major else a[None, :]
        T, C = a.shape if time_major else a.shape[::-1]
        out = torch.empty((C, T), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            ctx.ingest_binary_host(a, T, C, time_major, out.data_ptr(), max(T, 1), scale, offset)
        name = "array"
    else:
        path = os.fspath(source)
        if path.endswith(".npy"):
            dtype, shape, fortran, byte_offset = _npy_header(path)
            if len(shape) == 1:
                shape = (shape[0], 1) if time_major else (1, shape[0])
            if fortran:  # column-major [a, b] is row-major [b, a]
                shape, time_major = shape[::-1], not time_major
            T, C = shape if time_major else shape[::-1]
        else:
            if dtype is None or channels is None:
                raise ValueError("a raw binary file needs dtype and channels")
            dtype = np.dtype(dtype).name
            C = int(channels)
            T = (os.path.getsize(path) - byte_offset) // (np.dtype(dtype).itemsize * C)
        out = torch.empty((C, T), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            ctx.ingest_binary_file(path, byte_offset, dtype, T, C, time_major, out.data_ptr(), max(T, 1), scale, offset)
        name = path
    if labels is None:
        labels = [f"{name}_ch{c}" for c in range(C)]
    return [DeepRawObject(device_data=out[c], column=f"ch{c}", f_samp=f_samp, f_mod=f_mod, label=labels[c], t0=t0)
            for c in range(C)]
