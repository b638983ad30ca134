"""Post-fit step of the readout on the GPU.

``vectorized_downsample`` keeps the name, arguments and edge behaviour of dsp.py:3-56 (block means; the ragged tail
is dropped; a bad ``R`` or a signal shorter than ``R`` gives an empty array).  ``lpsd`` takes the arguments the
reference passes to ``spectools.lpsd.lpsd`` (core.py:590-609, data.py:239-244) and returns, like its ``'legacy'``
form, a 6-tuple whose first element is the frequency vector and whose third is the one-sided power spectral density
the reference stores as ``Sxx``.  Both run on the device only.
"""
from __future__ import annotations

import numpy as np

from . import _lib


def _as_device_vector(signal, device):
    import torch
    if isinstance(signal, torch.Tensor) and signal.is_cuda:
        return signal.to(torch.float64).contiguous().view(-1), True
    return np.ascontiguousarray(np.asarray(signal, dtype=np.float64)).reshape(-1), False


def vectorized_downsample(signal, R, device=0):
    """Block means of ``signal`` over blocks of ``R`` samples (dsp.py:3-56).

    A numpy array (streamed through the library's staged host path) gives a numpy array; a CUDA tensor is used in
    place and gives a CUDA tensor."""
    if not isinstance(R, (int, np.integer)) or isinstance(R, bool) or R <= 0:
        print(f"Downsampling factor R must be a positive integer, but got {R}. Returning empty array.")
        return np.array([])
    R = int(R)
    x, on_device = _as_device_vector(signal, device)
    n = int(x.shape[0])
    if n // R == 0:
        return np.array([])
    ctx = _lib.get_context(x.device.index if on_device else device)
    if not on_device:
        return ctx.downsample_host(x, R)
    import torch
    out = torch.empty(n // R, dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        ctx.use_torch_stream()
        try:
            ctx.downsample_dev(x.data_ptr(), n, R, out.data_ptr())
        finally:
            ctx.use_default_stream()
    return out


def _window_code(win):
    if win is None or win is np.kaiser:
        return 0
    if win is np.hanning:
        return 1
    name = str(getattr(win, "__name__", win)).lower()
    if name.startswith("kaiser"):
        return 0
    if name in ("hann", "hanning"):
        return 1
    raise ValueError(f"window {win!r} is not available on the device (kaiser, hann)")


def lpsd_opts(olap="default", bmin=1, Lmin=0, Jdes=500, Kdes=100, order=0, win=np.kaiser, psll=200):
    o = _lib.default_lpsd_opts()
    o.olap = -1.0 if (olap is None or isinstance(olap, str)) else float(olap)
    o.bmin = float(bmin)
    o.lmin = int(Lmin)
    o.jdes = int(Jdes)
    o.kdes = int(Kdes)
    o.order = int(order)
    o.window = _window_code(win)
    o.psll = float(psll)
    return o


def lpsd(x, fs, olap="default", bmin=1, Lmin=0, Jdes=500, Kdes=100, order=0, win=np.kaiser, psll=200,
         return_type="legacy", device=0):
    """Log-frequency spectral estimate of one series (or of every row of a 2-D array / tensor).

    return_type 'legacy': ``(f, ps, psd, enbw, navs, plan)`` -- positions 0 and 2 are what the reference reads;
    'dict': the same by name.  For a 2-D input ``ps`` and ``psd`` are ``[C, nf]``."""
    import torch
    on_device = isinstance(x, torch.Tensor) and x.is_cuda
    if on_device:
        xt = x.to(torch.float64).contiguous()
        device = xt.device.index
    else:
        xh = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
        xt = torch.from_numpy(xh if xh.flags.writeable else xh.copy()).to(torch.device("cuda", device))
    one = xt.dim() == 1
    if one:
        xt = xt[None, :]
    C, N = xt.shape
    opts = lpsd_opts(olap, bmin, Lmin, Jdes, Kdes, order, win, psll)
    ctx = _lib.get_context(device)
    with torch.cuda.device(xt.device):
        ctx.use_torch_stream()
        try:
            out = ctx.lpsd_dev(xt.data_ptr(), N, 1, C, N, float(fs), opts)
        finally:
            ctx.use_default_stream()
    if one:
        out["ps"], out["psd"] = out["ps"][0], out["psd"][0]
    if return_type == "dict":
        return out
    plan = _lib.lpsd_plan(N, fs, opts)
    return out["f"], out["ps"], out["psd"], out["enbw"], out["navs"], plan
