"""Modulation waveforms g(theta) of the reference (waveforms.py), with what the device generator needs to know about
them: the harmonic-series ones (a sum of cosines of multiples of the modulation phase) are evaluated on the GPU per
trial; any other callable is evaluated once on the host and shipped as a table (experiments.Experiment)."""
import numpy as np


def cosine(t_phase):
    """The default waveform of LaserConfig (physics.py:44)."""
    return np.cos(t_phase)


def second_harmonic_distortion(t_phase, distortion_amp=0.0, distortion_phase=0.0):
    """Fundamental plus a second harmonic (waveforms.py:4-25)."""
    return np.cos(t_phase) + distortion_amp * np.cos(2 * t_phase + distortion_phase)


def triangle_wave(t_phase, width=0.5):
    """Sawtooth of period 2 pi rising from -1 to 1 over the fraction ``width`` of the period and falling back over the
    rest; 0.5 gives the symmetric triangle (waveforms.py:26-32, where a signal-processing library supplies it).
    Table-evaluated."""
    t = np.asarray(t_phase, dtype=np.float64)
    w = float(width)
    if not 0.0 <= w <= 1.0:
        raise ValueError("width must be in the interval [0, 1].")
    tmod = np.mod(t, 2 * np.pi)
    y = np.empty_like(tmod)
    rising = tmod < w * 2 * np.pi
    if w > 0:
        y[rising] = tmod[rising] / (np.pi * w) - 1
    if w < 1:
        y[~rising] = (np.pi * (w + 1) - tmod[~rising]) / (np.pi * (1 - w))
    return y


def square_wave(t_phase, duty=0.5):
    """+1 for the first ``duty`` fraction of each 2 pi period, -1 for the rest (waveforms.py:34-43).  Table-evaluated."""
    t = np.asarray(t_phase, dtype=np.float64)
    d = float(duty)
    if not 0.0 <= d <= 1.0:
        raise ValueError("duty must be in the interval [0, 1].")
    return np.where(np.mod(t, 2 * np.pi) < d * 2 * np.pi, 1.0, -1.0)


def dfm_like_wave(t_phase, harmonics=None):
    """Fundamental plus in-phase harmonics ``{n: amplitude}`` (waveforms.py:47-64)."""
    if harmonics is None:
        harmonics = {2: 0.1, 3: 0.05}
    y = np.cos(t_phase)
    for n, amp in harmonics.items():
        y = y + amp * np.cos(n * t_phase)
    return y


def dfm_wave(t_phase, m=1.0, phi=0.0):
    """AC shape of an ideal DFMI signal, cos(phi + m cos(theta)) (waveforms.py:66-89).  Table-evaluated."""
    return np.cos(phi + m * np.cos(t_phase))


def harmonic_terms(func, kwargs):
    """``[(h, a, p), ...]`` with g = sum a cos(h theta + p) when ``func`` is one of the harmonic-series waveforms
    (matched by name, so the reference's own waveforms module works too), else None."""
    name = getattr(func, "__name__", "")
    kwargs = kwargs or {}
    if name in ("cosine", "<lambda>") and not kwargs:
        # LaserConfig's default is `lambda t_phase: np.cos(t_phase)`; any other lambda must not be guessed at
        if name == "<lambda>":
            probe = np.array([0.0, 0.3, 1.7, 4.0])
            try:
                if not np.array_equal(func(probe), np.cos(probe)):
                    return None
            except Exception:
                return None
        return [(1.0, 1.0, 0.0)]
    if name == "second_harmonic_distortion" and set(kwargs) <= {"distortion_amp", "distortion_phase"}:
        return [(1.0, 1.0, 0.0), (2.0, float(kwargs.get("distortion_amp", 0.0)), float(kwargs.get("distortion_phase", 0.0)))]
    if name == "dfm_like_wave" and set(kwargs) <= {"harmonics"}:
        harmonics = kwargs.get("harmonics")
        harmonics = {2: 0.1, 3: 0.05} if harmonics is None else harmonics
        if len(harmonics) <= 5:
            return [(1.0, 1.0, 0.0)] + [(float(n), float(a), 0.0) for n, a in harmonics.items()]
    return None
