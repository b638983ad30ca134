"""Solver tunables of the reference's ``fit.py`` (fit.py:5-16), kept as module globals.

Users of the reference patch these at run time (notebook 0.0_benchmark cell 1 sets
``fit.MAX_LMA_STEPS``, ``fit.FITOK_THRESHOLD``, ``fit.M_GRID_*`` ...).  The CUDA solver takes them as
kernel arguments, so they are read *at call time* by :func:`current_lm_opts` -- patching this module
(or passing a patched reference ``fit`` module via ``tunables_from=``) keeps working.
"""
NPARM = 4
MAXDATA = 40                    # dead in the reference as well (fit.py:6); N up to 64 is supported
MAX_LMA_STEPS = 100
LMA_CONVERGENCE_IMPROVE = 1e-9
LMA_CONVERGENCE_PARAM_CHANGE = 1e-9
FITOK_THRESHOLD = 1e-3

M_GRID_MIN = 5.0
M_GRID_MAX = 30.0
M_GRID_STEP = 0.5
BESSEL_AMP_THRESHOLD = 0.05
SINCOS_AMP_THRESHOLD = 0.1

# not in the reference: lanes cooperating on one fit in the LM kernel (0 = choose from the batch size)
LANES_PER_FIT = 0


def current_lm_opts(tunables_from=None):
    """Snapshot the tunables into the C struct the library takes."""
    import sys
    from . import _lib
    mod = tunables_from if tunables_from is not None else sys.modules[__name__]
    o = _lib.LmOpts()
    o.max_lma_steps = int(getattr(mod, "MAX_LMA_STEPS", MAX_LMA_STEPS))
    o.lanes_per_fit = int(getattr(mod, "LANES_PER_FIT", 0))
    o.conv_improve = float(getattr(mod, "LMA_CONVERGENCE_IMPROVE", LMA_CONVERGENCE_IMPROVE))
    o.conv_param = float(getattr(mod, "LMA_CONVERGENCE_PARAM_CHANGE", LMA_CONVERGENCE_PARAM_CHANGE))
    o.fitok_threshold = float(getattr(mod, "FITOK_THRESHOLD", FITOK_THRESHOLD))
    o.m_grid_min = float(getattr(mod, "M_GRID_MIN", M_GRID_MIN))
    o.m_grid_max = float(getattr(mod, "M_GRID_MAX", M_GRID_MAX))
    o.m_grid_step = float(getattr(mod, "M_GRID_STEP", M_GRID_STEP))
    o.bessel_amp_threshold = float(getattr(mod, "BESSEL_AMP_THRESHOLD", BESSEL_AMP_THRESHOLD))
    o.sincos_amp_threshold = float(getattr(mod, "SINCOS_AMP_THRESHOLD", SINCOS_AMP_THRESHOLD))
    return o
