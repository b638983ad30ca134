"""B200-native DFMI readout hot path: lock-in demodulation, batched LM NLS fit, EKF (sm_100a CUDA)."""
from . import fit  # noqa: F401  (module-level solver tunables, patched like the reference's fit.py)
from .core import DeepFitFramework, DeepFitObject, DeepRawObject  # noqa: F401
from .fitters import (BaseFitter, EKFFitter, StandardNLSFitter, ekf_fit_batch, nls_fit_batch,  # noqa: F401
                      rows_to_frame)

__all__ = ["BaseFitter", "StandardNLSFitter", "EKFFitter", "DeepFitFramework", "DeepRawObject", "DeepFitObject",
           "nls_fit_batch", "ekf_fit_batch", "rows_to_frame", "fit"]
