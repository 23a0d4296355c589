"""tsadar_b200 -- B200-native (sm_100a) kernels for TSADAR's Thomson-scattering form-factor hot path.

Host-side mirror of the reference interface for that path only:
    FormFactor, FitModel, ThomsonScatteringDiagnostic, LossFunction   (same names / argument meaning as
    tsadar.core.physics.form_factor, generate_spectra, tsadar.core.thomson_diagnostic, tsadar.inverse.loss_function)
All arithmetic runs in hand-written CUDA behind the C ABI of include/tsff.h; there is no CPU fallback.
"""
from . import _ffi  # noqa: F401

__all__ = ["_ffi"]
