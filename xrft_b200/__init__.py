"""xrft_b200 -- B200-native spectral engine behind the xrft API.

Drop-in for the data-parallel hot path of xgcm/xrft (detrend + window + FFT + spectrum /
cross-spectrum / phase / isotropic binning): the public functions below keep xrft's names,
arguments, coordinate bookkeeping, warnings and errors (xrft/__init__.py:1-8), while the
arithmetic runs in hand-written sm_100a CUDA kernels behind a C-ABI (include/xrft_b200.h).
See DESIGN.md and INTEGRATION.md.
"""
from .dataarray import DataArray, set_options  # noqa: F401
from .api import (  # noqa: F401
    fft, ifft, dft, idft, power_spectrum, cross_spectrum, cross_phase, cross_spectrum_and_phase, isotropize,
    isotropic_power_spectrum, isotropic_cross_spectrum, fit_loglog, detrend, pad, unpad,
)

from .stream import stream  # noqa: F401,E402

__version__ = "0.2.0"
__all__ = [
    "DataArray", "fft", "ifft", "dft", "idft", "power_spectrum", "cross_spectrum", "cross_phase", "cross_spectrum_and_phase",
    "isotropize",
    "isotropic_power_spectrum", "isotropic_cross_spectrum", "fit_loglog", "detrend", "pad", "unpad", "stream",
]
