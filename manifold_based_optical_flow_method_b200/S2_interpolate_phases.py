"""Drop-in for the numerical core of the reference's ``S2_interpolate_phases.py``: the phase variant
of the electrode -> surface interpolation (see S2_interpolate.py in this package).

    compute_phase_from_potentials(potentials)                                   # reference :58-68
    interpolation(surface, ieeg_data_array, coordinates, start, end, save_path, ifsave)   # reference :22-56

``ieeg_data_array`` is the complex array exp(1j * phase) the reference's main block builds
(:181); the complex interpolant's angle is returned / saved.
"""
import numpy as np

from .S2_interpolate import _save, rbf_interpolate


def compute_phase_from_potentials(potentials):
    """Instantaneous phase of the analytic signal.  Like the reference, which calls
    ``scipy.signal.hilbert(potentials)`` without an axis on a (time, electrodes) array, the transform
    runs along the LAST axis.  Host-side numpy (a (T, m) array, upstream of the GPU path)."""
    x = np.asarray(potentials, dtype=np.float64)
    n = x.shape[-1]
    h = np.zeros(n)
    if n % 2 == 0:
        h[0] = h[n // 2] = 1
        h[1:n // 2] = 2
    else:
        h[0] = 1
        h[1:(n + 1) // 2] = 2
    return np.angle(np.fft.ifft(np.fft.fft(x, axis=-1) * h, axis=-1))


def interpolation(surface_path, ieeg_data_array, coordinates, start_sample, end_sample, save_path, ifsave):
    data = np.asarray(ieeg_data_array, dtype=np.complex128)[start_sample:end_sample]
    values = rbf_interpolate(data, coordinates, surface_path, phase=True)
    print(f"interpolated shape (t, vertices): {values.shape}")
    if ifsave is True:
        _save(values, save_path)
    return values


__all__ = ["interpolation", "compute_phase_from_potentials"]
