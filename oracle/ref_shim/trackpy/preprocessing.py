"""Stand-ins for the trackpy functions the reference's preprocessing imports (trackpy is a
third-party dependency of the reference, unpinned in its setup.py, and absent here: no network).
Restated from trackpy's published algorithm (v0.3 series, preprocessing.py): PARITY UNPINNED --
no trackpy output can be generated in this container; goldens made through these functions pin the
REFERENCE's own arithmetic around them, not trackpy's."""
import numpy as np
from scipy.ndimage import correlate1d, uniform_filter1d

from .masks import gaussian_kernel
from .utils import validate_tuple


def boxcar(image, size):
    size = validate_tuple(size, image.ndim)
    if not np.all([x & 1 for x in size]):
        raise ValueError("Smoothing size must be an odd integer. Round up.")
    result = np.array(image, dtype=float)
    for axis, _size in enumerate(size):
        if _size > 1:
            uniform_filter1d(result, _size, axis, output=result, mode='nearest', cval=0)
    return result


def bandpass(image, lshort, llong, threshold=None, truncate=4):
    lshort = validate_tuple(lshort, image.ndim)
    llong = validate_tuple(llong, image.ndim)
    if np.any([x >= y for (x, y) in zip(lshort, llong)]):
        raise ValueError("The smoothing length scale must be larger than the noise length scale.")
    if threshold is None:
        threshold = 1 if np.issubdtype(image.dtype, np.integer) else 1 / 256.
    result = np.array(image, dtype=float)
    for axis, sigma in enumerate(lshort):
        correlate1d(result, gaussian_kernel(sigma, truncate), axis, output=result, mode='constant',
                    cval=0.0)
    result -= boxcar(image, llong)
    return np.where(result > threshold, result, 0)


def scalefactor_to_gamut(image, original_dtype):
    return np.iinfo(original_dtype).max / image.max()


def scale_to_gamut(image, original_dtype, scale_factor=None):
    if scale_factor is None:
        scale_factor = scalefactor_to_gamut(image, original_dtype)
    return (scale_factor * image.clip(min=0.)).astype(original_dtype)
