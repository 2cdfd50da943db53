"""Image preparation in front of the feature finding (reference: clustertracking/preprocessing.py).

``lowpass`` (preprocessing.py:12-49) and ``preprocess`` (preprocessing.py:52-75) are the reference's
own; ``bandpass``, ``boxcar``, ``scalefactor_to_gamut`` and ``scale_to_gamut`` are trackpy functions
the reference imports (preprocessing.py:9).  trackpy is a third-party dependency of the reference,
unpinned in its setup.py and absent from this image, so those four are restated from trackpy's
published algorithm (v0.3 series) with PARITY UNPINNED against trackpy itself; the goldens
(tests/golden/preprocess_*.npz) pin the reference's ``preprocess`` around the same restatement.

Host code (numpy / scipy.ndimage, like the reference): these run once per frame in front of the
hot path and are not part of it.  The separable gaussian that ``refine_leastsq`` applies per cluster
when ``noise_size`` is set runs on the device (csrc/ctk_solver.cuh ``lowpass_value``).
"""
import numpy as np
from scipy.ndimage import correlate1d, uniform_filter1d

from .utils import validate_tuple


class Frame(np.ndarray):
    """ndarray that carries ``frame_no`` and ``metadata`` (the part of ``pims.Frame`` the
    reference's pipeline uses: preprocessing.py:45-49, 70-75, find_link.py:46-48)."""

    def __new__(cls, array, frame_no=None, metadata=None):
        obj = np.asarray(array).view(cls)
        obj.frame_no = frame_no
        obj.metadata = dict(metadata) if metadata is not None else {}
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.frame_no = getattr(obj, 'frame_no', None)
        self.metadata = getattr(obj, 'metadata', {})


def gaussian_kernel(sigma, truncate=4.0):
    """trackpy.masks.gaussian_kernel: normalised gaussian on [-lw, lw], lw = int(truncate sigma + .5)."""
    lw = int(truncate * sigma + 0.5)
    x = np.arange(-lw, lw + 1)
    w = np.exp(x ** 2 / (-2 * sigma ** 2))
    return w / np.sum(w)


def _default_threshold(image, threshold):
    if threshold is not None:
        return threshold
    return 1 if np.issubdtype(image.dtype, np.integer) else 1 / 256.


def lowpass(image, lshort, threshold=None):
    """Gaussian lowpass, zero beyond the edge, values <= threshold set to 0 (preprocessing.py:12-49)."""
    lshort = validate_tuple(lshort, image.ndim)
    threshold = _default_threshold(image, threshold)
    result = np.array(image, dtype=np.float64)
    for axis, size in enumerate(lshort):
        if size > 0:
            correlate1d(result, gaussian_kernel(size, 4), axis, output=result, mode='constant', cval=0.0)
    return Frame(np.where(result > threshold, result, 0), getattr(image, 'frame_no', None))


def boxcar(image, size):
    """trackpy.preprocessing.boxcar: running mean of odd width ``size`` per axis, edge replicated."""
    size = validate_tuple(size, image.ndim)
    if not all(int(x) & 1 for x in size):
        raise ValueError("Smoothing size must be an odd integer. Round up.")
    result = np.array(image, dtype=np.float64)
    for axis, width in enumerate(size):
        if width > 1:
            uniform_filter1d(result, int(width), axis, output=result, mode='nearest', cval=0)
    return result


def bandpass(image, lshort, llong, threshold=None, truncate=4):
    """trackpy.preprocessing.bandpass: gaussian lowpass minus boxcar, values <= threshold -> 0."""
    lshort = validate_tuple(lshort, image.ndim)
    llong = validate_tuple(llong, image.ndim)
    if any(x >= y for x, y in zip(lshort, llong)):
        raise ValueError("The smoothing length scale must be larger than the noise length scale.")
    threshold = _default_threshold(image, threshold)
    result = np.array(image, dtype=np.float64)
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


def preprocess(raw_image, noise_size=None, smoothing_size=None, threshold=None):
    """preprocessing.py:52-75: bandpass + rescaling to the full range of the integer type; integer
    images pass untouched without ``noise_size``; float images are scaled to uint8.  The scale
    factor travels in ``metadata['scale_factor']`` (``characterize`` divides it out again)."""
    raw = np.asarray(raw_image)
    if noise_size is not None:
        image = bandpass(raw, noise_size, smoothing_size, threshold)
        dtype = raw.dtype if np.issubdtype(raw.dtype, np.integer) else np.uint8
        scale_factor = scalefactor_to_gamut(image, dtype)
        image = scale_to_gamut(image, dtype, scale_factor)
    elif np.issubdtype(raw.dtype, np.integer):
        scale_factor, image = 1., raw
    else:
        scale_factor = scalefactor_to_gamut(raw, np.uint8)
        image = scale_to_gamut(raw, np.uint8, scale_factor)
    return Frame(image, getattr(raw_image, 'frame_no', None), metadata=dict(scale_factor=scale_factor))


# ---------------------------------------------------------------------------------------------------
# characterize                                                          find_link.py:44-79
# ---------------------------------------------------------------------------------------------------
def _mask_grids(radius):
    """Offsets of the (2 r + 1)-box and the ellipse test of trackpy.masks.binary_mask."""
    points = [np.arange(-r, r + 1) for r in radius]
    grids = np.meshgrid(*points, indexing='ij')
    inside = sum((g / r) ** 2 for g, r in zip(grids, radius)) <= 1
    return grids, inside


def characterize(coords, image, radius, isotropic=True, scale_factor=None):
    """mass, signal and size of every feature from the masked image around it (find_link.py:44-79):
    the box ``slice_pad`` cuts (masks.py:8-27: corner round(c - r), zero padding outside the image),
    masked by the ellipse around the TRUE coordinate (``mask_image``, masks.py:101-118, edge
    included); size = radius of gyration with trackpy's integer r^2 masks.  Vectorised over the
    features; returns the reference's dict."""
    image_arr = np.asarray(image)
    coords = np.atleast_2d(np.asarray(coords, dtype=np.float64))
    ndim = len(radius)
    radius = tuple(int(r) for r in radius)
    if scale_factor is None:
        scale_factor = getattr(image, 'metadata', {}).get('scale_factor', 1.)
    n = len(coords)
    grids, inside = _mask_grids(radius)
    # python's round(): half to even, like np.rint
    corner = np.rint(coords - np.asarray(radius)).astype(np.int64)                 # masks.py:13
    pix = [corner[:, k].reshape((n,) + (1,) * ndim) + (grids[k] + radius[k])[None] for k in range(ndim)]
    valid = np.ones(pix[0].shape, dtype=bool)
    for k in range(ndim):
        valid &= (pix[k] >= 0) & (pix[k] < image_arr.shape[k])
    idx = tuple(np.clip(pix[k], 0, image_arr.shape[k] - 1) for k in range(ndim))
    patch = np.where(valid, image_arr[idx], 0)                                     # zero padding
    # ellipse around the true coordinate, in box coordinates (masks.py:88-90, include_edge)
    rel = coords - corner
    dist = sum(((grids[k] + radius[k])[None] - rel[:, k].reshape((n,) + (1,) * ndim)) ** 2 / radius[k] ** 2
               for k in range(ndim))
    patch = patch * (dist <= 1)
    axes = tuple(range(1, ndim + 1))
    mass = patch.sum(axis=axes)
    signal = patch.max(axis=axes)
    result = dict(mass=mass / scale_factor, signal=signal / scale_factor)
    with np.errstate(invalid='ignore', divide='ignore'):
        if isotropic:
            r2 = (sum(g ** 2 for g in grids)).astype(int) * inside                 # r_squared_mask
            result['size'] = np.sqrt((r2[None] * patch).sum(axis=axes) / mass)
        else:
            for k, key in enumerate(['size_z', 'size_y', 'size_x'][-ndim:]):
                x2 = (grids[k] ** 2).astype(int) * inside                           # x_squared_masks
                result[key] = np.sqrt(ndim * (x2[None] * patch).sum(axis=axes) / mass)
    return result
