"""CPU restatement of the reference's feature finding (clustertracking/find.py:166-277).

TEST INFRASTRUCTURE: only tests/ may import it.  It calls the same third-party routines the
reference calls (scipy.ndimage.grey_dilation, scipy.spatial.cKDTree, numpy.percentile) and is pinned
to the unmodified reference by tests/golden/find_*.npz (oracle/make_golden.py find).
"""
import numpy as np
from scipy import ndimage
from scipy.spatial import cKDTree


def _per_axis(value, ndim):
    if np.ndim(value) == 0:
        return (value,) * ndim
    value = tuple(value)
    if len(value) != ndim:
        raise ValueError("expected %d values" % ndim)
    return value


def where_close(pos, separation, intensity=None):
    """find.py:166-199: of every pair closer than `separation` (scaled distance < 1 - 1e-7) the
    dimmer feature goes, ties and the intensity-free case by the smaller coordinate sum."""
    if len(pos) == 0:
        return []
    pos = np.asarray(pos)
    separation = _per_axis(separation, pos.shape[1])
    if any(s == 0 for s in separation):
        return []
    scaled = pos / separation
    pairs = cKDTree(scaled, 30).query_pairs(1 - 1e-7, output_type='ndarray')
    if len(pairs) == 0:
        return []
    first, second = pairs[:, 0], pairs[:, 1]
    by_position = np.where(scaled[first].sum(1) > scaled[second].sum(1), second, first)
    if intensity is None:
        return np.unique(by_position)
    intensity = np.asarray(intensity)
    drop = np.where(intensity[first] > intensity[second], second, first)
    tie = intensity[first] == intensity[second]
    drop[tie] = by_position[tie]
    return np.unique(drop)


def grey_dilation(image, separation, percentile=64, margin=None, precise=True):
    """find.py:219-277."""
    image = np.asarray(image)
    ndim = image.ndim
    separation = _per_axis(separation, ndim)
    if margin is None:
        margin = tuple(int(s / 2) for s in separation)
    bright = image[image != 0]
    if len(bright) == 0:
        return np.empty((0, ndim))
    threshold = np.percentile(bright, percentile)
    box = [int(2 * s / np.sqrt(ndim)) for s in separation]
    peaks = (image == ndimage.grey_dilation(image, box, mode='constant')) & (image > threshold)
    if not peaks.any():
        return np.empty((0, ndim))
    pos = np.argwhere(peaks)
    inner = ~np.any((pos < margin) | (pos > np.array(image.shape) - margin - 1), axis=1)
    pos = pos[inner]
    if len(pos) == 0:
        return np.empty((0, ndim))
    if precise:
        pos = np.delete(pos, where_close(pos, separation, image[peaks][inner]), axis=0)
    return pos
