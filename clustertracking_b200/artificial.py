"""Synthetic test and benchmark inputs (host side, numpy).

Restates the drawing semantics of the reference's generator (clustertracking/artificial.py:12-141:
radial profile evaluated at pixel centres inside a box of 8 sizes, truncated to the image dtype and
added) and provides the workload generators BASELINE.json names.  Frames are
always PRE-RENDERED into arrays: the reference's ``CoordinateReader`` re-renders with fresh noise on
every access (artificial.py:417-422), which makes it unusable for parity work.
Nothing here is accelerated; it only feeds the hot path.
"""
import numpy as np
import pandas as pd

from .utils import validate_tuple


def feat_gauss(r, ndim):
    """artificial.py:12-14."""
    return np.exp(r ** 2 * ndim / -2)


def feat_ring(r, ndim, thickness):
    """artificial.py:17-19."""
    return np.exp(((r - 1 + thickness) / thickness) ** 2 * ndim / -2)


def feat_disc(r, ndim, disc_size):
    """artificial.py:22-28 (``feat_hat``)."""
    out = np.ones_like(r)
    outer = r > disc_size
    out[outer] = np.exp(((r[outer] - disc_size) / (1 - disc_size)) ** 2 * ndim / -2)
    return out


FEATURES = dict(gauss=feat_gauss, ring=feat_ring, disc=feat_disc)


def draw_feature(image, position, size, max_value, feat_func='gauss', mask_diameter=None, **kwargs):
    """Add one feature to ``image`` in place (artificial.py:81-141)."""
    func = FEATURES[feat_func] if not callable(feat_func) else feat_func
    ndim = image.ndim
    size = validate_tuple(size, ndim)
    extent = [8 * s for s in size] if mask_diameter is None else validate_tuple(mask_diameter, ndim)
    box, axes = [], []
    for c, s, m, lim in zip(position, size, extent, image.shape):
        if c >= lim or c < 0:
            raise ValueError("Position outside of image.")
        lo = max(int(np.floor(c - m / 2)), 0)
        hi = min(int(np.ceil(c + m / 2 + 1)), lim)
        box.append(slice(lo, hi))
        axes.append(np.arange(lo - c, hi - c) / s)
    grids = np.meshgrid(*axes, indexing='ij', sparse=True)
    r = np.sqrt(sum(g ** 2 for g in grids))
    image[tuple(box)] += (max_value * func(r, ndim=ndim, **kwargs)).astype(image.dtype)


def draw_features(shape, positions, size, signals, feat_func='gauss', noise=0, rng=None,
                  dtype=np.uint8, **kwargs):
    """Render many features of one frame at once: every spot is truncated to an integer like
    ``draw_feature`` does, spots are summed, Poisson noise of mean ``noise`` is added and the result
    is clipped to the dtype (artificial.py:368-378)."""
    func = FEATURES[feat_func]
    ndim = len(shape)
    positions = np.atleast_2d(np.asarray(positions, dtype=np.float64))
    size = np.asarray(validate_tuple(size, ndim), dtype=np.float64)
    signals = np.broadcast_to(np.asarray(signals, dtype=np.float64), (len(positions),))
    half = np.ceil(4 * size).astype(int) + 1
    offs = np.meshgrid(*[np.arange(-h, h + 1) for h in half], indexing='ij')
    offs = np.stack([o.ravel() for o in offs], axis=1)                       # (K, ndim)
    base = np.floor(positions).astype(np.int64)                              # (F, ndim)
    pix = base[:, None, :] + offs[None, :, :]                                # (F, K, ndim)
    rel = (pix - positions[:, None, :])
    inside = np.all(np.abs(rel) <= (4 * size + 1), axis=2)
    inside &= np.all((pix >= 0) & (pix < np.asarray(shape)), axis=2)
    r = np.sqrt(np.sum((rel / size) ** 2, axis=2))
    spot = np.floor(signals[:, None] * func(r, ndim=ndim, **kwargs)).astype(np.int64)
    acc = np.zeros(int(np.prod(shape)), dtype=np.int64)
    flat = np.ravel_multi_index(tuple(np.clip(pix[..., k], 0, shape[k] - 1) for k in range(ndim)),
                                shape)
    np.add.at(acc, flat[inside], spot[inside])
    acc = acc.reshape(shape)
    if noise > 0:
        rng = np.random.default_rng(0) if rng is None else rng
        acc = acc + rng.poisson(noise, shape)
    info = np.iinfo(dtype)
    return np.clip(acc, info.min, info.max).astype(dtype)


def jittered_grid(shape, pitch, margin, jitter, rng):
    axes = [np.arange(margin, s - margin + 1e-9, pitch) for s in shape]
    pos = np.array([g.ravel() for g in np.meshgrid(*axes, indexing='ij')], dtype=np.float64).T
    return pos + rng.uniform(-jitter, jitter, pos.shape)


def grow_clusters(rng, centres, counts, bond, max_reach=None):
    """Clusters of ``counts[c]`` members: each new member sits at distance ``bond`` (per-axis
    ellipsoid) from a random earlier member, no two members closer than 0.999 bond, and no member
    farther than ``max_reach`` from the centre."""
    centres = np.atleast_2d(centres)
    ndim = centres.shape[1]
    bond = np.asarray(validate_tuple(bond, ndim), dtype=np.float64)
    pos, owner = [], []
    for c_id, (c, k) in enumerate(zip(centres, counts)):
        members = [np.array(c, dtype=np.float64)]
        tries = 0
        while len(members) < k and tries < 1000:
            tries += 1
            start = members[rng.integers(len(members))]
            v = rng.normal(size=ndim)
            cand = start + v / np.linalg.norm(v) * bond
            if max_reach is not None and np.linalg.norm(cand - c) > max_reach:
                continue
            if all(np.sum(((cand - m) / bond) ** 2) >= 0.998 for m in members):
                members.append(cand)
        pos.extend(members)
        owner.extend([c_id] * len(members))
    return np.array(pos), np.array(owner)


def _start_frame(pos, rng, pos_err, columns, **const):
    f0 = pd.DataFrame(pos + rng.uniform(-pos_err, pos_err, pos.shape), columns=columns)
    for key, val in const.items():
        f0[key] = float(val)
    return f0


def isolated_frame(shape=(512, 512), count=200, size=2.75, spacing=24, margin=12, noise=8,
                   seed=0, pos_err=0.5, signal_range=(80, 160)):
    """BASELINE config 1: isolated gaussian features on one uint8 frame.
    -> (frame, start DataFrame, true positions)."""
    rng = np.random.default_rng(seed)
    pos = []
    tries = 0
    while len(pos) < count and tries < 100 * count:
        tries += 1
        cand = rng.uniform(margin, np.asarray(shape) - margin)
        if all(np.linalg.norm(cand - p) >= spacing for p in pos):
            pos.append(cand)
    pos = np.array(pos)
    signal = rng.uniform(*signal_range, len(pos))
    frame = draw_features(shape, pos, size, signal, noise=noise, rng=rng)
    f0 = _start_frame(pos, rng, pos_err, ['y', 'x'], signal=120., size=size, background=noise / 2.)
    return frame, f0, pos


def clustered_positions(shape, pitch, size, rng, k_range=(2, 6)):
    """Cluster members on a jittered grid of centres."""
    ndim = len(shape)
    size_t = np.asarray(validate_tuple(size, ndim), dtype=np.float64)
    centres = jittered_grid(shape, pitch, pitch / 2., 3., rng)
    counts = rng.integers(k_range[0], k_range[1] + 1, len(centres))
    reach = pitch / 2. - 2 * float(size_t.max()) - 1.
    return grow_clusters(rng, centres, counts, 2 * size_t, max_reach=reach)


def clustered_frame(shape=(1024, 1024), pitch=44, size=2.75, noise=8, seed=0, pos_err=0.5,
                    k_range=(2, 6), signal_range=(80, 160)):
    """BASELINE config 2, one frame: clusters of 2-6 overlapping gaussians, Poisson(noise).
    -> (frame, start DataFrame, true positions)."""
    rng = np.random.default_rng(seed)
    pos, _ = clustered_positions(shape, pitch, size, rng, k_range)
    signal = rng.uniform(*signal_range, len(pos))
    frame = draw_features(shape, pos, size, signal, noise=noise, rng=rng)
    cols = ['z', 'y', 'x'][-len(shape):]
    f0 = _start_frame(pos, rng, pos_err, cols, signal=120., size=size, background=noise / 2.)
    return frame, f0, pos


class FrameStack(object):
    """Minimal reader over a pre-rendered stack: what ``refine_leastsq`` needs from a
    pims.FramesSequence (``frame_shape`` and ``__getitem__``, refine.py:252-256)."""

    def __init__(self, stack, first_frame=0):
        self.stack = stack
        self.first_frame = first_frame
        self.frame_shape = tuple(stack.shape[1:])

    def __len__(self):
        return len(self.stack)

    def __getitem__(self, i):
        return self.stack[int(i) - self.first_frame]


def clustered_video(n_frames, shape=(1024, 1024), pitch=44, size=2.75, noise=8, seed=0,
                    pos_err=0.5, k_range=(2, 6)):
    """BASELINE config 2: ``n_frames`` independent clustered frames (per-frame seed = seed + index).
    -> (FrameStack, start DataFrame with a 'frame' column)."""
    stack = np.empty((n_frames,) + tuple(shape), dtype=np.uint8)
    rows = []
    for t in range(n_frames):
        frame, f0, _ = clustered_frame(shape, pitch, size, noise, seed + t, pos_err, k_range)
        stack[t] = frame
        f0['frame'] = t
        rows.append(f0)
    return FrameStack(stack), pd.concat(rows, ignore_index=True)


def _video(n_frames, shape, make_positions, draw_size, noise, columns, seed, signal_range=(100., 180.),
           **const):
    """``n_frames`` independent frames: positions from ``make_positions(rng)``, start coordinates
    within +-0.5 px of the truth.  -> (FrameStack, start DataFrame with a 'frame' column)."""
    rng = np.random.default_rng(seed)
    stack, rows = [], []
    for t in range(n_frames):
        pos = make_positions(rng)
        signal = rng.uniform(*signal_range, len(pos))
        stack.append(draw_features(shape, pos, draw_size, signal, noise=noise, rng=rng))
        f0 = _start_frame(pos, rng, 0.5, columns, **const)
        f0['frame'] = t
        rows.append(f0)
    return FrameStack(np.ascontiguousarray(np.array(stack))), pd.concat(rows, ignore_index=True)


def dimer_trimer_video(n_frames, shape=(512, 512), size=4.0, bond=8.0, n_dimers=60, n_trimers=40,
                       noise=6, seed=3):
    """BASELINE config 3: rigid dimers (two features one bond apart) and equilateral trimers at
    random angles, like the templates of artificial.py:144-185 -- the shapes ``constraints.dimer`` /
    ``constraints.trimer`` describe.  Fit with diameter 16."""
    def positions(rng):
        centres = jittered_grid(shape, 48, 30, 4, rng)
        centres = centres[rng.permutation(len(centres))[:n_dimers + n_trimers]]
        pos = []
        for k, c in enumerate(centres):
            theta = rng.uniform(0, 2 * np.pi)
            if k < n_dimers:
                offs = [(bond / 2., theta), (bond / 2., theta + np.pi)]
            else:
                offs = [(bond / np.sqrt(3.), theta + 2 * np.pi * m / 3) for m in range(3)]
            pos.extend([c + r * np.array([np.sin(a), np.cos(a)]) for r, a in offs])
        return np.array(pos)

    return _video(n_frames, shape, positions, size, noise, ['y', 'x'], seed,
                  signal=150., size=size, background=noise / 2.)


def confocal_video(n_stacks, shape=(64, 256, 256), size=(2.25, 3.25, 3.25), n_clusters=60,
                   noise=4, seed=4):
    """BASELINE config 4: anisotropic 3D stacks with clusters of 1-4 features at a bond of two
    sizes per axis.  Fit with diameter (9, 13, 13) and ``param_mode=dict(size='var')``."""
    bond = tuple(2 * s for s in size)

    def positions(rng):
        centres = jittered_grid(shape, 30, 16, 3, rng)
        centres = centres[rng.permutation(len(centres))[:n_clusters]]
        counts = rng.integers(1, 5, len(centres))
        return grow_clusters(rng, centres, counts, bond, max_reach=None)[0]

    return _video(n_stacks, shape, positions, size, noise, ['z', 'y', 'x'], seed, signal=150.,
                  size_z=size[0], size_y=size[1], size_x=size[2], background=noise / 2.)
