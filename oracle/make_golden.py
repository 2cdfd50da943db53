"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference in the build container.

TEST INFRASTRUCTURE.  Run here (``python oracle/make_golden.py``); it needs ``/root/reference`` and
therefore cannot run on the GPU box -- the fixtures it writes are committed and travel instead.
Every case is seeded; inputs and the reference's outputs are stored side by side, so that
``tests/test_oracle_golden.py`` can pin ``oracle/cluster_oracle.py`` to the reference and the
``-m gpu`` tests can compare the CUDA path with the reference's own numbers.

Cases
-----
fitfunc_*   residual / jacobian values of ``FitFunctions.get_residual`` (fitfunc.py:421-489)
pixels_*    ``prepare_subimage`` pixel sets (refine.py:28-58) and ``slices_multiple`` boxes
clusters_*  ``find_clusters`` labels (find.py:132-163)
refine_*    ``refine_leastsq`` end to end (refine.py:82-452), default ``tol`` and ``tol=1e-12``
lowpass_*   ``prepare_subimage`` with ``noise_size`` (refine.py:36-40, preprocessing.py:12-49)
find_*      ``grey_dilation`` local maxima (find.py:166-277)
refine_ring* / refine_disc*  ring and disc in 2D / 3D, isotropic / anisotropic, shape parameter free
fuzz_*      fixed-seed subset of the randomised option sweep (tests/fuzz_cases.py)
"""
import json
import os
import sys
import warnings

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote %-40s %7.1f KiB" % (name + ".npz", os.path.getsize(path) / 1024.))


def frame_to_arrays(prefix, df):
    out = {prefix + "columns": np.array(list(df.columns)), prefix + "index": df.index.values}
    for col in df.columns:
        out[prefix + "col_" + col] = df[col].values
    return out


# ------------------------------------------------------------------------------------------------
def golden_fitfunc(ct):
    from clustertracking.fitfunc import FitFunctions, vect_from_params
    rng = np.random.RandomState(11)
    cases = [
        ("gauss", 2, True, 1, {}), ("gauss", 2, False, 1, {}), ("gauss", 3, True, 1, {}),
        ("gauss", 3, False, 1, {}), ("ring", 2, True, 1, {}), ("ring", 3, False, 2, {}),
        ("gauss", 2, True, 2, {}), ("gauss", 2, True, 3, dict(signal='cluster')),
        ("gauss", 2, True, 4, dict(size='cluster')), ("disc", 2, True, 2, {}),
        ("disc", 3, False, 1, {}),
    ]
    n_pix = 100
    for k, (family, ndim, iso, n, custom) in enumerate(cases):
        ff = FitFunctions(family, ndim, iso)
        param_mode = {p: 'var' for p in ff.params}
        param_mode['background'] = 'cluster'
        param_mode.update(custom)
        ff = FitFunctions(family, ndim, iso, param_mode=param_mode)
        params = rng.random_sample((n, len(ff.params))) * 10
        if family != 'gauss':
            params[:, -1] = rng.uniform(0.2, 0.8, n)
        image = rng.random_sample(n_pix) * 200
        mesh = rng.random_sample((ndim, n_pix)) * 10
        masks = rng.random_sample((n, n_pix)) > 0.5
        norm = 3.7
        residual, jacobian = ff.get_residual([image], [mesh], [masks], params, None, norm)
        vect = vect_from_params(params, ff.modes, None, operation=np.mean)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            fun = residual(vect)
            jac = jacobian(vect) if jacobian is not None else np.zeros(0)
        save("fitfunc_%02d" % k, family=np.array(family), ndim=ndim, isotropic=iso, n=n,
             param_mode=np.array(json.dumps(param_mode)), params=params, image=image, mesh=mesh,
             masks=masks, norm=norm, vect=vect, fun=fun, jac=jac, modes=np.array(ff.modes),
             param_names=np.array(ff.params))


def golden_pixels(ct):
    from clustertracking.refine import prepare_subimage
    from clustertracking.masks import slices_multiple
    rng = np.random.RandomState(5)
    # bounding boxes, incl. edges / out of image / half-integer coordinates (round-half-even)
    boxes = []
    for shape, radius in (((9, 9), 2), ((20, 31), (3, 5)), ((9, 9, 9), 2), ((12, 20, 20), (2, 4, 4))):
        ndim = len(shape)
        for _ in range(12):
            n = rng.randint(1, 5)
            coords = rng.uniform(-4, max(shape) + 4, (n, ndim))
            if rng.rand() < 0.4:
                coords = np.round(coords * 2) / 2.         # exact .5 values
            slices, origin = slices_multiple(coords, shape, radius)
            lo = [-1] * ndim if origin is None else list(origin)
            hi = [-1] * ndim if origin is None else [s.stop for s in slices]
            boxes.append((shape, radius, coords, lo, hi))
    save("pixels_boxes", n_cases=len(boxes),
         **{"shape_%d" % i: np.array(b[0]) for i, b in enumerate(boxes)},
         **{"radius_%d" % i: np.array(b[1]) for i, b in enumerate(boxes)},
         **{"coords_%d" % i: b[2] for i, b in enumerate(boxes)},
         **{"lo_%d" % i: np.array(b[3]) for i, b in enumerate(boxes)},
         **{"hi_%d" % i: np.array(b[4]) for i, b in enumerate(boxes)})
    # pixel sets; integer-valued coordinates put pixels exactly on the mask boundary
    k = 0
    for shape, radius in (((40, 48), 5), ((40, 48), (3, 6)), ((16, 30, 30), (3, 5, 5))):
        ndim = len(shape)
        image = rng.randint(0, 255, shape).astype(np.uint8)
        for variant in range(4):
            n = variant + 1
            centre = np.array([rng.uniform(r + 1, s - r - 1) for r, s in
                               zip(np.broadcast_to(radius, ndim), shape)])
            coords = centre + rng.uniform(-1, 1, (n, ndim)) * np.broadcast_to(radius, ndim) * 1.2
            if variant == 1:
                coords = np.round(coords)                  # boundary pixels with dist == 1 exactly
            if variant == 3:
                coords[0] = 0.4                            # close to the image corner
            vals, mesh, masks = prepare_subimage(coords, image, radius)
            save("pixels_%02d" % k, image=image, radius=np.array(radius), coords=coords, values=vals,
                 mesh=mesh, masks=masks)
            k += 1


def golden_clusters(ct):
    rng = np.random.RandomState(21)
    for k, (ndim, sep) in enumerate(((2, 11), (2, (8, 12)), (3, (6, 10, 10)))):
        rows = []
        for frame in range(3):
            n = 150
            pos = rng.uniform(0, 200 if ndim == 2 else 60, (n, ndim))
            df = pd.DataFrame(pos, columns=['z', 'y', 'x'][-ndim:])
            df['frame'] = frame
            rows.append(df)
        f = pd.concat(rows, ignore_index=True)
        f = f.sample(frac=1, random_state=3)               # shuffled index order
        res = ct.find_clusters(f, sep)
        save("clusters_%02d" % k, separation=np.array(sep), **frame_to_arrays("in_", f),
             **frame_to_arrays("out_", res))


# ------------------------------------------------------------------------------------------------
def _draw(shape, pos, size, signal, feat_func, noise, rng, **feat_kwargs):
    from clustertracking.artificial import draw_feature
    image = np.zeros(shape, dtype=np.uint8)
    for p, s in zip(pos, np.broadcast_to(signal, len(pos))):
        draw_feature(image, p, size, float(s), feat_func, **feat_kwargs)
    if noise > 0:
        image = np.clip(image + rng.poisson(noise, shape), 0, 255).astype(np.uint8)
    return image


def _grow_clusters(rng, centres, sizes_k, bond, ndim):
    """clusters of k members, each attached at distance `bond` to a random earlier member."""
    pos, member_of = [], []
    for c_id, (c, k) in enumerate(zip(centres, sizes_k)):
        members = [np.array(c, dtype=float)]
        while len(members) < k:
            base = members[rng.randint(len(members))]
            v = rng.normal(size=ndim)
            v *= np.broadcast_to(bond, ndim) / np.linalg.norm(v)
            cand = base + v
            if all(np.sum(((cand - m) / np.broadcast_to(bond, ndim)) ** 2) >= 0.998 for m in members):
                members.append(cand)
        pos.extend(members)
        member_of.extend([c_id] * k)
    return np.array(pos), np.array(member_of)


def _refine_case(ct, name, image, f0, diameter, call_kwargs, frames=None, constraint=None,
                 watch_loop=False, first_frame=0):
    """Run the reference with default tol and tol=1e-12 and store everything.
    ``constraint`` = (kind, dist) describes ``call_kwargs['constraints']`` for the fixture.
    ``watch_loop``: also store the clusters whose re-mask loop did not settle (_OuterLoopSpy)."""
    import clustertracking.refine as ref_refine
    reader = image if frames is None else frames
    outs, unsettled = {}, set()
    cols = [c for c in ('z', 'y', 'x') if c in f0.columns]
    for tag, extra in (("ref_", {}), ("tight_", dict(tol=1e-12, options=dict(maxiter=1000)))):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if watch_loop:
                with _OuterLoopSpy(ref_refine) as spy:
                    res = ct.refine_leastsq(f0.copy(), reader, diameter, **dict(call_kwargs, **extra))
                unsettled.update(spy.unsettled(f0, res, cols))
            else:
                res = ct.refine_leastsq(f0.copy(), reader, diameter, **dict(call_kwargs, **extra))
        outs.update(frame_to_arrays(tag, res))
    meta = dict(diameter=diameter, constraint=constraint,
                kwargs={k: v for k, v in call_kwargs.items() if k != 'constraints'})
    if first_frame:
        meta['first_frame'] = first_frame
    if watch_loop:
        meta['unsettled'] = sorted(unsettled)
    save(name, image=np.asarray(image), meta=np.array(json.dumps(meta)),
         **frame_to_arrays("in_", f0), **outs)


class _Video(object):
    """FramesSequence-like reader over a pre-rendered stack (refine.py:252-256 only needs
    ``frame_shape`` and ``__getitem__``)."""

    def __init__(self, stack):
        self.stack = stack
        self.frame_shape = stack.shape[1:]

    def __getitem__(self, i):
        return self.stack[i]

    def __len__(self):
        return len(self.stack)


def golden_refine(ct):
    from clustertracking.constraints import dimer, trimer, tetramer
    from clustertracking.artificial import feat_gauss, feat_ring, feat_disc, draw_cluster
    rng = np.random.RandomState(1234)

    def grid_positions(shape, pitch, margin, jitter):
        axes = [np.arange(margin, s - margin + 1e-9, pitch) for s in shape]
        pos = np.array([g.ravel() for g in np.meshgrid(*axes, indexing='ij')], float).T
        return pos + rng.uniform(-jitter, jitter, pos.shape)

    def start_frame(pos, err, cols, **const):
        f0 = pd.DataFrame(pos + rng.uniform(-err, err, pos.shape), columns=cols)
        for k, v in const.items():
            f0[k] = v
        return f0

    # 1. config-1-like: isolated gaussians, Poisson(8) noise, default modes
    pos = grid_positions((128, 128), 24, 14, 3)
    signal = rng.uniform(80, 160, len(pos))
    image = _draw((128, 128), pos, 2.75, signal, feat_gauss, 8, rng)
    f0 = start_frame(pos, 0.5, ['y', 'x'], signal=120., size=2.75, background=4.)
    _refine_case(ct, "refine_gauss2d_isolated", image, f0, 11, {})
    _refine_case(ct, "refine_gauss2d_isolated_sizevar", image, f0, 11,
                 dict(param_mode=dict(size='var')))

    # 2. config-2-like: clusters of 2-6 at bond length 2*size
    centres = grid_positions((176, 176), 44, 22, 3)
    ks = rng.randint(2, 7, len(centres))
    pos, _ = _grow_clusters(rng, centres, ks, 5.5, 2)
    signal = rng.uniform(80, 160, len(pos))
    image = _draw((176, 176), pos, 2.75, signal, feat_gauss, 8, rng)
    f0 = start_frame(pos, 0.5, ['y', 'x'], signal=120., size=2.75, background=4.)
    _refine_case(ct, "refine_gauss2d_clusters", image, f0, 11, {})
    _refine_case(ct, "refine_gauss2d_clusters_sizevar", image, f0, 11,
                 dict(param_mode=dict(size='var')))
    _refine_case(ct, "refine_gauss2d_clusters_sigcluster", image, f0, 11,
                 dict(param_mode=dict(signal='cluster', size='cluster')))
    # no background column / signal constant / user bounds
    f1 = f0.drop(columns=['background'])
    _refine_case(ct, "refine_gauss2d_clusters_bounds", image, f1, 11,
                 dict(param_mode=dict(size='var'),
                      bounds=dict(signal=(20, 2000), size=(1.0, 9), pos_diff=2.0,
                                  signal_rel_diff=0.5)))

    # 3. noise free, integer start coordinates (mask-boundary pixels, active background bound)
    pos = grid_positions((96, 96), 24, 14, 3)
    image = _draw((96, 96), pos, 2.75, 150., feat_gauss, 0, rng)
    f0 = pd.DataFrame(np.round(pos), columns=['y', 'x'])
    f0['signal'] = 120.
    f0['size'] = 2.75
    _refine_case(ct, "refine_gauss2d_integer_start", image, f0, 11, {})

    # 4. anisotropic 2D, size var per axis
    pos = grid_positions((120, 100), 32, 18, 2)
    image = _draw((120, 100), pos, (5., 3.), 160., feat_gauss, 4, rng)
    f0 = start_frame(pos, 1.0, ['y', 'x'], signal=140., size_y=4.5, size_x=3.3, background=2.)
    _refine_case(ct, "refine_gauss2d_aniso", image, f0, (20, 12), dict(param_mode=dict(size='var')))

    # 5. 3D isotropic and anisotropic dimers
    centres = grid_positions((24, 56, 56), 28, 12, 1)
    pos, _ = _grow_clusters(rng, centres, [2] * len(centres), 5.0, 3)
    image = _draw((24, 56, 56), pos, 2.5, 140., feat_gauss, 4, rng)
    f0 = start_frame(pos, 0.5, ['z', 'y', 'x'], signal=120., size=2.5, background=2.)
    _refine_case(ct, "refine_gauss3d_iso", image, f0, 10, {})
    centres = grid_positions((28, 72, 72), 36, 14, 1)
    ks = rng.randint(1, 4, len(centres))
    pos, _ = _grow_clusters(rng, centres, ks, (4.5, 6.5, 6.5), 3)
    image = _draw((28, 72, 72), pos, (2.25, 3.25, 3.25), 140., feat_gauss, 4, rng)
    f0 = start_frame(pos, 0.5, ['z', 'y', 'x'], signal=120., size_z=2.4, size_y=3.1, size_x=3.4,
                     background=2.)
    _refine_case(ct, "refine_gauss3d_aniso", image, f0, (9, 13, 13),
                 dict(param_mode=dict(size='var')))

    # 6. ring and disc
    pos = grid_positions((120, 120), 32, 18, 2)
    image = _draw((120, 120), pos, 4., 160., feat_ring, 4, rng, thickness=0.2)
    f0 = start_frame(pos, 0.8, ['y', 'x'], signal=150., size=4., background=2.)
    _refine_case(ct, "refine_ring2d", image, f0, 16, dict(fit_function='ring',
                                                         param_val=dict(thickness=0.2)))
    _refine_case(ct, "refine_ring2d_sizevar", image, f0, 16,
                 dict(fit_function='ring', param_val=dict(thickness=0.2),
                      param_mode=dict(size='var')))
    image = _draw((120, 120), pos, 4., 160., feat_disc, 4, rng, disc_size=0.5)
    _refine_case(ct, "refine_disc2d", image, f0, 16, dict(fit_function='disc',
                                                         param_val=dict(disc_size=0.5)))

    # 7. constrained dimers / trimers (one constraint kind per call)
    for k, cname, maker in ((2, "dimer", dimer), (3, "trimer", trimer)):
        shape = (150, 150)
        centres = grid_positions(shape, 40, 24, 2)
        image = np.zeros(shape, dtype=np.uint8)
        pos = []
        for c in centres:
            pos.extend(draw_cluster(image, c, (4., 4.), k, 1., rng.uniform(0, 2 * np.pi),
                                    max_value=float(rng.uniform(128, 192)), feat_func=feat_gauss))
        pos = np.array(pos)
        image = np.clip(image + rng.poisson(6, shape), 0, 255).astype(np.uint8)
        f0 = start_frame(pos, 1.0, ['y', 'x'], signal=160., size=4., background=3.)
        cons = maker(8.0, 2)
        _refine_case(ct, "refine_%s2d_constrained" % cname, image, f0, 16, dict(constraints=cons),
                     constraint=(cname, 8.0))
        _refine_case(ct, "refine_%s2d_free" % cname, image, f0, 16, {})
    shape = (32, 64, 64)
    centres = grid_positions(shape, 32, 16, 1)
    image = np.zeros(shape, dtype=np.uint8)
    pos = []
    for c in centres:
        pos.extend(draw_cluster(image, c, (3., 4., 4.), 2, 1., rng.uniform(0, 2 * np.pi, 3),
                                max_value=160., feat_func=feat_gauss))
    pos = np.array(pos)
    f0 = start_frame(pos, 0.7, ['z', 'y', 'x'], signal=150., size_z=3., size_y=4., size_x=4.)
    _refine_case(ct, "refine_dimer3d_constrained", image, f0, (12, 16, 16),
                 dict(constraints=dimer((6., 8., 8.), 3)), constraint=("dimer", (6., 8., 8.)))

    # 8. large initial error: the outer re-mask loop runs more than once (tests/test_refine.py:913)
    from clustertracking.artificial import SimulatedImage
    np.random.seed(7)
    im = SimulatedImage((160, 160), 5.25, dtype=np.uint8, signal=200, feat_func=feat_gauss, noise=0)
    im.draw_features(40, 15, 21)
    f0 = im.f(noise=7)
    _refine_case(ct, "refine_gauss2d_overlap_7px", np.asarray(im()), f0, 21, dict(separation=24))

    # 9. failure path: rms_dev above max_rms_dev -> cost NaN, parameters unchanged
    pos = grid_positions((96, 96), 24, 14, 3)
    image = _draw((96, 96), pos, 2.75, 150., feat_gauss, 8, rng)
    f0 = start_frame(pos, 0.5, ['y', 'x'], signal=120., size=2.75)
    _refine_case(ct, "refine_failure_rms", image, f0, 11, dict(max_rms_dev=1e-4))

    # 10. multi-frame reader, shuffled row order, frame numbers not starting at 0
    stack, rows = [], []
    for t in range(3):
        centres = grid_positions((96, 96), 44, 24, 3)
        ks = rng.randint(1, 5, len(centres))
        pos, _ = _grow_clusters(rng, centres, ks, 5.5, 2)
        stack.append(_draw((96, 96), pos, 2.75, rng.uniform(80, 160, len(pos)), feat_gauss, 8, rng))
        df = start_frame(pos, 0.5, ['y', 'x'], signal=120., size=2.75, background=4.)
        df['frame'] = t
        rows.append(df)
    f0 = pd.concat(rows, ignore_index=True).sample(frac=1, random_state=9)
    video = _Video(np.array(stack))
    _refine_case(ct, "refine_gauss2d_video", np.array(stack), f0, 11, {}, frames=video)


def golden_tetramer(ct):
    """Tetramers: square in 2D (four shortest distances), tetrahedron in 3D (all six)
    (constraints.py:102-137, tests/test_refine.py:753-765)."""
    from clustertracking.constraints import tetramer
    from clustertracking.artificial import feat_gauss, draw_cluster
    rng = np.random.RandomState(4321)

    def grid_positions(shape, pitch, margin, jitter):
        axes = [np.arange(margin, s - margin + 1e-9, pitch) for s in shape]
        pos = np.array([g.ravel() for g in np.meshgrid(*axes, indexing='ij')], float).T
        return pos + rng.uniform(-jitter, jitter, pos.shape)

    def start_frame(pos, err, cols, **const):
        f0 = pd.DataFrame(pos + rng.uniform(-err, err, pos.shape), columns=cols)
        for k, v in const.items():
            f0[k] = v
        return f0

    shape = (150, 150)
    centres = grid_positions(shape, 48, 28, 2)
    image = np.zeros(shape, dtype=np.uint8)
    pos = []
    for c in centres:
        pos.extend(draw_cluster(image, c, (4., 4.), 4, 1., rng.uniform(0, 2 * np.pi),
                                max_value=float(rng.uniform(128, 192)), feat_func=feat_gauss))
    pos = np.array(pos)
    image = np.clip(image + rng.poisson(6, shape), 0, 255).astype(np.uint8)
    f0 = start_frame(pos, 0.8, ['y', 'x'], signal=160., size=4., background=3.)
    _refine_case(ct, "refine_tetramer2d_constrained", image, f0, 16,
                 dict(constraints=tetramer(8.0, 2)), constraint=("tetramer", 8.0))
    shape = (40, 72, 72)
    centres = grid_positions(shape, 36, 18, 1)
    image = np.zeros(shape, dtype=np.uint8)
    pos = []
    for c in centres:
        pos.extend(draw_cluster(image, c, (3., 4., 4.), 4, 1., rng.uniform(0, 2 * np.pi, 3),
                                max_value=160., feat_func=feat_gauss))
    pos = np.array(pos)
    f0 = start_frame(pos, 0.7, ['z', 'y', 'x'], signal=150., size_z=3., size_y=4., size_x=4.)
    _refine_case(ct, "refine_tetramer3d_constrained", image, f0, (12, 16, 16),
                 dict(constraints=tetramer((6., 8., 8.), 3)), constraint=("tetramer", (6., 8., 8.)))

def golden_lowpass(ct):
    """``noise_size`` / ``threshold``: lowpass of the cluster sub-image (refine.py:36-40,
    preprocessing.py:12-49), pixel sets and end-to-end fits."""
    from clustertracking.refine import prepare_subimage
    from clustertracking.artificial import feat_gauss
    rng = np.random.RandomState(99)
    # pixel values after the filter; boxes touching the image edge (zero padding at the BOX edge)
    k = 0
    for shape, radius, noise_size, threshold in (((40, 48), 5, 1, None), ((40, 48), (3, 6), (1, 1.5), 20),
                                                 ((40, 48), 4, (0, 2), None),
                                                 ((16, 30, 30), (3, 5, 5), (0.8, 1, 1), 5)):
        ndim = len(shape)
        image = rng.randint(0, 255, shape).astype(np.uint8)
        for variant in range(3):
            n = variant + 1
            centre = np.array([rng.uniform(r + 1, s - r - 1) for r, s in
                               zip(np.broadcast_to(radius, ndim), shape)])
            coords = centre + rng.uniform(-1, 1, (n, ndim)) * np.broadcast_to(radius, ndim) * 1.2
            if variant == 2:
                coords[0] = 1.3                            # close to the image corner
            vals, mesh, masks = prepare_subimage(coords, image, radius, noise_size, threshold)
            save("lowpass_pixels_%02d" % k, image=image, radius=np.array(radius), coords=coords,
                 values=np.asarray(vals), mesh=mesh, masks=masks,
                 noise_size=np.array(noise_size, dtype=float),
                 threshold=np.array(np.nan if threshold is None else threshold, dtype=float))
            k += 1

    def grid_positions(shape, pitch, margin, jitter):
        axes = [np.arange(margin, s - margin + 1e-9, pitch) for s in shape]
        pos = np.array([g.ravel() for g in np.meshgrid(*axes, indexing='ij')], float).T
        return pos + rng.uniform(-jitter, jitter, pos.shape)

    def start_frame(pos, err, cols, **const):
        f0 = pd.DataFrame(pos + rng.uniform(-err, err, pos.shape), columns=cols)
        for key, v in const.items():
            f0[key] = v
        return f0

    centres = grid_positions((176, 176), 44, 22, 3)
    ks = rng.randint(1, 6, len(centres))
    pos, _ = _grow_clusters(rng, centres, ks, 5.5, 2)
    signal = rng.uniform(80, 160, len(pos))
    image = _draw((176, 176), pos, 2.75, signal, feat_gauss, 16, rng)
    f0 = start_frame(pos, 0.5, ['y', 'x'], signal=120., size=2.75, background=4.)
    _refine_case(ct, "refine_lowpass2d", image, f0, 11, dict(noise_size=1))
    _refine_case(ct, "refine_lowpass2d_threshold", image, f0, 11,
                 dict(noise_size=(1, 0.7), threshold=12, param_mode=dict(size='var')))
    centres = grid_positions((28, 72, 72), 36, 14, 1)
    ks = rng.randint(1, 3, len(centres))
    pos, _ = _grow_clusters(rng, centres, ks, (4.5, 6.5, 6.5), 3)
    image = _draw((28, 72, 72), pos, (2.25, 3.25, 3.25), 140., feat_gauss, 8, rng)
    f0 = start_frame(pos, 0.5, ['z', 'y', 'x'], signal=120., size_z=2.25, size_y=3.25, size_x=3.25,
                     background=2.)
    _refine_case(ct, "refine_lowpass3d", image, f0, (9, 13, 13), dict(noise_size=(0.6, 1, 1)))
def golden_find(ct):
    """``grey_dilation`` (find.py:219-277): local maxima above a percentile, margin, drop_close."""
    from clustertracking.find import grey_dilation
    rng = np.random.RandomState(77)

    def blobs(shape, n, sigma, amp, noise, dtype):
        """smooth random blobs + noise; plateaus and ties included on purpose (integer images)"""
        img = np.zeros(shape, dtype=float)
        grids = np.meshgrid(*[np.arange(s) for s in shape], indexing='ij')
        for _ in range(n):
            c = [rng.uniform(0, s) for s in shape]
            r2 = sum(((g - ci) / sg) ** 2 for g, ci, sg in zip(grids, c, np.broadcast_to(sigma, len(shape))))
            img += rng.uniform(0.4, 1.0) * amp * np.exp(-r2)
        img += rng.poisson(noise, shape)
        info = np.iinfo(dtype)
        return np.clip(img, 0, info.max).astype(dtype)

    cases = [
        ("2d_u8", blobs((200, 240), 60, 3.0, 180, 6, np.uint8), dict(separation=11)),
        ("2d_u8_even", blobs((160, 200), 50, 3.0, 200, 4, np.uint8), dict(separation=12, percentile=80)),
        ("2d_u8_aniso", blobs((180, 150), 40, (4.0, 2.5), 180, 5, np.uint8),
         dict(separation=(14, 9), margin=(3, 20))),
        ("2d_u8_fast", blobs((150, 150), 40, 3.0, 160, 8, np.uint8), dict(separation=9, precise=False)),
        ("2d_u16", blobs((128, 160), 30, 3.5, 3000, 40, np.uint16), dict(separation=13, percentile=50)),
        ("2d_black", np.zeros((64, 64), np.uint8), dict(separation=9)),
        ("2d_sparse", (rng.uniform(size=(90, 90)) > 0.995).astype(np.uint8) * 200, dict(separation=7)),
        ("3d_u8", blobs((24, 64, 72), 25, (2.0, 3.0, 3.0), 180, 3, np.uint8), dict(separation=(7, 11, 11))),
        ("3d_u8_margin", blobs((20, 50, 60), 20, 2.5, 200, 2, np.uint8),
         dict(separation=9, margin=(2, 6, 6), percentile=90)),
    ]
    for name, image, kwargs in cases:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            pos = grey_dilation(image, **kwargs)
        save("find_" + name, image=image, kwargs=np.array(json.dumps(kwargs)),
             pos=np.asarray(pos, dtype=np.int64).reshape(-1, image.ndim))


def golden_ringdisc(ct):
    """Ring and disc in the geometries the reference's own suite runs them in
    (tests/test_refine.py:768-881: 2D, 2D anisotropic, 3D, 3D anisotropic), with the default modes
    and with the shape parameter (``thickness`` / ``disc_size``) free.  refine_ring2d /
    refine_ring2d_sizevar / refine_disc2d (golden_refine) hold the 2D isotropic default cases."""
    from clustertracking.artificial import feat_ring, feat_disc
    rng = np.random.RandomState(2468)

    def grid_positions(shape, pitch, margin, jitter):
        axes = [np.arange(margin, s - margin + 1e-9, pitch) for s in shape]
        pos = np.array([g.ravel() for g in np.meshgrid(*axes, indexing='ij')], float).T
        return pos + rng.uniform(-jitter, jitter, pos.shape)

    geometries = [
        ("2d", (120, 120), 32, 18, 4., 16, ['y', 'x']),
        ("2d_aniso", (140, 110), 36, 20, (5., 3.), (20, 12), ['y', 'x']),
        ("3d", (28, 60, 60), 28, 14, 2.5, 10, ['z', 'y', 'x']),
        ("3d_aniso", (30, 72, 72), 34, 15, (2.25, 3.25, 3.25), (9, 13, 13), ['z', 'y', 'x']),
    ]
    for family, feat, extra in (("ring", feat_ring, dict(thickness=0.25)),
                                ("disc", feat_disc, dict(disc_size=0.5))):
        for tag, shape, pitch, margin, size, diameter, cols in geometries:
            pos = grid_positions(shape, pitch, margin, 2)
            image = _draw(shape, pos, size, rng.uniform(120, 180, len(pos)), feat, 4, rng, **extra)
            f0 = pd.DataFrame(pos + rng.uniform(-0.5, 0.5, pos.shape), columns=cols)
            f0['signal'] = 150.
            f0['background'] = 2.
            if np.isscalar(size):
                f0['size'] = size
            else:
                for c, sz in zip(cols, size):
                    f0['size_' + c] = sz
            free = dict(param_mode={list(extra)[0]: 'var'})
            if not (tag == "2d"):
                _refine_case(ct, "refine_%s%s" % (family, tag), image, f0, diameter,
                             dict(fit_function=family, param_val=extra), watch_loop=True)
            _refine_case(ct, "refine_%s%s_free" % (family, tag), image, f0, diameter,
                         dict(fit_function=family, param_val=extra, **free), watch_loop=True)


# (seed, case) of the randomised sweep (tests/fuzz_cases.py, seeds 1-4 of
# profiles/tools/fuzz_parity.py) kept as fixtures: every ring / disc case with a free shape
# parameter or with ``noise_size``, plus every case in which the two solvers end further than
# 1e-3 px apart or the reference's re-mask loop cycles.
FUZZ_SUBSET = [
    (1, 2), (1, 3), (1, 5), (1, 10), (1, 31), (1, 33), (1, 35), (2, 6), (2, 10), (2, 22), (2, 45),
    (2, 50), (2, 52), (2, 57), (3, 6), (3, 16), (3, 17), (3, 18), (3, 19), (3, 20), (3, 22), (3, 24),
    (3, 26), (3, 27), (3, 29), (3, 41), (3, 48), (3, 51), (3, 54), (4, 9), (4, 10), (4, 17), (4, 22),
    (4, 27), (4, 34), (4, 37), (4, 40), (4, 42), (4, 43), (4, 45), (4, 47), (4, 56)]


class _OuterLoopSpy(object):
    """Observes (does not alter) the reference's re-mask loop: records the mask centres handed to
    ``prepare_subimages`` (refine.py:366), from which ``unsettled`` derives the clusters whose loop
    ran out of ``max_iter`` while the mask was still moving (refine.py:383-388)."""

    def __init__(self, module):
        self.module, self.inner, self.calls = module, module.prepare_subimages, []

    def __enter__(self):
        def spy(coords, *args, **kwargs):
            self.calls.append(np.array(coords, dtype=float))
            return self.inner(coords, *args, **kwargs)
        self.module.prepare_subimages = spy
        return self

    def __exit__(self, *exc):
        self.module.prepare_subimages = self.inner

    def unsettled(self, f0, res, cols, max_shift=1.):
        """Cluster ids whose final positions are further than max_shift from the last mask centre."""
        groups = [(int(cid), group) for (_, cid), group in res.groupby(['frame', 'cluster'])]
        starts = [f0.loc[group.index, cols].values.astype(float) for _, group in groups]
        out, at = [], 0
        for k, (cid, group) in enumerate(groups):
            assert at < len(self.calls) and np.array_equal(self.calls[at], starts[k]), "spy lost track"
            last = at
            at += 1
            # later calls of this cluster: until the next cluster's start vector shows up
            while at < len(self.calls) and not (k + 1 < len(groups) and
                                                np.array_equal(self.calls[at], starts[k + 1])):
                last = at
                at += 1
            if np.isnan(group['cost'].values).any():
                continue
            moved = np.sum((group[cols].values - self.calls[last]) ** 2, axis=1)
            if not np.all(moved < max_shift ** 2):
                out.append(cid)
        assert at == len(self.calls), "spy lost track"
        return out


def golden_fuzz(ct):
    """The fixed-seed subset of the randomised option sweep, answered by the unmodified reference at
    its default tolerance and at tol=1e-12, plus the clusters whose re-mask loop did not settle."""
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
    sys.path.insert(0, os.path.dirname(HERE))
    import fuzz_cases
    import clustertracking.refine as ref_refine
    by_seed = {}
    for seed, case in FUZZ_SUBSET:
        by_seed.setdefault(seed, []).append(case)
    for seed, wanted in sorted(by_seed.items()):
        for case in fuzz_cases.cases(seed, max(wanted) + 1, only=wanted):
            kwargs = fuzz_cases.bind(case, ct.constraints)
            f0, cols = case['f0'], case['cols']
            outs, unsettled = {}, set()
            for tag, extra in (("ref_", {}), ("tight_", dict(tol=1e-12, options=dict(maxiter=1000)))):
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    with _OuterLoopSpy(ref_refine) as spy:
                        res = ct.refine_leastsq(f0.copy(), case['frame'], case['diameter'],
                                                **dict(kwargs, **extra))
                unsettled.update(spy.unsettled(f0, res, cols))
                outs.update(frame_to_arrays(tag, res))
            meta = dict(diameter=case['diameter'], constraint=case['constraint'], seed=seed,
                        case=case['case'], kwargs=case['kwargs'], unsettled=sorted(unsettled),
                        **case['meta'])
            save("fuzz_%d_%02d" % (seed, case['case']), image=case['frame'],
                 meta=np.array(json.dumps(meta)), **frame_to_arrays("in_", f0), **outs)


def golden_global(ct):
    """Global-mode fits (refine.py:319-332): one problem over the whole table, some columns shared
    by ALL features.  The combinations the reference's suite exercises (tests/test_refine.py:924-938:
    ``signal='global'`` with free positions) plus a shared size, single- and multi-frame, 3D."""
    rng = np.random.RandomState(97531)

    def grid_positions(shape, pitch, margin, jitter):
        axes = [np.arange(margin, s - margin + 1e-9, pitch) for s in shape]
        pos = np.array([g.ravel() for g in np.meshgrid(*axes, indexing='ij')], float).T
        return pos + rng.uniform(-jitter, jitter, pos.shape)

    def start_frame(pos, err, cols, **const):
        f0 = pd.DataFrame(pos + rng.uniform(-err, err, pos.shape), columns=cols)
        for k, v in const.items():
            f0[k] = v
        return f0

    shape = (128, 128)
    centres = grid_positions(shape, 40, 24, 3)
    pos, _ = _grow_clusters(rng, centres, rng.randint(1, 4, len(centres)), 5.5, 2)
    image = _draw(shape, pos, 2.75, 160., 'gauss', 6, rng)
    f0 = start_frame(pos, 0.4, ['y', 'x'], signal=130., size=2.75, background=3.)
    _refine_case(ct, "refine_global_signal2d", image, f0, 11,
                 dict(param_mode=dict(signal='global')))
    f0 = start_frame(pos, 0.4, ['y', 'x'], signal=130., size=3.2, background=3.)
    _refine_case(ct, "refine_global_size2d", image, f0, 11,
                 dict(param_mode=dict(signal='var', size='global')))
    # the train_leastsq pattern (refine.py:494-512): positions, signal, background fixed, shape global
    f0 = start_frame(pos, 0.0, ['y', 'x'], signal=150., size=3.3, background=6.)
    _refine_case(ct, "refine_global_train2d", image, f0, 11,
                 dict(param_mode=dict(signal='const', background='const', pos='const', size='global'),
                      bounds=dict(size_rel_diff=(0.9, 9))))
    # three frames, signal and size global
    stack, rows = [], []
    for t in range(3):
        c = grid_positions((96, 96), 44, 26, 3)
        p, _ = _grow_clusters(rng, c, rng.randint(1, 3, len(c)), 5.5, 2)
        stack.append(_draw((96, 96), p, 2.75, 150., 'gauss', 5, rng))
        ft = start_frame(p, 0.4, ['y', 'x'], signal=120., size=3.0, background=2.)
        ft['frame'] = t + 2
        rows.append(ft)
    f0 = pd.concat(rows, ignore_index=True)
    frames = {t + 2: img for t, img in enumerate(stack)}
    _refine_case(ct, "refine_global_video2d", np.array(stack), f0, 11,
                 dict(param_mode=dict(signal='global', size='global')), frames=_VideoAt(np.array(stack), 2),
                 first_frame=2)
    shape = (32, 64, 64)
    centres = grid_positions(shape, 30, 15, 2)
    pos, _ = _grow_clusters(rng, centres, rng.randint(1, 3, len(centres)), (4.5, 6.5, 6.5), 3)
    image = _draw(shape, pos, (2.25, 3.25, 3.25), 150., 'gauss', 4, rng)
    f0 = start_frame(pos, 0.4, ['z', 'y', 'x'], signal=120., size_z=2.5, size_y=3.0, size_x=3.0,
                     background=2.)
    _refine_case(ct, "refine_global_size3d", image, f0, (9, 13, 13),
                 dict(param_mode=dict(signal='var', size='global')))


def golden_preprocess(ct):
    """The reference's ``preprocess`` (preprocessing.py:52-75) and ``characterize``
    (find_link.py:44-79).  Both lean on trackpy functions that are absent here; the shim restates
    those from the published algorithm (oracle/ref_shim/trackpy: PARITY UNPINNED against trackpy),
    so these fixtures pin the reference's own arithmetic around them."""
    from clustertracking.preprocessing import preprocess, lowpass
    from clustertracking.find_link import characterize
    rng = np.random.RandomState(2468)
    img = (rng.poisson(12, (96, 128)) + 90 * np.exp(-((np.indices((96, 128)) - np.array([40, 70])[:, None, None]) ** 2).sum(0) / 18.)).astype(np.uint8)
    cases = dict(u8_band=(img, dict(noise_size=1, smoothing_size=11)),
                 u8_band_aniso=(img, dict(noise_size=(1, 1.5), smoothing_size=(9, 13), threshold=2)),
                 u8_plain=(img, dict()),
                 f64_plain=(img.astype(np.float64) / 7.3, dict()),
                 u16_band=((img.astype(np.uint16) * 37), dict(noise_size=1.2, smoothing_size=9)))
    vol = (rng.poisson(6, (24, 40, 40)) + 80 * np.exp(-((np.indices((24, 40, 40)) - np.array([12, 20, 22])[:, None, None, None]) ** 2 / np.array([8., 18., 18.])[:, None, None, None]).sum(0))).astype(np.uint8)
    cases['u8_band_3d'] = (vol, dict(noise_size=1, smoothing_size=(5, 9, 9)))
    for name, (image, kw) in cases.items():
        out = preprocess(image, **kw)
        save("preprocess_" + name, image=image, kwargs=np.array(json.dumps(kw)), out=np.asarray(out),
             scale_factor=np.float64(out.metadata['scale_factor']))
    lp = lowpass(img, (1, 1.5), threshold=3)
    save("preprocess_lowpass", image=img, kwargs=np.array(json.dumps(dict(lshort=(1, 1.5), threshold=3))),
         out=np.asarray(lp), scale_factor=np.float64(1.))
    # characterize: features inside, at the edge (zero padding) and at half-integer coordinates
    coords = np.array([[40.3, 70.2], [3.2, 5.7], [92.5, 120.5], [50.5, 64.0], [41.0, 69.5]])
    pre = preprocess(img, noise_size=1, smoothing_size=11)
    for name, image, radius, iso, cds in (
            ("iso2d", pre, (5, 5), True, coords), ("aniso2d", pre, (4, 6), False, coords),
            ("raw2d", img, (5, 5), True, coords),
            ("aniso3d", vol, (3, 5, 5), False, np.array([[12.2, 20.4, 21.7], [1.5, 3.0, 36.5], [20.6, 30.1, 8.8]]))):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            res = characterize(cds, image, radius, isotropic=iso)
        save("characterize_" + name, image=np.asarray(image), coords=cds, radius=np.array(radius),
             isotropic=np.array(iso), scale_factor=np.float64(getattr(image, 'metadata', {}).get('scale_factor', 1.)),
             **{"out_" + k: np.asarray(v) for k, v in res.items()})


class _VideoAt(object):
    """_Video whose first frame has a number other than 0."""

    def __init__(self, stack, first):
        self.stack, self.first = stack, first
        self.frame_shape = stack.shape[1:]

    def __getitem__(self, i):
        return self.stack[int(i) - self.first]

    def __len__(self):
        return len(self.stack)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    ct = ref_loader.load()
    only = sys.argv[1:] or ["fitfunc", "pixels", "clusters", "refine", "tetramer", "lowpass", "find",
                            "ringdisc", "fuzz"]
    for part in only:
        globals()["golden_" + part](ct)
