"""TEST INFRASTRUCTURE: point sets for the cluster-labelling tests (device kernel / its one-lane host
build against the host restatement)."""
import numpy as np


def label_cases():
    rng = np.random.RandomState(1)
    cases = {}

    def one(name, pos, separation, cuts=None):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        n = len(pos)
        cuts = [0, n] if cuts is None else cuts
        cases[name] = (pos, np.asarray(cuts[:-1], np.int64), np.asarray(cuts[1:], np.int64),
                       np.asarray(separation, np.float64))

    for n, extent, m in [(900, 30, 2), (700, 9, 3), (2100, 45, 2), (5000, 40, 2), (40, 0.6, 2), (1, 1, 2),
                         (17, 3, 2), (3000, 14, 3), (500, 40, 1)]:
        one("uniform_%dd_%d" % (m, n), rng.uniform(0, extent, (n, m)), np.ones(m))
    one("grid_2d", rng.randint(0, 200, (800, 2)), [11., 11.])
    one("grid_3d", rng.randint(0, 60, (600, 3)), [9., 13., 13.])
    one("coarse_grid", rng.randint(0, 40, (300, 2)) * 0.5, [1., 1.])
    one("identical", np.zeros((100, 2)), [1., 1.])
    one("repeated", np.repeat(rng.uniform(0, 30, (50, 2)), 20, axis=0), [1., 1.])
    one("dense_integers", rng.randint(0, 60, (4000, 2)), [1.5, 1.5])
    one("frames", rng.uniform(0, 60, (6000, 2)), [2., 2.], cuts=[0, 2000, 2000, 4100, 6000])
    one("many_frames", rng.uniform(0, 100, (64 * 500, 2)), [3., 2.], cuts=list(range(0, 64 * 500 + 1, 500)))
    one("very_dense", rng.uniform(0, 4, (600, 2)), [1., 1.])            # more pairs than the device keeps
    return cases


label_cases.expected_flagged = {"very_dense": 1, "identical": 1, "repeated": 1, "uniform_1d_500": 1}
