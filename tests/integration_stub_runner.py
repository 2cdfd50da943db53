"""Runs the binding of INTEGRATION.md section 2 VERBATIM (both ``python`` code blocks, extracted from
the file) on one golden fixture, in a process that never imports ``clustertracking_b200``:

    LD_LIBRARY_PATH=clustertracking_b200 python tests/integration_stub_runner.py <fixture> <out.npz>

The names the second block expects are those in scope at clustertracking/refine.py:336 (``f`` after
``find_clusters`` and the default columns, ``ff``, ``bounds`` = the validated tables, ``radius`` ...);
they are provided here by the CPU oracle's restatement of refine.py:242-315 (test infrastructure)."""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)


def stub_blocks():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    section = text.split("## 2. Bind the C ABI inside the reference")[1].split("\n## ")[0]
    blocks = re.findall(r"```python\n(.*?)```", section, flags=re.S)
    assert len(blocks) == 2, len(blocks)
    return blocks


def main(name, out_path):
    import golden_io
    from oracle import cluster_oracle as oracle
    d = golden_io.load(name)
    f0, reader, diameter, kwargs = golden_io.refine_inputs(d, oracle)
    assert not kwargs or list(kwargs) == ['constraints'], kwargs
    ndim = 3 if 'z' in f0.columns else 2
    # ---- what refine.py:242-315 leaves in scope -------------------------------------------------
    ns = dict(np=np)
    ns['_kwargs'] = dict(method='SLSQP', tol=1e-6, options=dict(maxiter=100, disp=False))
    ns['max_iter'], ns['max_shift'], ns['max_rms_dev'], ns['residual_factor'] = 10, 1, 1., 100000.
    ns['constraints'] = kwargs.get('constraints')
    diameter = oracle.as_ndim_tuple(diameter, ndim)
    ns['ndim'], ns['radius'] = ndim, tuple(int(x // 2) for x in diameter)
    ns['isotropic'], ns['name'] = oracle.all_equal(diameter), 'gauss'
    ff = oracle.ModelSpec('gauss', ndim, ns['isotropic'], None)
    ns['ff'] = ff
    if 'frame' not in f0:
        f0['frame'] = 0
    f = oracle.find_clusters(f0, diameter, None, 'frame')                      # refine.py:297
    for col in [p for p in ff.params if p not in f.columns]:                   # refine.py:303-305
        f[col] = ff.default[col]
    ns['f'] = f
    ns['bounds'] = ff.bounds_tables(None, ns['radius'])                        # refine.py:315
    ns['reader'] = reader if hasattr(reader, 'frame_shape') else {0: reader}
    # ---- the documented binding, verbatim ---------------------------------------------------------
    first, second = stub_blocks()
    exec(compile(first, "INTEGRATION.md:_ctk.py", "exec"), ns)
    second = second.replace("from ._ctk import lib, Problem\n", "")            # same namespace here
    exec(compile(second, "INTEGRATION.md:refine.py", "exec"), ns)
    assert not any(m.startswith("clustertracking_b200") for m in sys.modules), "stub must stand alone"
    f = ns['f']
    status = ns['status'].cpu().numpy()
    if (status != 0).any():
        print("cluster status codes:", status.tolist(), "sizes:", ns['sizes'].tolist(), file=sys.stderr)
    np.savez(out_path, index=f.index.values, columns=np.array(list(f.columns)),
             **{"col_" + c: f[c].values for c in f.columns})


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
