"""Build ``libctk.so`` in-tree for sm_100a (``python -m clustertracking_b200.build``).

One nvcc invocation per (arithmetic, model family, flavour) instance file plus the API and host files, run
in parallel, then one link.  Objects are cached under ``csrc/_build`` keyed by a hash of the sources
and flags, so repeated calls are cheap.
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libctk.so")
BUILD = os.path.join(CSRC, "_build")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
          "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build libctk.so")
    return exe


def _source_hash(extra):
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        path = os.path.join(CSRC, name)
        if os.path.isfile(path):
            h.update(name.encode())
            with open(path, "rb") as fh:
                h.update(fh.read())
    with open(os.path.join(ROOT, "include", "ctk.h"), "rb") as fh:
        h.update(fh.read())
    h.update(repr(extra).encode())
    return h.hexdigest()[:16]


def _jobs():
    jobs = [("api", "ctk_api.cu", []), ("host", "ctk_host.cpp", []), ("thread", "ctk_thread.cu", []),
            ("find", "ctk_find.cu", [])]
    for real in ("float", "double"):
        for fam in (0, 1, 2):
            for extra in (0, 1):       # lean / full flavour of every instance (ctk_solver.cuh)
                jobs.append(("inst_%s_%d_%d" % (real, fam, extra), "ctk_inst.cu",
                             ["-DCTK_INST_REAL=%s" % real, "-DCTK_INST_FAM=%d" % fam,
                              "-DCTK_INST_EXTRA=%d" % extra]))
    return jobs


def _compile(job):
    name, src, defs = job
    obj = os.path.join(BUILD, name + ".o")
    cmd = [_nvcc()] + ARCH + COMMON + defs + ["-c", os.path.join(CSRC, src), "-o", obj]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (name, " ".join(cmd), proc.stderr))
    return obj


def build_variant(out, defines, only=None):
    """Experiment helper: build a library variant with extra -D flags into ``out`` (objects in a
    scratch directory).  ``only`` = iterable of job names to compile with the flags (all if None)."""
    import tempfile
    scratch = tempfile.mkdtemp(prefix="ctk_variant_")
    objs = []

    def one(job):
        name, src, defs = job
        obj = os.path.join(scratch, name + ".o")
        cmd = [_nvcc()] + ARCH + COMMON + defs + list(defines) + ["-c", os.path.join(CSRC, src), "-o", obj]
        subprocess.run(cmd, check=True, capture_output=True)
        return obj

    with concurrent.futures.ThreadPoolExecutor(os.cpu_count() or 1) as pool:
        objs = list(pool.map(one, _jobs()))
    subprocess.run([_nvcc()] + ARCH + ["-shared", "-cudart", "static", "-o", out] + objs, check=True)
    shutil.rmtree(scratch, ignore_errors=True)
    return out


def build(force=False, verbose=True):
    """Compile and link ``libctk.so``; returns its path."""
    stamp = os.path.join(BUILD, "stamp")
    want = _source_hash((ARCH, [c for c in COMMON if ROOT not in c]))
    if not force and os.path.exists(OUT) and os.path.exists(stamp):
        with open(stamp) as fh:
            if fh.read().strip() == want:
                return OUT
    os.makedirs(BUILD, exist_ok=True)
    jobs = _jobs()
    if verbose:
        print("building libctk.so: %d translation units for sm_100a ..." % len(jobs), flush=True)
    workers = max(1, min(len(jobs), os.cpu_count() or 1))
    with concurrent.futures.ThreadPoolExecutor(workers) as pool:
        objs = list(pool.map(_compile, jobs))
    cmd = [_nvcc()] + ARCH + ["-shared", "-cudart", "static", "-o", OUT] + objs
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (" ".join(cmd), proc.stderr))
    with open(stamp, "w") as fh:
        fh.write(want)
    if verbose:
        print("built", OUT, flush=True)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv)
