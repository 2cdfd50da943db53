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


_INCLUDE = None


def _deps(path, seen=None):
    """The file and every quoted include reachable from it (csrc/ and include/)."""
    import re
    global _INCLUDE
    if _INCLUDE is None:
        _INCLUDE = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)
    seen = set() if seen is None else seen
    if path in seen or not os.path.isfile(path):
        return seen
    seen.add(path)
    with open(path) as fh:
        text = fh.read()
    for name in _INCLUDE.findall(text):
        for base in (os.path.dirname(path), CSRC, os.path.join(ROOT, "include")):
            _deps(os.path.join(base, name), seen)
    return seen


def _job_hash(job):
    """Hash of one translation unit: its source, the headers it includes, its flags."""
    name, src, defs = job
    h = hashlib.sha256()
    for path in sorted(_deps(os.path.join(CSRC, src))):
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as fh:
            h.update(fh.read())
    h.update(repr((ARCH, [c for c in COMMON if ROOT not in c], defs)).encode())
    return h.hexdigest()[:16]


def _jobs():
    jobs = [("api", "ctk_api.cu", []), ("host", "ctk_host.cpp", []),
            ("find", "ctk_find.cu", []), ("label", "ctk_label.cu", [])]
    for real in ("float", "double"):
        for fam in (0, 1, 2):
            for extra in (0, 1):       # lean / full flavour of every instance (ctk_solver.cuh)
                jobs.append(("inst_%s_%d_%d" % (real, fam, extra), "ctk_inst.cu",
                             ["-DCTK_INST_REAL=%s" % real, "-DCTK_INST_FAM=%d" % fam,
                              "-DCTK_INST_EXTRA=%d" % extra]))
    return jobs


def _compile(job):
    """Compile one translation unit unless its cached object is current (per-object stamp)."""
    name, src, defs = job
    obj = os.path.join(BUILD, name + ".o")
    stamp = obj + ".hash"
    want = _job_hash(job)
    if os.path.exists(obj) and os.path.exists(stamp):
        with open(stamp) as fh:
            if fh.read().strip() == want:
                return obj, False
    cmd = [_nvcc()] + ARCH + COMMON + defs + ["-c", os.path.join(CSRC, src), "-o", obj]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (name, " ".join(cmd), proc.stderr))
    with open(stamp, "w") as fh:
        fh.write(want)
    return obj, True


def build_variant(out, defines, only=None):
    """Experiment helper: build a library variant with extra -D flags into ``out`` (objects in a
    scratch directory).  ``only`` = iterable of job names to compile with the flags (all if None)."""
    import tempfile
    scratch = tempfile.mkdtemp(prefix="ctk_variant_")
    objs = []

    def one(job):
        name, src, defs = job
        if only is not None and name not in only:      # unchanged unit: the cached object of build()
            return _compile(job)[0]
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
    """Compile (only the translation units whose sources changed) and link ``libctk.so``."""
    os.makedirs(BUILD, exist_ok=True)
    jobs = _jobs()
    if force:
        for name, _, _ in jobs:
            path = os.path.join(BUILD, name + ".o.hash")
            if os.path.exists(path):
                os.remove(path)
    stamp = os.path.join(BUILD, "stamp")
    want = hashlib.sha256("".join(_job_hash(j) for j in jobs).encode()).hexdigest()[:16]
    if not force and os.path.exists(OUT) and os.path.exists(stamp):
        with open(stamp) as fh:
            if fh.read().strip() == want:
                return OUT
    workers = max(1, min(len(jobs), os.cpu_count() or 1))
    with concurrent.futures.ThreadPoolExecutor(workers) as pool:
        done = list(pool.map(_compile, jobs))
    if verbose:
        print("libctk.so: compiled %d of %d translation units for sm_100a"
              % (sum(1 for _, fresh in done if fresh), len(jobs)), flush=True)
    cmd = [_nvcc()] + ARCH + ["-shared", "-cudart", "static", "-o", OUT] + [obj for obj, _ in done]
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
