"""The device labelling source (csrc/ctk_label.cuh) compiled as plain C++ with a one-lane warp, held
to the host restatement (which is verified against scipy): same label values; and its restated
std::nth_element against the real one, including inputs that reach the heap-select fallback."""
import ctypes
import hashlib
import os
import subprocess

import numpy as np
import pytest

from clustertracking_b200 import _lib
from label_cases import label_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "emul", "ctk_label_emul.cpp")
BUILD = os.path.join(ROOT, "tests", "emul", "_build")
_handle = None


def lib():
    global _handle
    if _handle is None:
        h = hashlib.sha256()
        for path in (SRC, os.path.join(ROOT, "clustertracking_b200", "csrc", "ctk_label.cuh")):
            with open(path, "rb") as fh:
                h.update(fh.read())
        out = os.path.join(BUILD, "liblabel_emul_%s.so" % h.hexdigest()[:16])
        if not os.path.exists(out):
            os.makedirs(BUILD, exist_ok=True)
            for stale in os.listdir(BUILD):
                if stale.startswith("liblabel_emul_") and stale.endswith(".so"):
                    os.remove(os.path.join(BUILD, stale))
            subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                                   "-I", os.path.join(ROOT, "include"),
                                   "-I", os.path.join(ROOT, "clustertracking_b200", "csrc"), SRC, "-o", out])
        _handle = ctypes.CDLL(out)
        _handle.ctk_emul_label_frames.restype = ctypes.c_int
        _handle.ctk_emul_nth_element_check.restype = ctypes.c_int
        _handle.ctk_emul_query_pairs.restype = ctypes.c_int
    return _handle


def emulated_labels(pos, starts, stops, separation, pair_factor=6, window=-1):
    n, m = pos.shape
    cols = [np.ascontiguousarray(pos[:, k]) for k in range(m)]
    ptrs = (ctypes.c_void_p * 3)(*([c.ctypes.data for c in cols] + [None] * (3 - m)))
    labels = np.full(n, -7, np.int32)
    sizes = np.full(n, -7, np.int32)
    flags = np.full(len(starts), -7, np.int32)
    lib().ctk_emul_label_frames(ptrs, ctypes.c_int32(m), ctypes.c_void_p(starts.ctypes.data),
                                ctypes.c_void_p(stops.ctypes.data), ctypes.c_int64(len(starts)),
                                ctypes.c_void_p(separation.ctypes.data), ctypes.c_int64(pair_factor), ctypes.c_int64(window),
                                ctypes.c_void_p(labels.ctypes.data), ctypes.c_void_p(sizes.ctypes.data),
                                ctypes.c_void_p(flags.ctypes.data))
    return labels, sizes, flags


@pytest.mark.parametrize("window", [-1, 0, 20000])
@pytest.mark.parametrize("name", sorted(label_cases()))
def test_emulated_device_labels_equal_host_labels(name, window):
    """window: the fast (shared-memory) window as the kernel sizes it / none / one that only some of
    the arrays fit."""
    pos, starts, stops, separation = label_cases()[name]
    want_l, want_s, _, _ = _lib.cluster_frames(pos, starts, stops, separation, 1)
    got_l, got_s, flags = emulated_labels(pos, starts, stops, separation, pair_factor=400, window=window)
    for f, (a, b) in enumerate(zip(starts, stops)):
        if flags[f] == 1 and window != 0:
            continue                   # more tree nodes than the window holds: flagged, never mislabelled
        assert flags[f] == 0
        assert np.array_equal(got_l[a:b], want_l[a:b])
        assert np.array_equal(got_s[a:b], want_s[a:b])


@pytest.mark.parametrize("name", sorted(label_cases()))
def test_emulated_pairs_come_in_scipy_order(name):
    """The pair LIST itself (the labels depend on its order only through hash collisions)."""
    pos, starts, stops, separation = label_cases()[name]
    data = np.ascontiguousarray(pos[starts[0]:stops[0]] / separation)
    n, m = data.shape
    want = _lib.query_pairs(data)
    cols = [np.ascontiguousarray(data[:, k]) for k in range(m)]
    ptrs = (ctypes.c_void_p * 3)(*([c.ctypes.data for c in cols] + [None] * (3 - m)))
    got = np.empty((len(want) + 1, 2), np.int64)
    count = ctypes.c_int64(0)
    rc = lib().ctk_emul_query_pairs(ptrs, ctypes.c_int32(m), ctypes.c_int64(n), ctypes.c_int64(400),
                                    ctypes.c_int64(-1), ctypes.c_void_p(got.ctypes.data),
                                    ctypes.c_int64(len(got)), ctypes.byref(count))
    assert rc == 0 and count.value == len(want)
    assert np.array_equal(got[:len(want)], want)


def test_capacity_is_flagged_not_mislabelled():
    pos, starts, stops, separation = label_cases()["very_dense"]
    _, _, flags = emulated_labels(pos, starts, stops, separation, pair_factor=6)
    assert flags.tolist() == [1]


@pytest.mark.parametrize("n", [4, 7, 16, 17, 33, 100, 1000, 5000])
def test_restated_nth_element_equals_libstdcxx(n):
    hits = ctypes.c_int32(0)
    rc = lib().ctk_emul_nth_element_check(ctypes.c_int32(n), ctypes.c_int32(n // 2), ctypes.c_int32(200),
                                          ctypes.c_uint32(n), ctypes.byref(hits))
    assert rc == 0
    if n >= 33:
        assert hits.value >= 1          # the adversarial input reached the heap-select fallback


def test_pack_labelled_equals_pack_columns():
    """ctk_cluster_pack_labelled with given labels (and a flagged frame) == the all-host call."""
    pos, starts, stops, separation = label_cases()["frames"]
    n = len(pos)
    cols = [np.ascontiguousarray(pos[:, k]) for k in range(2)]
    sources = [cols[0], cols[1], 3.5]
    out_a = np.empty((n, 3))
    want = _lib.cluster_pack_frames(cols, starts, stops, separation, 2, sources, 0, out_a)
    labels = want[0].astype(np.int32)
    flags = np.zeros(len(starts), np.int32)
    flags[2] = 1                                           # this frame is labelled on the host
    labels[starts[2]:stops[2]] = -1
    out_b = np.empty((n, 3))
    got = _lib.cluster_pack_frames(cols, starts, stops, separation, 2, sources, 0, out_b, labels=labels, flags=flags)
    for w, g in zip(want[:5], got[:5]):                     # labels, sizes, by_cluster, spans, group counts
        assert np.array_equal(w, g)
    for f, (a, count) in enumerate(zip(starts, want[4])):  # group starts: the first `count` entries of a frame
        assert np.array_equal(want[5][a:a + count], got[5][a:a + count])
    assert np.array_equal(out_a, out_b)
