"""ctypes loader for the CPU oracle (oracle/dte_oracle.c).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product package.
PARITY STATUS: parity unpinned by the reference (it ships no tests or vectors); see dte_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdte_oracle.so")


class Cfg(C.Structure):
    _fields_ = [("num_levels", C.c_uint32), ("clusters", C.c_uint32), ("trees_per_pu", C.c_uint32),
                ("missing_value", C.c_uint32), ("tree_w_cls", C.c_uint32), ("tree_f_cls", C.c_uint32),
                ("tuple_cls", C.c_uint32), ("num_trees", C.c_uint32)]


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("dte_oracle.c", "dte_oracle.h", "Makefile")]
    if not force and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in src):
        return _SO
    res = subprocess.run(["make", "-C", _HERE, "-B"], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)
    return _SO


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(_SO)
        L.dteo_fpadd.restype = C.c_uint32
        L.dteo_fpadd.argtypes = [C.c_uint32, C.c_uint32]
        L.dteo_fpadd_literal.restype = C.c_uint32
        L.dteo_fpadd_literal.argtypes = [C.c_uint32, C.c_uint32]
        L.dteo_leaf.restype = C.c_uint32
        L.dteo_leaf.argtypes = [C.POINTER(Cfg), C.c_void_p, C.c_void_p, C.c_void_p]
        L.dteo_scores.restype = C.c_int
        L.dteo_scores.argtypes = [C.POINTER(Cfg), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int]
        L.dteo_scores_literal.restype = C.c_int
        L.dteo_scores_literal.argtypes = [C.POINTER(Cfg), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.dteo_ring_combine.restype = None
        L.dteo_ring_combine.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_size_t, C.c_void_p]
        L.dteo_labels.restype = None
        L.dteo_labels.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.dteo_max_threads.restype = C.c_int
        L.dteo_online_cpus.restype = C.c_int
        L.dteo_scores_blocked.restype = C.c_int
        L.dteo_scores_blocked.argtypes = [C.POINTER(Cfg), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
        L.dteo_result_lines.restype = C.c_size_t
        L.dteo_result_lines.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        _LIB = L
    return _LIB


def make_cfg(D, K, S, missing, w_cls, f_cls, tuple_cls, T):
    return Cfg(int(D), int(K), int(S), int(missing) & 0xFFFFFFFF, int(w_cls), int(f_cls), int(tuple_cls), int(T))


def fpadd(a, b, literal=False):
    L = lib()
    return (L.dteo_fpadd_literal if literal else L.dteo_fpadd)(int(a) & 0xFFFFFFFF, int(b) & 0xFFFFFFFF)


def fpadd_many(a, b, literal=False):
    a = np.asarray(a, dtype=np.uint32)
    b = np.asarray(b, dtype=np.uint32)
    f = lib().dteo_fpadd_literal if literal else lib().dteo_fpadd
    return np.fromiter((f(int(x), int(y)) for x, y in zip(a.ravel(), b.ravel())), dtype=np.uint32, count=a.size).reshape(a.shape)


def scores(cfg, weights_cls, findex_cls, tuples, literal_adder=False, threads=1):
    """Raw uint32 score words for tuples [n, F] (uint32 bit patterns or float32)."""
    w = np.ascontiguousarray(weights_cls)
    f = np.ascontiguousarray(findex_cls)
    t = np.ascontiguousarray(tuples)
    n = t.shape[0]
    out = np.empty(n, dtype=np.uint32)
    rc = lib().dteo_scores(C.byref(cfg), w.ctypes.data, f.ctypes.data, t.ctypes.data, n, out.ctypes.data,
                           1 if literal_adder else 0, int(threads))
    if rc:
        raise RuntimeError("dteo_scores failed: %d" % rc)
    return out


def scores_blocked(cfg, weights_cls, findex_cls, tuples, threads=1):
    """Same words as scores(), computed tree-group by tree-group over blocks of tuples (cache-friendly)."""
    w = np.ascontiguousarray(weights_cls)
    f = np.ascontiguousarray(findex_cls)
    t = np.ascontiguousarray(tuples)
    n = t.shape[0]
    out = np.empty(n, dtype=np.uint32)
    rc = lib().dteo_scores_blocked(C.byref(cfg), w.ctypes.data, f.ctypes.data, t.ctypes.data, n, out.ctypes.data, int(threads))
    if rc:
        raise RuntimeError("dteo_scores_blocked failed: %d" % rc)
    return out


def result_lines(score_words):
    s = np.ascontiguousarray(score_words, dtype=np.uint32)
    out = np.empty((s.size // 4, 4), dtype=np.uint32)
    n = lib().dteo_result_lines(s.ctypes.data, s.size, out.ctypes.data)
    return out[:n]


def online_cpus():
    return int(lib().dteo_online_cpus())


def scores_literal(cfg, weights_cls, findex_cls, tuples):
    w = np.ascontiguousarray(weights_cls)
    f = np.ascontiguousarray(findex_cls)
    t = np.ascontiguousarray(tuples)
    n = t.shape[0]
    out = np.empty(n, dtype=np.uint32)
    rc = lib().dteo_scores_literal(C.byref(cfg), w.ctypes.data, f.ctypes.data, t.ctypes.data, n, out.ctypes.data)
    if rc:
        raise RuntimeError("dteo_scores_literal failed: %d" % rc)
    return out


def ring_combine(partials):
    ps = [np.ascontiguousarray(p, dtype=np.uint32) for p in partials]
    n = ps[0].size
    arr = (C.c_void_p * len(ps))(*[p.ctypes.data for p in ps])
    out = np.empty(n, dtype=np.uint32)
    lib().dteo_ring_combine(arr, len(ps), n, out.ctypes.data)
    return out


def labels(score_words):
    s = np.ascontiguousarray(score_words, dtype=np.uint32)
    out = np.empty(s.size, dtype=np.uint8)
    lib().dteo_labels(s.ctypes.data, s.size, out.ctypes.data)
    return out


def max_threads():
    return int(lib().dteo_max_threads())
