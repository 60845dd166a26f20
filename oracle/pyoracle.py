"""ctypes binding of oracle/liboracle.so — the CPU restatement of the reference's hot path.

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never from pyrope_b200/ (the product fails loudly without
its CUDA library instead of falling back to this).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

L2, IP, COSINE = 0, 1, 2
METRICS = {"L2": L2, "InnerProduct": IP, "IP": IP, "Cosine": COSINE}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    try:
        build()
    except Exception:
        if not os.path.exists(_SO):
            raise
    L = C.CDLL(_SO)
    f32p, i64p, i32p, u8p = (C.POINTER(C.c_float), C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                             C.POINTER(C.c_uint8))
    vp = C.c_void_p

    def sig(name, res, *args):
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = list(args)

    sig("orc_random_new", vp, C.c_int32)
    sig("orc_random_free", None, vp)
    sig("orc_random_next", C.c_int32, vp)
    sig("orc_random_next_double", C.c_double, vp)
    sig("orc_random_fill", None, C.c_int32, C.c_int64, f32p)
    for n in ("orc_dot", "orc_l2sq", "orc_dot_unsafe", "orc_l2sq_unsafe"):
        sig(n, C.c_float, f32p, f32p, C.c_int)
    sig("orc_norm", C.c_float, f32p, C.c_int)
    sig("orc_cosine", C.c_float, f32p, f32p, C.c_float, C.c_float, C.c_int)
    sig("orc_find_nearest_centroid", C.c_int, f32p, f32p, f32p, C.c_int, C.c_int, C.c_int)
    sig("orc_kmeans_train", C.c_int, f32p, C.c_int64, C.c_int, C.c_int64, C.c_int, C.c_int,
        C.c_int, C.c_int32, f32p, i32p)
    sig("orc_pq_new", vp, C.c_int, C.c_int, C.c_int)
    sig("orc_pq_free", None, vp)
    sig("orc_pq_train", None, vp, f32p, C.c_int64)
    sig("orc_pq_ksub", C.c_int, vp, C.c_int)
    sig("orc_pq_get_codebook", None, vp, f32p)
    sig("orc_pq_set_codebook", None, vp, f32p, i32p)
    sig("orc_pq_encode", C.c_int, vp, f32p, u8p)
    sig("orc_pq_distance_table", None, vp, f32p, f32p)
    sig("orc_flat_new", vp, C.c_int, C.c_int)
    sig("orc_flat_free", None, vp)
    sig("orc_flat_add", C.c_int, vp, C.c_int64, f32p)
    sig("orc_flat_upsert", None, vp, C.c_int64, f32p)
    sig("orc_flat_delete", C.c_int, vp, C.c_int64)
    sig("orc_flat_count", C.c_int, vp)
    sig("orc_flat_add_batch", None, vp, C.c_int64, i64p, f32p)
    sig("orc_flat_search", C.c_int, vp, f32p, C.c_int, C.c_int64, i64p, f32p)
    sig("orc_ivfflat_new", vp, C.c_int, C.c_int, C.c_int)
    sig("orc_ivfflat_free", None, vp)
    sig("orc_ivfflat_set_nprobe", None, vp, C.c_int)
    sig("orc_ivfflat_add", None, vp, C.c_int64, f32p)
    sig("orc_ivfflat_add_batch", None, vp, C.c_int64, i64p, f32p)
    sig("orc_ivfflat_delete", C.c_int, vp, C.c_int64)
    sig("orc_ivfflat_build", None, vp)
    sig("orc_ivfflat_is_built", C.c_int, vp)
    sig("orc_ivfflat_ncentroids", C.c_int, vp)
    sig("orc_ivfflat_get_centroids", None, vp, f32p)
    sig("orc_ivfflat_count", C.c_int, vp)
    sig("orc_ivfflat_adopt", None, vp, C.c_int, f32p, i64p, i64p, f32p)
    sig("orc_ivfflat_list_size", C.c_int, vp, C.c_int)
    sig("orc_ivfflat_get_list", None, vp, C.c_int, i64p)
    sig("orc_ivfflat_search", C.c_int, vp, f32p, C.c_int, C.c_int64, C.c_int, i64p, f32p)
    sig("orc_ivfpq_new", vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)
    sig("orc_ivfpq_free", None, vp)
    sig("orc_ivfpq_add", None, vp, C.c_int64, f32p)
    sig("orc_ivfpq_add_batch", None, vp, C.c_int64, i64p, f32p)
    sig("orc_ivfpq_delete", C.c_int, vp, C.c_int64)
    sig("orc_ivfpq_build", None, vp)
    sig("orc_ivfpq_is_built", C.c_int, vp)
    sig("orc_ivfpq_ncentroids", C.c_int, vp)
    sig("orc_ivfpq_get_centroids", None, vp, f32p)
    sig("orc_ivfpq_pq", vp, vp)
    sig("orc_ivfpq_list_size", C.c_int, vp, C.c_int)
    sig("orc_ivfpq_get_list", None, vp, C.c_int, i64p, u8p)
    sig("orc_ivfpq_adopt", None, vp, C.c_int, f32p, f32p, i64p, i64p, u8p)
    sig("orc_ivfpq_search", C.c_int, vp, f32p, C.c_int, C.c_int, i64p, f32p)
    sig("orc_delta_merge", C.c_int, i64p, f32p, C.c_int, i64p, f32p, C.c_int, C.c_int, i64p, f32p)
    sig("orc_search_batch", None, C.c_int, vp, f32p, C.c_int64, C.c_int, C.c_int64, C.c_int,
        C.c_int, i64p, f32p, i32p)
    sig("orc_max_threads", C.c_int)
    _lib = L
    return L


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def _i64(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(C.POINTER(C.c_int64))


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


# ---------------------------------------------------------------- System.Random
class DotNetRandom:
    def __init__(self, seed: int):
        self._h = lib().orc_random_new(seed)

    def next(self) -> int:
        return lib().orc_random_next(self._h)

    def next_double(self) -> float:
        return lib().orc_random_next_double(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_random_free(self._h)
            self._h = None


def random_vectors(count: int, dim: int, seed: int) -> np.ndarray:
    """Pyrope.Benchmarks/Program.cs:251-263 GenerateRandomVectors."""
    out = np.empty((count, dim), dtype=np.float32)
    lib().orc_random_fill(seed, count * dim, _ptr(out, C.c_float))
    return out


# ---------------------------------------------------------------- VectorMath
def _bin(name, a, b):
    a, pa = _f32(a)
    b, pb = _f32(b)
    if a.shape != b.shape:
        raise ValueError("Vector dimension mismatch")  # VectorMath.cs:421-426
    return float(getattr(lib(), name)(pa, pb, a.size))


def dot(a, b):
    return _bin("orc_dot", a, b)


def l2sq(a, b):
    return _bin("orc_l2sq", a, b)


def dot_unsafe(a, b):
    return _bin("orc_dot_unsafe", a, b)


def l2sq_unsafe(a, b):
    return _bin("orc_l2sq_unsafe", a, b)


def norm(v):
    v, pv = _f32(v)
    return float(lib().orc_norm(pv, v.size))


def cosine(a, b, an=None, bn=None):
    a, pa = _f32(a)
    b, pb = _f32(b)
    if a.shape != b.shape:
        raise ValueError("Vector dimension mismatch")
    an = norm(a) if an is None else an
    bn = norm(b) if bn is None else bn
    return float(lib().orc_cosine(pa, pb, an, bn, a.size))


# ---------------------------------------------------------------- KMeans / PQ
def find_nearest_centroid(vec, centroids, metric=L2):
    v, pv = _f32(vec)
    c, pc = _f32(centroids)
    cn = np.array([norm(r) for r in c], dtype=np.float32) if metric == COSINE else np.zeros(len(c), np.float32)
    return lib().orc_find_nearest_centroid(pv, pc, _ptr(cn, C.c_float), c.shape[0], c.shape[1], metric)


def kmeans_train(data, k, metric=L2, max_iter=10, seed=42):
    d, pd = _f32(data)
    n, dim = d.shape
    kk = max(1, min(k if k > 0 else 1, n))
    out = np.zeros((kk, dim), dtype=np.float32)
    iters = C.c_int32(0)
    got = lib().orc_kmeans_train(pd, n, dim, dim, k, metric, max_iter, seed, _ptr(out, C.c_float),
                                 C.byref(iters))
    return out[:got], iters.value


class ProductQuantizer:
    def __init__(self, dim, m, k):
        if dim % m != 0:
            raise ValueError("Dimension must be divisible by M")
        if k > 256:
            raise ValueError("K must be <= 256 for byte encoding")
        self.dim, self.m, self.k, self.sub = dim, m, k, dim // m
        self._h = lib().orc_pq_new(dim, m, k)
        self._own = True

    @classmethod
    def _borrow(cls, h, dim, m, k):
        o = cls.__new__(cls)
        o.dim, o.m, o.k, o.sub, o._h, o._own = dim, m, k, dim // m, h, False
        return o

    def train(self, data):
        d, pd = _f32(data)
        lib().orc_pq_train(self._h, pd, d.shape[0])

    def ksub(self):
        return [lib().orc_pq_ksub(self._h, i) for i in range(self.m)]

    def codebook(self):
        out = np.zeros((self.m, self.k, self.sub), dtype=np.float32)
        lib().orc_pq_get_codebook(self._h, _ptr(out, C.c_float))
        return out

    def set_codebook(self, cb, ksub=None):
        cb, pcb = _f32(cb)
        ks = None
        if ksub is not None:
            ksa = np.ascontiguousarray(ksub, dtype=np.int32)
            ks = _ptr(ksa, C.c_int32)
        lib().orc_pq_set_codebook(self._h, pcb, ks)

    def encode(self, vec):
        v, pv = _f32(vec)
        if v.size != self.dim:
            raise ValueError("Vector dimension mismatch")
        code = np.zeros(self.m, dtype=np.uint8)
        if lib().orc_pq_encode(self._h, pv, _ptr(code, C.c_uint8)) != 0:
            raise RuntimeError("PQ not trained")
        return code

    def distance_table(self, q):
        v, pv = _f32(q)
        out = np.zeros((self.m, self.k), dtype=np.float32)
        lib().orc_pq_distance_table(self._h, pv, _ptr(out, C.c_float))
        return out

    def __del__(self):
        if getattr(self, "_own", False) and getattr(self, "_h", None):
            lib().orc_pq_free(self._h)
            self._h = None


# ---------------------------------------------------------------- indexes
class _Index:
    kind = -1

    def _out(self, topk):
        k = max(topk, 1)
        return np.zeros(k, dtype=np.int64), np.zeros(k, dtype=np.float32)

    def search_batch(self, Q, topk, max_scans=-1, nprobe=-1, nthreads=0):
        Q, pq = _f32(Q)
        nq = Q.shape[0]
        ids = np.full((nq, topk), -1, dtype=np.int64)
        sc = np.zeros((nq, topk), dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.int32)
        lib().orc_search_batch(self.kind, self._h, pq, nq, topk, max_scans, nprobe, nthreads,
                               _ptr(ids, C.c_int64), _ptr(sc, C.c_float), _ptr(cnt, C.c_int32))
        return ids, sc, cnt


class FlatIndex(_Index):
    """BruteForceVectorIndex.cs"""
    kind = 0

    def __init__(self, dim, metric=L2):
        if dim <= 0:
            raise ValueError("Dimension must be positive.")
        self.dim, self.metric = dim, metric
        self._h = lib().orc_flat_new(dim, metric)

    def _vec(self, v):
        v, pv = _f32(v)
        if v.size != self.dim:
            raise ValueError("Vector dimension mismatch.")
        return v, pv

    def add(self, id, vec):
        v, pv = self._vec(vec)
        if lib().orc_flat_add(self._h, id, pv) != 0:
            raise KeyError(f"Vector with id '{id}' already exists.")

    def add_batch(self, X, ids=None):
        X, px = _f32(X)
        pi = None
        if ids is not None:
            ids, pi = _i64(ids)
        lib().orc_flat_add_batch(self._h, X.shape[0], pi, px)

    def upsert(self, id, vec):
        v, pv = self._vec(vec)
        lib().orc_flat_upsert(self._h, id, pv)

    def delete(self, id):
        return bool(lib().orc_flat_delete(self._h, id))

    def count(self):
        return lib().orc_flat_count(self._h)

    def search(self, q, topk, max_scans=-1):
        v, pv = self._vec(q)
        if topk <= 0:
            raise IndexError("topK must be positive.")
        ids, sc = self._out(topk)
        n = lib().orc_flat_search(self._h, pv, topk, max_scans, _ptr(ids, C.c_int64), _ptr(sc, C.c_float))
        return ids[:n].copy(), sc[:n].copy()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_flat_free(self._h)
            self._h = None


class IvfFlatIndex(_Index):
    """IvfFlatVectorIndex.cs"""
    kind = 1

    def __init__(self, dim, metric=L2, nlist=100):
        if dim <= 0:
            raise ValueError("dimension")
        self.dim, self.metric, self.nlist = dim, metric, nlist
        self._h = lib().orc_ivfflat_new(dim, metric, nlist)

    def set_nprobe(self, n):
        lib().orc_ivfflat_set_nprobe(self._h, n)

    def _vec(self, v):
        v, pv = _f32(v)
        if v.size != self.dim:
            raise ValueError("Vector dimension mismatch")
        return v, pv

    def add(self, id, vec):
        v, pv = self._vec(vec)
        lib().orc_ivfflat_add(self._h, id, pv)

    upsert = add

    def add_batch(self, X, ids=None):
        X, px = _f32(X)
        pi = None
        if ids is not None:
            ids, pi = _i64(ids)
        lib().orc_ivfflat_add_batch(self._h, X.shape[0], pi, px)

    def delete(self, id):
        return bool(lib().orc_ivfflat_delete(self._h, id))

    def build(self):
        lib().orc_ivfflat_build(self._h)

    def is_built(self):
        return bool(lib().orc_ivfflat_is_built(self._h))

    def centroids(self):
        if not self.is_built():
            return None
        nc = lib().orc_ivfflat_ncentroids(self._h)
        out = np.zeros((nc, self.dim), dtype=np.float32)
        lib().orc_ivfflat_get_centroids(self._h, _ptr(out, C.c_float))
        return out

    def count(self):
        return lib().orc_ivfflat_count(self._h)

    def adopt(self, centroids, list_offsets, ids, vecs):
        """Take over an index built elsewhere (the GPU index under test): searches then compare the scan alone."""
        c, pc = _f32(centroids)
        off, poff = _i64(list_offsets)
        ids, pids = _i64(ids)
        v, pv = _f32(vecs)
        assert v.shape == (len(ids), self.dim) and off[-1] == len(ids)
        lib().orc_ivfflat_adopt(self._h, c.shape[0], pc, poff, pids, pv)

    def lists(self):
        nc = lib().orc_ivfflat_ncentroids(self._h)
        out = []
        for c in range(nc):
            n = lib().orc_ivfflat_list_size(self._h, c)
            ids = np.zeros(max(n, 1), dtype=np.int64)
            lib().orc_ivfflat_get_list(self._h, c, _ptr(ids, C.c_int64))
            out.append(ids[:n].copy())
        return out

    def search(self, q, topk, max_scans=-1, nprobe=-1):
        v, pv = self._vec(q)
        ids, sc = self._out(topk)
        n = lib().orc_ivfflat_search(self._h, pv, topk, max_scans, nprobe, _ptr(ids, C.c_int64),
                                     _ptr(sc, C.c_float))
        return ids[:n].copy(), sc[:n].copy()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_ivfflat_free(self._h)
            self._h = None


class IvfPqIndex(_Index):
    """IvfPqVectorIndex.cs"""
    kind = 2

    def __init__(self, dim, metric=L2, m=4, k=256, nlist=100):
        if dim % m != 0:
            raise ValueError("Dimension must be divisible by M")
        if k > 256:
            raise ValueError("K must be <= 256 for byte encoding")
        self.dim, self.metric, self.m, self.k, self.nlist = dim, metric, m, k, nlist
        self._h = lib().orc_ivfpq_new(dim, metric, m, k, nlist)

    def add(self, id, vec):
        v, pv = _f32(vec)
        lib().orc_ivfpq_add(self._h, id, pv)

    upsert = add

    def add_batch(self, X, ids=None):
        X, px = _f32(X)
        pi = None
        if ids is not None:
            ids, pi = _i64(ids)
        lib().orc_ivfpq_add_batch(self._h, X.shape[0], pi, px)

    def delete(self, id):
        return bool(lib().orc_ivfpq_delete(self._h, id))

    def build(self):
        lib().orc_ivfpq_build(self._h)

    def is_built(self):
        return bool(lib().orc_ivfpq_is_built(self._h))

    def centroids(self):
        nc = lib().orc_ivfpq_ncentroids(self._h)
        out = np.zeros((nc, self.dim), dtype=np.float32)
        lib().orc_ivfpq_get_centroids(self._h, _ptr(out, C.c_float))
        return out

    def pq(self):
        return ProductQuantizer._borrow(lib().orc_ivfpq_pq(self._h), self.dim, self.m, self.k)

    def lists(self):
        nc = lib().orc_ivfpq_ncentroids(self._h)
        out = []
        for c in range(nc):
            n = lib().orc_ivfpq_list_size(self._h, c)
            ids = np.zeros(max(n, 1), dtype=np.int64)
            codes = np.zeros((max(n, 1), self.m), dtype=np.uint8)
            lib().orc_ivfpq_get_list(self._h, c, _ptr(ids, C.c_int64), _ptr(codes, C.c_uint8))
            out.append((ids[:n].copy(), codes[:n].copy()))
        return out

    def adopt(self, centroids, codebook, list_offsets, ids, codes):
        c, pc = _f32(centroids)
        cb, pcb = _f32(codebook)
        off, poff = _i64(list_offsets)
        ids, pids = _i64(ids)
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        lib().orc_ivfpq_adopt(self._h, c.shape[0], pc, pcb, poff, pids, _ptr(codes, C.c_uint8))

    def search(self, q, topk, nprobe=-1):
        v, pv = _f32(q)
        ids, sc = self._out(topk)
        n = lib().orc_ivfpq_search(self._h, pv, topk, nprobe, _ptr(ids, C.c_int64), _ptr(sc, C.c_float))
        return ids[:n].copy(), sc[:n].copy()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_ivfpq_free(self._h)
            self._h = None


def delta_merge(head, tail, topk):
    """DeltaVectorIndex.cs:95-121; head/tail are (ids, scores)."""
    hi, phi = _i64(head[0])
    hs, phs = _f32(head[1])
    ti, pti = _i64(tail[0])
    ts, pts = _f32(tail[1])
    n = len(hi) + len(ti)
    ids = np.zeros(max(n, 1), dtype=np.int64)
    sc = np.zeros(max(n, 1), dtype=np.float32)
    got = lib().orc_delta_merge(phi, phs, len(hi), pti, pts, len(ti), topk, _ptr(ids, C.c_int64),
                                _ptr(sc, C.c_float))
    return ids[:got].copy(), sc[:got].copy()


def max_threads():
    return lib().orc_max_threads()
