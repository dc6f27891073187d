"""ctypes front-end of the CPU oracle (oracle/pcc_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (pointcloudcomparator_b200) never
imports this module.  Parity status: see the header of pcc_oracle.c ("parity unpinned" against
the reference itself; pinned against OpenCV-FLANN / scipy stand-ins).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpcc_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile libpcc_oracle.so with oracle/Makefile (gcc only, no deps)."""
    src = os.path.join(_HERE, "pcc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], env={**os.environ, "CC": "/usr/bin/gcc"})
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, i64, i32, f32p, i32p, i64p, dbl = C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.c_double
        L.orc_tree_build.restype = vp
        L.orc_tree_build.argtypes = [f32p, i64, i32, i32]
        L.orc_tree_free.argtypes = [vp]
        L.orc_tree_size.restype = i64
        L.orc_tree_size.argtypes = [vp]
        L.orc_tree_knn.argtypes = [vp, f32p, i64, i32, i32, i32p, f32p, i32]
        L.orc_tree_radius_count.argtypes = [vp, f32p, i64, i32, dbl, C.c_uint, i64p, i32]
        L.orc_tree_radius_fill.argtypes = [vp, f32p, i64, i32, dbl, C.c_uint, i64p, i32p, f32p, i32]
        L.orc_brute_knn.argtypes = [f32p, i64, i32, f32p, i64, i32, i32, i32p, f32p, i32]
        L.orc_brute_radius.argtypes = [f32p, i64, i32, f32p, i64, i32, dbl, C.c_uint, i64p, i32p, f32p]
        L.orc_normals_from_lists.argtypes = [f32p, i32, f32p, i64, i32, i64p, i32p, C.c_float, C.c_float, C.c_float, f32p]
        L.orc_sor_mean_dist.argtypes = [f32p, i64, i32, i32, f32p]
        L.orc_sor_threshold.restype = i64
        L.orc_sor_threshold.argtypes = [f32p, C.POINTER(C.c_uint8), i64, dbl, C.POINTER(dbl), C.POINTER(dbl), C.POINTER(dbl), C.POINTER(C.c_uint8)]
        L.orc_ece.restype = i64
        L.orc_ece.argtypes = [vp, f32p, i64, i32, dbl, i64, i64, i32p, i64p, i64]
        L.orc_umeyama_from_sums.argtypes = [C.POINTER(dbl), C.POINTER(dbl), C.POINTER(dbl), dbl, f32p]
        L.orc_icp_pass.restype = i64
        L.orc_icp_pass.argtypes = [vp, f32p, i32, f32p, i64, C.POINTER(dbl), i32p, f32p, i32]
        L.orc_icp.argtypes = [f32p, i64, i32, f32p, i64, i32, i32, f32p, C.POINTER(i32), C.POINTER(dbl), C.POINTER(i32), C.POINTER(dbl), i32]
        L.orc_first_within.argtypes = [f32p, i64, i32, f32p, i64, i32, dbl, i32p]
        L.orc_voxel_grid.restype = i64
        L.orc_voxel_grid.argtypes = [f32p, i64, i32, i32, f32p, i32, f32p]
        L.orc_descriptor_nn.argtypes = [f32p, i64, f32p, i64, i32, i32p, f32p]
        L.orc_region_growing.restype = i64
        L.orc_region_growing.argtypes = [i32p, i64, i32, f32p, C.c_float, C.c_float, i64, i64, i32p]
        L.orc_num_threads.restype = i32
        _lib = L
    return _lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] >= 3
    return a


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def num_threads() -> int:
    return int(lib().orc_num_threads())


class KdTree:
    """Restated pcl::KdTreeFLANN / flann::KDTreeSingleIndex (leaf 15) over the finite rows of pts[n, >=3]."""

    def __init__(self, pts, leaf_max: int = 15):
        self.pts = _f32(pts)
        self._h = lib().orc_tree_build(_p(self.pts, C.c_float), self.pts.shape[0], self.pts.shape[1], leaf_max)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:          # _lib is None during interpreter shutdown
            _lib.orc_tree_free(h)

    @property
    def size(self) -> int:
        return int(lib().orc_tree_size(self._h))

    def knn(self, q, k: int, threads: int = 0):
        q = _f32(q)
        idx = np.empty((q.shape[0], k), np.int32)
        d2 = np.empty((q.shape[0], k), np.float32)
        keff = lib().orc_tree_knn(self._h, _p(q, C.c_float), q.shape[0], q.shape[1], k, _p(idx, C.c_int32), _p(d2, C.c_float), threads)
        return idx, d2, keff

    def radius(self, q, radius: float, max_nn: int = 0, threads: int = 0):
        q = _f32(q)
        off = np.zeros(q.shape[0] + 1, np.int64)
        lib().orc_tree_radius_count(self._h, _p(q, C.c_float), q.shape[0], q.shape[1], radius, max_nn, _p(off, C.c_int64), threads)
        idx = np.empty(max(int(off[-1]), 1), np.int32)
        d2 = np.empty(max(int(off[-1]), 1), np.float32)
        lib().orc_tree_radius_fill(self._h, _p(q, C.c_float), q.shape[0], q.shape[1], radius, max_nn, _p(off, C.c_int64), _p(idx, C.c_int32), _p(d2, C.c_float), threads)
        return off, idx[: off[-1]], d2[: off[-1]]

    def ece(self, tolerance: float, min_size: int, max_size: int):
        n = self.pts.shape[0]
        labels = np.empty(n, np.int32)
        sizes = np.zeros(max(n, 1), np.int64)
        nc = lib().orc_ece(self._h, _p(self.pts, C.c_float), n, self.pts.shape[1], tolerance, min_size, max_size, _p(labels, C.c_int32), _p(sizes, C.c_int64), sizes.shape[0])
        return labels, sizes[:nc].copy()

    def icp_pass(self, cur, threads: int = 0):
        cur = np.ascontiguousarray(cur, np.float32)
        assert cur.shape[1] == 3
        sums = np.zeros(16, np.float64)
        ci = np.empty(cur.shape[0], np.int32)
        cd = np.empty(cur.shape[0], np.float32)
        cnt = lib().orc_icp_pass(self._h, _p(self.pts, C.c_float), self.pts.shape[1], _p(cur, C.c_float), cur.shape[0], _p(sums, C.c_double), _p(ci, C.c_int32), _p(cd, C.c_float), threads)
        return int(cnt), sums, ci, cd


def brute_knn(pts, q, k: int, threads: int = 0):
    pts, q = _f32(pts), _f32(q)
    idx = np.empty((q.shape[0], k), np.int32)
    d2 = np.empty((q.shape[0], k), np.float32)
    keff = lib().orc_brute_knn(_p(pts, C.c_float), pts.shape[0], pts.shape[1], _p(q, C.c_float), q.shape[0], q.shape[1], k, _p(idx, C.c_int32), _p(d2, C.c_float), threads)
    return idx, d2, keff


def brute_radius(pts, q, radius: float, max_nn: int = 0):
    pts, q = _f32(pts), _f32(q)
    off = np.zeros(q.shape[0] + 1, np.int64)
    a = (_p(pts, C.c_float), pts.shape[0], pts.shape[1], _p(q, C.c_float), q.shape[0], q.shape[1], radius, max_nn, _p(off, C.c_int64))
    lib().orc_brute_radius(*a, None, None)
    idx = np.empty(max(int(off[-1]), 1), np.int32)
    d2 = np.empty(max(int(off[-1]), 1), np.float32)
    lib().orc_brute_radius(*a, _p(idx, C.c_int32), _p(d2, C.c_float))
    return off, idx[: off[-1]], d2[: off[-1]]


def normals_from_lists(pts, qpts, offsets, nbr, viewpoint=(0.0, 0.0, 0.0)):
    pts, qpts = _f32(pts), _f32(qpts)
    offsets = np.ascontiguousarray(offsets, np.int64)
    nbr = np.ascontiguousarray(nbr, np.int32).reshape(-1)
    out = np.empty((qpts.shape[0], 4), np.float32)
    lib().orc_normals_from_lists(_p(pts, C.c_float), pts.shape[1], _p(qpts, C.c_float), qpts.shape[0], qpts.shape[1], _p(offsets, C.c_int64), _p(nbr, C.c_int32), *[float(v) for v in viewpoint], _p(out, C.c_float))
    return out


def normals_knn(pts, k: int, viewpoint=(0.0, 0.0, 0.0), tree: KdTree | None = None):
    """NormalEstimation with setKSearch(k) over the cloud itself (src/segmentation.cpp:236-240)."""
    tree = tree or KdTree(pts)
    idx, _, _ = tree.knn(pts, k)
    off = np.arange(idx.shape[0] + 1, dtype=np.int64) * k
    return normals_from_lists(pts, pts, off, idx, viewpoint)


def normals_radius(pts, radius: float, viewpoint=(0.0, 0.0, 0.0), tree: KdTree | None = None):
    """NormalEstimation with setRadiusSearch(r) (src/comparator.cpp:628-635)."""
    tree = tree or KdTree(pts)
    off, idx, _ = tree.radius(pts, radius)
    return normals_from_lists(pts, pts, off, idx, viewpoint)


def sor_mean_dist(d2_rows, mean_k: int):
    d2_rows = np.ascontiguousarray(d2_rows, np.float32)
    out = np.empty(d2_rows.shape[0], np.float32)
    lib().orc_sor_mean_dist(_p(d2_rows, C.c_float), d2_rows.shape[0], d2_rows.shape[1], mean_k, _p(out, C.c_float))
    return out


def sor_threshold(distances, std_mul: float, valid=None):
    distances = np.ascontiguousarray(distances, np.float32)
    keep = np.empty(distances.shape[0], np.uint8)
    m, s, t = C.c_double(), C.c_double(), C.c_double()
    v = None if valid is None else _p(np.ascontiguousarray(valid, np.uint8), C.c_uint8)
    kept = lib().orc_sor_threshold(_p(distances, C.c_float), v, distances.shape[0], std_mul, C.byref(m), C.byref(s), C.byref(t), _p(keep, C.c_uint8))
    return dict(mean=m.value, stddev=s.value, threshold=t.value, kept=int(kept), keep=keep.astype(bool))


def sor(pts, mean_k: int = 50, std_mul: float = 1.5, tree: KdTree | None = None):
    """StatisticalOutlierRemoval as configured at src/comparator.cpp:1523-1527."""
    tree = tree or KdTree(pts)
    _, d2, _ = tree.knn(pts, mean_k + 1)
    dist = sor_mean_dist(d2, mean_k)
    valid = np.isfinite(np.asarray(pts, np.float32)[:, :3]).all(1).astype(np.uint8)
    r = sor_threshold(dist, std_mul, valid)
    r["distances"] = dist
    return r


def umeyama_from_sums(sums, n):
    sums = np.ascontiguousarray(sums, np.float64)
    T = np.empty(16, np.float32)
    lib().orc_umeyama_from_sums(_p(sums[0:3].copy(), C.c_double), _p(sums[3:6].copy(), C.c_double), _p(sums[6:15].copy(), C.c_double), float(n), _p(T, C.c_float))
    return T.reshape(4, 4)


def icp(src, tgt, max_iter: int = 20, threads: int = 0):
    """IterativeClosestPoint as configured at src/comparator.cpp:1089-1110."""
    src, tgt = _f32(src), _f32(tgt)
    T = np.empty(16, np.float32)
    conv, it, fit = C.c_int(), C.c_int(), C.c_double()
    mse = np.zeros(max_iter, np.float64)
    lib().orc_icp(_p(src, C.c_float), src.shape[0], src.shape[1], _p(tgt, C.c_float), tgt.shape[0], tgt.shape[1], max_iter, _p(T, C.c_float), C.byref(conv), C.byref(fit), C.byref(it), _p(mse, C.c_double), threads)
    return dict(T=T.reshape(4, 4), converged=bool(conv.value), fitness=fit.value, iterations=it.value, mse=mse[: it.value])


def first_within(pts, q, thr: float):
    pts, q = _f32(pts), _f32(q)
    out = np.empty(q.shape[0], np.int32)
    lib().orc_first_within(_p(pts, C.c_float), pts.shape[0], pts.shape[1], _p(q, C.c_float), q.shape[0], q.shape[1], thr, _p(out, C.c_int32))
    return out


def voxel_grid(rows, leaf, rgb_offset_floats: int = -1, min_points: int = 0):
    """pcl::VoxelGrid::applyFilter over rows[n, stride] (float32; colour as a packed BGRA word at rgb_offset_floats)."""
    rows = np.ascontiguousarray(rows, np.float32)
    leaf = np.ascontiguousarray(np.broadcast_to(np.asarray(leaf, np.float32), (3,)))
    out = np.zeros_like(rows)
    n = lib().orc_voxel_grid(_p(rows, C.c_float), rows.shape[0], rows.shape[1], rgb_offset_floats, _p(leaf, C.c_float), min_points, _p(out, C.c_float))
    return out[:n]


def descriptor_nn(ref, qry, dims: int | None = None):
    """matchRIFTFeaturesKnn's inner search: nearest reference descriptor of every query descriptor (index or -1, d2), over the
    first `dims` floats of a row (None = all; 3 = PCL 1.7's DefaultPointRepresentation clamp for an unregistered Histogram<32>)."""
    ref, qry = np.asarray(ref, np.float32), np.asarray(qry, np.float32)
    if dims is not None:
        ref, qry = ref[:, :dims], qry[:, :dims]
    ref, qry = np.ascontiguousarray(ref), np.ascontiguousarray(qry)
    idx, d2 = np.empty(qry.shape[0], np.int32), np.empty(qry.shape[0], np.float32)
    lib().orc_descriptor_nn(_p(ref, C.c_float), ref.shape[0], _p(qry, C.c_float), qry.shape[0], ref.shape[1], _p(idx, C.c_int32), _p(d2, C.c_float))
    return idx, d2


def region_growing(neighbours, normals, smoothness_rad=3.0 / 180.0 * np.pi, curvature_threshold=1.0, min_size=50, max_size=1000000):
    """pcl::RegionGrowing::extract over a neighbour table [n, k] and normals [n, 4] (src/segmentation.cpp:249-271)."""
    nb = np.ascontiguousarray(neighbours, np.int32)
    nm = np.ascontiguousarray(normals, np.float32)
    labels = np.empty(nb.shape[0], np.int32)
    nc = lib().orc_region_growing(_p(nb, C.c_int32), nb.shape[0], nb.shape[1], _p(nm, C.c_float), smoothness_rad, curvature_threshold, min_size, max_size, _p(labels, C.c_int32))
    return labels, int(nc)
