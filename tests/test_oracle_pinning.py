"""Pins the CPU oracle (oracle/pcc_oracle.c) before anything is compared against it.

The reference has no tests or golden vectors for this path and PCL/FLANN are not
installable here, so the pins are: (1) committed golden vectors produced by OpenCV's bundled
FLANN KDTreeSingleIndex (tests/golden/make_golden.py), (2) the same library called live when
cv2 is importable, (3) scipy cKDTree at set level, (4) hand-derived known answers.
"""
import numpy as np
import pytest

import oracle
from conftest import bits
from pointcloudcomparator_b200 import synth


def test_golden_knn(golden):
    ref, qry = golden["ref"], golden["qry"]
    tree = oracle.KdTree(ref)
    for k in (1, 16, 50):
        for idx, d2, _ in (tree.knn(qry, k), oracle.brute_knn(ref, qry, k)):
            assert np.array_equal(idx, golden[f"knn{k}_idx"])
            assert np.array_equal(bits(d2), bits(golden[f"knn{k}_d2"]))
    ti, td, _ = oracle.KdTree(golden["uref"]).knn(golden["uqry"], 16)
    assert np.array_equal(ti, golden["uknn16_idx"]) and np.array_equal(bits(td), bits(golden["uknn16_d2"]))


def test_golden_radius(golden):
    ref, qry, r = golden["ref"], golden["qry"], float(golden["radius"])
    for off, idx, d2 in (oracle.KdTree(ref).radius(qry, r), oracle.brute_radius(ref, qry, r)):
        assert np.array_equal(off, golden["rad_off"])
        assert np.array_equal(idx, golden["rad_idx"])
        assert np.array_equal(bits(d2), bits(golden["rad_d2"]))


def test_live_opencv_flann():
    cv2 = pytest.importorskip("cv2")
    ref = synth.room(60000, 1001)
    qry = synth.sweep_queries(ref, 3000, seed=9)
    fl = cv2.flann_Index(ref, dict(algorithm=4, leaf_max_size=15))
    ci, cd = fl.knnSearch(qry, 16, params=dict(checks=-1, eps=0.0, sorted=True))
    ti, td, _ = oracle.KdTree(ref).knn(qry, 16)
    assert np.array_equal(ci, ti) and np.array_equal(bits(cd), bits(td))


def test_lattice_ties_distance_multiset_matches_flann():
    """FLANN keeps visit order among equal d2; the canonical order is (d2, idx).  Distances agree always."""
    cv2 = pytest.importorskip("cv2")
    g = np.arange(12, dtype=np.float32)
    ref = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    qry = np.ascontiguousarray(ref[::7])
    fl = cv2.flann_Index(ref, dict(algorithm=4, leaf_max_size=15))
    _, cd = fl.knnSearch(qry, 8, params=dict(checks=-1, eps=0.0, sorted=True))
    ti, td, _ = oracle.KdTree(ref).knn(qry, 8)
    bi, bd, _ = oracle.brute_knn(ref, qry, 8)
    assert np.array_equal(bits(cd), bits(td))
    assert np.array_equal(ti, bi) and np.array_equal(bits(td), bits(bd))
    # canonical order: rows sorted by (d2, idx)
    key = td.astype(np.float64) * 1e7 + ti
    assert (np.diff(key, axis=1) > 0).all()


def test_scipy_set_level():
    from scipy.spatial import cKDTree
    ref = synth.uniform(30000, 5001, extent=2.0)
    qry = synth.sweep_queries(ref, 1000, seed=3, sigma=0.05)
    ti, _, _ = oracle.KdTree(ref).knn(qry, 10)
    _, si = cKDTree(ref.astype(np.float64)).query(qry.astype(np.float64), 10)
    same = np.mean([set(a) == set(b) for a, b in zip(ti, si)])
    assert same > 0.999   # float64 vs fp32 rounding may swap a near-tie at rank 10


def test_nan_rows_k_gt_n_and_duplicates():
    ref = np.array([[0, 0, 0], [np.nan, 0, 0], [1, 0, 0], [1, 0, 0], [0, 2, 0], [np.inf, 1, 1]], np.float32)
    qry = np.array([[0.9, 0, 0], [np.nan, 0, 0]], np.float32)
    for idx, d2, keff in (oracle.KdTree(ref).knn(qry, 6), oracle.brute_knn(ref, qry, 6)):
        assert keff == 4
        assert idx[0].tolist() == [2, 3, 0, 4, -1, -1]          # duplicates tie -> lower index first
        assert np.isinf(d2[0, 4:]).all() and (idx[1] == -1).all()
    assert oracle.KdTree(ref).size == 4


def test_radius_strict_boundary_and_max_nn():
    ref = np.array([[0, 0, 0], [0.5, 0, 0], [1.0, 0, 0], [0, 0.25, 0]], np.float32)
    qry = np.zeros((1, 3), np.float32)
    for fn in (lambda r, m: oracle.KdTree(ref).radius(qry, r, m), lambda r, m: oracle.brute_radius(ref, qry, r, m)):
        off, idx, d2 = fn(0.5, 0)            # d2 == r2 exactly -> excluded (strict <)
        assert idx.tolist() == [0, 3]
        off, idx, d2 = fn(0.5000001, 0)
        assert idx.tolist() == [0, 3, 1]
        off, idx, d2 = fn(2.0, 2)            # max_nn keeps the closest
        assert idx.tolist() == [0, 3]
        off, idx, d2 = fn(1e-3, 0)
        assert idx.tolist() == [0]


def test_normals_planar_patch_and_degenerate():
    rng = np.random.default_rng(0)
    p = np.zeros((400, 3), np.float32)
    p[:, :2] = rng.random((400, 2)) * 0.2 + 1.0
    p[:, 2] = 3.0
    n = oracle.normals_knn(p, 20)
    assert np.allclose(np.abs(n[:, 2]), 1.0, atol=1e-3) and (n[:, 2] < 0).all()      # flipped towards the origin; fp32 single-pass covariance is only ~1e-3 accurate here (PCL behaviour)
    assert np.all(n[:, 3] < 5e-3)
    # fewer than 3 neighbours -> NaN
    off, idx, _ = oracle.KdTree(p[:2]).radius(p[:2], 10.0)
    out = oracle.normals_from_lists(p[:2], p[:2], off, idx)
    assert np.isnan(out).all()


def test_sor_known_answer():
    g = np.arange(10, dtype=np.float32)
    ref = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    ref = np.concatenate([ref, np.array([[30, 30, 30]], np.float32)])
    r = oracle.sor(ref, mean_k=6, std_mul=1.5)
    assert not r["keep"][-1] and r["keep"][555]
    assert np.isclose(r["distances"][555], 1.0)       # interior lattice point: 6 neighbours at distance 1


def test_ece_two_blobs_gap_and_size_filter():
    rng = np.random.default_rng(1)
    a = rng.random((300, 3)).astype(np.float32) * 0.1
    b = a + np.array([0.1 + 0.06, 0, 0], np.float32)          # nearest gap 0.06
    small = np.array([[5, 5, 5], [5.01, 5, 5]], np.float32)
    pts = np.concatenate([a, b, small])
    lab, sizes = oracle.KdTree(pts).ece(0.05, 100, 250000)
    assert sizes.tolist() == [300, 300] and set(lab[:300]) == {0} and set(lab[300:600]) == {1} and (lab[600:] == -1).all()
    lab, sizes = oracle.KdTree(pts).ece(0.07, 100, 250000)
    assert sizes.tolist() == [600]
    lab, sizes = oracle.KdTree(pts).ece(0.07, 100, 500)        # over max -> dropped whole
    assert sizes.tolist() == [] and (lab == -1).all()


def test_ece_scene_known_cluster_count():
    pts, ids = synth.scene(80000, 3001, extent=6.0, n_objects=25)
    lab, sizes = oracle.KdTree(pts).ece(0.05, 100, 250000)
    assert len(sizes) == 25
    # one label per object
    for o in range(25):
        assert len(set(lab[ids == o])) == 1
    assert (np.diff(sizes) <= 0).all()


def test_icp_recovers_rigid_transform():
    src, tgt, T = synth.icp_pair(40000, 4001, size=(5, 5, 3))
    r = oracle.icp(src, tgt, 20)
    assert r["converged"] and r["iterations"] <= 20
    Tinv = np.linalg.inv(T)
    assert np.allclose(r["T"], Tinv, atol=2e-4)
    assert r["fitness"] < 1e-5


def test_umeyama_reflection_guard():
    rng = np.random.default_rng(5)
    s = rng.normal(size=(50, 3))
    R = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]])
    t = s @ R.T + np.array([1, 2, 3.0])
    sums = np.zeros(16)
    sums[0:3], sums[3:6], sums[6:15] = s.sum(0), t.sum(0), (t.T @ s).reshape(-1)
    T = oracle.umeyama_from_sums(sums, 50)
    assert np.allclose(T[:3, :3], R, atol=1e-6) and np.allclose(T[:3, 3], [1, 2, 3], atol=1e-5)


def test_first_within():
    pts = np.array([[0, 0, 0], [0.04, 0, 0], [0.01, 0, 0]], np.float32)
    q = np.array([[0.03, 0, 0], [1, 1, 1]], np.float32)
    assert oracle.first_within(pts, q, 0.05).tolist() == [0, -1]


def test_voxel_grid_known_answers():
    # two points in one 1 m voxel, one alone; colour averaged and truncated; output ordered by voxel index (x fastest)
    rows = np.zeros((3, 8), np.float32)
    rows[:, :3] = [[2.25, 0.5, 0.5], [0.25, 0.25, 0.25], [0.75, 0.5, 0.75]]
    rows[:, 3] = 1.0
    col = rows[:, 4:5].view(np.uint8)                       # BGRA bytes
    col[0] = [10, 20, 30, 0]; col[1] = [0, 100, 255, 0]; col[2] = [1, 101, 0, 0]
    out = oracle.voxel_grid(rows, 1.0, rgb_offset_floats=4)
    assert out.shape[0] == 2
    assert np.allclose(out[0, :3], [0.5, 0.375, 0.5]) and np.allclose(out[1, :3], [2.25, 0.5, 0.5])
    assert out[0, 4:5].view(np.uint8).tolist() == [0, 100, 127, 0] and out[1, 4:5].view(np.uint8).tolist() == [10, 20, 30, 0]
    assert (out[:, 3] == 1.0).all()
    # distinct-voxel count equals numpy's, NaN rows are ignored
    p = synth.room(50000, 1001, stride4=True)
    p[7, 0] = np.nan
    o = oracle.voxel_grid(p, 0.025)
    fin = np.isfinite(p[:, :3]).all(1)
    ijk = np.floor(p[fin, :3] * np.float32(1 / np.float32(0.025))).astype(np.int64)
    assert o.shape[0] == len(np.unique(ijk, axis=0))
    vox = np.floor(o[:, :3] * np.float32(1 / np.float32(0.025)) + 1e-4).astype(np.int64)
    assert len(np.unique(vox, axis=0)) >= 0.999 * o.shape[0]      # every centroid stays inside its own voxel (up to fp32 rounding on faces)


def test_descriptor_nn_matches_opencv_flann_32d():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    ref = rng.random((700, 32), dtype=np.float32) * 0.3
    qry = (ref[rng.integers(0, 700, 400)] + rng.normal(0, 0.02, (400, 32))).astype(np.float32)
    fl = cv2.flann_Index(ref, dict(algorithm=4, leaf_max_size=15))          # same index family KdTreeFLANN<Histogram<32>> builds
    ci, cd = fl.knnSearch(qry, 1, params=dict(checks=-1, eps=0.0, sorted=True))
    oi, od = oracle.descriptor_nn(ref, qry)
    # OpenCV's FLANN uses the 4-way unrolled cvflann::L2 functor (sums four squared differences per step); PCL's KdTreeFLANN
    # uses L2_Simple (one dimension at a time), which the oracle restates.  Same neighbours, distances equal to a few ulp.
    assert np.array_equal(ci[:, 0], oi) and np.allclose(cd[:, 0], od, rtol=1e-6, atol=0)
    ref[5, 7] = np.nan; qry[3, 0] = np.inf
    oi, od = oracle.descriptor_nn(ref, qry)
    assert (oi != 5).all() and oi[3] == -1 and np.isinf(od[3])
