"""Region-growing grow phase: the product's host C++ (pcc_region_growing, pure host code inside the
C-ABI library -- callable without a GPU) against the oracle's independent C restatement, on the same neighbour table and normals."""
import numpy as np
import pytest

import oracle
from pointcloudcomparator_b200 import synth
from pointcloudcomparator_b200.search import region_growing


def _inputs(n, k, seed):
    pts = synth.room(n, seed)
    tree = oracle.KdTree(pts)
    nbr, _, _ = tree.knn(pts, k)
    off = np.arange(n + 1, dtype=np.int64) * k
    normals = oracle.normals_from_lists(pts, pts, off, nbr)
    return pts, nbr, normals


def test_region_growing_matches_oracle_reference_configuration():
    pts, nbr, normals = _inputs(30000, 100, 1001)                              # k = 100, 3 degrees, curvature 1, sizes 50..1e6
    lab, nc = region_growing(nbr, normals)
    olab, onc = oracle.region_growing(nbr, normals)
    assert nc == onc and nc >= 6 and np.array_equal(lab, olab)
    big = np.bincount(lab[lab >= 0])
    assert big.max() > 2000                                                     # walls / floor come out as large smooth regions


@pytest.mark.parametrize("theta_deg,curv,k,mn", [(1.0, 0.02, 30, 10), (8.0, 1.0, 20, 1), (3.0, 0.0005, 50, 50)])
def test_region_growing_parameter_sweep(theta_deg, curv, k, mn):
    pts, nbr, normals = _inputs(12000, k, 7)
    normals[::997] = np.nan                                                     # NaN normals: |dot| < cos is false, as in PCL
    a = region_growing(nbr, normals, theta_deg / 180 * np.pi, curv, mn, 5000)
    b = oracle.region_growing(nbr, normals, theta_deg / 180 * np.pi, curv, mn, 5000)
    assert a[1] == b[1] and np.array_equal(a[0], b[0])


def test_region_growing_short_rows_and_single_point():
    nbr = np.array([[0, 1, -1], [1, 0, 2], [2, 1, -1], [3, -1, -1]], np.int32)
    normals = np.array([[0, 0, 1, 0.1], [0, 0, 1, 0.0], [0, 0.01, 1, 0.2], [1, 0, 0, 0.05]], np.float32)
    normals[:, :3] /= np.linalg.norm(normals[:, :3], axis=1, keepdims=True)
    lab, nc = region_growing(nbr, normals, 0.1, 1.0, 1, 10)
    olab, onc = oracle.region_growing(nbr, normals, 0.1, 1.0, 1, 10)
    assert nc == onc == 2 and lab.tolist() == olab.tolist() == [0, 0, 0, 1]


@pytest.mark.gpu
def test_region_growing_pipeline_on_gpu_tables():
    """region_growing_segmentation (src/segmentation.cpp:232-271) end to end: N x 100 table and 50-NN normals from the GPU,
    grow phase in the C-ABI library; the table must equal the oracle's bit for bit, and the grow phase must agree with the
    oracle's on the same (GPU-made) inputs."""
    from pointcloudcomparator_b200.search import GridSearch
    pts = synth.room(40000, 31)
    s = GridSearch().setInputCloud(pts, k_hint=100)
    nbr, _, _ = s.nearestKSearch(None, 100)
    onbr, _, _ = oracle.KdTree(pts).knn(pts, 100)
    assert np.array_equal(nbr, onbr)
    normals = s.normalsKnn(None, 50)
    lab, nc = region_growing(nbr, normals)
    olab, onc = oracle.region_growing(nbr, normals)
    assert nc == onc and nc >= 6 and np.array_equal(lab, olab)
    # smooth regions do not straddle perpendicular room faces: inside a kept region normals agree up to sign within ~45 degrees of its seed
    for c in range(min(nc, 5)):
        m = normals[lab == c, :3]
        assert (np.abs(m @ m[0]) > 0.7).mean() > 0.95


def test_region_growing_known_answer_two_perpendicular_planes():
    """Hand-derived answer: a floor (normal +z) and a wall (normal +x) meeting in an edge, exact normals, zero curvature.
    The 3-degree smoothness test never lets a region cross the edge, every point has in-plane neighbours, so both
    restatements must return exactly the two planes, the floor (lowest indices, first seed) as cluster 0."""
    g = np.arange(40, dtype=np.float32) * np.float32(0.05)
    floor = np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)
    pts = np.concatenate([np.c_[floor, np.zeros(len(floor))], np.c_[np.zeros(len(floor)), floor[:, 0], floor[:, 1] + 0.05]]).astype(np.float32)
    normals = np.zeros((len(pts), 4), np.float32)
    normals[: len(floor), 2] = 1.0
    normals[len(floor):, 0] = 1.0
    nbr, _, _ = oracle.KdTree(pts).knn(pts, 12)
    for fn in (region_growing, oracle.region_growing):
        lab, nc = fn(nbr, normals, 3.0 / 180 * np.pi, 1.0, 50, 100000)
        assert nc == 2
        assert (lab[: len(floor)] == 0).all() and (lab[len(floor):] == 1).all()


# ---------------------------------------------------------------------------------------------------------------------
# RegionGrowingRGB (color_growing_segmentation, src/segmentation.cpp:161-216; per matched cluster at src/comparator.cpp:1457-1460)

def _coloured_patches(n, seed, n_patches=6, jitter=2, speckle=0.0):
    """A flat 1 x 1 m patch-work: vertical stripes of distinct base colours (+- jitter per channel) and optional random speckle."""
    rng = np.random.default_rng(seed)
    p = np.zeros((n, 3), np.float32)
    p[:, :2] = rng.random((n, 2))
    base = rng.integers(30, 226, (n_patches, 3))
    stripe = np.minimum((p[:, 0] * n_patches).astype(int), n_patches - 1)
    col = np.clip(base[stripe] + rng.integers(-jitter, jitter + 1, (n, 3)), 0, 255)
    sp = rng.random(n) < speckle
    col[sp] = rng.integers(0, 256, (int(sp.sum()), 3))
    rgba = (col[:, 0].astype(np.uint32) << 16) | (col[:, 1].astype(np.uint32) << 8) | col[:, 2].astype(np.uint32)
    return p, rgba, stripe


def _table(p, k):
    idx, d2, _ = oracle.KdTree(p).knn(p, k)
    return idx, d2


def test_region_growing_rgb_matches_python_oracle_reference_configuration():
    from oracle import rgb_region_growing
    from pointcloudcomparator_b200.search import region_growing_rgb
    p, rgba, stripe = _coloured_patches(6000, 3, n_patches=5, jitter=1)
    nb, nd = _table(p, 100)
    lab, nc = region_growing_rgb(nb, nd, rgba)                               # distance 10, point colour 6, region colour 5, min 200
    olab, onc = rgb_region_growing.extract(nb, nd, rgba)
    assert nc == onc and np.array_equal(lab, olab)
    assert nc == 5                                                            # one cluster per stripe: the known answer
    for s in range(5):
        assert len(set(lab[stripe == s].tolist())) == 1


@pytest.mark.parametrize("jitter,speckle,pthr,rthr,dthr,mn,grow", [(4, 0.0, 6.0, 5.0, 10.0, 200, 30), (3, 0.05, 6.0, 5.0, 10.0, 50, 30), (6, 0.02, 9.0, 12.0, 0.02, 20, 10),
                                                              (2, 0.10, 4.0, 40.0, 10.0, 400, 100)])
def test_region_growing_rgb_parameter_sweep(jitter, speckle, pthr, rthr, dthr, mn, grow):
    """Speckle makes hundreds of tiny segments, so the neighbour heaps, the colour merge and the fold of small regions all run."""
    from oracle import rgb_region_growing
    from pointcloudcomparator_b200.search import region_growing_rgb
    p, rgba, _ = _coloured_patches(3000, 11 + jitter, n_patches=7, jitter=jitter, speckle=speckle)
    nb, nd = _table(p, 40)
    a = region_growing_rgb(nb, nd, rgba, dthr, pthr, rthr, grow, mn)
    b = rgb_region_growing.extract(nb, nd, rgba, dthr, pthr, rthr, grow, None, mn)
    assert a[1] == b[1] and np.array_equal(a[0], b[0])
    assert (np.bincount(a[0][a[0] >= 0]) >= mn).all() if a[1] else True


def test_region_growing_rgb_small_known_answers():
    from pointcloudcomparator_b200.search import region_growing_rgb
    # four points on a line, two colours; k = 3; everything within distance 10 -> colours decide
    nb = np.array([[0, 1, 2], [1, 0, 2], [2, 3, 1], [3, 2, 1]], np.int32)
    nd = np.array([[0, 1, 4], [0, 1, 1], [0, 1, 1], [0, 1, 4]], np.float32)
    red, blue = 0xC80000, 0x0000C8
    lab, nc = region_growing_rgb(nb, nd, np.array([red, red, blue, blue], np.uint32), 10.0, 6.0, 5.0, 30, 1)
    assert nc == 2 and lab.tolist() == [0, 0, 1, 1]
    # min size 3: each 2-point region folds into its neighbour -> one cluster of 4
    lab, nc = region_growing_rgb(nb, nd, np.array([red, red, blue, blue], np.uint32), 10.0, 6.0, 5.0, 30, 3)
    assert nc == 1 and lab.tolist() == [0, 0, 0, 0]
    # a colour step of exactly the point threshold is accepted (PCL rejects on ">"): 6^2 = 36 = (6, 0, 0) squared
    lab, nc = region_growing_rgb(nb, nd, np.array([0x640000, 0x6A0000, 0x700000, 0x760000], np.uint32), 10.0, 6.0, 5.0, 30, 1)
    assert nc == 1
    lab, nc = region_growing_rgb(nb, nd, np.array([0x640000, 0x6B0000, 0x720000, 0x790000], np.uint32), 10.0, 6.0, 0.5, 30, 1)
    assert nc == 4


@pytest.mark.gpu
def test_region_growing_rgb_on_gpu_table():
    """The N x 100 table and its squared distances come from the GPU (pcc_knn with q == NULL); clusters must equal the oracle's on
    the oracle's own table, i.e. the table is bit-exact and the host grow / merge agrees."""
    from oracle import rgb_region_growing
    from pointcloudcomparator_b200.search import GridSearch, region_growing_rgb
    p, rgba, _ = _coloured_patches(5000, 21, n_patches=4, jitter=2, speckle=0.01)
    s = GridSearch().setInputCloud(p, k_hint=100)
    nb, nd, _ = s.nearestKSearch(None, 100)
    onb, ond = _table(p, 100)
    assert np.array_equal(nb, onb) and np.array_equal(nd.view(np.uint32), ond.view(np.uint32))
    lab, nc = region_growing_rgb(nb, nd, rgba)
    olab, onc = rgb_region_growing.extract(onb, ond, rgba)
    assert nc == onc and np.array_equal(lab, olab)
