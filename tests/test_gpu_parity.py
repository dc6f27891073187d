"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded
inputs and against the committed golden vectors.  Bar: bit-exact indices / fp32 squared distances / labels;
1e-5 relative (stated per test) for floating-point reductions.  Run on the B200 box with `-m gpu`."""
import numpy as np
import pytest

import oracle
from conftest import bits
from pointcloudcomparator_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def GS():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pointcloudcomparator_b200.search import GridSearch
    return GridSearch


def assert_knn_equal(gi, gd, oi, od):
    assert np.array_equal(gi, oi), f"index mismatch in {np.argwhere(gi != oi)[:5].tolist()}"
    assert np.array_equal(bits(gd), bits(od))


# ---------------------------------------------------------------- kNN
def test_golden_knn_and_radius(GS, golden):
    s = GS().setInputCloud(golden["ref"])
    for k in (1, 16, 50):
        gi, gd, keff = s.nearestKSearch(golden["qry"], k)
        assert keff == k
        assert_knn_equal(gi, gd, golden[f"knn{k}_idx"], golden[f"knn{k}_d2"])
    off, idx, d2 = s.radiusSearch(golden["qry"], float(golden["radius"]))
    assert np.array_equal(off, golden["rad_off"]) and np.array_equal(idx, golden["rad_idx"]) and np.array_equal(bits(d2), bits(golden["rad_d2"]))
    u = GS().setInputCloud(golden["uref"])
    gi, gd, _ = u.nearestKSearch(golden["uqry"], 16)
    assert_knn_equal(gi, gd, golden["uknn16_idx"], golden["uknn16_d2"])


@pytest.mark.parametrize("k", [1, 2, 3, 4, 8, 16, 17, 32, 33, 50, 51, 100, 128])
def test_knn_room_vs_oracle(GS, k):
    ref = synth.room(100000, 1001)
    qry = synth.noisy_copy(ref, 1002, 0.002)[:20000]
    s = GS().setInputCloud(ref, k_hint=k)
    gi, gd, keff = s.nearestKSearch(qry, k)
    oi, od, okeff = oracle.KdTree(ref).knn(qry, k)
    assert keff == okeff == k
    assert_knn_equal(gi, gd, oi, od)


@pytest.mark.parametrize("k", [1, 16, 50])
def test_knn_self_query_all_points(GS, k):
    ref = synth.room(60000, 2001, stride4=True)
    s = GS().setInputCloud(ref, k_hint=k)
    gi, gd, _ = s.nearestKSearch(None, k)                      # RegionGrowing::findPointNeighbours shape: N x k table
    oi, od, _ = oracle.KdTree(ref).knn(ref, k)
    assert_knn_equal(gi, gd, oi, od)
    assert (gi[:, 0] == np.arange(ref.shape[0])).all()         # tie-free cloud: nearest is the point itself


def test_knn_uniform_volume_and_mismatched_k_hint(GS):
    ref = synth.uniform(200000, 5001, extent=4.0)
    qry = synth.sweep_queries(ref, 30000, seed=5002, sigma=0.05)
    s = GS().setInputCloud(ref, k_hint=4)                     # grid tuned for k=4, queried at k=32 -> ring expansion
    for k in (1, 32):
        gi, gd, _ = s.nearestKSearch(qry, k)
        oi, od, _ = oracle.KdTree(ref).knn(qry, k)
        assert_knn_equal(gi, gd, oi, od)


def test_grid_autotune_tells_surface_from_volume(GS):
    """The cell size is chosen from the measured cloud: ~8 points per occupied cell on a surface (9-13 of the 27 cells of a block are
    occupied), about half of that in a filled volume (all 27 are) -- csrc/pcc_build.cu occupancy_target / occupancy_target_volume.
    Results stay exact either way (every kNN test); this pins the choice itself."""
    vol = GS().setInputCloud(synth.uniform(300000, 5001, extent=4.0), k_hint=16).grid_info()["occupancy"]
    surf = GS().setInputCloud(synth.room(300000, 4001, size=(10.0, 10.0, 3.0)), k_hint=16).grid_info()["occupancy"]
    assert 3.5 <= vol <= 5.7, vol
    assert 7.0 <= surf <= 11.5, surf
    vol4 = GS().setInputCloud(synth.uniform(300000, 5001, extent=4.0), k_hint=4).grid_info()["occupancy"]
    assert 1.7 <= vol4 <= 2.7, vol4


def test_knn_lattice_ties_everywhere(GS):
    g = np.arange(16, dtype=np.float32) * 0.25
    ref = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    ref = ref[np.random.default_rng(0).permutation(len(ref))]
    s = GS().setInputCloud(ref)
    for k in (1, 7, 8, 27, 50):
        gi, gd, _ = s.nearestKSearch(ref, k)
        oi, od, _ = oracle.brute_knn(ref, ref, k)
        assert_knn_equal(gi, gd, oi, od)


def test_knn_duplicates_nan_rows_k_gt_n(GS):
    ref = np.array([[0, 0, 0], [np.nan, 0, 0], [1, 0, 0], [1, 0, 0], [0, 2, 0], [np.inf, 1, 1]], np.float32)
    qry = np.array([[0.9, 0, 0], [np.nan, 0, 0], [100, -50, 3]], np.float32)
    s = GS().setInputCloud(ref)
    assert s.size == 4
    gi, gd, keff = s.nearestKSearch(qry, 6)
    oi, od, okeff = oracle.brute_knn(ref, qry, 6)
    assert keff == okeff == 4
    assert_knn_equal(gi, gd, oi, od)
    assert gi[0].tolist() == [2, 3, 0, 4, -1, -1] and (gi[1] == -1).all()
    gi, gd, _ = s.nearestKSearch(None, 3)                       # self query: skipped rows stay empty
    oi, od, _ = oracle.brute_knn(ref, ref, 3)
    assert_knn_equal(gi, gd, oi, od)


def test_knn_outliers_negative_coords_cell_faces(GS):
    rng = np.random.default_rng(3)
    ref = np.concatenate([rng.normal(0, 0.05, (20000, 3)), rng.normal(0, 0.05, (20000, 3)) + [3, -2, 1], [[40, 40, 40], [-35, 0, 7]]]).astype(np.float32)
    s = GS().setInputCloud(ref, cell_hint=0.0625)
    cell = np.float32(0.0625)
    origin = ref.min(0)
    on_faces = (origin + cell * rng.integers(0, 40, (2000, 3)).astype(np.float32)).astype(np.float32)    # exactly on cell faces
    far = rng.uniform(-60, 60, (500, 3)).astype(np.float32)                                              # far outside the bbox
    qry = np.concatenate([on_faces, far, ref[::97]])
    for k in (1, 16):
        gi, gd, _ = s.nearestKSearch(qry, k)
        oi, od, _ = oracle.KdTree(ref).knn(qry, k)
        assert_knn_equal(gi, gd, oi, od)


def test_knn_indices_subset_and_stride32(GS):
    pts = np.zeros((30000, 8), np.float32)
    pts[:, :3] = synth.room(30000, 77)
    pts[:, 3] = 1.0
    sub = np.random.default_rng(1).choice(30000, 9000, replace=False).astype(np.int32)
    s = GS().setInputCloud(pts, indices=sub)
    qry = pts[:4000]
    gi, gd, _ = s.nearestKSearch(qry, 8)
    oi, od, _ = oracle.brute_knn(pts[sub], qry, 8)
    assert_knn_equal(gi, gd, sub[oi], od)                       # returned indices address the ORIGINAL cloud


def test_knn_tiny_clouds(GS):
    for n in (1, 2, 5, 33):
        ref = np.random.default_rng(n).random((n, 3)).astype(np.float32)
        s = GS().setInputCloud(ref)
        gi, gd, keff = s.nearestKSearch(ref, 4)
        oi, od, okeff = oracle.brute_knn(ref, ref, 4)
        assert keff == okeff
        assert_knn_equal(gi, gd, oi, od)
    same = np.ones((50, 3), np.float32)                         # zero-extent cloud
    gi, gd, _ = GS().setInputCloud(same).nearestKSearch(same[:3], 5)
    assert gi.tolist() == [[0, 1, 2, 3, 4]] * 3 and (gd == 0).all()


# ---------------------------------------------------------------- radius
@pytest.mark.parametrize("radius,hint", [(0.05, 0.05), (0.03, 0.0), (0.12, 0.05)])
def test_radius_vs_oracle(GS, radius, hint):
    ref = synth.room(80000, 1001)
    qry = synth.noisy_copy(ref, 5, 0.003)[:10000]
    s = GS().setInputCloud(ref, cell_hint=hint)
    off, idx, d2 = s.radiusSearch(qry, radius)
    ooff, oidx, od2 = oracle.KdTree(ref).radius(qry, radius)
    assert np.array_equal(off, ooff) and np.array_equal(idx, oidx) and np.array_equal(bits(d2), bits(od2))


def test_radius_boundary_max_nn_empty_unsorted(GS):
    ref = np.array([[0, 0, 0], [0.5, 0, 0], [1.0, 0, 0], [0, 0.25, 0]], np.float32)
    qry = np.zeros((1, 3), np.float32)
    s = GS().setInputCloud(ref)
    assert s.radiusSearch(qry, 0.5)[1].tolist() == [0, 3]                 # d2 == r2 is excluded
    assert s.radiusSearch(qry, 0.5000001)[1].tolist() == [0, 3, 1]
    assert s.radiusSearch(qry, 2.0, max_nn=2)[1].tolist() == [0, 3]
    off, idx, _ = s.radiusSearch(np.array([[9, 9, 9]], np.float32), 0.1)
    assert off.tolist() == [0, 0] and idx.size == 0
    big = synth.room(20000, 9)
    s2 = GS(sorted=False).setInputCloud(big, cell_hint=0.1)
    off, idx, d2 = s2.radiusSearch(big[:3000], 0.1)
    ooff, oidx, od2 = oracle.KdTree(big).radius(big[:3000], 0.1)
    assert np.array_equal(off, ooff)
    for i in range(0, 3000, 50):                                            # unsorted rows: same sets
        assert sorted(idx[off[i]:off[i + 1]].tolist()) == sorted(oidx[off[i]:off[i + 1]].tolist())
    s3 = GS().setInputCloud(big, cell_hint=0.1)
    off, idx, d2 = s3.radiusSearch(big[:3000], 0.1, max_nn=5)
    ooff, oidx, od2 = oracle.KdTree(big).radius(big[:3000], 0.1, max_nn=5)
    assert np.array_equal(off, ooff) and np.array_equal(idx, oidx) and np.array_equal(bits(d2), bits(od2))


# ---------------------------------------------------------------- fused consumers
@pytest.mark.parametrize("mean_k", [16, 50, 7])
def test_sor_mean_distance_and_threshold(GS, mean_k):
    ref = synth.room(100000, 2001)
    s = GS().setInputCloud(ref, k_hint=mean_k + 1)
    dist = s.meanNeighbourDistance(None, mean_k)
    o = oracle.sor(ref, mean_k, 1.5)
    assert np.array_equal(bits(dist), bits(o["distances"]))                 # fp64 sqrt/sum in the same order: bit-exact
    r = s.sorThreshold(dist, len(ref), 1.5)
    for key in ("mean", "stddev", "threshold"):
        assert abs(r[key] - o[key]) <= 1e-9 * abs(o[key])                   # parallel fp64 reduction order, tolerance 1e-9 rel
    assert r["kept"] == o["kept"] and np.array_equal(r["keep"].astype(bool), o["keep"])


def _normal_err(a, b):
    return np.linalg.norm(a[:, :3] - b[:, :3], axis=1)


def test_normals_knn_vs_oracle(GS):
    ref = synth.room(60000, 1001)
    s = GS().setInputCloud(ref, k_hint=50)
    n = s.normalsKnn(None, 50)
    o = oracle.normals_knn(ref, 50)
    # eigen33 uses atan2f/cosf/sinf whose last-ulp rounding differs between glibc and CUDA; north_star's tolerance is 1e-5 on the unit
    # normal and EVERY point must meet it (measured on B200: max 3.0e-7 over all 60000 points, gap (l1 - l0) / max|cov| >= 0.27 for 99 %
    # of them -- profiles/r2/normals_error_vs_gap.txt); the oracle's own eigen33 is pinned against LAPACK in test_oracle_pinning.py.
    err = _normal_err(n, o)
    assert np.isfinite(n).all()
    assert err.max() < 1e-5, err.max()
    assert np.allclose(n[:, 3], o[:, 3], rtol=1e-3, atol=5e-7)           # curvature = |l0 / trace|: absolute error 8e-8 measured; tiny curvatures make the relative one meaningless
    assert (np.einsum("ij,ij->i", -ref[:, :3], n[:, :3]) >= -1e-6).all()   # flipped towards the viewpoint (origin)


def test_normals_radius_and_degenerate(GS):
    ref = synth.room(40000, 1001)
    s = GS().setInputCloud(ref, cell_hint=0.03)
    n = s.normalsRadius(None, 0.03)
    o = oracle.normals_radius(ref, 0.03)
    assert np.array_equal(np.isnan(n[:, 0]), np.isnan(o[:, 0]))             # < 3 neighbours -> NaN on both sides
    ok = ~np.isnan(o[:, 0])
    err = _normal_err(n[ok], o[ok])
    assert err.max() < 1e-5, err.max()                                      # every normal, also the ill-conditioned 3- and 4-point neighbourhoods (measured max 2.0e-7)
    assert np.allclose(n[ok, 3], o[ok, 3], rtol=1e-3, atol=5e-7)
    lone = np.array([[0, 0, 0], [5, 5, 5]], np.float32)
    assert np.isnan(GS().setInputCloud(lone).normalsRadius(None, 0.1)).all()


def test_euclidean_clusters(GS):
    rng = np.random.default_rng(1)
    a = rng.random((300, 3)).astype(np.float32) * 0.1
    b = a + np.array([0.16, 0, 0], np.float32)
    pts = np.concatenate([a, b, np.array([[5, 5, 5], [5.01, 5, 5]], np.float32)])
    for tol, mn, mx in ((0.05, 100, 250000), (0.07, 100, 250000), (0.07, 100, 500), (0.05, 1, 250000)):
        lab, sizes = GS().setInputCloud(pts, cell_hint=tol).euclideanClusters(tol, mn, mx)
        olab, osizes = oracle.KdTree(pts).ece(tol, mn, mx)
        assert np.array_equal(lab, olab) and np.array_equal(sizes, osizes)


def test_euclidean_clusters_scene_known_count(GS):
    pts, ids = synth.scene(400000, 3001, extent=12.0, n_objects=80)
    s = GS().setInputCloud(pts, cell_hint=0.05)
    lab, sizes = s.euclideanClusters(0.05, 100, 250000)
    olab, osizes = oracle.KdTree(pts).ece(0.05, 100, 250000)
    assert len(sizes) == 80
    assert np.array_equal(sizes, osizes) and np.array_equal(lab, olab)      # bit-exact labels incl. the PCL ordering


def test_icp_step_and_align(GS):
    src, tgt, T = synth.icp_pair(60000, 4001, size=(5, 5, 3))
    s = GS().setInputCloud(tgt, k_hint=1)
    cnt, sums, ci, cd = s.icpStep(src.copy(), None, want_correspondences=True)
    ocnt, osums, oci, ocd = oracle.KdTree(tgt).icp_pass(src)
    assert cnt == ocnt and np.array_equal(ci, oci) and np.array_equal(bits(cd), bits(ocd))
    assert np.allclose(sums, osums, rtol=1e-12, atol=1e-9)                  # fp64 sums, different association only
    r = s.icpAlign(src, 20)
    o = oracle.icp(src, tgt, 20)
    assert r["converged"] == o["converged"] and r["iterations"] == o["iterations"]
    assert np.allclose(r["T"], o["T"], rtol=1e-5, atol=1e-6)                # tolerance 1e-5 (north_star) on the 4x4
    assert abs(r["fitness"] - o["fitness"]) <= 1e-5 * o["fitness"]
    assert np.allclose(r["T"], np.linalg.inv(T), atol=2e-4)


def test_first_within(GS):
    ref = synth.room(30000, 5)
    q = synth.noisy_copy(ref, 6, 0.02)[:2000]
    got = GS().setInputCloud(ref, cell_hint=0.05).firstWithin(q, 0.05)
    assert np.array_equal(got, oracle.first_within(ref, q, 0.05))


# ---------------------------------------------------------------- device-pointer path + properties at full size
def test_device_pointer_path_matches_host_path(GS):
    import torch
    ref = synth.room(50000, 11, stride4=True)
    qry = synth.noisy_copy(ref, 12, 0.002, stride4=True)[:8000]
    hi, hd, _ = GS().setInputCloud(ref).nearestKSearch(qry, 16)
    s = GS().setInputCloud(torch.from_numpy(ref).cuda())
    di, dd, _ = s.nearestKSearch(torch.from_numpy(qry).cuda(), 16)
    assert np.array_equal(di.cpu().numpy(), hi) and np.array_equal(bits(dd.cpu().numpy()), bits(hd))
    off, idx, d2 = s.radiusSearch(torch.from_numpy(qry).cuda(), 0.05)
    ooff, oidx, od2 = oracle.KdTree(ref).radius(qry, 0.05)
    assert np.array_equal(off.cpu().numpy(), ooff) and np.array_equal(idx.cpu().numpy(), oidx)


def test_full_size_properties_10m(GS):
    """BASELINE headline size (10 M reference points, k = 16): size-independent properties + a sampled oracle check."""
    import torch
    ref = synth.room(10_000_000, 4001, size=(10.0, 10.0, 3.0), stride4=True)
    dref = torch.from_numpy(ref).cuda()
    s = GS().setInputCloud(dref, k_hint=16)
    idx, d2, keff = s.nearestKSearch(None, 16)
    assert keff == 16
    assert bool((d2[:, 1:] >= d2[:, :-1]).all())                                            # sorted rows
    assert bool((idx[:, 0] == torch.arange(ref.shape[0], device="cuda", dtype=torch.int32)).all()) or bool((d2[:, 0] == 0).all())
    assert bool((idx >= 0).all()) and bool((idx < ref.shape[0]).all())
    srt = torch.sort(idx.long(), dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())                                          # no duplicates in a row
    # recompute the distances from the returned indices (fp32, same association)
    sel = torch.randint(0, ref.shape[0], (200000,), device="cuda")
    p = dref[sel, :3].unsqueeze(1)
    nb = dref[idx[sel].long(), :3]
    df = p - nb
    rec = (df[..., 0] * df[..., 0] + df[..., 1] * df[..., 1]) + df[..., 2] * df[..., 2]
    assert bool(((rec - d2[sel]).abs() <= 1e-6 * rec.abs() + 1e-12).all())
    sample = np.random.default_rng(0).choice(ref.shape[0], 20000, replace=False)
    oi, od, _ = oracle.KdTree(ref).knn(ref[sample], 16)
    assert np.array_equal(idx[torch.from_numpy(sample).cuda()].cpu().numpy(), oi)
    assert np.array_equal(bits(d2[torch.from_numpy(sample).cuda()].cpu().numpy()), bits(od))


# ---------------------------------------------------------------- randomized stress (fp32 margins, odd shapes)
def _stress_cloud(rng, kind, n):
    if kind == "big_offset":                      # coordinates ~1e4 with centimetre structure: fp32 grid-coordinate rounding
        return (rng.random((n, 3)) * [3.0, 2.0, 1.0] + [12345.0, -9876.0, 4321.0]).astype(np.float32)
    if kind == "flat":                            # degenerate z extent
        p = rng.random((n, 3)) * [5.0, 5.0, 0.0]
        return p.astype(np.float32)
    if kind == "line":                            # 1-D manifold
        t = rng.random(n)
        return np.stack([t * 7.0, 0.3 * np.sin(t * 20), 0.1 * t], 1).astype(np.float32)
    if kind == "anisotropic":                     # 1e3 : 1 : 1e-3 extents
        return (rng.random((n, 3)) * [1000.0, 1.0, 1e-3]).astype(np.float32)
    if kind == "clustered":                       # very uneven density + far outliers
        c = rng.normal(0, 1.0, (8, 3))
        p = c[rng.integers(0, 8, n)] + rng.normal(0, 0.01, (n, 3)) * rng.choice([1.0, 10.0], (n, 1))
        p[: n // 100] = rng.uniform(-30, 30, (n // 100, 3))
        return p.astype(np.float32)
    if kind == "quantized":                       # millimetre-quantized scanner output: many exact ties and duplicates
        return (np.round(rng.random((n, 3)) * [2.0, 2.0, 0.5] * 200) / 200).astype(np.float32)
    if kind == "tiny_scale":                      # micrometre-scale cloud
        return (rng.random((n, 3)) * 1e-4).astype(np.float32)
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["big_offset", "flat", "line", "anisotropic", "clustered", "quantized", "tiny_scale"])
def test_randomized_stress_knn_and_radius(GS, kind):
    import zlib
    rng = np.random.default_rng(zlib.crc32(kind.encode()))          # stable per-kind seed
    ref = _stress_cloud(rng, kind, 30000)
    spread = np.maximum(ref.max(0) - ref.min(0), 1e-9)
    qry = np.concatenate([ref[rng.integers(0, len(ref), 1500)] + rng.normal(0, 1e-3, (1500, 3)) * spread,
                          ref[:500],
                          ref.min(0) + rng.uniform(-0.5, 1.5, (500, 3)) * spread]).astype(np.float32)
    tree = oracle.KdTree(ref)
    for k in (1, 5, 16, 32, 40):
        s = GS().setInputCloud(ref, k_hint=k)
        gi, gd, _ = s.nearestKSearch(qry, k)
        oi, od, _ = tree.knn(qry, k)
        assert_knn_equal(gi, gd, oi, od)
    # radius: pick r so rows hold ~20 neighbours on average
    _, d2_20, _ = tree.knn(qry[:200], 20)
    r = float(np.sqrt(np.median(d2_20[:, -1])))
    s = GS().setInputCloud(ref, cell_hint=r)
    off, idx, d2 = s.radiusSearch(qry, r)
    ooff, oidx, od2 = tree.radius(qry, r)
    assert np.array_equal(off, ooff) and np.array_equal(idx, oidx) and np.array_equal(bits(d2), bits(od2))
    md = GS().setInputCloud(ref, k_hint=9).meanNeighbourDistance(None, 8)
    assert np.array_equal(bits(md), bits(oracle.sor(ref, 8, 1.0, tree=tree)["distances"]))


def test_sharded_clustering_two_emulated_ranks(GS):
    """pcc_ece_link_range / absorb / finish with two ranks emulated in one process (sequential kernels over both ranks'
    forests; the element-wise MIN stands in for the NCCL all-reduce of shard.euclidean_clusters_sharded)."""
    import torch
    pts, ids = synth.scene(300000, 3001, extent=10.0, n_objects=60)
    s = GS().setInputCloud(torch.from_numpy(pts).cuda(), cell_hint=0.05)
    n = s.size
    ranges = [(0, n // 3), (n // 3, n)]                                       # uneven shards
    forests = []
    for b, e in ranges:
        f = s.eceNewForest(); s.eceLinkRange(f, 0.05, b, e); s.eceAbsorb(f, f); forests.append(f)
    assert not torch.equal(forests[0], forests[1])
    rounds = 0
    while True:
        merged = torch.minimum(forests[0], forests[1])
        if all(torch.equal(merged, f) for f in forests):
            break
        for f in forests:
            s.eceAbsorb(f, merged)
        rounds += 1
        assert rounds < 20
    lab, sizes = s.eceFinish(forests[0], 100, 250000)
    olab, osizes = oracle.KdTree(pts).ece(0.05, 100, 250000)
    assert np.array_equal(lab.cpu().numpy(), olab) and np.array_equal(sizes.cpu().numpy(), osizes)
    lab1, sizes1 = s.eceFinish(forests[1], 100, 250000)
    assert torch.equal(lab, lab1) and torch.equal(sizes, sizes1)


def test_voxel_grid_vs_oracle(GS):
    """VoxelGrid leaf 0.025 on the coloured room (src/segmentation.cpp:69-74); bit-exact vs the oracle."""
    import torch
    from pointcloudcomparator_b200.search import voxel_grid
    n = 200000
    rows = np.zeros((n, 8), np.float32)
    rows[:, :3] = synth.room(n, 1001)
    rows[:, 3] = 1.0
    rgb = synth.rgb_for(rows, 7)
    rows[:, 4:5].view(np.uint8)[:, :3] = rgb[:, ::-1]          # BGRA
    rows[11, 1] = np.nan
    for leaf in (0.025, 0.1, (0.05, 0.025, 0.2)):
        got = voxel_grid(rows, leaf, rgb_offset_bytes=16)
        ref = oracle.voxel_grid(rows, leaf, rgb_offset_floats=4)
        assert got.shape == ref.shape and np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    got = voxel_grid(rows, 0.1, rgb_offset_bytes=16, min_points=5)
    ref = oracle.voxel_grid(rows, 0.1, rgb_offset_floats=4, min_points=5)
    assert got.shape == ref.shape and np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    dev = voxel_grid(torch.from_numpy(rows).cuda(), 0.025, rgb_offset_bytes=16)
    assert np.array_equal(dev.cpu().numpy().view(np.uint32), oracle.voxel_grid(rows, 0.025, rgb_offset_floats=4).view(np.uint32))
    xyz = synth.uniform(100000, 5, 2.0)                        # packed xyz rows, no colour
    assert np.array_equal(voxel_grid(xyz, 0.05).view(np.uint32), oracle.voxel_grid(xyz, 0.05).view(np.uint32))


def test_tma_cell_kernel_variant_is_exact(GS, monkeypatch):
    """The opt-in warp-owns-a-cell variant (cp.async.bulk-staged stencil tiles, PCC_CELL_KERNEL=1) returns the same rows."""
    monkeypatch.setenv("PCC_CELL_KERNEL", "1")
    ref = synth.room(150000, 1001)
    qry = np.concatenate([synth.sweep_queries(ref, 400000, seed=3, sigma=0.01), ref[:1000], np.full((3, 3), np.nan, np.float32),
                          np.random.default_rng(1).uniform(-2, 8, (500, 3)).astype(np.float32)])
    tree = oracle.KdTree(ref)
    for k, hint in ((16, 16), (5, 16), (32, 32), (16, 64)):          # hint 64: dense cells -> stencils larger than one 320-point tile
        s = GS().setInputCloud(ref, k_hint=hint)
        gi, gd, _ = s.nearestKSearch(qry, k)
        oi, od, _ = tree.knn(qry, k)
        assert_knn_equal(gi, gd, oi, od)
    g = np.arange(12, dtype=np.float32) * 0.5                          # lattice: ties -> fix-up path from the cell kernel
    lat = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    gi, gd, _ = GS().setInputCloud(lat).nearestKSearch(lat, 8)
    oi, od, _ = oracle.brute_knn(lat, lat, 8)
    assert_knn_equal(gi, gd, oi, od)


def test_pipelined_host_path_large_batch(GS):
    """Host-buffer batches >= 2 Mi queries go through the two-slot chunked pipeline; rows must equal the one-shot device path."""
    import torch
    ref = synth.room(300000, 1001, stride4=True)
    qry = synth.sweep_queries(ref, 2_600_001, seed=4, sigma=0.01, stride4=True)
    s = GS().setInputCloud(ref, k_hint=16)
    hi, hd, keff = s.nearestKSearch(qry, 16)                                   # numpy -> PCC_HOST -> pipelined
    di, dd, _ = s.nearestKSearch(torch.from_numpy(qry).cuda(), 16)             # one launch
    assert keff == 16 and np.array_equal(hi, di.cpu().numpy()) and np.array_equal(bits(hd), bits(dd.cpu().numpy()))
    sel = np.random.default_rng(2).choice(len(qry), 20000, replace=False)
    oi, od, _ = oracle.KdTree(ref).knn(qry[sel], 16)
    assert_knn_equal(hi[sel], hd[sel], oi, od)


def test_radius_rows_of_every_length(GS):
    """Row sort tiers: <= 32 keys (shuffles), <= 1024 (shared-memory bitonic), longer (library fallback) -- all (d2, idx)-sorted."""
    rng = np.random.default_rng(9)
    ref = np.concatenate([rng.normal(0, 0.02, (6000, 3)), rng.random((20000, 3)) * 2.0 + 1.0]).astype(np.float32)   # a dense blob + sparse volume
    qry = np.concatenate([ref[:300], ref[6000:9000], rng.random((200, 3)).astype(np.float32) * 3])
    tree = oracle.KdTree(ref)
    for r in (0.04, 0.12, 0.3):
        s = GS().setInputCloud(ref, cell_hint=r)
        off, idx, d2 = s.radiusSearch(qry, r)
        ooff, oidx, od2 = tree.radius(qry, r)
        lens = np.diff(ooff)
        assert np.array_equal(off, ooff) and np.array_equal(idx, oidx) and np.array_equal(bits(d2), bits(od2)), (r, lens.max())
    assert lens.max() > 1024 and (lens <= 32).any() and ((lens > 32) & (lens <= 1024)).any()


def test_indices_with_duplicates_leave_unindexed_rows_empty(GS):
    """an `indices` list may repeat rows, so n_indexed >= n_input does not mean every input row is indexed; the
    self-query outputs of the rows that are not must still be pre-filled ((-1, +inf) / 0 / NaN)."""
    pts = np.array([[0, 0, 0], [1, 0, 0], [5, 5, 5], [6, 6, 6]], np.float32)
    s = GS().setInputCloud(pts, indices=np.array([0, 0, 1, 1], np.int32))
    assert s.size == 4
    gi, gd, _ = s.nearestKSearch(None, 2)
    assert (gi[2:] == -1).all() and np.isinf(gd[2:]).all()
    assert set(gi[0].tolist()) == {0} and set(gi[1].tolist()) == {1} and (gd[:2] == 0).all()      # each indexed twice: its two nearest are its own two copies
    md = s.meanNeighbourDistance(None, 1)
    assert (md[2:] == 0).all()
    nrm = s.normalsKnn(None, 3)
    assert np.isnan(nrm[2:]).all()
    nr = s.normalsRadius(None, 10.0)
    assert np.isnan(nr[2:]).all()


def test_query_batch_size_guard(GS):
    """Row counts are handed to CUB as 32-bit ints: a batch of 2^31 rows must be refused loudly, before any memory is touched."""
    from pointcloudcomparator_b200 import _lib
    import ctypes as C
    s = GS().setInputCloud(synth.room(2000, 1001))
    L = _lib.lib()
    keff = C.c_int()
    rc = L.pcc_knn(s._h, C.c_void_p(16), 1 << 31, 16, 4, C.c_void_p(16), C.c_void_p(16), C.byref(keff), _lib.DEVICE, None)
    assert rc == -1 and b"nq" in L.pcc_last_error()


def test_tma_staged_block_kernel_parity(GS, monkeypatch):
    """knn_thr_staged_kernel (PCC_THR_STAGED=1): the warp's union stencil pulled into shared memory with cp.async.bulk + mbarrier, lanes walk
    their windows of the tile.  Opt-in (measured slower than the per-lane walk), but exact: same rows as the oracle on a surface cloud with
    external noisy queries (mixed warps: lanes outside the leader's row keep the global path), on self-queries, and on a clustered cloud
    whose dense unions do not fit the tile."""
    monkeypatch.setenv("PCC_THR_STAGED", "1")
    ref = synth.room(120000, 1001)
    qry = synth.sweep_queries(ref, 50000, seed=11, sigma=0.01)
    tree = oracle.KdTree(ref)
    s = GS().setInputCloud(ref, k_hint=16)
    for k in (16, 12):                                      # 12: K = 16 template with k < K takes the non-staged kernel
        gi, gd, _ = s.nearestKSearch(qry, k)
        oi, od, _ = tree.knn(qry, k)
        assert_knn_equal(gi, gd, oi, od)
    gi, gd, _ = s.nearestKSearch(None, 16)
    oi, od, _ = tree.knn(ref, 16)
    assert_knn_equal(gi, gd, oi, od)
    rng = np.random.default_rng(5)
    c = rng.normal(0, 1.0, (6, 3))
    dense = (c[rng.integers(0, 6, 60000)] + rng.normal(0, 0.01, (60000, 3)) * rng.choice([1.0, 10.0], (60000, 1))).astype(np.float32)
    q2 = (dense[::3] + rng.normal(0, 0.002, (20000, 3))).astype(np.float32)
    gi, gd, _ = GS().setInputCloud(dense, k_hint=16).nearestKSearch(q2, 16)
    oi, od, _ = oracle.KdTree(dense).knn(q2, 16)
    assert_knn_equal(gi, gd, oi, od)


@pytest.mark.parametrize("k", [9, 13, 15, 16])
def test_threshold_block_kernel_paths(GS, k):
    """knn_thr_kernel<16> serves 8 < k <= 16 on batches >= 1024: k < 16 takes its runtime-k instantiation; a reference cloud smaller than k
    sends every query to the wide pass; far-away and clamped queries, exact duplicates (ties -> 64-bit selection from the log in the retry
    pass) and a very dense clump (log overflow -> in-place compression in the retry pass) all have to come out bit-exact."""
    rng = np.random.default_rng(40 + k)
    ref = synth.room(80000, 1001, size=(3.0, 2.0, 1.5))
    dup = ref[rng.integers(0, len(ref), 3000)]                                         # exact duplicates: d2 ties at every rank
    clump = (ref[123] + rng.normal(0, 2e-4, (4000, 3))).astype(np.float32)             # 4000 points inside one cell
    ref2 = np.concatenate([ref, dup, clump]).astype(np.float32)
    ref2 = ref2[rng.permutation(len(ref2))]
    qry = np.concatenate([synth.sweep_queries(ref2, 6000, seed=k, sigma=0.01), dup[:500], clump[:500] + np.float32(1e-4),
                          rng.uniform(-5, 8, (300, 3)).astype(np.float32), np.array([[1e3, -1e3, 5e2], [np.nan, 0, 0]], np.float32)]).astype(np.float32)
    tree = oracle.KdTree(ref2)
    gi, gd, keff = GS().setInputCloud(ref2, k_hint=k).nearestKSearch(qry, k)
    oi, od, _ = tree.knn(qry, k)
    assert keff == k
    assert_knn_equal(gi, gd, oi, od)
    tiny = ref[:k - 3]                                                                  # fewer points than k: (-1, +inf) padding, every query "wide"
    gi, gd, keff = GS().setInputCloud(tiny, k_hint=k).nearestKSearch(qry[:2000], k)
    oi, od, okeff = oracle.KdTree(tiny).knn(qry[:2000], k)
    assert keff == okeff == k - 3
    assert_knn_equal(gi, gd, oi, od)


def test_voxel_grid_leaf_too_small_passes_input_through(GS):
    """PCL 1.7 VoxelGrid::applyFilter warns and returns the input unfiltered when the voxel index would overflow int32
    (src/segmentation.cpp:69-74 would then cluster the full cloud); product and oracle both mirror that instead of failing."""
    from pointcloudcomparator_b200.search import voxel_grid
    p = synth.room(5000, 1001, stride4=True)
    p[17, 2] = np.nan
    out = voxel_grid(p, 1e-5)
    ref = oracle.voxel_grid(p, 1e-5)
    assert out.shape == p.shape and np.array_equal(bits(out), bits(p)) and np.array_equal(bits(ref), bits(p))


def test_descriptor_nn_vs_oracle(GS):
    """RIFT32 1-NN (src/comparator.cpp:560-588), bit-exact vs the oracle incl. the reference's vector shape."""
    import torch
    from pointcloudcomparator_b200.search import descriptor_nn, match_rift_features_knn
    rng = np.random.default_rng(5)
    for n1, n2 in ((1, 3), (63, 200), (700, 400), (5000, 3000)):
        ref = rng.random((n1, 32), dtype=np.float32) * 0.3
        qry = (ref[rng.integers(0, n1, n2)] + rng.normal(0, 0.03, (n2, 32))).astype(np.float32)
        if n1 > 10:
            ref[3, 9] = np.nan; qry[1, 31] = np.inf; ref[7] = ref[2]           # skipped row, empty query, exact duplicate (tie -> lower index)
        gi, gd = descriptor_nn(ref, qry)
        oi, od = oracle.descriptor_nn(ref, qry)
        assert np.array_equal(gi, oi) and np.array_equal(bits(gd), bits(od))
        di, dd = descriptor_nn(torch.from_numpy(ref).cuda(), torch.from_numpy(qry).cuda())
        assert np.array_equal(di.cpu().numpy(), oi) and np.array_equal(bits(dd.cpu().numpy()), bits(od))
        corr = match_rift_features_knn(ref, qry, match_dims=32)
        assert corr == [0] + oi[(oi >= 0) & (od < np.float32(0.05))].tolist()
        # reference-exact mode: PCL 1.7 compares only the first 3 floats of an unregistered Histogram<32> (distance AND validity)
        g3, d3 = descriptor_nn(ref, qry, dims=3)
        o3, od3 = oracle.descriptor_nn(ref, qry, dims=3)
        assert np.array_equal(g3, o3) and np.array_equal(bits(d3), bits(od3))
        t3, td3 = descriptor_nn(torch.from_numpy(ref).cuda(), torch.from_numpy(qry).cuda(), dims=3)
        assert np.array_equal(t3.cpu().numpy(), o3) and np.array_equal(bits(td3.cpu().numpy()), bits(od3))
        if n1 > 10:
            assert o3[1] >= 0 and oi[1] == -1            # the query whose bin 31 is inf is valid in 3-D, empty in 32-D
            assert (o3 == 3).sum() >= 0 and not (oi == 3).any()       # reference row 3 (NaN in bin 9) can only match in 3-D
        assert match_rift_features_knn(ref, qry) == [0] + o3[(o3 >= 0) & (od3 < np.float32(0.05))].tolist()      # default = reference-exact


# ---------------------------------------------------------------- the passes behind the 3x3x3 block (rings / wide / tied)
@pytest.mark.parametrize("k", [2, 5, 8, 16, 31, 32])
def test_knn_density_gradient_rings_and_wide_passes(GS, k):
    """A dense core with a halo 30x and 1000x sparser: core queries settle in the 3x3x3 block, halo queries need one or two
    more rings (knn_rings_kernel), and queries out in the thin halo or beyond it see fewer than k points in their block
    (knn_wide_kernel, one warp per query, rings until covered)."""
    rng = np.random.default_rng(100 + k)
    ref = np.concatenate([rng.normal(0, 0.2, (60000, 3)), rng.normal(0, 0.8, (8000, 3)), rng.uniform(-6, 6, (1500, 3))]).astype(np.float32)
    qry = np.concatenate([ref[::7] + rng.normal(0, 0.01, (len(ref[::7]), 3)), rng.uniform(-7, 7, (3000, 3)), rng.uniform(-40, 40, (200, 3))]).astype(np.float32)
    s = GS().setInputCloud(ref, k_hint=k)
    gi, gd, _ = s.nearestKSearch(qry, k)
    oi, od, _ = oracle.KdTree(ref).knn(qry, k)
    assert_knn_equal(gi, gd, oi, od)


def test_knn_few_and_many_tied_queries(GS):
    """More than K candidates tied at the k-th distance: a short list of such queries is finished by the warp-per-query
    kernel, a long one (> 4096) by the per-thread exact kernel; both must give the canonical (d2, index) order."""
    rng = np.random.default_rng(5)
    g = np.arange(6, dtype=np.float32) * 0.125
    patch = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3) + np.float32(8.0)             # 216 lattice points, exact in fp32
    ref = np.concatenate([rng.uniform(0, 4, (30000, 3)).astype(np.float32), patch])
    ref = ref[rng.permutation(len(ref))]
    s = GS().setInputCloud(ref, cell_hint=0.3)
    tree = oracle.KdTree(ref)
    for k in (4, 7, 16):
        gi, gd, _ = s.nearestKSearch(ref, k)                                                              # 216 tied queries
        oi, od, _ = tree.knn(ref, k)
        assert_knn_equal(gi, gd, oi, od)
    g = np.arange(20, dtype=np.float32) * 0.25
    lattice = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)                            # 8000 tied queries
    lattice = lattice[rng.permutation(len(lattice))]
    s = GS().setInputCloud(lattice)
    for k in (7, 16):
        gi, gd, _ = s.nearestKSearch(lattice, k)
        oi, od, _ = oracle.KdTree(lattice).knn(lattice, k)
        assert_knn_equal(gi, gd, oi, od)


def test_knn_adopted_grid_rebuilds_occupancy(GS):
    """pcc_export / pcc_adopt (the multi-GPU broadcast path): the adopting index derives its occupancy bitmap on the first query."""
    import torch
    ref = synth.room(50000, 12)
    qry = np.concatenate([ref[::5] + np.float32(0.004), np.random.default_rng(2).uniform(-1, 6, (2000, 3)).astype(np.float32)])
    a = GS().setInputCloud(torch.from_numpy(ref).cuda(), k_hint=8)
    meta, a_pts, a_cells = a.export()
    b = GS()
    _, b_pts, b_cells = b.adopt(meta)

    def view(ptr, nbytes):
        class _Mem:
            pass
        m = _Mem()
        m.__cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<f4", "data": (int(ptr), False), "version": 3}
        return torch.as_tensor(m, device="cuda:0")

    view(b_pts, int(meta[0]) * 16).copy_(view(a_pts, int(meta[0]) * 16))
    view(b_cells, (int(meta[11]) + 1) * 4).copy_(view(a_cells, (int(meta[11]) + 1) * 4))
    torch.cuda.synchronize()
    gi, gd, _ = b.nearestKSearch(qry, 8)
    oi, od, _ = oracle.KdTree(ref).knn(qry, 8)
    assert_knn_equal(gi, gd, oi, od)


@pytest.mark.parametrize("k", [40, 64, 120, 200, 257, 512])
def test_knn_large_k_selection_path(GS, k):
    """k > 32: the k-th distance is bracketed by a histogram, the rows are filled like a radius search and sorted in
    registers (<= 256 keys) or by the generic row sort; far queries, duplicates and a cloud smaller than some blocks."""
    rng = np.random.default_rng(300 + k)
    ref = np.concatenate([synth.room(20000, 5), rng.normal(2.0, 0.02, (600, 3)), np.repeat(rng.uniform(0, 3, (5, 3)), 70, axis=0)]).astype(np.float32)
    qry = np.concatenate([ref[::11] + rng.normal(0, 0.01, (len(ref[::11]), 3)), rng.uniform(-3, 8, (300, 3)), [[np.nan, 0, 0], [60, 60, 60]]]).astype(np.float32)
    s = GS().setInputCloud(ref, k_hint=k)
    gi, gd, _ = s.nearestKSearch(qry, k)
    oi, od, _ = oracle.brute_knn(ref, qry, k)          # canonical (d2, index) order: 70-fold duplicates tie far beyond k
    assert_knn_equal(gi, gd, oi, od)
    small = ref[:100]
    gi, gd, keff = GS().setInputCloud(small, k_hint=k).nearestKSearch(qry[:50], k)       # k > cloud size: padded rows
    oi, od, okeff = oracle.brute_knn(small, qry[:50], k)
    assert keff == okeff == min(k, 100)
    assert_knn_equal(gi, gd, oi, od)


def test_icp_previous_match_bound_does_not_change_results(GS):
    """pcc_icp_step starts every search from the target point matched in the previous call.  The bound must be invisible:
    the same pass answered cold (fresh index), warm (bounds from an identical pass) and with misleading bounds (left by a
    very different source cloud of the same size) gives identical correspondences, and they equal the oracle's 1-NN."""
    rng = np.random.default_rng(77)
    tgt = synth.room(60000, 9)
    src = (tgt[rng.permutation(len(tgt))[:20000]] + rng.normal(0, 0.03, (20000, 3))).astype(np.float32)
    other = rng.uniform(-2, 8, (20000, 3)).astype(np.float32)
    oi, od, _ = oracle.KdTree(tgt).knn(src, 1)
    s = GS().setInputCloud(tgt, k_hint=16)
    cold = s.icpStep(src.copy(), None, want_correspondences=True)
    warm = s.icpStep(src.copy(), None, want_correspondences=True)
    s.icpStep(other.copy(), None)                                   # leaves bounds that have nothing to do with `src`
    misled = s.icpStep(src.copy(), None, want_correspondences=True)
    for cnt, sums, ci, cd in (cold, warm, misled):
        assert cnt == 20000
        assert np.array_equal(ci, oi[:, 0]) and np.array_equal(bits(cd), bits(od[:, 0]))
        assert np.array_equal(sums, cold[1])                        # deterministic reduction: bit-identical sums


def test_large_k_properties_2m(GS):
    """k = 50 (the reference's NormalEstimation / SOR neighbourhood) at 2 M points through the selection path: size-
    independent properties plus a sampled oracle check."""
    import torch
    ref = synth.room(2_000_000, 4001, size=(6.0, 6.0, 3.0), stride4=True)
    dref = torch.from_numpy(ref).cuda()
    s = GS().setInputCloud(dref, k_hint=50)
    idx, d2, keff = s.nearestKSearch(None, 50)
    assert keff == 50
    assert bool((d2[:, 1:] >= d2[:, :-1]).all()) and bool((d2[:, 0] == 0).all())
    assert bool((idx >= 0).all()) and bool((idx < ref.shape[0]).all())
    srt = torch.sort(idx.long(), dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())
    sample = np.random.default_rng(1).choice(ref.shape[0], 5000, replace=False)
    oi, od, _ = oracle.KdTree(ref).knn(ref[sample], 50)
    assert np.array_equal(idx[torch.from_numpy(sample).cuda()].cpu().numpy(), oi)
    assert np.array_equal(bits(d2[torch.from_numpy(sample).cuda()].cpu().numpy()), bits(od))
