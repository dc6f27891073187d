"""results.txt writer: the text and the return code of computeSimilarity (src/comparator.cpp:1112-1636)
for hand-built inputs.  The expected strings below were typed from the reference's `myfile << ...` statements; every quirk asserted
here (integer divisions, the dummy correspondence, the colour line that prints "pcl1 over pcl2" in both branches) is cited in
pointcloudcomparator_b200/report.py."""
import numpy as np

from pointcloudcomparator_b200 import report


def _cluster(centre, n, seed):
    rng = np.random.default_rng(seed)
    return (np.asarray(centre, np.float32) + rng.normal(0, 0.01, (n, 3))).astype(np.float32)


def test_icp_failure_report_and_return_code():
    txt, rc = report.write_results("a.ply", "b.ply", 1000, 900, [], [], [], [], None, None, icp=False)
    assert rc == -1
    assert txt == ("Results of comparison between a.ply and b.ply\n" + "-" * 80 + "\n\n" + "----------------------------\n\n"
                   "ICP could not match the point clouds. They are probably too dissimilar.\n Brief comparison:\n"
                   "PCL1 has more points: 1000 over: 900\n")
    assert report.write_results("a", "b", 5, 5, [], [], [], [], None, None, icp=False)[0].endswith("Both PCL have the same number of points\n")
    assert "PCL2 has more points: 9 over: 5\n" in report.write_results("a", "b", 5, 9, [], [], [], [], None, None, icp=False)[0]


def test_matching_rule_integer_divisions():
    cen = [[0, 0, 0]]
    # coef = size2 / size1 in integers: 1999 / 1000 = 1 passes, 2000 / 1000 = 2 and 999 / 1000 = 0 do not
    for s2, ok in ((1000, True), (1999, True), (2000, False), (999, False)):
        m = report.match_clusters([1000], [s2], [10], [10], cen, cen, lambda i, j: 11)
        assert (m[0] == 0) == ok, s2
    # correspondences.size() / max(d1, d2) in integers: the vector has one leading dummy, so 9 matches + 1 = 10 of 10 descriptors passes, 8 + 1 does not
    assert report.match_clusters([1000], [1000], [10], [10], cen, cen, lambda i, j: 10)[0] == 0
    assert report.match_clusters([1000], [1000], [10], [10], cen, cen, lambda i, j: 9)[0] == -1
    assert report.match_clusters([1000], [1000], [10], [12], cen, cen, lambda i, j: 11)[0] == -1       # 11 / 12 = 0
    # more than 3 descriptors on both sides
    assert report.match_clusters([1000], [1000], [3], [10], cen, cen, lambda i, j: 99)[0] == -1
    # only the 3 closest centroids are tried, and the one with the most correspondences wins
    cen2 = [[5, 0, 0], [1, 0, 0], [2, 0, 0], [3, 0, 0]]
    sizes2, nd2 = [1000] * 4, [10] * 4
    assert report.match_clusters([1000], sizes2, [10], nd2, cen, cen2, lambda i, j: {0: 50, 1: 10, 2: 12, 3: 11}[j])[0] == 2
    assert report.closest_centroid([0, 0, 0], cen2, {1}) == 2 and report.closest_centroid([0, 0, 0], [], set()) == -1


def test_full_report_text():
    c1 = [_cluster((0, 0, 0), 1200, 1), _cluster((2, 0, 0), 800, 2)]
    c2 = [_cluster((2.01, 0, 0), 900, 3), _cluster((0.01, 0, 0), 1500, 4), _cluster((9, 9, 9), 60, 5)]
    nd1, nd2 = [40, 30], [28, 40, 5]
    corr = {(0, 1): 41, (1, 0): 30}                           # cluster 0 <-> 1 and 1 <-> 0 reach the integer-division bar
    colours = {(0, 1): (3, 5), (1, 0): (4, 4)}
    txt, rc = report.write_results("A.ply", "B.ply", 2100, 2500, c1, c2, nd1, nd2, lambda i, j: corr.get((i, j), 1), lambda i, j: colours[(i, j)],
                                   icp=True, noise_kept=(2000, 2400))
    cen = [report.centroid_f32(c) for c in c2]
    exp_head = ("Results of comparison between A.ply and B.ply\n" + "-" * 80 + "\n\n"
                "ICP has converged. Point clouds segmentation is as follows: \n"
                "Number of points of PCL 1: 2100\nNumber of points of PCL 2: 2500\n" + "+" * 40 + "\n"
                "Number of clusters of PCL 1: 2\nNumber of clusters of PCL 2: 3\n"
                "\n------------------------------------\nInformation of clusters of PCL2:\n------------------------------------\n"
                "PCL2 cluster 0:\n\tNumber of points: 900\n\tNumber of descriptors: 28\n"
                f"\tCoordinates of centroid: [{'%g' % cen[0][0]},{'%g' % cen[0][1]},{'%g' % cen[0][2]}]\n")
    assert txt.startswith(exp_head)
    sep = "      " + "+" * 58 + "\t\n"
    exp_tail = ("\n------------------------------------\nInformation of matches of clusters of PCL 1 and PCL 2:\n------------------------------------\n"
                "\tMatched cluster 0 of PCL 1 with cluster 1 of PCL 2:\n"
                "\t\tSegment of PCL 2 has more points: 1500 over: 1200\n"
                "\t\tBoth segments have the same number of descriptors: 40\n"
                "\t\tSegment of PCL 2 has more elements based on color differences: 3 over 5\n" + sep +
                "\tMatched cluster 1 of PCL 1 with cluster 0 of PCL 2:\n"
                "\t\tSegment of PCL 2 has more points: 900 over: 800\n"
                "\t\tSegment of PCL 1 has more descriptors: 30 over: 28\n"
                "\t\tSegment of PCL 1 and segment of PCL 2 have the same number of elements based on color differences: 4\n" + sep +
                "Total number of matches found: 2\n\n"
                "----------------------------------------\n Noise analysis: \n"
                "Both pcl have the same percentage of noisy points: 0\n"          # (2100 - 2000) / 2100 == 0 in size_t
                "\n----------------------------\n\n"
                "points score pcl1: 2000\npoints score pcl2: 2400\n\n"
                "descriptors score pcl1: 70\ndescriptors score pcl2: 68\n\n"
                "color elements score pcl1: 7\ncolor elements score pcl2: 9\n"
                "\n----------------------------\n\n"
                f"Ratio of similarity over the 2 matches: {'%g' % ((2000 / 2400 + 70 / 68 + 7 / 9) / 3)}\n"
                f"Ratio of general similarity of pcl 1 over pcl 2: {'%g' % ((2000 / 2400 + 70 / 68 + 7 / 9) / 3 * (2 / 3))}\n")
    assert txt.endswith(exp_tail), txt[-len(exp_tail):]
    assert rc == 2                                            # points: 2, descriptors: 1, colour: 2


def test_no_match_no_clusters_and_number_format():
    c1 = [_cluster((0, 0, 0), 100, 1)]
    txt, rc = report.write_results("A", "B", 100, 0, c1, [], [10], [], lambda i, j: 0, None)
    assert "\t\tCluster 0 of PCL 1 has no match in PCL 2\n" in txt and "Total number of matches found: 0\n" in txt
    assert txt.endswith("Ratio of similarity over the 0 matches: 0\nRatio of general similarity of pcl 1 over pcl 2: -nan\n") and rc == 0
    assert "Noise analysis" not in txt and "ICP" not in txt
    assert report._g(1234567.0) == "1.23457e+06" and report._g(0.000012345678) == "1.23457e-05" and report._g(346911.0) == "346911"
    # every point removed is the only way to a non-zero noise figure
    t2, _ = report.write_results("A", "B", 100, 50, c1, [], [10], [], lambda i, j: 0, None, noise_kept=(0, 25))
    assert "\tPCL1 has more noisy points: (%) 100 over: (%) 0\n" in t2


# ---- the C++ twin (include/pcc/report.hpp): identical text and return code on the same inputs ----------------------------------
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def report_exe(tmp_path_factory):
    cxx = shutil.which("g++")
    if not cxx:
        pytest.skip("no g++")
    exe = str(tmp_path_factory.mktemp("report") / "test_report")
    subprocess.run([cxx, "-O1", "-std=c++17", "-ffp-contract=off", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_report.cpp")], check=True)
    return exe


def _run_cpp(exe, name1, name2, n1, n2, c1, c2, nd1, nd2, corr, colours, icp, noise_kept):
    lines = [f"{name1} {name2} {n1} {n2} {-1 if icp is None else int(icp)} {0 if noise_kept is None else 1} "
             f"{0 if noise_kept is None else noise_kept[0]} {0 if noise_kept is None else noise_kept[1]} {len(c1)} {len(c2)}"]
    for c, nd in list(zip(c1, nd1)) + list(zip(c2, nd2)):
        lines.append(f"{len(c)} {nd} " + " ".join("%.9g" % v for v in np.asarray(c, np.float32)[:, :3].ravel()))
    lines.append(" ".join(str(corr(i, j)) for i in range(len(c1)) for j in range(len(c2))))
    lines.append(" ".join("%d %d" % colours(i, j) for i in range(len(c1)) for j in range(len(c2))))
    out = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    head, _, text = out.stdout.partition("\n")
    return text, int(head[3:])


def test_cpp_report_writer_matches_python_twin(report_exe):
    rng = np.random.default_rng(11)
    cases = []
    c1 = [_cluster((0, 0, 0), 1200, 1), _cluster((2, 0, 0), 800, 2)]
    c2 = [_cluster((2.01, 0, 0), 900, 3), _cluster((0.01, 0, 0), 1500, 4), _cluster((9, 9, 9), 60, 5)]
    corr = {(0, 1): 41, (1, 0): 30}
    cases.append(("A.ply", "B.ply", 2100, 2500, c1, c2, [40, 30], [28, 40, 5], lambda i, j: corr.get((i, j), 1), lambda i, j: ((3, 5), (4, 4))[i], True, (2000, 2400)))
    cases.append(("a", "b", 1000, 900, [], [], [], [], lambda i, j: 0, lambda i, j: (0, 0), False, None))
    cases.append(("a", "b", 100, 0, [_cluster((0, 0, 0), 100, 1)], [], [10], [], lambda i, j: 0, lambda i, j: (0, 0), None, None))
    cases.append(("a", "b", 100, 50, [_cluster((0, 0, 0), 100, 1)], [], [10], [], lambda i, j: 0, lambda i, j: (0, 0), None, (0, 25)))
    for seed in range(6):                                     # random scenes: 1-6 clusters per cloud at shared centres, random counts on both sides
        k1, k2 = int(rng.integers(1, 7)), int(rng.integers(1, 7))
        centres = rng.uniform(-3, 3, (8, 3))
        a = [_cluster(centres[i], int(rng.integers(300, 1500)), 100 + seed * 10 + i) for i in range(k1)]
        b = [_cluster(centres[j] + rng.normal(0, 0.02, 3), int(rng.integers(300, 1500)), 200 + seed * 10 + j) for j in range(k2)]
        nd1, nd2 = [int(v) for v in rng.integers(2, 60, k1)], [int(v) for v in rng.integers(2, 60, k2)]
        cm = rng.integers(1, 80, (k1, k2)); col = rng.integers(1, 9, (k1, k2, 2))
        na, nb = sum(len(x) for x in a) + 50, sum(len(x) for x in b) + 70
        cases.append((f"s{seed}a", f"s{seed}b", na, nb, a, b, nd1, nd2, (lambda i, j, cm=cm: int(cm[i, j])), (lambda i, j, col=col: (int(col[i, j, 0]), int(col[i, j, 1]))),
                      (None, True)[seed % 2], (None, (na - 7, nb))[seed % 3 == 0]))
    matched = 0
    for name1, name2, n1, n2, a, b, nd1, nd2, corr_fn, col_fn, icp, kept in cases:
        want, want_rc = report.write_results(name1, name2, n1, n2, a, b, nd1, nd2, corr_fn, col_fn, icp=icp, noise_kept=kept)
        got, got_rc = _run_cpp(report_exe, name1, name2, n1, n2, a, b, nd1, nd2, corr_fn, col_fn, icp, kept)
        assert got == want and got_rc == want_rc, (name1, got[-400:], want[-400:])
        matched += want.count("Matched cluster")
    assert matched >= 4                                       # the random scenes do exercise the match lines
