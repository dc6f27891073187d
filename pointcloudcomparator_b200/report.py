"""results.txt writer of computeSimilarity (src/comparator.cpp:1112-1636), restated so that the scores the reference prints can be
diffed against a run of the GPU pipeline.

Only the report logic lives here: cluster centroids, the 3-closest-centroid candidate search, the acceptance rule of a cluster match,
the per-match comparison lines, the noise lines, the score / ratio footer and the function's return code.  The quantities it consumes
(cluster point arrays, descriptor counts, the size of matchRIFTFeaturesKnn's correspondence vector, colour-segment counts, the
StatisticalOutlierRemoval survivor counts) come from the stages the C ABI serves.

The reference's arithmetic quirks are kept on purpose, because they decide which lines are printed:
  * `coef = size2 / size1` is an integer division assigned to a double (src/comparator.cpp:1315-1317): only 1 passes 0.5 < coef < 2;
  * `correspondences.size() / descriptors.size()` is an integer division too (:1333-1353): "> 0.5" means ">= 1", and the
    correspondence vector carries one leading dummy element (:568);
  * the noise percentages are `(n - kept) / n` in size_t (:1533-1535, :1547-1549): 0 unless every point was removed;
  * doubles and floats go through operator<< with the default precision (6 significant digits, %g).
"""
from __future__ import annotations

import math

import numpy as np

RULE = "\n--------------------------------------------------------------------------------\n\n"


def _g(x) -> str:
    """std::ostream << double / float with default flags."""
    x = float(x)
    if math.isnan(x):
        return "-nan" if math.copysign(1.0, x) < 0 else "nan"
    return "%g" % x


def centroid_f32(points) -> list:
    """The reference's centroid: float accumulators, points added in order, divided by the (integer) count (:1235-1248, :1275-1288)."""
    p = np.asarray(points, np.float32)[:, :3]
    c = [np.float32(0), np.float32(0), np.float32(0)]
    for row in p:
        c[0] = np.float32(c[0] + row[0]); c[1] = np.float32(c[1] + row[1]); c[2] = np.float32(c[2] + row[2])
    n = np.float32(len(p))
    return [np.float32(c[0] / n), np.float32(c[1] / n), np.float32(c[2] / n)]


def distance_centroids(c1, c2) -> np.float32:
    """distanceCentroids (:1047-1054): the differences are float subtractions, pow(float, 2) promotes each to double, the sum is stored
    in a float, sqrt of that float."""
    dx, dy, dz = (float(np.float32(np.float32(a) - np.float32(b))) for a, b in zip(c1[:3], c2[:3]))
    d = np.float32(dx ** 2 + dy ** 2 + dz ** 2)
    return np.float32(math.sqrt(float(d)))


def closest_centroid(centroid, centroids2, used) -> int:
    """closestCentroid (:1070-1087): nearest centroid of cloud 2 not in `used`; -1 when none is nearer than 1e12."""
    best, min_dis = -1, np.float32(1000000000000.0)
    for i, c in enumerate(centroids2):
        if i in used:
            continue
        d = distance_centroids(centroid, c)
        if d < min_dis:
            min_dis, best = d, i
    return best


def match_clusters(sizes1, sizes2, ndesc1, ndesc2, centroids1, centroids2, correspondence_size):
    """The matching loop of computeSimilarity (:1290-1366).  correspondence_size(i, j) = len(matchRIFTFeaturesKnn(desc1[i], desc2[j]))
    INCLUDING its leading dummy element.  Returns matches[i] = cluster of cloud 2 or -1."""
    matches = []
    for i in range(len(sizes1)):
        matches.append(-1)
        max_cor2 = 0
        used = set()
        closest = []
        c1 = closest_centroid(centroids1[i], centroids2, used); used.add(c1); closest.append(c1)
        c2 = closest_centroid(centroids1[i], centroids2, used); used.add(c2); closest.append(c2)
        closest.append(closest_centroid(centroids1[i], centroids2, used))
        for j in closest:
            if j == -1:
                continue
            d1, d2 = int(ndesc1[i]), int(ndesc2[j])
            if not (d2 > 3 and d1 > 3):                       # !empty() is implied
                continue
            coef = float(int(sizes2[j]) // int(sizes1[i]))     # integer division, then double
            if not (0.5 < coef < 2):
                continue
            cor = int(correspondence_size(i, j))
            denom = d1 if d1 > d2 else d2
            if (cor // denom) > 0.5 and cor > max_cor2:
                max_cor2 = cor
                matches[i] = j
    return matches


def write_results(name1: str, name2: str, n1: int, n2: int, clusters1, clusters2, ndesc1, ndesc2, correspondence_size, colour_segments,
                  icp: bool | None = None, noise_kept: tuple | None = None):
    """The text of results.txt and computeSimilarity's return value.

    clusters1 / clusters2: point arrays [m, >=3] per cluster; ndesc*: descriptors per cluster; correspondence_size(i, j) as in
    match_clusters; colour_segments(i, j) -> (count1, count2) = sizes of color_growing_segmentation's outputs for a matched pair;
    icp: None = -i not given, True / False = performICP's result; noise_kept = (kept1, kept2) survivors of StatisticalOutlierRemoval
    when -n is given."""
    out = ["Results of comparison between " + name1 + " and " + name2 + RULE]
    if icp is not None:
        if not icp:
            out.append("----------------------------\n\n")
            out.append("ICP could not match the point clouds. They are probably too dissimilar.\n Brief comparison:\n")
            if n1 > n2:
                out.append(f"PCL1 has more points: {n1} over: {n2}\n")
            elif n2 > n1:
                out.append(f"PCL2 has more points: {n2} over: {n1}\n")
            else:
                out.append("Both PCL have the same number of points\n")
            return "".join(out), -1
        out.append("ICP has converged. Point clouds segmentation is as follows: \n")
    sizes1, sizes2 = [len(c) for c in clusters1], [len(c) for c in clusters2]
    out.append(f"Number of points of PCL 1: {n1}\n")
    out.append(f"Number of points of PCL 2: {n2}\n")
    out.append("++++++++++++++++++++++++++++++++++++++++\n")
    out.append(f"Number of clusters of PCL 1: {len(clusters1)}\n")
    out.append(f"Number of clusters of PCL 2: {len(clusters2)}\n")
    out.append("\n------------------------------------\n")
    out.append("Information of clusters of PCL2:\n")
    out.append("------------------------------------\n")
    cen2 = []
    for j, c in enumerate(clusters2):
        cen = centroid_f32(c); cen2.append(cen)
        out.append(f"PCL2 cluster {j}:\n\tNumber of points: {sizes2[j]}\n\tNumber of descriptors: {int(ndesc2[j])}\n")
        out.append(f"\tCoordinates of centroid: [{_g(cen[0])},{_g(cen[1])},{_g(cen[2])}]\n")
    out.append("\n------------------------------------\n")
    out.append("Information of clusters of PCL 1:\n")
    out.append("------------------------------------\n")
    cen1 = []
    for i, c in enumerate(clusters1):
        cen = centroid_f32(c); cen1.append(cen)
        out.append(f"PCL1 cluster {i}:\n\tNumber of points: {sizes1[i]}\n\tNumber of descriptors: {int(ndesc1[i])}\n")
        out.append(f"\tCoordinates of centroid: [{_g(cen[0])},{_g(cen[1])},{_g(cen[2])}]\n")
    matches = match_clusters(sizes1, sizes2, ndesc1, ndesc2, cen1, cen2, correspondence_size)
    out.append("\n------------------------------------\n")
    out.append("Information of matches of clusters of PCL 1 and PCL 2:\n")
    out.append("------------------------------------\n")
    p1 = p2 = d1 = d2 = c1 = c2 = 0.0
    num_matches = 0.0
    for i, m in enumerate(matches):
        if m != -1:
            num_matches += 1
            out.append(f"\tMatched cluster {i} of PCL 1 with cluster {m} of PCL 2:\n")
            a, b = sizes1[i], sizes2[m]
            p1 += a; p2 += b
            if a > b:
                out.append(f"\t\tSegment of PCL 1 has more points: {a} over: {b}\n")
            elif a < b:
                out.append(f"\t\tSegment of PCL 2 has more points: {b} over: {a}\n")
            else:
                out.append(f"\t\tBoth segments have the same number of points: {a}\n")
            a, b = int(ndesc1[i]), int(ndesc2[m])
            d1 += a; d2 += b
            if a > b:
                out.append(f"\t\tSegment of PCL 1 has more descriptors: {a} over: {b}\n")
            elif a < b:
                out.append(f"\t\tSegment of PCL 2 has more descriptors: {b} over: {a}\n")
            else:
                out.append(f"\t\tBoth segments have the same number of descriptors: {a}\n")
            a, b = colour_segments(i, m)
            c1 += a; c2 += b
            if a > b:
                out.append(f"\t\tSegment of PCL 1 has more elements based on color differences: {a} over {b}\n")
            elif a < b:      # the reference prints the two counts in the SAME order here (pcl1 over pcl2), src/comparator.cpp:1475-1478
                out.append(f"\t\tSegment of PCL 2 has more elements based on color differences: {a} over {b}\n")
            else:
                out.append(f"\t\tSegment of PCL 1 and segment of PCL 2 have the same number of elements based on color differences: {a}\n")
        else:
            out.append(f"\t\tCluster {i} of PCL 1 has no match in PCL 2\n")
        out.append("      ++++++++++++++++++++++++++++++++++++++++++++++++++++++++++\t\n")
    out.append(f"Total number of matches found: {_g(num_matches)}\n\n")
    if noise_kept is not None:
        noise1 = float((n1 - int(noise_kept[0])) // n1) if n1 else float("nan")      # size_t / size_t
        noise2 = float((n2 - int(noise_kept[1])) // n2) if n2 else float("nan")
        out.append("----------------------------------------\n Noise analysis: \n")
        if noise1 > noise2:
            out.append(f"\tPCL1 has more noisy points: (%) {_g(noise1 * 100)} over: (%) {_g(noise2 * 100)}\n")
        elif noise1 < noise2:
            out.append(f"\tPCL2 has more noisy points: (%) {_g(noise2 * 100)} over: (%) {_g(noise1 * 100)}\n")
        else:
            out.append(f"Both pcl have the same percentage of noisy points: {_g(noise1 * 100)}\n")
    out.append("\n----------------------------\n\n")
    out.append(f"points score pcl1: {_g(p1)}\npoints score pcl2: {_g(p2)}\n\n")
    out.append(f"descriptors score pcl1: {_g(d1)}\ndescriptors score pcl2: {_g(d2)}\n\n")
    out.append(f"color elements score pcl1: {_g(c1)}\ncolor elements score pcl2: {_g(c2)}\n")
    out.append("\n----------------------------\n\n")
    r_points = p1 / p2 if p2 != 0 else 0.0
    r_des = d1 / d2 if d2 != 0 else 0.0
    r_color = c1 / c2 if c2 != 0 else 0.0
    ratio = (r_points + r_des + r_color) / 3
    out.append(f"Ratio of similarity over the {_g(num_matches)} matches: {_g(ratio)}\n")
    n_c2 = len(clusters2)
    # numMatches / clusters_pcl_2.size(): with no cluster in cloud 2 nothing can match, 0.0 / 0 is x86's default NaN (sign bit set), printed "-nan"
    out.append(f"Ratio of general similarity of pcl 1 over pcl 2: {_g(ratio * (num_matches / n_c2)) if n_c2 else '-nan'}\n")
    t1 = (p1 > p2) + (d1 > d2) + (c1 > c2)
    t2 = (p1 < p2) + (d1 < d2) + (c1 < c2)
    return "".join(out), (1 if t1 > t2 else 2 if t1 < t2 else 0)
