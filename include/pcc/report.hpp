// pcc/report.hpp -- the results.txt writer of computeSimilarity (src/comparator.cpp:1112-1636), host-only C++ (no CUDA, no PCL).
//
// SURVEY.md section 8f row 4.  A maintainer who moves the expensive stages onto libpcc_search keeps the reference's report by
// feeding their outputs to pcc::report::write_results: cluster point arrays, descriptor counts, the size of matchRIFTFeaturesKnn's
// correspondence vector (pcc::matchRIFTFeaturesKnn, leading dummy element included), the colour-segment counts
// (pcc::RegionGrowingRGB) and the StatisticalOutlierRemoval survivor counts.  The text is produced with operator<< exactly as the
// reference does, so the number formatting is the reference's by construction.
//
// The reference's arithmetic quirks decide which lines are printed and are kept on purpose:
//   * `coef = size2 / size1` is an integer division assigned to a double (:1315-1317): only 1 passes 0.5 < coef < 2;
//   * `correspondences.size() / descriptors.size()` is an integer division too (:1333-1353): "> 0.5" means ">= 1";
//   * the noise percentages are `(n - kept) / n` in size_t (:1533-1535, :1547-1549): 0 unless every point was removed;
//   * the colour line prints "pcl1 over pcl2" in BOTH branches (:1475-1478);
//   * centroids use float accumulators in point order; distanceCentroids subtracts in float, squares in double, stores the sum
//     in a float (:1047-1054).
// The Python twin is pointcloudcomparator_b200/report.py; tests/test_report.py checks both against strings typed from the
// reference's `myfile <<` statements and against each other (tests/cpp/test_report.cpp).
#ifndef PCC_REPORT_HPP_
#define PCC_REPORT_HPP_

#include <cmath>
#include <cstddef>
#include <set>
#include <sstream>
#include <string>
#include <vector>

namespace pcc {
namespace report {

struct Cluster {              // m points, xyz in the first three floats of a row, rows `stride_floats` apart
    const float *xyz; std::size_t size; std::size_t stride_floats;
};
struct Result { std::string text; int code; };   // code: computeSimilarity's return value (-1 ICP failed, 0 tie, 1 / 2 = which cloud scores higher)

// float accumulators, points added in order, divided by the count (:1235-1248, :1275-1288)
inline std::vector<float> centroid(const Cluster &c) {
    std::vector<float> cen(3, 0.f);
    for (std::size_t l = 0; l < c.size; ++l) { const float *p = c.xyz + l * c.stride_floats; cen[0] += p[0]; cen[1] += p[1]; cen[2] += p[2]; }
    cen[0] = cen[0] / c.size; cen[1] = cen[1] / c.size; cen[2] = cen[2] / c.size;
    return cen;
}
// distanceCentroids (:1047-1054)
inline float distance_centroids(const std::vector<float> &c1, const std::vector<float> &c2) {
    const float dx = c1[0] - c2[0], dy = c1[1] - c2[1], dz = c1[2] - c2[2];
    const float distance = (float)(std::pow((double)dx, 2) + std::pow((double)dy, 2) + std::pow((double)dz, 2));
    return (float)std::sqrt((double)distance);
}
// closestCentroid (:1070-1087): nearest centroid of cloud 2 not in `used`; -1 when none is nearer than 1e12
inline int closest_centroid(const std::vector<float> &cen, const std::vector<std::vector<float> > &centroids2, const std::set<int> &used) {
    int best = -1; float min_dis = 1000000000000.0f;
    for (std::size_t i = 0; i < centroids2.size(); ++i) {
        if (used.count((int)i)) continue;
        const float d = distance_centroids(cen, centroids2[i]);
        if (d < min_dis) { min_dis = d; best = (int)i; }
    }
    return best;
}
// The matching loop of computeSimilarity (:1290-1366).  correspondence_size(i, j) = matchRIFTFeaturesKnn(desc1[i], desc2[j]).size(),
// leading dummy element included.  matches[i] = cluster of cloud 2, or -1.
template <class CorrFn>
inline std::vector<int> match_clusters(const std::vector<std::size_t> &sizes1, const std::vector<std::size_t> &sizes2, const std::vector<std::size_t> &ndesc1,
                                       const std::vector<std::size_t> &ndesc2, const std::vector<std::vector<float> > &cen1,
                                       const std::vector<std::vector<float> > &cen2, CorrFn correspondence_size) {
    std::vector<int> matches;
    for (std::size_t i = 0; i < sizes1.size(); ++i) {
        matches.push_back(-1);
        std::size_t max_cor2 = 0;
        std::set<int> used; int closest[3];
        closest[0] = closest_centroid(cen1[i], cen2, used); used.insert(closest[0]);
        closest[1] = closest_centroid(cen1[i], cen2, used); used.insert(closest[1]);
        closest[2] = closest_centroid(cen1[i], cen2, used);
        for (int c = 0; c < 3; ++c) {
            const int j = closest[c];
            if (j == -1) continue;
            const std::size_t d1 = ndesc1[i], d2 = ndesc2[(std::size_t)j];
            if (!(d2 > 3 && d1 > 3)) continue;
            const double coef = (double)(sizes2[(std::size_t)j] / sizes1[i]);      // integer division, then double
            if (!(0.5 < coef && coef < 2)) continue;
            const std::size_t cor = (std::size_t)correspondence_size((int)i, j);
            const std::size_t denom = d1 > d2 ? d1 : d2;
            if ((cor / denom) > 0.5 && cor > max_cor2) { max_cor2 = cor; matches[i] = j; }
        }
    }
    return matches;
}

// The text of results.txt and computeSimilarity's return value.
//   icp: -1 = `-i` not given, 0 = performICP failed, 1 = converged.   noise_kept: nullptr, or the two StatisticalOutlierRemoval
//   survivor counts when `-n` is given.   colour_segments(i, j, &count1, &count2): sizes of color_growing_segmentation's outputs
//   for the matched pair (cluster i of cloud 1, cluster j of cloud 2).
template <class CorrFn, class ColourFn>
inline Result write_results(const std::string &name1, const std::string &name2, std::size_t n1, std::size_t n2, const std::vector<Cluster> &clusters1,
                            const std::vector<Cluster> &clusters2, const std::vector<std::size_t> &ndesc1, const std::vector<std::size_t> &ndesc2,
                            CorrFn correspondence_size, ColourFn colour_segments, int icp = -1, const std::size_t *noise_kept = nullptr) {
    std::ostringstream myfile;
    myfile << "Results of comparison between " << name1 << " and " << name2 << "\n--------------------------------------------------------------------------------\n\n";
    if (icp >= 0) {
        if (icp == 0) {
            myfile << "----------------------------\n\n";
            myfile << "ICP could not match the point clouds. They are probably too dissimilar.\n Brief comparison:\n";
            if (n1 > n2) myfile << "PCL1 has more points: " << n1 << " over: " << n2 << "\n";
            else if (n2 > n1) myfile << "PCL2 has more points: " << n2 << " over: " << n1 << "\n";
            else myfile << "Both PCL have the same number of points\n";
            Result r; r.text = myfile.str(); r.code = -1; return r;
        }
        myfile << "ICP has converged. Point clouds segmentation is as follows: \n";
    }
    std::vector<std::size_t> sizes1, sizes2;
    for (std::size_t i = 0; i < clusters1.size(); ++i) sizes1.push_back(clusters1[i].size);
    for (std::size_t j = 0; j < clusters2.size(); ++j) sizes2.push_back(clusters2[j].size);
    myfile << "Number of points of PCL 1: " << n1 << "\n";
    myfile << "Number of points of PCL 2: " << n2 << "\n";
    myfile << "++++++++++++++++++++++++++++++++++++++++\n";
    myfile << "Number of clusters of PCL 1: " << clusters1.size() << "\n";
    myfile << "Number of clusters of PCL 2: " << clusters2.size() << "\n";
    myfile << "\n------------------------------------\n" << "Information of clusters of PCL2:\n" << "------------------------------------\n";
    std::vector<std::vector<float> > cen1, cen2;
    for (std::size_t j = 0; j < clusters2.size(); ++j) {
        const std::vector<float> cen = centroid(clusters2[j]); cen2.push_back(cen);
        myfile << "PCL2 cluster " << j << ":\n\tNumber of points: " << sizes2[j] << "\n\tNumber of descriptors: " << ndesc2[j] << "\n";
        myfile << "\tCoordinates of centroid: [" << cen[0] << "," << cen[1] << "," << cen[2] << "]\n";
    }
    myfile << "\n------------------------------------\n" << "Information of clusters of PCL 1:\n" << "------------------------------------\n";
    for (std::size_t i = 0; i < clusters1.size(); ++i) {
        const std::vector<float> cen = centroid(clusters1[i]); cen1.push_back(cen);
        myfile << "PCL1 cluster " << i << ":\n\tNumber of points: " << sizes1[i] << "\n\tNumber of descriptors: " << ndesc1[i] << "\n";
        myfile << "\tCoordinates of centroid: [" << cen[0] << "," << cen[1] << "," << cen[2] << "]\n";
    }
    const std::vector<int> matches = match_clusters(sizes1, sizes2, ndesc1, ndesc2, cen1, cen2, correspondence_size);
    myfile << "\n------------------------------------\n" << "Information of matches of clusters of PCL 1 and PCL 2:\n" << "------------------------------------\n";
    double p1 = 0, p2 = 0, d1 = 0, d2 = 0, c1 = 0, c2 = 0, num_matches = 0;
    for (std::size_t i = 0; i < matches.size(); ++i) {
        const int m = matches[i];
        if (m != -1) {
            num_matches += 1;
            myfile << "\tMatched cluster " << i << " of PCL 1 with cluster " << m << " of PCL 2:\n";
            std::size_t a = sizes1[i], b = sizes2[(std::size_t)m];
            p1 += (double)a; p2 += (double)b;
            if (a > b) myfile << "\t\tSegment of PCL 1 has more points: " << a << " over: " << b << "\n";
            else if (a < b) myfile << "\t\tSegment of PCL 2 has more points: " << b << " over: " << a << "\n";
            else myfile << "\t\tBoth segments have the same number of points: " << a << "\n";
            a = ndesc1[i]; b = ndesc2[(std::size_t)m];
            d1 += (double)a; d2 += (double)b;
            if (a > b) myfile << "\t\tSegment of PCL 1 has more descriptors: " << a << " over: " << b << "\n";
            else if (a < b) myfile << "\t\tSegment of PCL 2 has more descriptors: " << b << " over: " << a << "\n";
            else myfile << "\t\tBoth segments have the same number of descriptors: " << a << "\n";
            std::size_t s1 = 0, s2 = 0;
            colour_segments((int)i, m, &s1, &s2);
            c1 += (double)s1; c2 += (double)s2;
            if (s1 > s2) myfile << "\t\tSegment of PCL 1 has more elements based on color differences: " << s1 << " over " << s2 << "\n";
            else if (s1 < s2) myfile << "\t\tSegment of PCL 2 has more elements based on color differences: " << s1 << " over " << s2 << "\n";
            else myfile << "\t\tSegment of PCL 1 and segment of PCL 2 have the same number of elements based on color differences: " << s1 << "\n";
        } else {
            myfile << "\t\tCluster " << i << " of PCL 1 has no match in PCL 2\n";
        }
        myfile << "      ++++++++++++++++++++++++++++++++++++++++++++++++++++++++++\t\n";
    }
    myfile << "Total number of matches found: " << num_matches << "\n\n";
    if (noise_kept) {
        const double nan = std::nan("");
        const double noise1 = n1 ? (double)((n1 - noise_kept[0]) / n1) : nan;     // size_t / size_t
        const double noise2 = n2 ? (double)((n2 - noise_kept[1]) / n2) : nan;
        myfile << "----------------------------------------\n Noise analysis: \n";
        if (noise1 > noise2) myfile << "\tPCL1 has more noisy points: (%) " << noise1 * 100 << " over: (%) " << noise2 * 100 << "\n";
        else if (noise1 < noise2) myfile << "\tPCL2 has more noisy points: (%) " << noise2 * 100 << " over: (%) " << noise1 * 100 << "\n";
        else myfile << "Both pcl have the same percentage of noisy points: " << noise1 * 100 << "\n";
    }
    myfile << "\n----------------------------\n\n";
    myfile << "points score pcl1: " << p1 << "\npoints score pcl2: " << p2 << "\n\n";
    myfile << "descriptors score pcl1: " << d1 << "\ndescriptors score pcl2: " << d2 << "\n\n";
    myfile << "color elements score pcl1: " << c1 << "\ncolor elements score pcl2: " << c2 << "\n";
    myfile << "\n----------------------------\n\n";
    const double r_points = p2 != 0 ? p1 / p2 : 0.0, r_des = d2 != 0 ? d1 / d2 : 0.0, r_color = c2 != 0 ? c1 / c2 : 0.0;
    const double ratio = (r_points + r_des + r_color) / 3;
    myfile << "Ratio of similarity over the " << num_matches << " matches: " << ratio << "\n";
    // numMatches / clusters_pcl_2.size(): with no cluster in cloud 2 nothing can match and 0.0 / 0 is x86's default NaN (sign bit set), printed "-nan"
    myfile << "Ratio of general similarity of pcl 1 over pcl 2: ";
    if (clusters2.empty()) myfile << "-nan"; else myfile << ratio * (num_matches / clusters2.size());
    myfile << "\n";
    const int t1 = (p1 > p2) + (d1 > d2) + (c1 > c2), t2 = (p1 < p2) + (d1 < d2) + (c1 < c2);
    Result r; r.text = myfile.str(); r.code = t1 > t2 ? 1 : (t1 < t2 ? 2 : 0);
    return r;
}

}  // namespace report
}  // namespace pcc
#endif  // PCC_REPORT_HPP_
