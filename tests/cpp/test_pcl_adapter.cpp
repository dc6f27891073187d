// Compiles the PCC_HAVE_PCL branch of include/pcc/grid_search.hpp -- pcc::search::GridSearch<PointT> DERIVED from
// pcl::search::Search<PointT> -- against tests/cpp/pcl_stub (PCL 1.7's published interface; the image has no PCL), and drives it
// the way the reference's consumers do after setSearchMethod(tree) (src/segmentation.cpp:129,182,237,262): through a
// pcl::search::Search<PointT>::Ptr, one virtual call per point.  Every answer is checked against an in-test brute force.
// Built by __graft_entry__.build(); run on the GPU box by tests/test_gpu_cpp_host.py.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <type_traits>
#include <utility>

#include <pcl/search/search.h>                    // the stub: grid_search.hpp finds it through __has_include and takes the PCL branch
#include "../../include/pcc/grid_search.hpp"

#ifndef PCC_HAVE_PCL
#error "the PCL branch of grid_search.hpp was not selected"
#endif

typedef pcl::PointXYZRGB P;
typedef pcl::PointCloud<P> Cloud;
static_assert(std::is_base_of<pcl::search::Search<P>, pcc::search::GridSearch<P> >::value, "GridSearch must derive from pcl::search::Search");
static_assert(!std::is_abstract<pcc::search::GridSearch<P> >::value, "GridSearch must override both pure virtuals");
static_assert(sizeof(P) == 32 && alignof(P) == 16, "PointXYZRGB layout");

static unsigned long long rng_state = 0x9E3779B97F4A7C15ull;
static float frand() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (float)((rng_state >> 11) & 0xFFFFFF) / 16777216.0f; }
static float d2(const P &a, const P &b) { float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z; float r = dx * dx; r = r + dy * dy; r = r + dz * dz; return r; }
#define REQUIRE(cond) do { if (!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); return 1; } } while (0)

// what a PCL consumer looks like from the search object's side: it only knows the base class
static int consumer_mean_k(const pcl::search::Search<P>::Ptr &tree, const Cloud &cloud, int k, std::vector<float> &mean) {
    std::vector<int> idx; std::vector<float> dist;
    mean.assign(cloud.size(), 0.f);
    for (size_t i = 0; i < cloud.size(); ++i) {
        if (tree->nearestKSearch((int)i, k + 1, idx, dist) != k + 1) return 1;
        double s = 0; for (int j = 1; j <= k; ++j) s += std::sqrt((double)dist[j]);
        mean[i] = (float)(s / k);
    }
    return 0;
}

int main() {
    Cloud::Ptr cloud(new Cloud);
    for (int i = 0; i < 3000; ++i) { P p; p.x = frand(); p.y = frand(); p.z = 0.05f * frand(); cloud->push_back(p); }
    pcl::search::Search<P>::Ptr tree(new pcc::search::GridSearch<P>());       // the swap INTEGRATION.md shows
    tree->setInputCloud(cloud);
    REQUIRE(tree->getName() == "pcc::search::GridSearch");
    REQUIRE(tree->getInputCloud() == cloud);
    tree->setSortedResults(true);
    REQUIRE(tree->getSortedResults());

    std::vector<int> idx; std::vector<float> dist;
    for (int t = 0; t < 40; ++t) {
        P q; q.x = frand(); q.y = frand(); q.z = 0.05f * frand();
        std::vector<std::pair<float, int> > ref;
        for (size_t i = 0; i < cloud->size(); ++i) ref.push_back(std::make_pair(d2(q, (*cloud)[i]), (int)i));
        std::sort(ref.begin(), ref.end());
        const int k = 1 + t % 17;
        REQUIRE(tree->nearestKSearch(q, k, idx, dist) == k);                  // virtual dispatch into the CUDA engine
        for (int j = 0; j < k; ++j) { REQUIRE(idx[j] == ref[j].second); REQUIRE(dist[j] == ref[j].first); }
        const double r = 0.03 + 0.002 * t;
        const float r2 = (float)(r * r);
        const int m = tree->radiusSearch(q, r, idx, dist);
        int expect = 0; while (expect < (int)ref.size() && ref[expect].first < r2) ++expect;
        REQUIRE(m == expect);
        for (int j = 0; j < m; ++j) { REQUIRE(idx[j] == ref[j].second); REQUIRE(dist[j] == ref[j].first); }
        REQUIRE(tree->radiusSearch(q, r, idx, dist, 5) == std::min(expect, 5));
    }
    // index overloads and the batched overloads, all through the base pointer
    REQUIRE(tree->nearestKSearch(7, 4, idx, dist) == 4 && idx[0] == 7 && dist[0] == 0.f);
    REQUIRE(tree->nearestKSearch(*cloud, 9, 2, idx, dist) == 2 && idx[0] == 9);
    REQUIRE(tree->radiusSearch(11, 0.05, idx, dist) >= 1 && idx[0] == 11);
    std::vector<std::vector<int> > bi; std::vector<std::vector<float> > bd;
    tree->nearestKSearch(*cloud, std::vector<int>(), 8, bi, bd);
    REQUIRE(bi.size() == cloud->size() && bi[5].size() == 8 && bi[5][0] == 5);
    std::vector<int> some; some.push_back(3); some.push_back(2999);
    tree->radiusSearch(*cloud, some, 0.04, bi, bd, 0);
    REQUIRE(bi.size() == 2 && bi[0][0] == 3 && bi[1][0] == 2999);
    // a serial PCL-style consumer on top of the base pointer agrees with the fused batched call of the derived class
    std::vector<float> mean;
    REQUIRE(consumer_mean_k(tree, *cloud, 8, mean) == 0);
    std::vector<float> fused(cloud->size());
    REQUIRE(pcc_knn_mean_dist(static_cast<pcc::search::GridSearch<P> *>(tree.get())->handle(), nullptr, 0, (int)sizeof(P), 8, fused.data(), PCC_HOST, nullptr) == PCC_OK);
    for (size_t i = 0; i < mean.size(); ++i) REQUIRE(fused[i] == mean[i]);
    std::printf("PCL adapter ok: GridSearch<PointXYZRGB> behind pcl::search::Search<PointT>::Ptr (stub headers), %zu points\n", cloud->size());
    return 0;
}
