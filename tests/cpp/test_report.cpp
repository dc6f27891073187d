// Driver for include/pcc/report.hpp: reads one fixture from stdin, prints "rc=<code>" and the results.txt text.  tests/test_report.py feeds it
// the fixtures of the Python twin (pointcloudcomparator_b200/report.py) and requires identical text and return code.  Host-only: no GPU, no library.
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/pcc/report.hpp"

int main() {
    std::string name1, name2; std::size_t n1, n2, kept[2]; int icp, noise, nc1, nc2;
    if (!(std::cin >> name1 >> name2 >> n1 >> n2 >> icp >> noise >> kept[0] >> kept[1] >> nc1 >> nc2)) return 2;
    std::vector<std::vector<float> > pts((std::size_t)(nc1 + nc2));
    std::vector<pcc::report::Cluster> c1, c2; std::vector<std::size_t> nd1, nd2;
    for (int c = 0; c < nc1 + nc2; ++c) {
        std::size_t m, nd; std::cin >> m >> nd;
        pts[(std::size_t)c].resize(m * 3);
        for (std::size_t i = 0; i < m * 3; ++i) std::cin >> pts[(std::size_t)c][i];
        pcc::report::Cluster cl; cl.xyz = pts[(std::size_t)c].data(); cl.size = m; cl.stride_floats = 3;
        if (c < nc1) { c1.push_back(cl); nd1.push_back(nd); } else { c2.push_back(cl); nd2.push_back(nd); }
    }
    std::vector<long> corr((std::size_t)(nc1 * nc2)), col1(corr.size()), col2(corr.size());
    for (std::size_t i = 0; i < corr.size(); ++i) std::cin >> corr[i];
    for (std::size_t i = 0; i < corr.size(); ++i) std::cin >> col1[i] >> col2[i];
    if (!std::cin) return 2;
    const pcc::report::Result r = pcc::report::write_results(
        name1, name2, n1, n2, c1, c2, nd1, nd2, [&](int i, int j) { return corr[(std::size_t)(i * nc2 + j)]; },
        [&](int i, int j, std::size_t *a, std::size_t *b) { *a = (std::size_t)col1[(std::size_t)(i * nc2 + j)]; *b = (std::size_t)col2[(std::size_t)(i * nc2 + j)]; },
        icp, noise ? kept : nullptr);
    std::printf("rc=%d\n", r.code);
    std::fputs(r.text.c_str(), stdout);
    return 0;
}
