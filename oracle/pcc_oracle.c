/*
 * pcc_oracle.c -- CPU ORACLE for the batched kNN / radius hot path.
 *
 * THIS FILE IS TEST INFRASTRUCTURE.  It is a plain-C restatement of the CPU algorithms the
 * reference (adr-arroyo/PointCloudComparator) reaches through PCL 1.7 + FLANN, used ONLY as the
 * checker in tests/, in __graft_entry__.smoke() and as bench.py's cpu_baseline / --impl reference
 * leg.  Nothing under pointcloudcomparator_b200/ may import, link or call it.
 *
 * PARITY STATUS: "parity unpinned" against the reference itself.  The reference ships no tests or
 * golden vectors and its arithmetic lives in un-vendored PCL 1.7.x / FLANN
 * (1.8.4 on the author's Ubuntu 14.04) which cannot be built in this image.  What IS pinned
 * (tests/test_oracle_pinning.py, tests/golden/): the brute-force definition and the kd-tree
 * restatement below agree bit-for-bit (indices and fp32 squared distances) with OpenCV 4.13's
 * bundled FLANN KDTreeSingleIndex (same index family PCL drives) on tie-free data, and at set
 * level with scipy.spatial.cKDTree.
 *
 * Reference call sites restated here (file:line in /root/reference):
 *   search::KdTree::setInputCloud / nearestKSearch / radiusSearch  src/segmentation.cpp:120-122,169-171,232-234
 *   NormalEstimation (kNN k=50 / radius 0.03)                     src/segmentation.cpp:236-240, src/comparator.cpp:628-635,764-771
 *   RegionGrowing(RGB)::findPointNeighbours (k=100 table)         src/segmentation.cpp:249-271,179-190
 *   StatisticalOutlierRemoval (MeanK=50, 1.5 sigma)               src/comparator.cpp:1523-1527,1537-1541
 *   EuclideanClusterExtraction (tol 0.05, 100..250000)            src/segmentation.cpp:125-131
 *   IterativeClosestPoint (20 iterations) + getFitnessScore       src/comparator.cpp:1089-1110
 *   SIFT keypoint -> first cloud point within 0.05                src/comparator.cpp:696-713
 * Upstream units followed ([upstream] in SURVEY.md section 8c): FLANN kdtree_single_index.h, dist.h
 * (L2_Simple), result_set.h; PCL kdtree_flann.hpp, normal_3d.h, centroid.hpp, eigen.hpp,
 * statistical_outlier_removal.hpp, extract_clusters.hpp, icp.hpp, correspondence_estimation.hpp,
 * transformation_estimation_svd.hpp (Umeyama), default_convergence_criteria.hpp.
 *
 * Canonical tie rule (north_star): neighbours are ordered by (fp32 d2, original index).  FLANN's
 * kNN result set keeps visit order among equal d2; results therefore agree with FLANN wherever
 * d2 values are unique, and as distance multisets always.
 *
 * Build: see oracle/Makefile (gcc -O3 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* FLANN L2_Simple: result = 0; for each dim: diff = a-b; result += diff*diff  (fp32, no FMA)   */
static inline float orc_d2(const float *a, const float *b) {
    float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    float r = dx * dx;
    r = r + dy * dy;
    r = r + dz * dz;
    return r;
}
static inline int orc_finite3(const float *p) { return isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]); }
static inline int orc_less(float d, int32_t i, float wd, int32_t wi) { return d < wd || (d == wd && i < wi); }

/* insertion-sorted (d2, idx) list of capacity k, ascending lexicographic */
typedef struct { float *d; int32_t *i; int k, n; } orc_topk;
static inline void orc_topk_push(orc_topk *t, float d, int32_t idx) {
    if (t->n == t->k) {
        if (!orc_less(d, idx, t->d[t->k - 1], t->i[t->k - 1])) return;
    } else t->n++;
    int j = t->n - 1;
    while (j > 0 && orc_less(d, idx, t->d[j - 1], t->i[j - 1])) { t->d[j] = t->d[j - 1]; t->i[j] = t->i[j - 1]; --j; }
    t->d[j] = d; t->i[j] = idx;
}
static inline float orc_topk_worst(const orc_topk *t) { return t->n == t->k ? t->d[t->k - 1] : INFINITY; }

/* ------------------------------------------------------------------------------------------ */
/* kd-tree: restatement of flann::KDTreeSingleIndex (leaf_max_size 15, reorder = true)          */
typedef struct { int32_t left, right, child1, child2, divfeat; float divlow, divhigh; } orc_node;
typedef struct orc_tree {
    int64_t n;            /* number of indexed (finite) points */
    float *data;          /* n x 3, reordered */
    int32_t *vind;        /* position -> original index */
    orc_node *nodes; int64_t n_nodes, cap_nodes;
    float bb_lo[3], bb_hi[3];
    int leaf_max;
} orc_tree;

typedef struct { float lo[3], hi[3]; } orc_box;

static int32_t orc_new_node(orc_tree *t) {
    if (t->n_nodes == t->cap_nodes) { t->cap_nodes = t->cap_nodes ? t->cap_nodes * 2 : 1024; t->nodes = (orc_node *)realloc(t->nodes, sizeof(orc_node) * (size_t)t->cap_nodes); }
    return (int32_t)t->n_nodes++;
}
/* src = un-reordered staging coordinates addressed through ind[] */
static void orc_minmax(const float *src, const int32_t *ind, int64_t cnt, int dim, float *mn, float *mx) {
    float a = src[3 * (int64_t)ind[0] + dim], b = a;
    for (int64_t i = 1; i < cnt; ++i) { float v = src[3 * (int64_t)ind[i] + dim]; if (v < a) a = v; if (v > b) b = v; }
    *mn = a; *mx = b;
}
static void orc_plane_split(const float *src, int32_t *ind, int64_t cnt, int feat, float cutval, int64_t *lim1, int64_t *lim2) {
    int64_t left = 0, right = cnt - 1;
    for (;;) {
        while (left <= right && src[3 * (int64_t)ind[left] + feat] < cutval) ++left;
        while (left <= right && src[3 * (int64_t)ind[right] + feat] >= cutval) --right;
        if (left > right) break;
        int32_t tmp = ind[left]; ind[left] = ind[right]; ind[right] = tmp; ++left; --right;
    }
    *lim1 = left; right = cnt - 1;
    for (;;) {
        while (left <= right && src[3 * (int64_t)ind[left] + feat] <= cutval) ++left;
        while (left <= right && src[3 * (int64_t)ind[right] + feat] > cutval) --right;
        if (left > right) break;
        int32_t tmp = ind[left]; ind[left] = ind[right]; ind[right] = tmp; ++left; --right;
    }
    *lim2 = left;
}
static int32_t orc_divide(orc_tree *t, const float *src, int32_t *ind, int64_t left, int64_t right, orc_box *bbox) {
    int32_t id = orc_new_node(t);
    int64_t cnt = right - left;
    if (cnt <= t->leaf_max) {
        orc_node nd; nd.child1 = nd.child2 = -1; nd.left = (int32_t)left; nd.right = (int32_t)right; nd.divfeat = 0; nd.divlow = nd.divhigh = 0.f;
        t->nodes[id] = nd;
        for (int d = 0; d < 3; ++d) orc_minmax(src, ind + left, cnt, d, &bbox->lo[d], &bbox->hi[d]);
        return id;
    }
    /* middle split: widest bbox side (within 1e-5), among those the widest actual spread */
    const float EPS = 0.00001f;
    float max_span = bbox->hi[0] - bbox->lo[0];
    for (int d = 1; d < 3; ++d) { float s = bbox->hi[d] - bbox->lo[d]; if (s > max_span) max_span = s; }
    float max_spread = -1.f; int cutfeat = 0;
    for (int d = 0; d < 3; ++d) {
        float s = bbox->hi[d] - bbox->lo[d];
        if (s > (1 - EPS) * max_span) {
            float mn, mx; orc_minmax(src, ind + left, cnt, d, &mn, &mx);
            if (mx - mn > max_spread) { cutfeat = d; max_spread = mx - mn; }
        }
    }
    float split_val = (bbox->lo[cutfeat] + bbox->hi[cutfeat]) / 2;
    float mn, mx; orc_minmax(src, ind + left, cnt, cutfeat, &mn, &mx);
    float cutval = split_val < mn ? mn : (split_val > mx ? mx : split_val);
    int64_t lim1, lim2, idx;
    orc_plane_split(src, ind + left, cnt, cutfeat, cutval, &lim1, &lim2);
    if (lim1 > cnt / 2) idx = lim1; else if (lim2 < cnt / 2) idx = lim2; else idx = cnt / 2;
    orc_box lb = *bbox, rb = *bbox;
    lb.hi[cutfeat] = cutval; rb.lo[cutfeat] = cutval;
    int32_t c1 = orc_divide(t, src, ind, left, left + idx, &lb);
    int32_t c2 = orc_divide(t, src, ind, left + idx, right, &rb);
    orc_node nd; nd.left = nd.right = 0; nd.child1 = c1; nd.child2 = c2; nd.divfeat = cutfeat;
    nd.divlow = lb.hi[cutfeat]; nd.divhigh = rb.lo[cutfeat];
    t->nodes[id] = nd;
    for (int d = 0; d < 3; ++d) { bbox->lo[d] = lb.lo[d] < rb.lo[d] ? lb.lo[d] : rb.lo[d]; bbox->hi[d] = lb.hi[d] > rb.hi[d] ? lb.hi[d] : rb.hi[d]; }
    return id;
}

/* KdTreeFLANN::setInputCloud: non-finite points are skipped, original indices kept (index_mapping_) */
ORC_API orc_tree *orc_tree_build(const float *pts, int64_t n, int stride_f, int leaf_max) {
    orc_tree *t = (orc_tree *)calloc(1, sizeof(orc_tree));
    t->leaf_max = leaf_max > 0 ? leaf_max : 15;
    float *src = (float *)malloc(sizeof(float) * 3 * (size_t)(n > 0 ? n : 1));
    int32_t *ind = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    int32_t *orig = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float *p = pts + i * stride_f;
        if (!orc_finite3(p)) continue;
        src[3 * m] = p[0]; src[3 * m + 1] = p[1]; src[3 * m + 2] = p[2]; orig[m] = (int32_t)i; ind[m] = (int32_t)m; ++m;
    }
    t->n = m;
    if (m > 0) {
        orc_box bb;
        for (int d = 0; d < 3; ++d) { bb.lo[d] = bb.hi[d] = src[d]; }
        for (int64_t i = 1; i < m; ++i) for (int d = 0; d < 3; ++d) { float v = src[3 * i + d]; if (v < bb.lo[d]) bb.lo[d] = v; if (v > bb.hi[d]) bb.hi[d] = v; }
        for (int d = 0; d < 3; ++d) { t->bb_lo[d] = bb.lo[d]; t->bb_hi[d] = bb.hi[d]; }
        orc_divide(t, src, ind, 0, m, &bb);
    }
    t->data = (float *)malloc(sizeof(float) * 3 * (size_t)(m > 0 ? m : 1));
    t->vind = (int32_t *)malloc(sizeof(int32_t) * (size_t)(m > 0 ? m : 1));
    for (int64_t i = 0; i < m; ++i) { memcpy(t->data + 3 * i, src + 3 * (int64_t)ind[i], 3 * sizeof(float)); t->vind[i] = orig[ind[i]]; }
    free(src); free(ind); free(orig);
    return t;
}
ORC_API void orc_tree_free(orc_tree *t) { if (!t) return; free(t->data); free(t->vind); free(t->nodes); free(t); }
ORC_API int64_t orc_tree_size(const orc_tree *t) { return t->n; }

static void orc_search_knn(const orc_tree *t, const float *q, int32_t node, float mindistsq, float *dists, orc_topk *res) {
    const orc_node *nd = &t->nodes[node];
    if (nd->child1 < 0) {
        for (int32_t i = nd->left; i < nd->right; ++i) orc_topk_push(res, orc_d2(q, t->data + 3 * (int64_t)i), t->vind[i]);
        return;
    }
    int f = nd->divfeat; float val = q[f];
    float diff1 = val - nd->divlow, diff2 = val - nd->divhigh;
    int32_t best, other; float cut;
    if (diff1 + diff2 < 0) { best = nd->child1; other = nd->child2; cut = (val - nd->divhigh) * (val - nd->divhigh); }
    else { best = nd->child2; other = nd->child1; cut = (val - nd->divlow) * (val - nd->divlow); }
    orc_search_knn(t, q, best, mindistsq, dists, res);
    float dst = dists[f];
    mindistsq = mindistsq + cut - dst;
    dists[f] = cut;
    if (mindistsq <= orc_topk_worst(res)) orc_search_knn(t, q, other, mindistsq, dists, res);
    dists[f] = dst;
}
static float orc_init_dists(const orc_tree *t, const float *q, float *dists) {
    float s = 0.f;
    for (int d = 0; d < 3; ++d) {
        dists[d] = 0.f;
        if (q[d] < t->bb_lo[d]) { dists[d] = (q[d] - t->bb_lo[d]) * (q[d] - t->bb_lo[d]); s += dists[d]; }
        if (q[d] > t->bb_hi[d]) { dists[d] = (q[d] - t->bb_hi[d]) * (q[d] - t->bb_hi[d]); s += dists[d]; }
    }
    return s;
}

/* nearestKSearch for nq queries.  Output rows have stride k; row entries >= min(k, size) are (-1, +inf).
 * Non-finite queries return an empty row.  Returns k_eff = min(k, indexed points). */
ORC_API int orc_tree_knn(const orc_tree *t, const float *q, int64_t nq, int qstride_f, int k, int32_t *out_idx, float *out_d2, int threads) {
    int keff = (int64_t)k < t->n ? k : (int)t->n;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 256) num_threads(threads)
#endif
    for (int64_t i = 0; i < nq; ++i) {
        const float *qp = q + i * qstride_f;
        int32_t *oi = out_idx + i * k; float *od = out_d2 + i * k;
        for (int j = 0; j < k; ++j) { oi[j] = -1; od[j] = INFINITY; }
        if (keff == 0 || !orc_finite3(qp)) continue;
        orc_topk r; r.d = od; r.i = oi; r.k = keff; r.n = 0;
        float dists[3]; float s = orc_init_dists(t, qp, dists);
        orc_search_knn(t, qp, 0, s, dists, &r);
    }
    (void)threads;
    return keff;
}

/* radius result buffer (growable), RadiusResultSet: keep dist < radius (strict) */
typedef struct { float *d; int32_t *i; int64_t n, cap; float r2; } orc_rad;
static inline void orc_rad_push(orc_rad *r, float d, int32_t idx) {
    if (!(d < r->r2)) return;
    if (r->n == r->cap) { r->cap = r->cap ? 2 * r->cap : 64; r->d = (float *)realloc(r->d, sizeof(float) * (size_t)r->cap); r->i = (int32_t *)realloc(r->i, sizeof(int32_t) * (size_t)r->cap); }
    r->d[r->n] = d; r->i[r->n] = idx; r->n++;
}
static void orc_search_rad(const orc_tree *t, const float *q, int32_t node, float mindistsq, float *dists, orc_rad *res) {
    const orc_node *nd = &t->nodes[node];
    if (nd->child1 < 0) {
        for (int32_t i = nd->left; i < nd->right; ++i) orc_rad_push(res, orc_d2(q, t->data + 3 * (int64_t)i), t->vind[i]);
        return;
    }
    int f = nd->divfeat; float val = q[f];
    float diff1 = val - nd->divlow, diff2 = val - nd->divhigh;
    int32_t best, other; float cut;
    if (diff1 + diff2 < 0) { best = nd->child1; other = nd->child2; cut = (val - nd->divhigh) * (val - nd->divhigh); }
    else { best = nd->child2; other = nd->child1; cut = (val - nd->divlow) * (val - nd->divlow); }
    orc_search_rad(t, q, best, mindistsq, dists, res);
    float dst = dists[f];
    mindistsq = mindistsq + cut - dst;
    dists[f] = cut;
    if (mindistsq <= res->r2) orc_search_rad(t, q, other, mindistsq, dists, res);
    dists[f] = dst;
}
typedef struct { float d; int32_t i; } orc_pair;
static int orc_pair_cmp(const void *a, const void *b) {
    const orc_pair *x = (const orc_pair *)a, *y = (const orc_pair *)b;
    if (x->d < y->d) return -1; if (x->d > y->d) return 1;
    return (x->i > y->i) - (x->i < y->i);
}
/* one radius query -> canonical (d2, idx)-sorted list, capped to the max_nn smallest; returns count */
static int64_t orc_radius_one(const orc_tree *t, const float *qp, float r2, unsigned max_nn, orc_rad *buf, orc_pair **tmp, int64_t *tmpcap) {
    buf->n = 0; buf->r2 = r2;
    if (t->n == 0 || !orc_finite3(qp)) return 0;
    float dists[3]; float s = orc_init_dists(t, qp, dists);
    orc_search_rad(t, qp, 0, s, dists, buf);
    if (buf->n > *tmpcap) { *tmpcap = buf->n * 2; *tmp = (orc_pair *)realloc(*tmp, sizeof(orc_pair) * (size_t)*tmpcap); }
    for (int64_t j = 0; j < buf->n; ++j) { (*tmp)[j].d = buf->d[j]; (*tmp)[j].i = buf->i[j]; }
    qsort(*tmp, (size_t)buf->n, sizeof(orc_pair), orc_pair_cmp);
    int64_t m = buf->n;
    if (max_nn != 0 && (int64_t)max_nn < m) m = max_nn;
    return m;
}
/* KdTreeFLANN::radiusSearch: r2 = float(radius*radius) computed in double; CSR offsets[nq+1] */
ORC_API int orc_tree_radius_count(const orc_tree *t, const float *q, int64_t nq, int qstride_f, double radius, unsigned max_nn, int64_t *offsets, int threads) {
    float r2 = (float)(radius * radius);
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel num_threads(threads)
#endif
    {
        orc_rad buf = {0}; orc_pair *tmp = NULL; int64_t tmpcap = 0;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 256)
#endif
        for (int64_t i = 0; i < nq; ++i) offsets[i + 1] = orc_radius_one(t, q + i * qstride_f, r2, max_nn, &buf, &tmp, &tmpcap);
        free(buf.d); free(buf.i); free(tmp);
    }
    offsets[0] = 0;
    for (int64_t i = 0; i < nq; ++i) offsets[i + 1] += offsets[i];
    (void)threads;
    return 0;
}
ORC_API int orc_tree_radius_fill(const orc_tree *t, const float *q, int64_t nq, int qstride_f, double radius, unsigned max_nn, const int64_t *offsets, int32_t *out_idx, float *out_d2, int threads) {
    float r2 = (float)(radius * radius);
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel num_threads(threads)
#endif
    {
        orc_rad buf = {0}; orc_pair *tmp = NULL; int64_t tmpcap = 0;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 256)
#endif
        for (int64_t i = 0; i < nq; ++i) {
            int64_t m = orc_radius_one(t, q + i * qstride_f, r2, max_nn, &buf, &tmp, &tmpcap);
            int64_t o = offsets[i];
            for (int64_t j = 0; j < m; ++j) { out_idx[o + j] = tmp[j].i; out_d2[o + j] = tmp[j].d; }
        }
        free(buf.d); free(buf.i); free(tmp);
    }
    (void)threads;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* brute force: THE canonical definition ((fp32 d2, original index) order over finite points)    */
ORC_API int orc_brute_knn(const float *pts, int64_t n, int stride_f, const float *q, int64_t nq, int qstride_f, int k, int32_t *out_idx, float *out_d2, int threads) {
    int64_t nfin = 0;
    for (int64_t i = 0; i < n; ++i) nfin += orc_finite3(pts + i * stride_f);
    int keff = (int64_t)k < nfin ? k : (int)nfin;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads)
#endif
    for (int64_t i = 0; i < nq; ++i) {
        const float *qp = q + i * qstride_f;
        int32_t *oi = out_idx + i * k; float *od = out_d2 + i * k;
        for (int j = 0; j < k; ++j) { oi[j] = -1; od[j] = INFINITY; }
        if (keff == 0 || !orc_finite3(qp)) continue;
        orc_topk r; r.d = od; r.i = oi; r.k = keff; r.n = 0;
        for (int64_t j = 0; j < n; ++j) { const float *p = pts + j * stride_f; if (orc_finite3(p)) orc_topk_push(&r, orc_d2(qp, p), (int32_t)j); }
    }
    (void)threads;
    return keff;
}
/* brute radius, single call on a caller-provided CSR capacity: pass out_idx == NULL to only count */
ORC_API int orc_brute_radius(const float *pts, int64_t n, int stride_f, const float *q, int64_t nq, int qstride_f, double radius, unsigned max_nn, int64_t *offsets, int32_t *out_idx, float *out_d2) {
    float r2 = (float)(radius * radius);
    orc_pair *tmp = (orc_pair *)malloc(sizeof(orc_pair) * (size_t)(n > 0 ? n : 1));
    int64_t total = 0;
    if (!out_idx) offsets[0] = 0;
    for (int64_t i = 0; i < nq; ++i) {
        const float *qp = q + i * qstride_f; int64_t m = 0;
        if (orc_finite3(qp))
            for (int64_t j = 0; j < n; ++j) { const float *p = pts + j * stride_f; if (!orc_finite3(p)) continue; float d = orc_d2(qp, p); if (d < r2) { tmp[m].d = d; tmp[m].i = (int32_t)j; ++m; } }
        qsort(tmp, (size_t)m, sizeof(orc_pair), orc_pair_cmp);
        if (max_nn != 0 && (int64_t)max_nn < m) m = max_nn;
        if (out_idx) { int64_t o = offsets[i]; for (int64_t j = 0; j < m; ++j) { out_idx[o + j] = tmp[j].i; out_d2[o + j] = tmp[j].d; } }
        else offsets[i + 1] = (total += m);
    }
    free(tmp);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* NormalEstimation: computeMeanAndCovarianceMatrix (fp32 single pass) -> eigen33 -> flip        */
static void orc_roots2(float b, float c, float *roots) {
    roots[0] = 0.f;
    float d = (float)((double)b * (double)b - 4.0 * (double)c);   /* "Scalar (b * b - 4.0 * c)": double intermediate */
    if (d < 0.0f) d = 0.0f;
    float sd = sqrtf(d);
    roots[2] = 0.5f * (b + sd);
    roots[1] = 0.5f * (b - sd);
}
static void orc_swapf(float *a, float *b) { float t = *a; *a = *b; *b = t; }
/* m = symmetric 3x3 row-major, already scaled */
static void orc_roots(const float *m, float *roots) {
    float c0 = m[0] * m[4] * m[8] + 2.0f * m[1] * m[2] * m[5] - m[0] * m[5] * m[5] - m[4] * m[2] * m[2] - m[8] * m[1] * m[1];
    float c1 = m[0] * m[4] - m[1] * m[1] + m[0] * m[8] - m[2] * m[2] + m[4] * m[8] - m[5] * m[5];
    float c2 = m[0] + m[4] + m[8];
    if (fabsf(c0) < FLT_EPSILON) { orc_roots2(c2, c1, roots); return; }
    const float s_inv3 = (float)(1.0 / 3.0);
    const float s_sqrt3 = sqrtf(3.0f);
    float c2_over_3 = c2 * s_inv3;
    float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
    if (a_over_3 > 0.f) a_over_3 = 0.f;
    float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
    float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
    if (q > 0.f) q = 0.f;
    float rho = sqrtf(-a_over_3);
    float theta = atan2f(sqrtf(-q), half_b) * s_inv3;
    float cos_theta = cosf(theta), sin_theta = sinf(theta);
    roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
    roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
    roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
    if (roots[0] >= roots[1]) orc_swapf(&roots[0], &roots[1]);
    if (roots[1] >= roots[2]) { orc_swapf(&roots[1], &roots[2]); if (roots[0] >= roots[1]) orc_swapf(&roots[0], &roots[1]); }
    if (roots[0] <= 0.f) orc_roots2(c2, c1, roots);
}
static void orc_cross(const float *a, const float *b, float *c) {
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
/* pcl::eigen33 (smallest eigenpair overload) */
static void orc_eigen33(const float *cov, float *eval, float *evec) {
    float scale = 0.f;
    for (int i = 0; i < 9; ++i) { float a = fabsf(cov[i]); if (a > scale) scale = a; }
    if (scale <= FLT_MIN) scale = 1.0f;
    float m[9];
    for (int i = 0; i < 9; ++i) m[i] = cov[i] / scale;
    float roots[3]; orc_roots(m, roots);
    *eval = roots[0] * scale;
    m[0] -= roots[0]; m[4] -= roots[0]; m[8] -= roots[0];
    float v1[3], v2[3], v3[3];
    orc_cross(m + 0, m + 3, v1); orc_cross(m + 0, m + 6, v2); orc_cross(m + 3, m + 6, v3);
    float l1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];
    float l2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
    float l3 = v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2];
    const float *v; float l;
    if (l1 >= l2 && l1 >= l3) { v = v1; l = l1; } else if (l2 >= l1 && l2 >= l3) { v = v2; l = l2; } else { v = v3; l = l3; }
    float s = sqrtf(l);
    evec[0] = v[0] / s; evec[1] = v[1] / s; evec[2] = v[2] / s;
}
/* covariance block shared with tests: accu[9] in neighbour order, then /n, then E[pp^T]-mu mu^T */
ORC_API int orc_covariance(const float *pts, int stride_f, const int32_t *nbr, int64_t m, float *cov9, float *centroid3) {
    float a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int64_t cnt = 0;
    for (int64_t j = 0; j < m; ++j) {
        if (nbr[j] < 0) continue;
        const float *p = pts + (int64_t)nbr[j] * stride_f;
        if (!orc_finite3(p)) continue;
        a[0] += p[0] * p[0]; a[1] += p[0] * p[1]; a[2] += p[0] * p[2];
        a[3] += p[1] * p[1]; a[4] += p[1] * p[2]; a[5] += p[2] * p[2];
        a[6] += p[0]; a[7] += p[1]; a[8] += p[2];
        ++cnt;
    }
    if (cnt == 0) return 0;
    float fn = (float)cnt;
    for (int i = 0; i < 9; ++i) a[i] /= fn;
    centroid3[0] = a[6]; centroid3[1] = a[7]; centroid3[2] = a[8];
    cov9[0] = a[0] - a[6] * a[6]; cov9[1] = a[1] - a[6] * a[7]; cov9[2] = a[2] - a[6] * a[8];
    cov9[4] = a[3] - a[7] * a[7]; cov9[5] = a[4] - a[7] * a[8]; cov9[8] = a[5] - a[8] * a[8];
    cov9[3] = cov9[1]; cov9[6] = cov9[2]; cov9[7] = cov9[5];
    return (int)cnt;
}
/* out[nq*4] = nx, ny, nz, curvature.  Neighbour lists in CSR (offsets[nq+1]); entries < 0 are ignored.
 * qpts = the points the normals belong to (flip towards viewpoint uses them). */
ORC_API void orc_normals_from_lists(const float *pts, int stride_f, const float *qpts, int64_t nq, int qstride_f, const int64_t *offsets, const int32_t *nbr, float vpx, float vpy, float vpz, float *out) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int64_t i = 0; i < nq; ++i) {
        float *o = out + 4 * i;
        const int32_t *l = nbr + offsets[i]; int64_t m = offsets[i + 1] - offsets[i];
        int64_t valid = 0; for (int64_t j = 0; j < m; ++j) valid += (l[j] >= 0);
        float cov[9], cen[3];
        if (valid < 3 || orc_covariance(pts, stride_f, l, m, cov, cen) == 0) { o[0] = o[1] = o[2] = o[3] = NAN; continue; }
        float ev, n[3]; orc_eigen33(cov, &ev, n);
        float eig_sum = cov[0] + cov[4] + cov[8];
        float curv = eig_sum != 0.f ? fabsf(ev / eig_sum) : 0.f;
        const float *p = qpts + i * qstride_f;
        float vx = vpx - p[0], vy = vpy - p[1], vz = vpz - p[2];
        float cos_theta = (vx * n[0] + vy * n[1] + vz * n[2]);
        if (cos_theta < 0) { n[0] *= -1; n[1] *= -1; n[2] *= -1; }
        o[0] = n[0]; o[1] = n[1]; o[2] = n[2]; o[3] = curv;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* StatisticalOutlierRemoval: d2 rows of (mean_k+1) ascending; j = 0 assumed to be the query     */
ORC_API void orc_sor_mean_dist(const float *d2, int64_t nq, int row_stride, int mean_k, float *distances) {
    for (int64_t i = 0; i < nq; ++i) {
        const float *r = d2 + i * row_stride;
        if (!(r[0] < INFINITY)) { distances[i] = 0.0f; continue; }   /* nearestKSearch returned 0 */
        double s = 0.0;
        for (int k = 1; k < mean_k + 1; ++k) s += sqrt((double)r[k]);
        distances[i] = (float)(s / mean_k);
    }
}
/* mean / stddev (n-1) / threshold over valid entries; keep[i] = distances[i] <= thr.  Returns kept count. */
ORC_API int64_t orc_sor_threshold(const float *distances, const uint8_t *valid, int64_t n, double std_mul, double *mean_out, double *stddev_out, double *thr_out, uint8_t *keep) {
    double sum = 0, sq = 0; int64_t nv = 0;
    for (int64_t i = 0; i < n; ++i) { sum += distances[i]; sq += distances[i] * distances[i]; nv += valid ? valid[i] : 1; }
    double mean = sum / (double)nv;
    double var = (sq - sum * sum / (double)nv) / ((double)nv - 1);
    double sd = sqrt(var), thr = mean + std_mul * sd;
    int64_t kept = 0;
    for (int64_t i = 0; i < n; ++i) { uint8_t k = !(distances[i] > thr); if (keep) keep[i] = k; kept += k; }
    if (mean_out) *mean_out = mean; if (stddev_out) *stddev_out = sd; if (thr_out) *thr_out = thr;
    return kept;
}

/* ------------------------------------------------------------------------------------------ */
/* EuclideanClusterExtraction: BFS over the radius graph (connected components), size filter,
 * indices ascending inside a cluster, clusters by size descending (ties: smallest member first).
 * labels[i] = cluster rank in that order or -1.  Returns the number of kept clusters.
 * Note: PCL skips result j=0 ("the query itself"); with exact duplicates that entry can be the
 * duplicate instead.  The oracle skips by identity, i.e. it returns true connected components. */
typedef struct { int64_t size; int32_t first; int32_t raw; } orc_cl;
static int orc_cl_cmp(const void *a, const void *b) {
    const orc_cl *x = (const orc_cl *)a, *y = (const orc_cl *)b;
    if (x->size != y->size) return x->size > y->size ? -1 : 1;
    return (x->first > y->first) - (x->first < y->first);
}
ORC_API int64_t orc_ece(const orc_tree *t, const float *pts, int64_t n, int stride_f, double tolerance, int64_t min_size, int64_t max_size, int32_t *labels, int64_t *sizes_out, int64_t sizes_cap) {
    float r2 = (float)(tolerance * tolerance);
    int32_t *raw = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    int32_t *queue = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; ++i) raw[i] = -1;
    orc_cl *cls = NULL; int64_t ncl = 0, capcl = 0;
    orc_rad buf = {0};
    for (int64_t s = 0; s < n; ++s) {
        if (raw[s] >= 0 || !orc_finite3(pts + s * stride_f)) continue;
        int64_t head = 0, tail = 0; queue[tail++] = (int32_t)s; raw[s] = (int32_t)ncl;
        int32_t first = (int32_t)s;
        while (head < tail) {
            int32_t c = queue[head++];
            buf.n = 0; buf.r2 = r2;
            float dists[3]; float sd = orc_init_dists(t, pts + (int64_t)c * stride_f, dists);
            orc_search_rad(t, pts + (int64_t)c * stride_f, 0, sd, dists, &buf);
            for (int64_t j = 0; j < buf.n; ++j) { int32_t nb = buf.i[j]; if (nb == c || raw[nb] >= 0) continue; raw[nb] = (int32_t)ncl; queue[tail++] = nb; if (nb < first) first = nb; }
        }
        if (ncl == capcl) { capcl = capcl ? capcl * 2 : 256; cls = (orc_cl *)realloc(cls, sizeof(orc_cl) * (size_t)capcl); }
        cls[ncl].size = tail; cls[ncl].first = first; cls[ncl].raw = (int32_t)ncl; ++ncl;
    }
    /* size filter + ordering */
    int32_t *rank = (int32_t *)malloc(sizeof(int32_t) * (size_t)(ncl > 0 ? ncl : 1));
    orc_cl *kept = (orc_cl *)malloc(sizeof(orc_cl) * (size_t)(ncl > 0 ? ncl : 1)); int64_t nk = 0;
    for (int64_t c = 0; c < ncl; ++c) { rank[c] = -1; if (cls[c].size >= min_size && cls[c].size <= max_size) kept[nk++] = cls[c]; }
    qsort(kept, (size_t)nk, sizeof(orc_cl), orc_cl_cmp);
    for (int64_t c = 0; c < nk; ++c) { rank[kept[c].raw] = (int32_t)c; if (sizes_out && c < sizes_cap) sizes_out[c] = kept[c].size; }
    for (int64_t i = 0; i < n; ++i) labels[i] = raw[i] >= 0 ? rank[raw[i]] : -1;
    free(raw); free(queue); free(cls); free(rank); free(kept); free(buf.d); free(buf.i);
    return nk;
}

/* ------------------------------------------------------------------------------------------ */
/* 3x3 SVD (double, one-sided Jacobi) for Umeyama                                                */
static void orc_svd3(const double A[9], double U[9], double S[3], double V[9]) {
    /* eigen-decompose AtA by cyclic Jacobi -> V, S^2; U = A V S^-1 (with completion for rank deficiency) */
    double B[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += A[3 * k + i] * A[3 * k + j]; B[3 * i + j] = s; }
    for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = fabs(B[1]) + fabs(B[2]) + fabs(B[5]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p) for (int q = p + 1; q < 3; ++q) {
            double apq = B[3 * p + q]; if (fabs(apq) < 1e-300) continue;
            double theta = (B[3 * q + q] - B[3 * p + p]) / (2.0 * apq);
            double tt = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;
            for (int k = 0; k < 3; ++k) { double bkp = B[3 * k + p], bkq = B[3 * k + q]; B[3 * k + p] = c * bkp - s * bkq; B[3 * k + q] = s * bkp + c * bkq; }
            for (int k = 0; k < 3; ++k) { double bpk = B[3 * p + k], bqk = B[3 * q + k]; B[3 * p + k] = c * bpk - s * bqk; B[3 * q + k] = s * bpk + c * bqk; }
            for (int k = 0; k < 3; ++k) { double vkp = V[3 * k + p], vkq = V[3 * k + q]; V[3 * k + p] = c * vkp - s * vkq; V[3 * k + q] = s * vkp + c * vkq; }
        }
    }
    double ev[3] = {B[0], B[4], B[8]};
    int ord[3] = {0, 1, 2};
    for (int i = 0; i < 2; ++i) for (int j = i + 1; j < 3; ++j) if (ev[ord[j]] > ev[ord[i]]) { int t = ord[i]; ord[i] = ord[j]; ord[j] = t; }
    double Vs[9];
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) Vs[3 * r + c] = V[3 * r + ord[c]];
    memcpy(V, Vs, sizeof(Vs));
    for (int c = 0; c < 3; ++c) S[c] = sqrt(ev[ord[c]] > 0 ? ev[ord[c]] : 0);
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) { double s = 0; for (int k = 0; k < 3; ++k) s += A[3 * r + k] * V[3 * k + c]; U[3 * r + c] = s; }
    /* normalise columns; rebuild degenerate ones orthogonally */
    int good[3];
    for (int c = 0; c < 3; ++c) {
        double nrm = sqrt(U[c] * U[c] + U[3 + c] * U[3 + c] + U[6 + c] * U[6 + c]);
        good[c] = nrm > 1e-12 * (S[0] > 0 ? S[0] : 1.0) && nrm > 0;
        if (good[c]) for (int r = 0; r < 3; ++r) U[3 * r + c] /= nrm;
    }
    if (!good[0]) { U[0] = 1; U[3] = 0; U[6] = 0; }
    if (!good[1]) {
        double a[3] = {U[0], U[3], U[6]}; double e[3] = {0, 0, 0};
        int m = fabs(a[0]) < fabs(a[1]) ? (fabs(a[0]) < fabs(a[2]) ? 0 : 2) : (fabs(a[1]) < fabs(a[2]) ? 1 : 2); e[m] = 1;
        double d = a[m]; double v[3] = {e[0] - d * a[0], e[1] - d * a[1], e[2] - d * a[2]};
        double nrm = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        U[1] = v[0] / nrm; U[4] = v[1] / nrm; U[7] = v[2] / nrm;
    }
    if (!good[2]) {
        double a[3] = {U[0], U[3], U[6]}, b[3] = {U[1], U[4], U[7]};
        U[2] = a[1] * b[2] - a[2] * b[1]; U[5] = a[2] * b[0] - a[0] * b[2]; U[8] = a[0] * b[1] - a[1] * b[0];
    }
}
static double orc_det3(const double *m) {
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}
/* Umeyama without scaling from the 15 correspondence sums (double): sum_src[3], sum_tgt[3], sum_ts[9] (tgt x src^T), n.
 * Returns row-major 4x4 float.  TransformationEstimationSVD -> pcl::umeyama(src, tgt, false). */
ORC_API void orc_umeyama_from_sums(const double *sum_src, const double *sum_tgt, const double *sum_ts, double n, float *T16) {
    double ms[3], mt[3], sigma[9];
    for (int i = 0; i < 3; ++i) { ms[i] = sum_src[i] / n; mt[i] = sum_tgt[i] / n; }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) sigma[3 * i + j] = sum_ts[3 * i + j] / n - mt[i] * ms[j];
    double U[9], S[3], V[9]; orc_svd3(sigma, U, S, V);
    double Sg[3] = {1, 1, 1};
    if (orc_det3(sigma) < 0) Sg[2] = -1;
    int rank = 0; for (int i = 0; i < 3; ++i) if (!(fabs(S[i]) <= fabs(S[0]) * 1e-12)) ++rank;   /* isMuchSmallerThan */
    if (rank == 2) { if (orc_det3(U) * orc_det3(V) > 0) { Sg[2] = 1; } else { Sg[2] = -1; } }
    double R[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += U[3 * i + k] * Sg[k] * V[3 * j + k]; R[3 * i + j] = s; }
    for (int i = 0; i < 16; ++i) T16[i] = (i % 5 == 0) ? 1.f : 0.f;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) T16[4 * i + j] = (float)R[3 * i + j];
        T16[4 * i + 3] = (float)(mt[i] - (R[3 * i] * ms[0] + R[3 * i + 1] * ms[1] + R[3 * i + 2] * ms[2]));
    }
}
/* IterativeClosestPoint::transformCloud arithmetic: ((m0*x + m1*y) + m2*z) + m3, fp32 */
static inline void orc_xform(const float *T, const float *p, float *o) {
    for (int i = 0; i < 3; ++i) o[i] = ((T[4 * i] * p[0] + T[4 * i + 1] * p[1]) + T[4 * i + 2] * p[2]) + T[4 * i + 3];
}
static void orc_mat4_mul(const float *A, const float *B, float *C) {
    float r[16];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { float s = 0.f; for (int k = 0; k < 4; ++k) s += A[4 * i + k] * B[4 * k + j]; r[4 * i + j] = s; }
    memcpy(C, r, sizeof(r));
}
/* One correspondence pass (CorrespondenceEstimation::determineCorrespondences + the sums Umeyama needs).
 * cur = current (already transformed) source points, ns x 3 dense.  sums[16]: 0-2 sum src, 3-5 sum tgt,
 * 6-14 sum tgt_i*src_j (row-major), 15 sum d2 (all double).  Returns the number of correspondences. */
ORC_API int64_t orc_icp_pass(const orc_tree *t, const float *tgt, int tstride_f, const float *cur, int64_t ns, double *sums, int32_t *corr_idx, float *corr_d2, int threads) {
    int32_t *ci = corr_idx ? corr_idx : (int32_t *)malloc(sizeof(int32_t) * (size_t)(ns > 0 ? ns : 1));
    float *cd = corr_d2 ? corr_d2 : (float *)malloc(sizeof(float) * (size_t)(ns > 0 ? ns : 1));
    orc_tree_knn(t, cur, ns, 3, 1, ci, cd, threads);
    for (int i = 0; i < 16; ++i) sums[i] = 0;
    int64_t cnt = 0;
    for (int64_t i = 0; i < ns; ++i) {
        if (ci[i] < 0) continue;
        const float *s = cur + 3 * i; const float *g = tgt + (int64_t)ci[i] * tstride_f;
        for (int a = 0; a < 3; ++a) { sums[a] += s[a]; sums[3 + a] += g[a]; for (int b = 0; b < 3; ++b) sums[6 + 3 * a + b] += (double)g[a] * (double)s[b]; }
        sums[15] += cd[i]; ++cnt;
    }
    if (!corr_idx) free(ci); if (!corr_d2) free(cd);
    return cnt;
}
/* ICP driver as the reference configures it (src/comparator.cpp:1089-1110): max_iter iterations,
 * transformation_epsilon 0, euclidean_fitness_epsilon -DBL_MAX, no rejectors, no max distance.
 * Outputs: T16 final transform (row-major), converged flag, fitness (getFitnessScore), iterations, mse per iteration. */
ORC_API int orc_icp(const float *src, int64_t ns, int sstride_f, const float *tgt, int64_t nt, int tstride_f, int max_iter, float *T16, int *converged, double *fitness, int *iterations, double *mse_trace, int threads) {
    orc_tree *t = orc_tree_build(tgt, nt, tstride_f, 15);
    float *cur = (float *)malloc(sizeof(float) * 3 * (size_t)(ns > 0 ? ns : 1));
    for (int64_t i = 0; i < ns; ++i) memcpy(cur + 3 * i, src + i * sstride_f, 3 * sizeof(float));
    float Tfinal[16], Tstep[16];
    for (int i = 0; i < 16; ++i) Tfinal[i] = (i % 5 == 0) ? 1.f : 0.f;
    int it = 0, conv = 0; double prev_mse = DBL_MAX;
    for (;;) {
        double sums[16];
        int64_t cnt = orc_icp_pass(t, tgt, tstride_f, cur, ns, sums, NULL, NULL, threads);
        if (cnt < 3) { conv = 0; break; }
        orc_umeyama_from_sums(sums, sums + 3, sums + 6, (double)cnt, Tstep);
        for (int64_t i = 0; i < ns; ++i) { float o[3]; float *p = cur + 3 * i; if (!orc_finite3(p)) continue; orc_xform(Tstep, p, o); p[0] = o[0]; p[1] = o[1]; p[2] = o[2]; }
        orc_mat4_mul(Tstep, Tfinal, Tfinal);
        ++it;
        /* DefaultConvergenceCriteria::hasConverged */
        if (it >= max_iter) { conv = 1; if (mse_trace) mse_trace[it - 1] = sums[15] / (double)cnt; break; }
        double cos_angle = 0.5 * ((double)Tstep[0] + (double)Tstep[5] + (double)Tstep[10] - 1);
        double tr2 = (double)Tstep[3] * Tstep[3] + (double)Tstep[7] * Tstep[7] + (double)Tstep[11] * Tstep[11];
        if (cos_angle >= 1.0 && tr2 <= 0.0) { conv = 1; break; }
        double mse = sums[15] / (double)cnt;
        if (mse_trace) mse_trace[it - 1] = mse;
        if (fabs(mse - prev_mse) < 1e-12) { conv = 1; break; }
        if (fabs(mse - prev_mse) / prev_mse < -DBL_MAX) { conv = 1; break; }
        prev_mse = mse;
    }
    memcpy(T16, Tfinal, sizeof(Tfinal));
    /* getFitnessScore: transform the ORIGINAL source by the final transform, mean of 1-NN d2 */
    for (int64_t i = 0; i < ns; ++i) { const float *p = src + i * sstride_f; float o[3]; orc_xform(Tfinal, p, o); memcpy(cur + 3 * i, o, sizeof(o)); }
    int32_t *ci = (int32_t *)malloc(sizeof(int32_t) * (size_t)(ns > 0 ? ns : 1));
    float *cd = (float *)malloc(sizeof(float) * (size_t)(ns > 0 ? ns : 1));
    orc_tree_knn(t, cur, ns, 3, 1, ci, cd, threads);
    double fs = 0; int64_t nr = 0;
    for (int64_t i = 0; i < ns; ++i) if (ci[i] >= 0) { fs += cd[i]; ++nr; }
    *fitness = nr > 0 ? fs / (double)nr : DBL_MAX;
    *converged = conv; *iterations = it;
    free(ci); free(cd); free(cur); orc_tree_free(t);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* SIFT keypoint snap (src/comparator.cpp:696-713): first cloud point (index order) with
 * sqrt(pow(dx,2)+pow(dy,2)+pow(dz,2)) < thr, float differences promoted to double.  -1 if none. */
ORC_API void orc_first_within(const float *pts, int64_t n, int stride_f, const float *q, int64_t nq, int qstride_f, double thr, int32_t *out) {
    for (int64_t i = 0; i < nq; ++i) {
        const float *qp = q + i * qstride_f; out[i] = -1;
        for (int64_t j = 0; j < n; ++j) {
            const float *p = pts + j * stride_f;
            double dx = (double)(qp[0] - p[0]), dy = (double)(qp[1] - p[1]), dz = (double)(qp[2] - p[2]);
            if (sqrt(dx * dx + dy * dy + dz * dz) < thr) { out[i] = (int32_t)j; break; }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* pcl::VoxelGrid<PointT>::applyFilter [upstream filters/impl/voxel_grid.hpp] as the reference configures it
 * (src/segmentation.cpp:69-74, 223-228: leaf 0.025, downsample_all_data = true, min_points_per_voxel = 0).
 * Points of a voxel are summed in ascending row order (PCL's std::sort leaves the order unspecified).
 * rows: stride_f floats; rgb_off_f = float offset of the packed BGRA word or -1.  Returns the number of output rows. */
typedef struct { uint32_t key; int32_t row; } orc_vx;
static int orc_vx_cmp(const void *a, const void *b) {
    const orc_vx *x = (const orc_vx *)a, *y = (const orc_vx *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return (x->row > y->row) - (x->row < y->row);
}
ORC_API int64_t orc_voxel_grid(const float *pts, int64_t n, int stride_f, int rgb_off_f, const float *leaf, int min_points, float *out) {
    float mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0}; int64_t m = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float *p = pts + i * stride_f;
        if (!orc_finite3(p)) continue;
        for (int d = 0; d < 3; ++d) { if (m == 0 || p[d] < mn[d]) mn[d] = p[d]; if (m == 0 || p[d] > mx[d]) mx[d] = p[d]; }
        ++m;
    }
    if (m == 0) return 0;
    float inv[3]; int min_b[3], mul[3]; int64_t div[3];
    for (int d = 0; d < 3; ++d) {
        inv[d] = 1.0f / leaf[d];
        min_b[d] = (int)floorf(mn[d] * inv[d]);
        div[d] = (int64_t)((int)floorf(mx[d] * inv[d])) - min_b[d] + 1;
    }
    {   /* PCL 1.7 VoxelGrid::applyFilter [upstream]: dx*dy*dz > INT_MAX -> PCL_WARN("Leaf size is too small ...") and `output = *input_`:
         * the cloud passes through unfiltered, every row (also the non-finite ones) */
        double dchk = 1.0;
        for (int d = 0; d < 3; ++d) dchk *= (double)((int64_t)((mx[d] - mn[d]) * inv[d]) + 1);
        if (dchk > 2147483647.0 || (double)div[0] * (double)div[1] * (double)div[2] > 2147483647.0) {
            memcpy(out, pts, sizeof(float) * (size_t)n * (size_t)stride_f);
            return n;
        }
    }
    mul[0] = 1; mul[1] = (int)div[0]; mul[2] = (int)(div[0] * div[1]);
    orc_vx *v = (orc_vx *)malloc(sizeof(orc_vx) * (size_t)m);
    int64_t c = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float *p = pts + i * stride_f;
        if (!orc_finite3(p)) continue;
        int i0 = (int)(floorf(p[0] * inv[0]) - (float)min_b[0]);
        int i1 = (int)(floorf(p[1] * inv[1]) - (float)min_b[1]);
        int i2 = (int)(floorf(p[2] * inv[2]) - (float)min_b[2]);
        v[c].key = (uint32_t)(i0 * mul[0] + i1 * mul[1] + i2 * mul[2]); v[c].row = (int32_t)i; ++c;
    }
    qsort(v, (size_t)m, sizeof(orc_vx), orc_vx_cmp);
    int64_t total = 0, index = 0;
    while (index < m) {
        int64_t i = index + 1;
        while (i < m && v[i].key == v[index].key) ++i;
        if (i - index >= (int64_t)min_points) {
            float sx = 0, sy = 0, sz = 0, sr = 0, sg = 0, sb = 0;
            for (int64_t j = index; j < i; ++j) {
                const float *p = pts + (int64_t)v[j].row * stride_f;
                if (j == index) { sx = p[0]; sy = p[1]; sz = p[2]; } else { sx += p[0]; sy += p[1]; sz += p[2]; }
                if (rgb_off_f >= 0) { const uint8_t *col = (const uint8_t *)(p + rgb_off_f); sb += (float)col[0]; sg += (float)col[1]; sr += (float)col[2]; }
            }
            float cnt = (float)(i - index);
            float *o = out + total * stride_f;
            memset(o, 0, sizeof(float) * (size_t)stride_f);
            o[0] = sx / cnt; o[1] = sy / cnt; o[2] = sz / cnt;
            if (stride_f >= 4) o[3] = 1.0f;
            if (rgb_off_f >= 0) { int r = (int)(sr / cnt), g = (int)(sg / cnt), b = (int)(sb / cnt); int rgb = (r << 16) | (g << 8) | b; memcpy(o + rgb_off_f, &rgb, 4); }
            ++total;
        }
        index = i;
    }
    free(v);
    return total;
}

/* ------------------------------------------------------------------------------------------ */
/* matchRIFTFeaturesKnn (src/comparator.cpp:560-588): 1-NN in descriptor space, FLANN L2_Simple over `dim` floats.
 * Brute force restatement (the reference uses a kd-tree; exact search, same result wherever d2 is unique). */
ORC_API void orc_descriptor_nn(const float *ref, int64_t n_ref, const float *qry, int64_t n_qry, int dim, int32_t *out_idx, float *out_d2) {
    for (int64_t t = 0; t < n_qry; ++t) {
        const float *q = qry + t * dim; int qok = 1;
        for (int i = 0; i < dim; ++i) qok &= isfinite(q[i]) != 0;
        float best = INFINITY; int32_t bi = -1;
        if (qok) for (int64_t r = 0; r < n_ref; ++r) {
            const float *p = ref + r * dim; int ok = 1; float d2 = 0.f;
            for (int i = 0; i < dim; ++i) { ok &= isfinite(p[i]) != 0; float diff = q[i] - p[i]; d2 += diff * diff; }
            if (ok && d2 < best) { best = d2; bi = (int32_t)r; }
        }
        out_idx[t] = bi; out_d2[t] = best;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* pcl::RegionGrowing (segmentation/impl/region_growing.hpp [upstream]) as configured at src/segmentation.cpp:249-271:
 * applySmoothRegionGrowingAlgorithm (seeds by ascending curvature), growRegion (FIFO of seeds, neighbour lists of k),
 * validatePoint (smooth mode: |n_nghbr . n_current| >= cos(theta); curvature flag decides whether the point seeds on),
 * assembleRegions + the size filter of extract.  labels = index among kept clusters in creation order, or -1. */
typedef struct { float res; int32_t idx; } orc_resid;
static int orc_resid_cmp(const void *a, const void *b) {
    const orc_resid *x = (const orc_resid *)a, *y = (const orc_resid *)b;
    const int nx = isnan(x->res), ny = isnan(y->res);          /* NaN curvatures last */
    if (nx != ny) return nx - ny;
    if (!nx) { if (x->res < y->res) return -1; if (x->res > y->res) return 1; }
    return (x->idx > y->idx) - (x->idx < y->idx);
}
ORC_API int64_t orc_region_growing(const int32_t *nbr, int64_t n, int k, const float *normals4, float theta, float curv_thr, int64_t min_size, int64_t max_size, int32_t *labels) {
    orc_resid *res = (orc_resid *)malloc(sizeof(orc_resid) * (size_t)(n > 0 ? n : 1));
    int32_t *point_labels = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    int32_t *seeds = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    int64_t *num_pts = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; ++i) { res[i].res = normals4[4 * i + 3]; res[i].idx = (int32_t)i; point_labels[i] = -1; }
    qsort(res, (size_t)n, sizeof(orc_resid), orc_resid_cmp);
    const float cosine_threshold = cosf(theta);
    int64_t segmented = 0, n_seg = 0, seed_counter = 0;
    while (segmented < n) {
        while (point_labels[res[seed_counter].idx] != -1) ++seed_counter;      /* next point that is not segmented yet */
        const int32_t seed = res[seed_counter].idx;
        int64_t head = 0, tail = 0, in_seg = 1;
        seeds[tail++] = seed; point_labels[seed] = (int32_t)n_seg;
        while (head < tail) {
            const int32_t cur = seeds[head++];
            for (int i_n = 0; i_n < k; ++i_n) {
                const int32_t index = nbr[(int64_t)cur * k + i_n];
                if (index < 0) break;
                if (point_labels[index] != -1) continue;
                const float *a = normals4 + 4 * (int64_t)index, *b = normals4 + 4 * (int64_t)cur;
                float dot = a[0] * b[0]; dot = dot + a[1] * b[1]; dot = dot + a[2] * b[2];
                if (fabsf(dot) < cosine_threshold) continue;
                point_labels[index] = (int32_t)n_seg; ++in_seg;
                int is_a_seed = 1;
                if (a[3] > curv_thr) is_a_seed = 0;
                if (is_a_seed) seeds[tail++] = index;
            }
        }
        num_pts[n_seg++] = in_seg; segmented += in_seg;
    }
    int64_t kept = 0;
    for (int64_t s2 = 0; s2 < n_seg; ++s2) { int64_t sz = num_pts[s2]; num_pts[s2] = (sz >= min_size && sz <= max_size) ? kept++ : -1; }
    for (int64_t i = 0; i < n; ++i) labels[i] = (int32_t)num_pts[point_labels[i]];
    free(res); free(point_labels); free(seeds); free(num_pts);
    return kept;
}

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
