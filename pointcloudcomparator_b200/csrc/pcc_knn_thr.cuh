// pcc_knn_thr.cuh -- single-walk threshold selection for 2 <= k <= 32: the block kernel of pcc_knn (round 2).
//
// Why (profiles/r1/knn16_v8_sass_hist.txt): the round-1 block kernel kept the K smallest distances sorted in registers while it
// walked the 3x3x3 block; the guarded min/max sweep of that list was 34 % of all issued instructions and ran whenever ANY
// lane of the warp had something to insert (79 sweeps per warp for ~30 insertions per lane, 14 of 32 lanes active).
//
// Here nothing is kept sorted during the walk.  A candidate is LOGGED (its fp32 d2 and its position, 8 bytes, slot-major
// shared memory) when d2 <= tau, with
//        tau = dmin + T,      dmin = smallest d2 seen so far (one FMNMX per candidate),
//        T   = ratio[M] * k / M * cell^2,      M = population of the 3x3x3 block (known from the 9 run bounds before the walk).
// (k-th d2) - (1st d2) is what the local density predicts: for a query at height h over a surface of density rho,
// d2_j = h^2 + j / (pi rho), so the difference does not depend on h, while d2_k itself varies by 8x at a given M when the
// queries carry 1 cm of noise (measured on the headline workload: a threshold predicted from M alone needs 40 logged
// candidates for 95 % coverage; dmin + T needs 25 for 98 %).  `ratio` is a 64-entry table indexed by M / 4, calibrated on
// the device from a sample of the batch by knn_calib_kernel (quantile of (d2_k - d2_1) * M / k, exact search) and cached in
// the index; it only steers how much is logged -- the result is exact whatever it holds:
//   * tau only shrinks, so every candidate with d2 <= final tau is in the log;
//   * after the walk the K-wide sorting network below selects the k smallest logged keys; if the k-th of them has
//     d2 <= final tau the k smallest of the LOG are the k smallest of the BLOCK (every unlogged candidate is > final tau);
//   * otherwise (too tight a table: ~1 %) the query is listed and knn_thr_retry_kernel runs it again with 4 T; when the log
//     overflowed, or on a tie the 32-bit keys cannot order (below), the retry kernel runs the exact per-thread search (64-bit
//     keys, ring expansion) for it instead.
// Selection: the network sorts 32-bit keys = (d2 bits << 1, low bits replaced by the log slot): the payload rides along for
// free with two VIMNMX per compare-exchange and no 64-bit compares.  Dropping the low 6 (7) mantissa bits cannot misorder two
// candidates whose keys differ by >= 2^6 (2^7); adjacent keys closer than that (0.1 % of the queries; every query of a lattice
// cloud) are "ties" and take the exact path.  The smallest dropped key is tracked so the k-th / (k+1)-th boundary is checked too.
// The ring / wide passes behind this kernel are unchanged: it produces the same row + tau + proved flag as knn_fast_kernel.
#pragma once
#include <utility>

#include "pcc_device.cuh"

namespace pcc {

#ifndef PCC_THR_QAHEAD
#define PCC_THR_QAHEAD 0          // blocks of look-ahead for the query prefetch (0 = off)
#endif
#ifndef PCC_THR_EARLY
#define PCC_THR_EARLY 16
#endif
#ifndef PCC_THR_JOINT
#define PCC_THR_JOINT 1
#endif
#ifndef PCC_THR_SLOTS8
#define PCC_THR_SLOTS8 36
#endif
#ifndef PCC_THR_SLOTS16
#define PCC_THR_SLOTS16 48
#endif
#ifndef PCC_THR_SLOTS32
#define PCC_THR_SLOTS32 100
#endif
#ifndef PCC_THR_MB8
#define PCC_THR_MB8 6
#endif
#ifndef PCC_THR_MB16
#define PCC_THR_MB16 4
#endif
#ifndef PCC_THR_MB32
#define PCC_THR_MB32 4
#endif
// Launch shape.  Log entry = (d2 bits, position), 8 bytes.  Measured on B200, headline workload (profiles/r2/): the kernel is
// co-limited by issue slots (46-51 %) and the L1 data pipe (54-68 % of its wavefronts: a 16-byte load whose 32 lanes sit in ~8
// different cells costs ~8 sector wavefronts).  A 4-byte log (position only, d2 recomputed in the selection) doubles the
// resident warps but its recompute loads -- every lane a different address, ~20 sectors per request -- cost more wavefronts
// than the whole walk: 2.69 ms vs 2.50 ms.  Blocks per SM (4..8), log slots (48 / 64) and the shared-memory carve-out
// (44..100 %) moved the 4-byte variant by < 5 % (r2f_carve.log), so the shape below is simply the one without spills.
template <int K> struct ThrCfg {
    static constexpr int B = K <= 16 ? 16 : 32;                                          // width of the sorting network
    static constexpr int slots = K <= 8 ? PCC_THR_SLOTS8 : (K <= 16 ? PCC_THR_SLOTS16 : PCC_THR_SLOTS32);   // the last 4 are sacrificial
    static constexpr int slot_bits = slots <= 64 ? 6 : 7;
    static constexpr int threads = K <= 16 ? 128 : 64;
    static constexpr int min_blocks = K <= 8 ? PCC_THR_MB8 : (K <= 16 ? PCC_THR_MB16 : PCC_THR_MB32);
    static constexpr size_t smem = (size_t)slots * threads * sizeof(uint2);
    static_assert(slots <= (1 << slot_bits) && slots % 4 == 0, "log slots must fit the key's slot field");
};
constexpr int kCalibBuckets = 64, kCalibBins = 64;
constexpr float kCalibBinsPerOctave = 8.f;

// for (I = 0; I < N; ++I) f(integral_constant<I>) with the unrolling guaranteed (a "#pragma unroll" loop over a register array
// that also holds loads was re-rolled by the compiler, which put the array in local memory)
template <class F, int... I> __device__ __forceinline__ void static_for_impl(F &&f, std::integer_sequence<int, I...>) { (f(std::integral_constant<int, I>{}), ...); }
template <int N, class F> __device__ __forceinline__ void static_for(F &&f) { static_for_impl(f, std::make_integer_sequence<int, N>{}); }
__device__ __forceinline__ void ce_u32(uint32_t &a, uint32_t &b) { const uint32_t lo = min(a, b), hi = max(a, b); a = lo; b = hi; }
// Batcher's odd-even merge sort as a compile-time network (63 compare-exchanges for 16 keys, 191 for 32).  The pair list is
// built by a constexpr constructor and applied through an index pack, so every array index is a literal and the keys stay in registers.
template <int N> struct OemNet {
    int a[N * 8], b[N * 8], n;
    constexpr OemNet() : a(), b(), n(0) {
        for (int p = 1; p < N; p <<= 1)
            for (int k = p; k >= 1; k >>= 1)
                for (int j = k % p; j <= N - 1 - k; j += 2 * k)
                    for (int i = 0; i < k; ++i)
                        if (i + j + k < N && (i + j) / (2 * p) == (i + j + k) / (2 * p)) { a[n] = i + j; b[n] = i + j + k; ++n; }
    }
};
template <int N, int... I>
__device__ __forceinline__ void oem_apply_u32(uint32_t (&v)[N], std::integer_sequence<int, I...>) {
    constexpr OemNet<N> net{};
    (ce_u32(v[net.a[I]], v[net.b[I]]), ...);
}
template <int N>
__device__ __forceinline__ void oem_sort_u32(uint32_t (&v)[N]) {
    constexpr OemNet<N> net{};
    oem_apply_u32<N>(v, std::make_integer_sequence<int, net.n>{});
}
// half-cleaner stages that sort a bitonic sequence of N keys
template <int N> struct BitonicMergeNet {
    int a[N * 8], b[N * 8], n;
    constexpr BitonicMergeNet() : a(), b(), n(0) {
        for (int j = N >> 1; j > 0; j >>= 1)
            for (int i = 0; i < N; ++i)
                if ((i ^ j) > i) { a[n] = i; b[n] = i ^ j; ++n; }
    }
};
template <int N, int... I>
__device__ __forceinline__ void bitonic_merge_apply_u32(uint32_t (&v)[N], std::integer_sequence<int, I...>) {
    constexpr BitonicMergeNet<N> net{};
    (ce_u32(v[net.a[I]], v[net.b[I]]), ...);
}
// best (ascending) and blk (ascending) -> best = the N smallest of both, ascending; dropmin = smallest key that fell out
template <int N>
__device__ __forceinline__ void merge_prune_u32(uint32_t (&best)[N], const uint32_t (&blk)[N], uint32_t &dropmin) {
    static_for<N>([&](auto I) {
        constexpr int i = decltype(I)::value;
        const uint32_t a = best[i], b = blk[N - 1 - i];
        best[i] = min(a, b);                              // bitonic sequence holding the N smallest
        dropmin = min(dropmin, max(a, b));
    });
    constexpr BitonicMergeNet<N> net{};
    bitonic_merge_apply_u32<N>(best, std::make_integer_sequence<int, net.n>{});
}

// ---- calibration: exact search over the block for a strided sample of the batch -> histogram of (d2_k - d2_1) * M / k ----
template <int K>
__global__ void __launch_bounds__(128) knn_calib_kernel(Grid g, QueryView v, int k, int64_t stride, unsigned *__restrict__ hist) {
    const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * stride;
    float x, y, z; int64_t row; bool empty;
    if (!load_query(g, v, t, x, y, z, row, empty)) return;
    const QueryCell c = locate(g, x, y, z);
    RegDist<K> list; list.init();
    uint32_t M = 0;
    scan_shell(g, c, -1, 1, [&](uint32_t, float4 p) { list.offer(dist2(x, y, z, p.x, p.y, p.z)); ++M; });
    const float kth = (k == K) ? list.d[K - 1] : list.at(k - 1);
    if (kth == CUDART_INF_F) return;
    const float ratio = (kth - list.d[0]) * g.inv_cell * g.inv_cell * (float)M / (float)k;
    const int bin = min(max((int)floorf(log2f(fmaxf(ratio, 1e-6f)) * kCalibBinsPerOctave) + kCalibBins / 2, 0), kCalibBins - 1);
    atomicAdd(hist + min(M >> 2, (uint32_t)kCalibBuckets - 1u) * kCalibBins + bin, 1u);
}
// one thread per M-bucket: upper edge of the bin that holds the `quant` quantile (a sparse bucket takes the quantile of all samples)
__global__ void knn_calib_finish_kernel(const unsigned *__restrict__ hist, float *__restrict__ ratio, float quant) {
    __shared__ unsigned all[kCalibBins];
    __shared__ float all_r;
    const int b = threadIdx.x;
    unsigned s = 0;
    for (int i = 0; i < kCalibBuckets; ++i) s += hist[i * kCalibBins + b];
    all[b] = s;
    __syncthreads();
    if (b == 0) {
        unsigned tot = 0; for (int j = 0; j < kCalibBins; ++j) tot += all[j];
        float r = 8.f;                                    // no sample at all: a loose default (the exact path catches the rest)
        if (tot) { unsigned cum = 0; int j = 0; for (; j < kCalibBins; ++j) { cum += all[j]; if ((float)cum >= quant * (float)tot) break; } r = exp2f((float)(min(j, kCalibBins - 1) + 1 - kCalibBins / 2) / kCalibBinsPerOctave); }
        all_r = r;
    }
    __syncthreads();
    unsigned tot = 0; for (int j = 0; j < kCalibBins; ++j) tot += hist[b * kCalibBins + j];
    float r = all_r;
    if (tot >= 64) { unsigned cum = 0; int j = 0; for (; j < kCalibBins; ++j) { cum += hist[b * kCalibBins + j]; if ((float)cum >= quant * (float)tot) break; } r = exp2f((float)(min(j, kCalibBins - 1) + 1 - kCalibBins / 2) / kCalibBinsPerOctave); }
    ratio[b] = r;
}

// One contiguous run of the walk, four candidates per step, branch-free: every candidate's (d2, position) is STORED at the write
// address and the address only advances when the candidate passed (a later candidate overwrites a rejected one).  tau is
// refreshed once per step (a stale tau within a step only logs a little more).  The write address is clamped once per step to
// `cap`, which leaves four sacrificial slots behind it: an address AT cap after the walk means the log overflowed (first pass);
// the retry pass (COMPRESS) compresses the log instead of clamping, so nothing is ever lost there.
__device__ __forceinline__ void sts_v2(uint32_t addr, uint32_t a, uint32_t b) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory"); }
// wa += STRIDE when d <= tau (one FSETP + one predicated IADD; the C form compiled to an add plus a predicated copy)
template <uint32_t STRIDE> __device__ __forceinline__ void advance_if_le(uint32_t &wa, float d, float tau) {
    asm("{\n\t.reg .pred p;\n\tsetp.le.f32 p, %1, %2;\n\t@p add.u32 %0, %0, %3;\n\t}" : "+r"(wa) : "f"(d), "f"(tau), "n"(STRIDE));
}
template <uint32_t STRIDE> __device__ __forceinline__ void advance_if_le_and(uint32_t &wa, float d, float tau, bool ok) {
    asm("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %4, 0;\n\tsetp.le.and.f32 p, %1, %2, q;\n\t@p add.u32 %0, %0, %3;\n\t}" : "+r"(wa) : "f"(d), "f"(tau), "n"(STRIDE), "r"((uint32_t)ok));
}
// The walk is bound by the latency of its point loads, not by issue (ncu, 8-byte-log version: 46 % issue slots, 3.5 of 8 stall
// cycles per issue on the load scoreboard, L2 hit rate 49 %: the first warp to touch a cell's points pays DRAM latency, and a step of
// 4 x 32 lane loads touches ~10 lines).  Prefetches cost no registers and no scoreboard: all nine rows go to L2 as soon as their
// bounds are known, and row i+1 goes to L1 while row i is walked.
#ifndef PCC_THR_PREFETCH
#define PCC_THR_PREFETCH 3
#endif
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// up to LINES 128-byte lines of the run [j, e) (8 points per line; a 3-cell run holds ~25 points)
template <int LINES, bool L1>
__device__ __forceinline__ void prefetch_run(const float4 *pts, uint32_t j, uint32_t e) {
    if (j >= e) return;
    const char *a = (const char *)(pts + j), *last = (const char *)(pts + (e - 1));
#pragma unroll
    for (int i = 0; i < LINES; ++i) {
        const char *q = a + 128 * i;
        if (q <= last || i == 0) { if (L1) prefetch_l1(q); else prefetch_l2(q); }
    }
    if (L1) prefetch_l1(last); else prefetch_l2(last);
}
#ifndef PCC_THR_PIPE
#define PCC_THR_PIPE 0
#endif
struct Quad { float4 p0, p1, p2, p3; };
// GEN: `pts` may point into the warp's shared-memory tile (staged kernel) -- plain generic loads; otherwise read-only global loads
template <bool GEN> __device__ __forceinline__ float4 load_pt(const float4 *pts, uint32_t j) { if (GEN) return pts[j]; return __ldg(pts + j); }
template <bool GEN> __device__ __forceinline__ Quad load_quad(const float4 *pts, uint32_t j) { Quad q; q.p0 = load_pt<GEN>(pts, j); q.p1 = load_pt<GEN>(pts, j + 1); q.p2 = load_pt<GEN>(pts, j + 2); q.p3 = load_pt<GEN>(pts, j + 3); return q; }
template <uint32_t STRIDE, bool COMPRESS, class F>
__device__ __forceinline__ void thr_quad(const Quad &q, const uint32_t j, const float x, const float y, const float z, float &T, float &dmin, float &tau, uint32_t &wa, const uint32_t cap, F &&on_full) {
    const float d0 = dist2(x, y, z, q.p0.x, q.p0.y, q.p0.z), d1 = dist2(x, y, z, q.p1.x, q.p1.y, q.p1.z), d2 = dist2(x, y, z, q.p2.x, q.p2.y, q.p2.z), d3 = dist2(x, y, z, q.p3.x, q.p3.y, q.p3.z);
    sts_v2(wa, __float_as_uint(d0), j); advance_if_le<STRIDE>(wa, d0, tau);
    sts_v2(wa, __float_as_uint(d1), j + 1); advance_if_le<STRIDE>(wa, d1, tau);
    sts_v2(wa, __float_as_uint(d2), j + 2); advance_if_le<STRIDE>(wa, d2, tau);
    sts_v2(wa, __float_as_uint(d3), j + 3); advance_if_le<STRIDE>(wa, d3, tau);
    dmin = fminf(fminf(dmin, fminf(d0, d1)), fminf(d2, d3)); tau = dmin + T;
    if (!COMPRESS) wa = min(wa, cap);
    else if (PCC_THR_JOINT ? __any_sync(__activemask(), wa >= cap) : (wa >= cap)) on_full();      // not clamped first: the entries in the sacrificial slots are real
}
// PCC_THR_PIPE = 1 software-pipelines the walk (the four loads of step i+1 are issued before step i is processed, two register
// sets, loop unrolled by two).  Measured on B200: no gain (3.62 vs 3.51 ms for the stage, and the retry kernel doubles) -- the
// exposed time is not the latency of one step's loads (profiles/r2/blockkernel_load_experiments.txt), so it is off by default.
template <uint32_t STRIDE, bool COMPRESS, bool GEN, class F>
__device__ __forceinline__ void thr_walk_run(const float4 *pts, uint32_t j, const uint32_t e, const float x, const float y, const float z, float &T,
                                             float &dmin, float &tau, uint32_t &wa, const uint32_t cap, F &&on_full) {
#if PCC_THR_PIPE
    bool have_a = j + 4 <= e;
    Quad a, b;
    if (have_a) a = load_quad<GEN>(pts, j);
    while (have_a) {
        const bool have_b = j + 8 <= e;
        if (have_b) b = load_quad<GEN>(pts, j + 4);
        thr_quad<STRIDE, COMPRESS>(a, j, x, y, z, T, dmin, tau, wa, cap, on_full);
        j += 4;
        if (!have_b) break;
        have_a = j + 8 <= e;
        if (have_a) a = load_quad<GEN>(pts, j + 4);
        thr_quad<STRIDE, COMPRESS>(b, j, x, y, z, T, dmin, tau, wa, cap, on_full);
        j += 4;
    }
#else
#ifndef PCC_THR_EXP
#define PCC_THR_EXP 0
#endif
    // PCC_THR_EXP (measurement builds only, results are wrong): 1 = every step re-reads the run's first four points (same sectors per
    // lane: L1 hits, same wavefront count); 2 = every lane reads point 0 (one wavefront per load: pure issue cost)
    const uint32_t j0 = j;
    for (; j + 4 <= e; j += 4) {
        const Quad a = load_quad<GEN>(pts, PCC_THR_EXP == 1 ? j0 : (PCC_THR_EXP == 2 ? 0u : j));
        thr_quad<STRIDE, COMPRESS>(a, j, x, y, z, T, dmin, tau, wa, cap, on_full);
    }
#endif
    if (j < e) {                                          // 1..3 left: the loads are clamped to the run, the extra lanes never advance
        const float4 p0 = load_pt<GEN>(pts, j), p1 = load_pt<GEN>(pts, min(j + 1, e - 1)), p2 = load_pt<GEN>(pts, min(j + 2, e - 1));
        const float d0 = dist2(x, y, z, p0.x, p0.y, p0.z), d1 = dist2(x, y, z, p1.x, p1.y, p1.z), d2 = dist2(x, y, z, p2.x, p2.y, p2.z);
        sts_v2(wa, __float_as_uint(d0), j); advance_if_le<STRIDE>(wa, d0, tau);
        sts_v2(wa, __float_as_uint(d1), j + 1); advance_if_le_and<STRIDE>(wa, d1, tau, j + 1 < e);
        sts_v2(wa, __float_as_uint(d2), j + 2); advance_if_le_and<STRIDE>(wa, d2, tau, j + 2 < e);
        dmin = fminf(fminf(dmin, d0), fminf(d1, d2)); tau = dmin + T;
        if (!COMPRESS) wa = min(wa, cap);
        else if (PCC_THR_JOINT ? __any_sync(__activemask(), wa >= cap) : (wa >= cap)) on_full();      // not clamped first: the entries in the sacrificial slots are real
    }
}

// ---- the block kernel ----
// FULLK: k == K (every per-element "j < k" test folds away).
// RETRY: the second pass, over the compacted list of queries the first pass could not settle (~1 %):
//   * table entry too tight (the k-th logged d2 is beyond the final tau): walked again with 4 T;
//   * log overflow (a block denser than its table entry expects): walked again, and whenever the log fills up it is COMPRESSED
//     in place -- the k-th smallest logged d2 is an upper bound of the k-th distance (every entry is a real candidate): tau
//     drops to it, entries beyond tau are discarded and the walk goes on;
//   * a tie the 32-bit keys cannot order: the k smallest are picked from the log with exact 64-bit (d2, index) keys.
// Only what is still unsettled after that (more ties than the log holds) runs the exact per-thread search with ring expansion.
constexpr float kRetryScale = 4.f;
// STAGED (knn_thr_staged_kernel): the candidate points are not loaded lane by lane from global memory.  The lanes of a warp that share
// the (y, z) row of its first live query -- queries are in cell order, so that is nearly all of them -- span a few adjacent cells; the
// union of their nine stencil rows is nine contiguous runs of the sorted array (~250 points), which lanes 0..8 pull into the warp's
// shared-memory tile with cp.async.bulk (TMA bulk copy, completion on an mbarrier).  Every lane then walks its own window of the
// tile; lanes outside the leader's row, or a union that does not fit the tile, keep the global path.  Why: the per-lane loads are
// what the measurements leave as removable cost (0.8 of 2.5 ms in misses, ~8 L1 sector wavefronts per load).
#ifndef PCC_THR_TILE
#define PCC_THR_TILE 384
#endif
namespace thr_tma {
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
}  // namespace thr_tma
#ifndef PCC_THR_MAIN_COMPRESS
#define PCC_THR_MAIN_COMPRESS 0
#endif
// COMPRESS: a full log is compressed in place during the walk (always in the retry pass; in the first pass only in builds with
// PCC_THR_MAIN_COMPRESS=1, which trade a smaller log -- more resident warps -- for compressions inside the main kernel)
template <int K, bool FULLK, bool RETRY, bool STAGED = false, bool COMPRESS = RETRY>
__device__ __forceinline__ void knn_thr_body(const Grid &g, const QueryView &v, const int64_t t, const int k_rt, int32_t *__restrict__ out_idx, float *__restrict__ out_d2, int vec4,
                                             const FixList &fix, const float *__restrict__ ratio, uint2 *__restrict__ slog, const bool loose = false,
                                             float4 *tile = nullptr, unsigned long long *bar = nullptr) {
    constexpr int TH = ThrCfg<K>::threads, SLOTS = ThrCfg<K>::slots, B = ThrCfg<K>::B;
    constexpr uint32_t SMASK = (1u << ThrCfg<K>::slot_bits) - 1u;
    const int k = FULLK ? K : k_rt;
    float x = 0.f, y = 0.f, z = 0.f; int64_t row = 0; bool empty = false;
#if PCC_THR_QAHEAD
    // A block starts with two dependent cache misses (order[t], then the query it names) before it has anything to do: 8 % of the kernel's
    // stall samples.  Blocks run in launch order, so this block asks L2 for what the blocks ~one and ~two GPU-fulls later will read first.
    uint32_t qi_ahead = 0xFFFFFFFFu;
    if (!RETRY && !STAGED && v.order) {
        constexpr int64_t AHEAD = (int64_t)PCC_THR_QAHEAD * ThrCfg<K>::threads;
        if ((threadIdx.x & 31) == 0 && t + 2 * AHEAD < v.nq) prefetch_l2(v.order + t + 2 * AHEAD);
        if (t + AHEAD < v.nq) qi_ahead = __ldg(v.order + t + AHEAD);
    }
#endif
    const bool live = load_query(g, v, t, x, y, z, row, empty);
    auto leave_dead = [&]() {                             // tail thread or non-finite query
        if (t < v.nq) fix.ring_flag[t] = 0;
        if (empty) { nkey_t e[K];
#pragma unroll
            for (int j = 0; j < K; ++j) e[j] = PCC_EMPTY_KEY;
            write_row<K>(e, k, out_idx + row * k, out_d2 + row * k, vec4); }
    };
    if (!STAGED && !live) { leave_dead(); return; }       // the staged kernel keeps every lane until the tile is in (warp-wide shuffles)
    QueryCell c; c.cx = c.cy = c.cz = 0; c.ux = c.uy = c.uz = 0.f;
    if (live) c = locate(g, x, y, z);
    // the 9 run bounds of the block: its population M, and a first bound for dmin (the middle point of the query's own row)
    uint32_t M = 0, s0 = 0, e0 = 0, rowmask = 0;         // rowmask bit r: stencil row r holds a point (an empty row costs neither its bounds lookup nor its clipping arithmetic)
    if (live) {
        const int xa = max(c.cx - 1, 0), xb = min(c.cx + 1, g.nx - 1);
        uint32_t rs[9], re[9];
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            const int zz = c.cz + centre_out(r / 3), yy = c.cy + centre_out(r % 3);
            rs[r] = re[r] = 0;
            if (zz >= 0 && zz < g.nz && yy >= 0 && yy < g.ny) {
                const size_t rowb = ((size_t)zz * g.ny + yy) * g.nx;
                rs[r] = cell_begin(g, rowb + xa); re[r] = cell_begin(g, rowb + xb + 1);
            }
        }
#pragma unroll
        for (int r = 0; r < 9; ++r) { M += re[r] - rs[r]; rowmask |= (re[r] > rs[r] ? 1u : 0u) << r; if (!STAGED && (PCC_THR_PREFETCH & 1)) prefetch_run<3, false>(g.pts, rs[r], re[r]); }
        s0 = rs[0]; e0 = re[0];
    }
#if PCC_THR_QAHEAD
    if (!RETRY && !STAGED && qi_ahead != 0xFFFFFFFFu) prefetch_l2(v.q + qi_ahead);
#endif
    auto leave_wide = [&]() {                             // fewer than k points in the block: a wide query (no walk)
        fix.ring_flag[t] = 0;
        if (fix.stats) { atomicAdd(fix.stats + 0, 1ull); atomicAdd(fix.stats + 4, 1ull); }
        push_list(fix.wide_list, fix.wide_count, (uint32_t)t);
    };
    const bool walker = live && M >= (uint32_t)k;
    if (!STAGED && !walker) { leave_wide(); return; }
    // ---- staged kernel: pull the union of the leader row's nine runs into the warp's tile ----
    bool lane_staged = false;
    uint32_t my_start = 0, my_off = 0;                    // lane r < 9: first sorted position / tile offset of stencil row r of the union
    if (STAGED) {
        const unsigned full = 0xffffffffu;
        const int lane = threadIdx.x & 31;
        const unsigned wmask = __ballot_sync(full, walker);
        if (wmask) {
            const int leader = __ffs(wmask) - 1;
            const int lcy = __shfl_sync(full, c.cy, leader), lcz = __shfl_sync(full, c.cz, leader);
            bool inrow = walker && c.cy == lcy && c.cz == lcz;
            const int xmin = __reduce_min_sync(full, inrow ? c.cx : 0x7fffffff);
            inrow = inrow && c.cx <= xmin + 9;            // at most 12 cells per staged row
            const int xmax = __reduce_max_sync(full, inrow ? c.cx : xmin);
            uint32_t len = 0;
            if (lane < 9) {
                const int zz = lcz + centre_out(lane / 3), yy = lcy + centre_out(lane % 3);
                if (zz >= 0 && zz < g.nz && yy >= 0 && yy < g.ny) {
                    const uint32_t *rowp = g.cell_start + ((size_t)zz * g.ny + yy) * g.nx;
                    my_start = __ldg(rowp + max(xmin - 1, 0)); len = __ldg(rowp + min(xmax + 1, g.nx - 1) + 1) - my_start;
                }
            }
            uint32_t inc = len;
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) { const uint32_t w = __shfl_up_sync(full, inc, o); if (lane >= o) inc += w; }
            my_off = inc - len;
            const uint32_t total = __shfl_sync(full, inc, 8);
            if (total <= (uint32_t)PCC_THR_TILE && total > 0) {
                if (lane == 0) thr_tma::mbar_expect_tx(bar, total * 16u);
                __syncwarp();
                if (lane < 9 && len) thr_tma::bulk_g2s(tile + my_off, g.pts + my_start, len * 16u, bar);
                lane_staged = inrow;
                thr_tma::mbar_wait(bar, 0);
            }
        }
    }
    float T = __ldg(ratio + min(M >> 2, (uint32_t)kCalibBuckets - 1u)) * (float)k / (float)max(M, 1u) * (g.cell * g.cell) * ((RETRY && loose) ? kRetryScale : 1.f);
    float dmin = CUDART_INF_F;
    if (walker && e0 > s0) { const float4 p = __ldg(g.pts + ((s0 + e0) >> 1)); dmin = dist2(x, y, z, p.x, p.y, p.z); }
    float tau = dmin + T;
    constexpr uint32_t STRIDE = (uint32_t)TH * (uint32_t)sizeof(uint2);
    constexpr int LOGCAP = SLOTS - 4;
    const uint32_t wa0 = (uint32_t)__cvta_generic_to_shared(slog), cap = wa0 + (uint32_t)LOGCAP * STRIDE;
    uint32_t wa = wa0;
    // select: the B smallest of the first n logged keys, B log slots at a time.  key = (d2 bits << 1, low bits = slot)
    uint32_t best[B], dropmin;
    auto select_log = [&](const int n) {
        auto load_keys = [&](uint32_t (&key)[B], const int b0) {
            static_for<B>([&](auto I) {
                constexpr int i = decltype(I)::value;
                key[i] = b0 + i < n ? (((slog[(b0 + i) * TH].x << 1) & ~SMASK) | (uint32_t)(b0 + i)) : 0xFFFFFFFFu;
            });
        };
        dropmin = 0xFFFFFFFFu;
        load_keys(best, 0);
        oem_sort_u32<B>(best);
        for (int b0 = B; b0 < n; b0 += B) {
            uint32_t blk[B];
            load_keys(blk, b0);
            oem_sort_u32<B>(blk);
            merge_prune_u32<B>(best, blk, dropmin);
        }
    };
    auto kth_key = [&]() {
        uint32_t kth = best[K - 1];
        if (!FULLK) {                                     // best is ascending: the k-th key is the largest of the first k (written so that it cannot become a runtime index)
            kth = 0u;
            static_for<K - 1>([&](auto I) { constexpr int i = decltype(I)::value; kth = max(kth, i < k ? best[i] : 0u); });
        }
        return kth;
    };
    bool give_up = false;
    auto compress_if = [&](const int min_n) {             // RETRY only: compress the log in place (see above) if it holds >= min_n entries
        const int nf = (int)((wa - wa0) / STRIDE);        // LOGCAP .. LOGCAP + 3 entries when the log is full
        if (nf < min_n) return;
        select_log(nf);
        const float tau_new = __uint_as_float(slog[(kth_key() & SMASK) * TH].x) * (1.f + 1.f / 32768.f);     // 2^-15 covers the key truncation
        // The log is complete up to the CURRENT tau only (earlier entries passed looser thresholds), so tau may shrink to tau_new
        // but never grow: constant threshold from here on if the bound is the tighter one (d2 >= 0 keeps dmin at 0).
        if (tau_new < tau) { dmin = 0.f; T = tau_new; tau = tau_new; }
        int m = 0;
        for (int i = 0; i < nf; ++i) { const uint2 e = slog[i * TH]; if (__uint_as_float(e.x) <= tau) { slog[m * TH] = e; ++m; } }
        if (m >= LOGCAP - 4) { give_up = true; m = 0; }   // more candidates inside tau than the log holds (ties)
        wa = wa0 + (uint32_t)m * STRIDE;
    };
    // called when SOME lane of the warp is full (PCC_THR_JOINT): that lane must compress, the others join it if they have a
    // worthwhile number of entries to drop, so that they do not stop the warp again a few steps later
    auto on_full = [&]() { compress_if(wa >= cap ? 0 : K + 13); };
    // the walk: rows centre-out, each clipped to the ball of the current tau (bounds of row i+1 fetched before row i is walked)
    {
        const RowRuns none = {0u, 0u, 0u, 0u};
        int az = 0, ay = 0;
        if (!walker) rowmask = 0;
        RowRuns nxt = (rowmask & 1u) ? row_runs(g, c, -1, 1, to_cell_units(g, tau), 0, 0) : none;
        for (;;) {
            const RowRuns cur = nxt;
            const int r = az * 3 + ay;
            if (++ay == 3) { ay = 0; ++az; }
            const bool more = az < 3;
            nxt = none;
            if (more && ((rowmask >> (az * 3 + ay)) & 1u)) { nxt = row_runs(g, c, -1, 1, to_cell_units(g, tau), az, ay); if (!STAGED && (PCC_THR_PREFETCH & 2)) prefetch_run<3, true>(g.pts, nxt.j1, nxt.e1); }
            if (STAGED) {
                // sorted position j of stencil row r sits at tile[off_r + (j - start_r)]: a lane in the leader's row reads the tile
                // (its own cell lies inside the staged x-range, so its window does too), any other lane reads global memory
                const uint32_t st = __shfl_sync(0xffffffffu, my_start, r), of = __shfl_sync(0xffffffffu, my_off, r);
                const float4 *src = lane_staged ? (const float4 *)tile + of - st : g.pts;
                thr_walk_run<STRIDE, COMPRESS, true>(src, cur.j1, cur.e1, x, y, z, T, dmin, tau, wa, cap, on_full);
            } else thr_walk_run<STRIDE, COMPRESS, false>(g.pts, cur.j1, cur.e1, x, y, z, T, dmin, tau, wa, cap, on_full);
            (void)r;
            if (!more) break;
            if (COMPRESS && PCC_THR_EARLY > 0) {
                // Compress together.  A lane whose log fills inside a row compresses alone while the other 31 wait for its ~800
                // instructions (ncu, retry pass: 6 of 32 lanes active, half of the samples in compression code).  So at a row
                // boundary, once any lane is within PCC_THR_EARLY slots of full, every lane that has entries to drop compresses in
                // the same pass (legal at any time with >= k entries: tau only shrinks).
                const int nl = (int)((wa - wa0) / STRIDE);
                if (__any_sync(__activemask(), nl >= LOGCAP - PCC_THR_EARLY)) compress_if(K + 5);
            }
        }
    }
    if (STAGED && !live) { leave_dead(); return; }
    if (STAGED && !walker) { leave_wide(); return; }
    if (RETRY && give_up) { fix.ring_flag[t] = 0; knn_reg_body<K>(g, v, t, k, out_idx, out_d2, vec4); return; }
    if (!RETRY && COMPRESS && give_up) { fix.ring_flag[t] = 0; push_list(fix.retry_list, fix.retry_count, (uint32_t)t); return; }      // more ties than slots: the retry pass ends in the exact search
    const int n = (int)((wa - wa0) / STRIDE);             // == LOGCAP: the log (may have) overflowed
    select_log(n);
    // the k-th key, the smallest gap between neighbours among the first k + 1 keys
    const uint32_t kth = kth_key();
    uint32_t gap = 0xFFFFFFFFu;
    static_for<K>([&](auto I) {
        constexpr int i = decltype(I)::value;
        const uint32_t nxt = i + 1 < B ? best[i + 1] : dropmin;
        const uint32_t d = nxt - best[i];
        gap = min(gap, (FULLK || i < k) ? d : 0xFFFFFFFFu);
    });
    const bool overflow = n >= LOGCAP;
    const bool tie = kth == 0xFFFFFFFFu || gap <= SMASK;
    // exact d2 of the k-th logged candidate
    float tau_k = CUDART_INF_F;
    if (kth != 0xFFFFFFFFu) tau_k = __uint_as_float(slog[(kth & SMASK) * TH].x);
    const bool under = !(tau_k <= tau);                    // the threshold was too tight: an unlogged candidate may be nearer than the k-th logged one
    if (fix.stats && !RETRY) { atomicAdd(fix.stats + 0, 1ull); atomicAdd(fix.stats + 1, (unsigned long long)n); atomicAdd(fix.stats + 3, overflow ? 1ull : 0ull); atomicAdd(fix.stats + 5, (overflow || tie || under) ? 1ull : 0ull); atomicAdd(fix.stats + 6, (unsigned long long)M); }
    if (!RETRY && (overflow || tie || under)) {           // list it for the retry pass (bit 31: 4 T)
        fix.ring_flag[t] = 0;
        push_list(fix.retry_list, fix.retry_count, (uint32_t)t | ((under && !overflow) ? 0x80000000u : 0u));
        return;
    }
    int32_t *oi = out_idx + row * k; float *od = out_d2 + row * k;
    if (RETRY && (overflow || tie || under)) {
        // exact selection from the log: 64-bit (d2, index) keys, insertion-sorted.  Valid when the log is complete up to its k-th d2.
        RegList<K> list; list.init();
        if (!overflow) for (int i = 0; i < n; ++i) { const uint2 e = slog[i * TH]; list.offer(((nkey_t)e.x << 32) | __float_as_uint(__ldg(&g.pts[e.y].w))); }
        const nkey_t kk = FULLK ? list.key[K - 1] : list.at(k - 1);
        tau_k = key_d2(kk);
        if (overflow || kk == PCC_EMPTY_KEY || !(tau_k <= tau)) { fix.ring_flag[t] = 0; knn_reg_body<K>(g, v, t, k, out_idx, out_d2, vec4); return; }
        const float cov = covered_d2(g, c, 1);
        const bool proved = cov == CUDART_INF_F || tau_k < cov;
        const bool wide = !proved && next_ring(g, 1, tau_k) > kRingMaxR;
        fix.ring_flag[t] = (!proved && !wide) ? 1 : 0;
        if (wide) { push_list(fix.wide_list, fix.wide_count, (uint32_t)t); return; }
        write_row<K>(list.key, k, oi, od, vec4);
        return;
    }
    const float cov = covered_d2(g, c, 1);
    const bool proved = cov == CUDART_INF_F || tau_k < cov;
    const bool wide = !proved && next_ring(g, 1, tau_k) > kRingMaxR;
    if (fix.stats && !RETRY) atomicAdd(fix.stats + 4, proved ? 0ull : 1ull);
    fix.ring_flag[t] = (!proved && !wide) ? 1 : 0;
    if (wide) { push_list(fix.wide_list, fix.wide_count, (uint32_t)t); return; }
    // the row: exact d2 from the log, original index from the point's .w, four at a time
    static_for<K / 4>([&](auto Q) {
        constexpr int q = decltype(Q)::value * 4;
        uint2 le[4]; int32_t id[4];
        static_for<4>([&](auto I) { constexpr int i = decltype(I)::value; le[i] = (FULLK || q + i < k) ? slog[(best[q + i] & SMASK) * TH] : make_uint2(0u, 0u); });     // keys past k may be empty
        static_for<4>([&](auto I) { constexpr int i = decltype(I)::value; id[i] = __float_as_int(__ldg(&g.pts[le[i].y].w)); });
        if (FULLK && vec4) {
            // streaming stores: the 1.28 GB result table is written once and must not evict the index from L2
            __stcs(reinterpret_cast<int4 *>(oi) + (q >> 2), make_int4(id[0], id[1], id[2], id[3]));
            __stcs(reinterpret_cast<float4 *>(od) + (q >> 2), make_float4(__uint_as_float(le[0].x), __uint_as_float(le[1].x), __uint_as_float(le[2].x), __uint_as_float(le[3].x)));
        } else {
            static_for<4>([&](auto I) { constexpr int i = decltype(I)::value; if (q + i < k) { oi[q + i] = id[i]; od[q + i] = __uint_as_float(le[i].x); } });
        }
    });
}
template <int K, bool FULLK>
__global__ void __launch_bounds__(ThrCfg<K>::threads, ThrCfg<K>::min_blocks) knn_thr_kernel(Grid g, QueryView v, int k, int32_t *__restrict__ out_idx, float *__restrict__ out_d2, int vec4, FixList fix, const float *__restrict__ ratio) {
    extern __shared__ uint2 thr_log[];
    knn_thr_body<K, FULLK, false, false, PCC_THR_MAIN_COMPRESS != 0>(g, v, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, k, out_idx, out_d2, vec4, fix, ratio, thr_log + threadIdx.x);
}
#ifndef PCC_THR_STAGED_MB
#define PCC_THR_STAGED_MB 3
#endif
template <int K> struct ThrStagedCfg {
    static constexpr int warps = ThrCfg<K>::threads / 32;
    static constexpr size_t log_bytes = ThrCfg<K>::smem, tile_bytes = (size_t)warps * PCC_THR_TILE * sizeof(float4);
    static constexpr size_t smem = log_bytes + tile_bytes + warps * sizeof(unsigned long long);
};
template <int K, bool FULLK>
__global__ void __launch_bounds__(ThrCfg<K>::threads, PCC_THR_STAGED_MB) knn_thr_staged_kernel(Grid g, QueryView v, int k, int32_t *__restrict__ out_idx, float *__restrict__ out_d2, int vec4, FixList fix, const float *__restrict__ ratio) {
    extern __shared__ __align__(128) unsigned char thr_smem[];
    uint2 *log = reinterpret_cast<uint2 *>(thr_smem);
    float4 *tiles = reinterpret_cast<float4 *>(thr_smem + ThrStagedCfg<K>::log_bytes);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(thr_smem + ThrStagedCfg<K>::log_bytes + ThrStagedCfg<K>::tile_bytes);
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) thr_tma::mbar_init(bars + warp, 1);
    __syncwarp();
    knn_thr_body<K, FULLK, false, true>(g, v, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, k, out_idx, out_d2, vec4, fix, ratio, log + threadIdx.x, false,
                                        tiles + (size_t)warp * PCC_THR_TILE, bars + warp);
}
template <int K, bool FULLK>
__global__ void __launch_bounds__(ThrCfg<K>::threads, 2) knn_thr_retry_kernel(Grid g, QueryView v, int k, int32_t *__restrict__ out_idx, float *__restrict__ out_d2, int vec4, FixList fix, const float *__restrict__ ratio) {
    extern __shared__ uint2 thr_log[];
    const unsigned n = *fix.retry_count;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t e = fix.retry_list[i];
        knn_thr_body<K, FULLK, true>(g, v, (int64_t)(e & 0x7FFFFFFFu), k, out_idx, out_d2, vec4, fix, ratio, thr_log + threadIdx.x, (e >> 31) != 0u);
    }
}

}  // namespace pcc
