// pcc_fastknn.cuh -- sorting networks of the two-phase kNN selection used for k <= 32 (see pcc_knn.cu).
//
// Measured on the first (insertion-sorted 64-bit key) kernel: 68 % of all issued instructions were the
// ISETP/SEL insertion network on the half-rate integer pipe at 45 % SIMD efficiency, and nearly every warp paid
// for an unclipped ring-2 pass (profiles/r1/knn16_v1_*).  The two-phase kernel keeps only fp32 distances while it
// searches (min/max insertion, 2 ops per slot), clips every ring >= 2 pass to the ball of the current k-th
// distance, then re-walks the cells inside the final ball once to pick up the (d2, idx) pairs and sorts at most K
// packed keys with the network below.
#pragma once
#include "pcc_device.cuh"

namespace pcc {

template <int N>
__device__ __forceinline__ void bitonic_sort_f32(float (&v)[N]) {
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const int l = i ^ j;
                if (l > i) {
                    const float a = v[i], b = v[l];
                    const float lo = fminf(a, b), hi = fmaxf(a, b);
                    const bool up = (i & k) == 0;
                    v[i] = up ? lo : hi;
                    v[l] = up ? hi : lo;
                }
            }
        }
    }
}
// a: running K smallest (ascending).  b: K new values (ascending).  a <- K smallest of the union, ascending.
template <int K>
__device__ __forceinline__ void merge_keep_low(float (&a)[K], const float (&b)[K]) {
#pragma unroll
    for (int i = 0; i < K; ++i) a[i] = fminf(a[i], b[K - 1 - i]);      // bitonic sequence holding the K smallest
#pragma unroll
    for (int j = K >> 1; j > 0; j >>= 1) {
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const int l = i ^ j;
            if (l > i) { const float x = a[i], y = a[l]; a[i] = fminf(x, y); a[l] = fmaxf(x, y); }
        }
    }
}
template <int N>
__device__ __forceinline__ void bitonic_sort_key(nkey_t (&v)[N]) {
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const int l = i ^ j;
                if (l > i) {
                    const nkey_t a = v[i], b = v[l];
                    const bool swap = ((i & k) == 0) ? (b < a) : (a < b);
                    v[i] = swap ? b : a;
                    v[l] = swap ? a : b;
                }
            }
        }
    }
}

}  // namespace pcc
