// yb_sort.cuh — single-CTA stable LSD radix sort used by both NMS algorithms.
// One CTA (kSortThreads threads) sorts the candidates of one image.  Each warp owns a contiguous
// range of the input and a private 256-bin histogram row in shared memory, so ranks are exact
// without atomics and stability holds by construction (warp ranges in order, lanes in order).
#pragma once
#include "yb_common.cuh"

namespace yb {

typedef unsigned long long u64;
typedef uint32_t u32;

constexpr int kSortThreads = 1024;
constexpr int kSortPrefetch = 4;  // loads in flight per lane

// descending-score key: ascending radix order of this key == torch.sort(descending=True);
// NaN sorts as the largest score, -0.0 == +0.0.
__device__ __forceinline__ u32 desc_key(float s) {
    s = s + 0.0f;
    u32 u = __float_as_uint(s);
    u32 asc = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    if (s != s) asc = 0xffffffffu;
    return ~asc;
}

// Lanes of the warp holding the same 8-bit digit (valid lanes only).  Eight ballots instead of
// match.any: MATCH issues at roughly one warp instruction per ~40 cycles per SM on sm_100 (the
// round-1 profile of the sort kernel showed it bounding both phases), VOTE at the normal rate.
__device__ __forceinline__ unsigned digit_peers(u32 d, bool valid) {
    unsigned m = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int bit = 0; bit < 8; ++bit) {
        const bool p = (d >> bit) & 1u;
        const unsigned v = __ballot_sync(0xffffffffu, p);
        m &= p ? v : ~v;
    }
    return m;
}

// One pass on the 8-bit digit at `shift`.  Returns false (and writes nothing) when every key has
// the same digit — the caller then keeps using the input buffers.
__device__ inline bool radix_pass(const u32* __restrict__ kin, const u32* __restrict__ vin,
                                  u32* __restrict__ kout, u32* __restrict__ vout, int M, int shift,
                                  u32* hist /*[32][256]*/, u32* tot /*[256]*/, int* flag /*smem*/) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (((M + 31) / 32) + 31) & ~31;
    const int beg = warp * per;
    const int end = min(beg + per, M);
    for (int i = tid; i < 32 * 256; i += kSortThreads) hist[i] = 0;
    if (tid == 0) *flag = 0;
    __syncthreads();
    u32* h = hist + warp * 256;
    for (int i0 = beg; i0 < end; i0 += 32 * kSortPrefetch) {
        u32 key[kSortPrefetch];
#pragma unroll
        for (int u = 0; u < kSortPrefetch; ++u) {
            const int i = i0 + 32 * u + lane;
            key[u] = i < end ? kin[i] : 0u;
        }
#pragma unroll
        for (int u = 0; u < kSortPrefetch; ++u) {
            const bool valid = (i0 + 32 * u + lane) < end;
            const u32 d = (key[u] >> shift) & 255u;
            const unsigned peers = digit_peers(d, valid);
            if (valid && lane == (__ffs(peers) - 1)) h[d] += __popc(peers);
            __syncwarp();
        }
    }
    __syncthreads();
    if (tid < 256) {
        u32 run = 0;
        for (int w = 0; w < 32; ++w) {
            const u32 t = hist[w * 256 + tid];
            hist[w * 256 + tid] = run;
            run += t;
        }
        tot[tid] = run;
        if (run == (u32)M) *flag = 1;
    }
    __syncthreads();
    if (*flag) return false;
    if (warp == 0) {
        u32 v[8], s = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { v[k] = tot[lane * 8 + k]; s += v[k]; }
        u32 inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        u32 run = inc - s;
#pragma unroll
        for (int k = 0; k < 8; ++k) { tot[lane * 8 + k] = run; run += v[k]; }
    }
    __syncthreads();
    for (int i0 = beg; i0 < end; i0 += 32 * kSortPrefetch) {
        u32 key[kSortPrefetch], val[kSortPrefetch];
#pragma unroll
        for (int u = 0; u < kSortPrefetch; ++u) {
            const int i = i0 + 32 * u + lane;
            key[u] = i < end ? kin[i] : 0u;
            val[u] = i < end ? vin[i] : 0u;
        }
#pragma unroll
        for (int u = 0; u < kSortPrefetch; ++u) {
            const bool valid = (i0 + 32 * u + lane) < end;
            const u32 d = (key[u] >> shift) & 255u;
            const unsigned peers = digit_peers(d, valid);
            const int rank = __popc(peers & ((1u << lane) - 1u));
            u32 base = 0;
            if (valid) {
                base = h[d];
                const u32 dst = base + tot[d] + rank;
                kout[dst] = key[u];
                vout[dst] = val[u];
            }
            __syncwarp();
            if (valid && lane == (__ffs(peers) - 1)) h[d] = base + __popc(peers);
            __syncwarp();
        }
    }
    __syncthreads();
    return true;
}

// Sorts (key,val) pairs on the digits [lo_shift, hi_shift) step 8.  Buffers ping-pong between
// (*ka,*va) and (*kb,*vb); on return (*ka,*va) hold the sorted data.
__device__ inline void radix_sort(u32*& ka, u32*& va, u32*& kb, u32*& vb, int M, int lo_shift, int hi_shift,
                                  u32* hist, u32* tot, int* flag) {
    for (int s = lo_shift; s < hi_shift; s += 8) {
        if (radix_pass(ka, va, kb, vb, M, s, hist, tot, flag)) {
            u32* t = ka; ka = kb; kb = t;
            t = va; va = vb; vb = t;
        }
    }
}

}  // namespace yb
