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

// ------------------------------------------------------------------------------------------------
// Shared-memory blocked LSD radix sort (graph NMS, images of up to kS16MaxM candidates).
//
// The warp-ballot sort above spends ~250 instructions per key and pass; a first blocked variant
// with per-thread counters needed only ~20 but scattered every key to global memory and ran into
// the L2's sector-transaction rate (ncu, round 1: 23 sectors per store request).  Here the data
// never leave shared memory between passes: a 32-bit key is sorted as two 16-bit halves (LSD order:
// low half first), each element is ONE u32 (half key | index << 16), so two ping-pong buffers of
// 4*M bytes plus 16 x 512 u16 counters fit the SM's 227 KB for M <= 26,880.
// Thread t owns the K consecutive elements [t*K, t*K+K); counter (bin, t) lives at cnt[bin*T + t];
// an exclusive scan in bin-major, thread-minor order gives stable destinations; each thread then
// walks its elements again and scatters within shared memory.  4-bit digits, 8 passes, passes whose
// digit is uniform are skipped.
// ------------------------------------------------------------------------------------------------
constexpr int kS16Threads = 512;  // (768 threads measured 6 % slower: the scan and barriers grow with the thread count)
constexpr int kS16Bins = 16;
constexpr int kS16MaxM = 26880;   // 2 x 4 B x M + 16 KB of counters within 227 KB
typedef unsigned short u16;

__host__ __device__ inline size_t s16_smem_bytes(int cap) {
    const size_t n = ((size_t)cap + 3) / 4 * 4;
    return 2 * n * sizeof(u32) + (size_t)kS16Bins * kS16Threads * sizeof(u16) + 64 * sizeof(u32);
}

__device__ inline bool s16_pass(const u32* __restrict__ in, u32* __restrict__ out, int M, int shift, u16* cnt,
                                u32* wsum /*[32] warp totals + 2 flags*/, int pass) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = (((M + kS16Threads - 1) / kS16Threads) + 3) & ~3;
    const int beg = min(tid * K, M), end = min(beg + K, M);
    u32* flag = wsum + 32 + (pass & 1);  // alternating slots: no reset/read race between passes
#pragma unroll
    for (int b = 0; b < kS16Bins; ++b) cnt[b * kS16Threads + tid] = 0;
    if (tid == 0) *flag = 0u;
    u16* mine = cnt + tid;
    int i = beg;
    for (; i + 4 <= end; i += 4) {
        const uint4 x = *reinterpret_cast<const uint4*>(in + i);
        mine[((x.x >> shift) & 15u) * kS16Threads] += 1;
        mine[((x.y >> shift) & 15u) * kS16Threads] += 1;
        mine[((x.z >> shift) & 15u) * kS16Threads] += 1;
        mine[((x.w >> shift) & 15u) * kS16Threads] += 1;
    }
    for (; i < end; ++i) mine[((in[i] >> shift) & 15u) * kS16Threads] += 1;
    __syncthreads();
    // exclusive scan of the 16*512 u16 counters in linear order; thread t owns cnt[16t, 16t+16)
    uint4* seg = reinterpret_cast<uint4*>(cnt) + tid * 2;
    const uint4 c0 = seg[0], c1 = seg[1];
    const u32 w[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
    u32 v[16];
    u32 tot = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        v[2 * j] = tot; tot += w[j] & 0xffffu;
        v[2 * j + 1] = tot; tot += w[j] >> 16;
    }
    u32 inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    u32 base = inc - tot;
#pragma unroll
    for (int k = 0; k < kS16Threads / 32; ++k)
        if (k < warp) base += wsum[k];
    u32 o8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o8[j] = ((base + v[2 * j]) & 0xffffu) | ((base + v[2 * j + 1]) << 16);
    seg[0] = make_uint4(o8[0], o8[1], o8[2], o8[3]);
    seg[1] = make_uint4(o8[4], o8[5], o8[6], o8[7]);
    __syncthreads();
    if (tid < kS16Bins) {  // bin b holds every key <=> its start is 0 and the next bin starts at M
        const u32 lo = cnt[tid * kS16Threads];
        const u32 hi = tid == kS16Bins - 1 ? (u32)M : (u32)cnt[(tid + 1) * kS16Threads];
        // a count of exactly 65536 cannot occur (M <= kS16MaxM)
        if (hi - lo == (u32)M && lo == 0u) *flag = 1u;
    }
    __syncthreads();
    if (*flag) return false;
    i = beg;
    for (; i + 4 <= end; i += 4) {
        const uint4 x = *reinterpret_cast<const uint4*>(in + i);
        u16* p;
        u32 d;
        p = mine + ((x.x >> shift) & 15u) * kS16Threads; d = *p; *p = (u16)(d + 1u); out[d] = x.x;
        p = mine + ((x.y >> shift) & 15u) * kS16Threads; d = *p; *p = (u16)(d + 1u); out[d] = x.y;
        p = mine + ((x.z >> shift) & 15u) * kS16Threads; d = *p; *p = (u16)(d + 1u); out[d] = x.z;
        p = mine + ((x.w >> shift) & 15u) * kS16Threads; d = *p; *p = (u16)(d + 1u); out[d] = x.w;
    }
    for (; i < end; ++i) {
        const u32 x = in[i];
        u16* p = mine + ((x >> shift) & 15u) * kS16Threads;
        const u32 d = *p;
        *p = (u16)(d + 1u);
        out[d] = x;
    }
    __syncthreads();
    return true;
}

// Stable ascending sort of the indices 0..M-1 by key(i) (32 bits), entirely in shared memory.
// `key` is evaluated twice per element (low half at the start, high half gathered by index after the
// first four passes).  On return `*result` points at the buffer whose elements hold (index << 16 | .).
template <typename KeyFn>
__device__ inline u32* s16_sort(int M, int cap, KeyFn key, u32* smem) {
    const size_t n = ((size_t)cap + 3) / 4 * 4;
    const size_t m4 = ((size_t)M + 3) / 4 * 4;
    // When the image leaves room (3 M <= 2 cap) the full keys stay in shared memory next to the two buffers, and the
    // high halves come from there instead of a second, index-gathered pass over global memory (ncu: the two global
    // passes and the output were 45 % of the kernel's time at 12.6 K candidates, one CTA per image, 16 warps).
    const bool keep = 3 * m4 <= 2 * n;
    u32 *a = smem, *b = keep ? smem + m4 : smem + n;
    u32* kfull = smem + 2 * m4;
    u16* cnt = reinterpret_cast<u16*>(smem + 2 * n);
    u32* wsum = smem + 2 * n + (size_t)kS16Bins * kS16Threads / 2;
    const int tid = threadIdx.x;
    constexpr int UB = 4;   // independent global loads in flight per thread
    for (int i0 = tid; i0 < M; i0 += UB * kS16Threads) {
        u32 k[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const int i = i0 + u * kS16Threads;
            if (i < M) k[u] = key(i);
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const int i = i0 + u * kS16Threads;
            if (i < M) {
                a[i] = (k[u] & 0xffffu) | ((u32)i << 16);
                if (keep) kfull[i] = k[u];
            }
        }
    }
    __syncthreads();
    int pass = 0;
    for (int s = 0; s < 16; s += 4, ++pass)
        if (s16_pass(a, b, M, s, cnt, wsum, pass)) { u32* t = a; a = b; b = t; }
    if (keep) {
        for (int i = tid; i < M; i += kS16Threads) {
            const u32 idx = a[i] >> 16;
            a[i] = (kfull[idx] >> 16) | (idx << 16);
        }
    } else {
        for (int i0 = tid; i0 < M; i0 += UB * kS16Threads) {
            u32 idx[UB], k[UB];
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int i = i0 + u * kS16Threads;
                if (i < M) { idx[u] = a[i] >> 16; k[u] = key((int)idx[u]); }
            }
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int i = i0 + u * kS16Threads;
                if (i < M) a[i] = (k[u] >> 16) | (idx[u] << 16);
            }
        }
    }
    __syncthreads();
    for (int s = 0; s < 16; s += 4, ++pass)
        if (s16_pass(a, b, M, s, cnt, wsum, pass)) { u32* t = a; a = b; b = t; }
    return a;
}

}  // namespace yb
