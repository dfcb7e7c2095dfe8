// yb_nms_graph.cu — K5 (default algorithm): exact greedy NMS as a sparse suppression graph.
// Reference semantics: torchvision.ops.batched_nms as called by predict(), train.py:1232-1233 —
// same arithmetic, same tie-break, same kept set as yb_nms.cu (the dense bitmask algorithm, kept
// for dense graphs) and as torchvision's CUDA kernel; what changes is how much work is done.
//
// Greedy NMS keeps box j iff no KEPT box i with a higher score has IoU(i,j) > thr.  Two facts:
//   1. IoU(i,j) > thr >= 0 needs the boxes to overlap by at least thr of the wider/taller one
//      and their areas to be within a factor thr of each other, so after ordering the boxes by
//      (class, area octave-pair, Morton code of the centre) almost all 32x32 tile pairs can be
//      rejected from their bounding boxes alone, and most rows of the surviving pairs too;
//   2. the edges that remain are few (about 1-3 per box on dense random heads), and the greedy
//      result is the unique fixed point of  "kept <=> every higher-scored neighbour is
//      suppressed; suppressed <=> some higher-scored neighbour is kept", reached by parallel
//      rounds (8 rounds on 25,200 random boxes).
// Five launches for a whole batch:
//   graph_score_kernel    one CTA per image: stable score order (rank <-> index maps), shared-memory blocked
//                         radix sort (yb_sort.cuh); on a side stream, needed by the resolve kernel only
//   graph_spatial_kernel  one CTA per chunk of 26,880 candidates: 14-bit spatial key (class | area bucket |
//                         Hilbert index), one counting pass in shared memory -> position order
//   graph_gather_kernel   boxes / score keys / indices / classes in position order (coordinate-offset trick
//                         applied), bounding statistics per 32-box tile and per 8-box sub-tile (REDUX)
//   graph_edge_kernel     persistent warps, one row tile at a time, three culling levels (tile,
//                         sub-tile, row vs sub-tile), surviving (row, sub-tile) items packed so that
//                         every pair-loop iteration tests 4 items x 8 columns on all 32 lanes;
//                         IoU > thr decided with the division-free margin test (ambiguous lanes redo
//                         torchvision's exact fma/div arithmetic with the higher-scored box as `a`);
//                         edges (rank_hi -> rank_lo) buffered per warp, one atomic per 33..64 edges
//   graph_resolve_kernel  one CTA per image: fixed-point rounds over the edge list in shared-memory
//                         bitmaps, then ordered emission of the kept indices by score rank
// Bound: fp32 SIMT issue on the pair evaluations that survive culling.
#include "yb_common.cuh"
#include "yb_sort.cuh"
#include <cuda_fp16.h>
#include <atomic>
#include <mutex>
#include <vector>

namespace yb {

constexpr int kTile = 32;
constexpr int kSub = 8;             // columns per sub-tile (row culling granularity); 4 halves the pair tests
                                    // (46 M vs 84 M on configs[1]) but doubles the per-chunk tests: 524 vs 381 us
constexpr int kSubs = kTile / kSub;
constexpr uint32_t kNoSub = 0xffffffffu;
constexpr int kGatherThreads = 256;
constexpr int kEdgeThreads = 256;
#ifndef YB_EDGE_CTAS
#define YB_EDGE_CTAS 4
#endif
constexpr int kEdgeCtasPerSM = YB_EDGE_CTAS;
constexpr int kResolveThreads = 1024;

enum { G_PLAIN = 0, G_TRICK = 1, G_CLASS = 2 };

struct GImg {
    int M, mode, exact, overflow;
    int n_tiles;
    u32 n_edges;
    float s_off, t2;  // coordinate offset step; pruning threshold thr*(1-2^-10) (or -1: no pruning)
    unsigned long long n_evals;  // statistics: (row, sub-tile) items filtered (kSub pair tests each)
    unsigned long long n_cands;  // statistics: pairs that passed the half-precision filter (decided exactly in fp32)
};

struct GArgs {
    const float4* boxes;
    const float* scores;
    const int64_t* classes;
    const int* counts;
    int B, cap, tcap, scap;   // scap: per-image stride of sboxes (cap rounded up to a whole sub-tile)
    float thr;
    int thr_fast_ok;
    float t3;       // thr * (1 - 2^-6): threshold of the half-precision culling tests
    u32 k16x2;      // half2 {K, K}, K = (1+th)/th * (1+2^-6) rounded up, th = thr * (1 - 2^-12)
    int hexp;       // local frames map a row tile's half extent to [2^hexp, 2^(hexp+1))
    long long trick_max_numel;
    u32 *k0, *v0;                      // (B,cap) sorted runs (key, index) of large images, merged by graph_score_merge_kernel
    int n_score_chunks;
    u32* order;                        // (B,cap) score rank -> original index      (score kernel)
    u32* rinv;                         // (B,cap) original index -> score rank      (score kernel)
    u32* pos;                          // (B,cap) position -> original index        (spatial kernel)
    float4* sboxes;                    // (B,scap) boxes in position order (offset applied)
    u32* skey;                         // (B,cap) position -> descending-score key (ties: lower original index first)
    u32* scls;                         // (B,cap) position -> class id (per-class mode)
    int* iflags;                       // (B) per-image flags ORed by the chunks of the spatial kernel (zeroed per call)
    int n_chunks;                      // spatial chunks per image
    u32 resolve_smem_words;            // 32-bit words of shared memory the resolve kernel has for the edge list
    float4* tstat;                     // (B,tcap,2) tile bbox | {amin, amax, cmin, cmax}
    float4* sstat;                     // (B,tcap,kSubs,2) sub-tile bbox | {amin, amax, -, -}
    GImg* info;
    u32* ticket;
    u32 ticket_init;                   // warps of the edge kernel: each one's first work item needs no atomic
    uint2* edges;
    u64 edges_per_img;
    int64_t* keep;
    int* n_keep;
};

template <int NT>
__device__ __forceinline__ float block_reduce_minmax(float v, bool is_max, float* s32, int lane, int warp) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float n = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmaxf(v, n) : fminf(v, n);
    }
    __syncthreads();
    if (lane == 0) s32[warp] = v;
    __syncthreads();
    float r = s32[0];
    for (int w = 1; w < NT / 32; ++w) r = is_max ? fmaxf(r, s32[w]) : fminf(r, s32[w]);
    return r;
}

// Hilbert index of an 8-bit cell (x, y): a CONTINUOUS space-filling curve, so 32 consecutive boxes never jump
// across the image the way the Z-order's quadrant changes make them (those tiles had bounding boxes spanning
// half the image and did 3x the average work).  Evaluated a nibble at a time through a 4-state table
// (state = complement | swap << 1 of the remaining low bits) built in shared memory at kernel start:
// two lookups instead of an 8-level bit loop of ~80 instructions per box.
__device__ __forceinline__ unsigned short hilbert_lut_entry(int state, u32 x, u32 y) {   // x, y: 4-bit
    if (state & 1) { x = 15u - x; y = 15u - y; }
    if (state & 2) { const u32 t = x; x = y; y = t; }
    u32 d = 0, c = 0, sw = 0;
#pragma unroll
    for (u32 s = 8u; s > 0u; s >>= 1) {
        const u32 rx = (x & s) ? 1u : 0u, ry = (y & s) ? 1u : 0u;
        d += s * s * ((3u * rx) ^ ry);
        if (ry == 0u) {
            if (rx == 1u) { x = 15u - x; y = 15u - y; c ^= 1u; }
            const u32 t = x; x = y; y = t;
            sw ^= 1u;
        }
    }
    return (unsigned short)(d | ((((u32)state & 1u) ^ c) << 8) | (((((u32)state >> 1) & 1u) ^ sw) << 9));
}
__device__ __forceinline__ u32 hilbert8(const unsigned short* lut, u32 x, u32 y) {   // lut[state*256 + xn*16 + yn]
    const u32 e = lut[((x >> 4) << 4) | (y >> 4)];
    const u32 e2 = lut[((e >> 8) << 8) | ((x & 15u) << 4) | (y & 15u)];
    return ((e & 255u) << 8) | (e2 & 255u);
}

// area bucket of 1.5 octaves (tools/nms_cull_sim.py: fewest sub-tile pairs on the bench workload)
__device__ __forceinline__ u32 area_bucket(const float4 q) {
    const float area = (q.z - q.x) * (q.w - q.y);
    const float lg = fminf(fmaxf(__log2f(area), -44.0f), 120.0f);   // area <= 0 or NaN -> lowest bucket
    return (u32)(int)floorf(lg * 0.6666667f + 32.0f) & 0x7fu;
}

// 14-bit spatial key: class | area bucket (relative to the chunk's lowest) | Hilbert index of the centre, the
// curve's direction alternating with the bucket's parity so that a tile straddling two buckets stays in one corner
// of the image.  The field widths are chosen per chunk (KeyBits).  The order only has to be spatially coherent:
// tools/nms_cull_sim.py shows no loss down to 9 curve bits even with a random order inside a key, so the sort is ONE
// counting pass with shared-memory atomics.  Against round 1's (2-octave bucket | Z-order), per image of the bench
// workload: sub-tile pairs 30.3 K -> 18 K, (row, sub-tile) items 161 K -> 120 K, heaviest row tile 842 -> 160.
constexpr int kKeyBits = 14;
struct KeyBits {
    int cls_shift, bkt_shift, bkt_drop, curve_drop;
    u32 bmin;
    float cmin, qs;
};
__device__ __forceinline__ u32 spatial_key(const float4 q, const u32 cls, const KeyBits& kb, const unsigned short* lut) {
    const u32 bucket = area_bucket(q);
    const float fx = fminf(fmaxf(((q.x + q.z) * 0.5f - kb.cmin) * kb.qs, 0.0f), 255.0f);
    const float fy = fminf(fmaxf(((q.y + q.w) * 0.5f - kb.cmin) * kb.qs, 0.0f), 255.0f);
    u32 curve = hilbert8(lut, (u32)fx, (u32)fy);
    if (bucket & 1u) curve ^= 0xffffu;
    const u32 c = cls & 0x1ffu;
    return ((c << kb.cls_shift) | (((bucket - kb.bmin) >> kb.bkt_drop) << kb.bkt_shift) | (curve >> kb.curve_drop)) &
           ((1u << kKeyBits) - 1u);
}

// ------------------------------------------------------------------------------------------------
// score order: stable descending sort of one image, rank <-> index maps.  Its own launch on a side stream:
// only the resolve kernel needs it, so it overlaps with the spatial kernel and the edge discovery.
// One CTA per chunk of kScoreChunk candidates: shared-memory blocked radix sort (yb_sort.cuh).  An image that fits
// one chunk gets its final order here; larger images get sorted runs, merged by graph_score_merge_kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kScoreChunk = 26880;   // == kS16MaxM
__global__ void __launch_bounds__(kS16Threads) graph_score_kernel(const GArgs a) {
    extern __shared__ __align__(16) u32 s_dyn[];  // s16_smem_bytes(min(cap, kScoreChunk))
    const int chunk = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    constexpr int NT = kS16Threads;
    const size_t off = (size_t)b * a.cap;
    int M = a.counts ? a.counts[b] : a.cap;
    M = M < 0 ? 0 : (M > a.cap ? a.cap : M);
    const int c0 = chunk * kScoreChunk;
    const int Mc = min(M - c0, kScoreChunk);
    if (Mc <= 0) return;
    const float* scores = a.scores + off + c0;
    const u32* res = s16_sort(Mc, min(a.cap, kScoreChunk), [&](int i) { return desc_key(scores[i]); }, s_dyn);
    if (a.n_score_chunks == 1) {   // the whole image: final order
        u32* order = a.order + off;
        u32* rinv = a.rinv + off;
        for (int r = tid; r < Mc; r += NT) {
            const u32 idx = res[r] >> 16;
            order[r] = idx;
            rinv[idx] = (u32)r;
        }
    } else {                       // a sorted run for graph_score_merge_kernel
        u32* ckey = a.k0 + off + c0;
        u32* cidx = a.v0 + off + c0;
        for (int r = tid; r < Mc; r += NT) {
            const u32 idx = res[r] >> 16;
            ckey[r] = desc_key(scores[idx]);
            cidx[r] = (u32)c0 + idx;
        }
    }
}

// Large images (more candidates than one shared-memory sort holds, e.g. 100,800 at 1280^2): the runs of
// graph_score_kernel are merged by ranking.  A thread takes one element of one run; its final rank is its position
// in its own run plus, for every other run, the number of elements that go before it — upper bound in the runs of
// lower indices (they win ties: stable), lower bound in the runs of higher indices.  Neighbouring threads search
// neighbouring keys, so the binary searches stay in cache.
__global__ void __launch_bounds__(256) graph_score_merge_kernel(const GArgs a) {
    const int b = blockIdx.y;
    const int p = blockIdx.x * 256 + threadIdx.x;
    const size_t off = (size_t)b * a.cap;
    int M = a.counts ? a.counts[b] : a.cap;
    M = M < 0 ? 0 : (M > a.cap ? a.cap : M);
    if (p >= M) return;
    const u32* ckey = a.k0 + off;
    const u32 key = ckey[p];
    const int c = p / kScoreChunk;
    u32 rank = (u32)(p - c * kScoreChunk);
    for (int o = 0; o * kScoreChunk < M; ++o) {
        if (o == c) continue;
        const u32* run = ckey + (size_t)o * kScoreChunk;
        const int len = min(kScoreChunk, M - o * kScoreChunk);
        int lo = 0, hi = len;   // first position whose key is > key (o < c) or >= key (o > c)
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const u32 k = run[mid];
            const bool before = o < c ? k <= key : k < key;
            if (before) lo = mid + 1; else hi = mid;
        }
        rank += (u32)lo;
    }
    const u32 idx = a.v0[off + p];
    a.order[off + rank] = idx;
    a.rinv[off + idx] = rank;
}

// ------------------------------------------------------------------------------------------------
// spatial order + gather.  One CTA of 1024 threads per chunk of kChunk candidates of one image (one chunk at
// 640^2; a 1280^2 image is four chunks = horizontal bands of the P3 grid, because candidates arrive in row-major
// grid order): validity / range reductions, 14-bit spatial key, ONE counting pass (histogram and scatter with
// shared-memory atomics; the order inside a key is arbitrary and changes nothing but the work distribution), then
// — straight from shared memory — boxes (coordinate-offset trick applied, boxes.py:99-101), score keys, original
// indices and classes in position order, and the bounding statistics of every 32-box tile and 8-box sub-tile.
// ------------------------------------------------------------------------------------------------
constexpr int kChunk = 26880;        // 840 tiles; u16 indices
constexpr int kSpThreads = 1024;
constexpr int kSpBatch = 4;

__host__ __device__ inline size_t spatial_smem_bytes(int ccap) {
    return ((size_t)1 << kKeyBits) * sizeof(u32) + 2 * (((size_t)ccap + 7) / 8 * 8) * sizeof(unsigned short) +
           1024 * sizeof(unsigned short);
}

__global__ void __launch_bounds__(kSpThreads) graph_spatial_kernel(const GArgs a) {
    extern __shared__ __align__(16) u32 s_dyn[];
    __shared__ float s_f[32];
    __shared__ int s_flag[32];
    __shared__ int s_maxcls[32];
    __shared__ int s_bmin[32], s_bmax[32];
    __shared__ u32 s_scan[32];
    constexpr int NT = kSpThreads;
    constexpr int NB = 1 << kKeyBits;

    const int chunk = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t off = (size_t)b * a.cap;
    {   // lookup table of the curve, zeroed histogram
        const int ccap0 = (min(a.cap, kChunk) + 7) & ~7;
        unsigned short* lut0 = reinterpret_cast<unsigned short*>(s_dyn + NB) + 2 * ccap0;
        lut0[tid] = hilbert_lut_entry(tid >> 8, (u32)(tid >> 4) & 15u, (u32)tid & 15u);
#pragma unroll
        for (int k = 0; k < NB / NT; ++k) s_dyn[k * NT + tid] = 0u;
    }
    int M = a.counts ? a.counts[b] : a.cap;
    M = M < 0 ? 0 : (M > a.cap ? a.cap : M);
    GImg* info = a.info + b;
    if (b == 0 && chunk == 0 && tid == 0) *a.ticket = a.ticket_init;   // the edge kernel's warps own tickets 0..n_warps-1
    const int c0 = chunk * kChunk;
    const int Mc = min(M - c0, kChunk);
    if (Mc <= 0) {
        if (tid == 0 && chunk == 0) {
            GImg z = {0, G_PLAIN, 0, 0, 0, 0u, 0.0f, -1.0f, 0ull, 0ull};
            *info = z;
            if (a.n_chunks == 1) a.iflags[b] = 0;
        }
        return;
    }
    const float4* boxes = a.boxes + off + c0;
    const int64_t* classes = a.classes ? a.classes + off + c0 : nullptr;
    const int ccap = (min(a.cap, kChunk) + 7) & ~7;
    u32* hist = s_dyn;
    unsigned short* keys = reinterpret_cast<unsigned short*>(s_dyn + NB);
    unsigned short* oidx = keys + ccap;
    unsigned short* lut = oidx + ccap;


    // ---- reductions (torchvision's boxes.max(), validity, centre range, class range, bucket range) ----
    const int mode = !a.classes ? G_PLAIN : ((long long)M * 4 <= a.trick_max_numel ? G_TRICK : G_CLASS);
    float mx = -INFINITY, cmin = INFINITY, cmax = -INFINITY;
    int bad = 0, maxcls = 0, bmin = 127, bmax = 0;
    // kSpBatch independent loads in flight per thread.  (Under ncu, with flushed caches, the first use of a loaded box
    // is this kernel's top stall; in the pipeline the boxes are L2-resident and the kernel's time did not change with
    // the batching: 4.9 M warp instructions on 64 SMs are 40 % of it in issue alone.)
    for (int i0 = tid; i0 < Mc; i0 += kSpBatch * NT) {
        float4 qv[kSpBatch];
        long long cv[kSpBatch];
#pragma unroll
        for (int u = 0; u < kSpBatch; ++u) {
            const int i = i0 + u * NT;
            if (i < Mc) {
                qv[u] = __ldg(boxes + i);
                cv[u] = classes ? __ldg(classes + i) : 0ll;
            }
        }
#pragma unroll
        for (int u = 0; u < kSpBatch; ++u) {
            if (i0 + u * NT >= Mc) break;
            const float4 q = qv[u];
            mx = fmaxf(mx, fmaxf(fmaxf(q.x, q.y), fmaxf(q.z, q.w)));
            const bool nan = (q.x != q.x) || (q.y != q.y) || (q.z != q.z) || (q.w != q.w);
            const float big = fmaxf(fmaxf(fabsf(q.x), fabsf(q.y)), fmaxf(fabsf(q.z), fabsf(q.w)));
            bad |= (nan ? 4 : 0) | ((nan || !(big <= 1e17f) || !(q.z >= q.x) || !(q.w >= q.y)) ? 1 : 0);
            const float cx = (q.x + q.z) * 0.5f, cy = (q.y + q.w) * 0.5f;
            cmin = fminf(cmin, fminf(cx, cy));
            cmax = fmaxf(cmax, fmaxf(cx, cy));
            const int bk = (int)area_bucket(q);
            bmin = min(bmin, bk);
            bmax = max(bmax, bk);
            if (classes) {
                const long long c = cv[u];
                bad |= (c < 0 || c >= 512) ? 2 : 0;
                maxcls = max(maxcls, (int)(c & 0x1ff));
            }
        }
    }
    bad = __reduce_or_sync(0xffffffffu, bad);
    maxcls = __reduce_max_sync(0xffffffffu, maxcls);
    bmin = __reduce_min_sync(0xffffffffu, bmin);
    bmax = __reduce_max_sync(0xffffffffu, bmax);
    if (lane == 0) { s_flag[warp] = bad; s_maxcls[warp] = maxcls; s_bmin[warp] = bmin; s_bmax[warp] = bmax; }
    mx = block_reduce_minmax<NT>(mx, true, s_f, lane, warp);
    cmin = block_reduce_minmax<NT>(cmin, false, s_f, lane, warp);
    cmax = block_reduce_minmax<NT>(cmax, true, s_f, lane, warp);
    bad = 0; maxcls = 0; bmin = 127; bmax = 0;
    for (int w = 0; w < NT / 32; ++w) {
        bad |= s_flag[w]; maxcls = max(maxcls, s_maxcls[w]); bmin = min(bmin, s_bmin[w]); bmax = max(bmax, s_bmax[w]);
    }
    // torchvision's offset step (boxes.py:99); only single-chunk images can be in trick mode (4*M <= 100000)
    float s_off = mx + 1.0f;
    if (bad & 4) s_off = __int_as_float(0x7fc00000);    // torch max propagates NaN
    if (mode == G_TRICK && !((float)maxcls * s_off + fabsf(mx) <= 1e17f)) bad |= 1;
    const bool exact = (bad & 1) || !a.thr_fast_ok;
    // 2: class ids beyond the key; 4: boxes or threshold the half-precision filter cannot serve (NaN/Inf, inverted,
    // |coordinate| > 1e17, thr outside [0.03, 1e30]).  Both are decided by the resolve kernel's greedy pass.
    const int flags = ((bad & 2) ? 2 : 0) | (exact ? 4 : 0);
    if (tid == 0) {
        if (a.n_chunks == 1) a.iflags[b] = flags;          // the only chunk: a plain store, no memset node before the launch
        else if (flags) atomicOr(a.iflags + b, flags);     // several chunks OR into a word the host zeroed
    }

    // ---- key layout of this chunk ----
    KeyBits kb;
    {
        const int nb_cls = classes ? 32 - __clz(maxcls) : 0;                 // <= 9
        const int span_bits = 32 - __clz(max(bmax - bmin, 0));             // <= 7
        const int nb_bkt = min(span_bits, kKeyBits - nb_cls);
        const int nb_curve = kKeyBits - nb_cls - nb_bkt;
        kb.bkt_drop = span_bits - nb_bkt;
        kb.cls_shift = nb_bkt + nb_curve;
        kb.bkt_shift = nb_curve;
        kb.curve_drop = 16 - nb_curve;
        kb.bmin = (u32)bmin;
        kb.cmin = cmin;
        kb.qs = (cmax > cmin) ? 255.0f / (cmax - cmin) : 0.0f;
    }
    // (the block reductions above ended with a barrier: lut and the zeroed histogram are visible)

    // ---- counting sort by key: histogram, exclusive scan, scatter ----
    for (int i0 = tid; i0 < Mc; i0 += kSpBatch * NT) {
        float4 qv[kSpBatch];
        u32 cv[kSpBatch];
#pragma unroll
        for (int u = 0; u < kSpBatch; ++u) {
            const int i = i0 + u * NT;
            if (i < Mc) {
                qv[u] = __ldg(boxes + i);
                cv[u] = classes ? (u32)__ldg(classes + i) : 0u;
            }
        }
#pragma unroll
        for (int u = 0; u < kSpBatch; ++u) {
            const int i = i0 + u * NT;
            if (i >= Mc) break;
            const u32 k = spatial_key(qv[u], cv[u], kb, lut);
            keys[i] = (unsigned short)k;
            atomicAdd(&hist[k], 1u);
        }
    }
    __syncthreads();
    {
        constexpr int PER = NB / NT;   // 16 consecutive bins per thread
        u32 v[PER], sum = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) { v[k] = hist[tid * PER + k]; sum += v[k]; }
        u32 inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        if (lane == 31) s_scan[warp] = inc;
        __syncthreads();
        u32 base = inc - sum;
        for (int w = 0; w < warp; ++w) base += s_scan[w];
#pragma unroll
        for (int k = 0; k < PER; ++k) { hist[tid * PER + k] = base; base += v[k]; }
    }
    __syncthreads();
    for (int i0 = tid; i0 < Mc; i0 += kSpBatch * NT) {   // the returning atomics of a batch are in flight together
        u32 d[kSpBatch];
#pragma unroll
        for (int u = 0; u < kSpBatch; ++u) {
            const int i = i0 + u * NT;
            if (i < Mc) d[u] = atomicAdd(&hist[keys[i]], 1u);
        }
#pragma unroll
        for (int u = 0; u < kSpBatch; ++u) {
            const int i = i0 + u * NT;
            if (i < Mc) oidx[d[u]] = (unsigned short)i;
        }
    }
    __syncthreads();

    // ---- position -> original index (the gather runs as its own, machine-wide launch) ----
    u32* pos = a.pos + off + c0;
    for (int p = tid; p < Mc; p += NT) pos[p] = (u32)c0 + oidx[p];
    if (tid == 0 && chunk == 0) {
        GImg o;
        o.M = M; o.mode = mode; o.exact = 0;
        o.overflow = 0;   // the edge and resolve kernels look at iflags[b]
        o.n_tiles = (M + kTile - 1) / kTile; o.n_edges = 0u; o.s_off = s_off;
        o.t2 = a.thr * (1.0f - 9.765625e-4f);
        o.n_evals = 0ull;
        o.n_cands = 0ull;
        *info = o;
    }
}

// ------------------------------------------------------------------------------------------------
// gather in position order (coordinate-offset trick, boxes.py:99-101) + tile / sub-tile statistics: a warp per
// 32-box tile, the whole machine.  Minima / maxima go through the integer warp-reduce unit (REDUX) on
// order-preserving integer images of the floats: 12 reductions per tile instead of 80 shuffle + min/max steps.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 f2ord(const float f) {
    const u32 u = __float_as_uint(f);
    return u ^ ((u32)((int)u >> 31) | 0x80000000u);
}
__device__ __forceinline__ float ord2f(const u32 o) {
    return __uint_as_float(o ^ ((o >> 31) ? 0x80000000u : 0xffffffffu));
}

__global__ void __launch_bounds__(kGatherThreads) graph_gather_kernel(const GArgs a) {
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const GImg info = a.info[b];
    const int t = blockIdx.x * (kGatherThreads / 32) + warp;
    if (t >= info.n_tiles) return;
    const int M = info.M;
    const size_t off = (size_t)b * a.cap;
    const float4* boxes = a.boxes + off;
    const int64_t* classes = a.classes ? a.classes + off : nullptr;
    const int p = t * kTile + lane;
    u32 x1 = 0xffffffffu, y1 = 0xffffffffu, x2 = 0u, y2 = 0u, amin = 0xffffffffu, amax = 0u, c0 = 0xffffffffu, c1 = 0u;
    float4* sbo = a.sboxes + (size_t)b * a.scap;
    if (p >= M && p < ((M + kSub - 1) & ~(kSub - 1)))   // the edge kernel stages whole sub-tiles: far-away padding boxes
        sbo[p] = make_float4(1e18f, 1e18f, 2e18f, 2e18f);
    if (p < M) {
        const u32 idx = a.pos[off + p];
        float4 q = boxes[idx];
        const float sc = a.scores[off + idx];
        u32 c = 0;
        if (classes) c = (u32)classes[idx];
        if (info.mode == G_TRICK) {
            const float o = (float)classes[idx] * info.s_off;
            q.x += o; q.y += o; q.z += o; q.w += o;
        }
        const u32 cls = (info.mode == G_CLASS) ? c : 0u;
        sbo[p] = q;
        a.skey[off + p] = desc_key(sc);
        if (info.mode == G_CLASS) a.scls[off + p] = cls;   // read only in per-class mode
        x1 = f2ord(q.x); y1 = f2ord(q.y); x2 = f2ord(q.z); y2 = f2ord(q.w);
        amin = amax = f2ord((q.z - q.x) * (q.w - q.y));
        c0 = c1 = cls;
    }
    // this lane's 8-box sub-tile: three butterfly steps (a REDUX per sub-tile mask runs four times per warp, once per
    // distinct mask: ncu counted 216 of the kernel's 380 instructions per tile there)
    u32 sx1 = x1, sy1 = y1, sx2 = x2, sy2 = y2, samin = amin, samax = amax;
#pragma unroll
    for (int o = 1; o < kSub; o <<= 1) {
        sx1 = min(sx1, __shfl_xor_sync(0xffffffffu, sx1, o));
        sy1 = min(sy1, __shfl_xor_sync(0xffffffffu, sy1, o));
        sx2 = max(sx2, __shfl_xor_sync(0xffffffffu, sx2, o));
        sy2 = max(sy2, __shfl_xor_sync(0xffffffffu, sy2, o));
        samin = min(samin, __shfl_xor_sync(0xffffffffu, samin, o));
        samax = max(samax, __shfl_xor_sync(0xffffffffu, samax, o));
    }
    if ((lane & (kSub - 1)) == 0) {
        float4* ss = a.sstat + (((size_t)b * a.tcap + t) * kSubs + (lane / kSub)) * 2;
        // an empty sub-tile (beyond M) keeps +inf/-inf style bounds: it never passes a test (and is never visited)
        ss[0] = make_float4(ord2f(sx1), ord2f(sy1), ord2f(sx2), ord2f(sy2));
        ss[1] = make_float4(ord2f(samin), ord2f(samax), 0.0f, 0.0f);
    }
    const u32 tx1 = __reduce_min_sync(0xffffffffu, sx1), ty1 = __reduce_min_sync(0xffffffffu, sy1);
    const u32 tx2 = __reduce_max_sync(0xffffffffu, sx2), ty2 = __reduce_max_sync(0xffffffffu, sy2);
    const u32 tamin = __reduce_min_sync(0xffffffffu, samin), tamax = __reduce_max_sync(0xffffffffu, samax);
    c0 = __reduce_min_sync(0xffffffffu, c0);
    c1 = __reduce_max_sync(0xffffffffu, c1);
    if (lane == 0) {
        float4* ts = a.tstat + ((size_t)b * a.tcap + t) * 2;
        ts[0] = make_float4(ord2f(tx1), ord2f(ty1), ord2f(tx2), ord2f(ty2));
        ts[1] = make_float4(ord2f(tamin), ord2f(tamax), __uint_as_float(c0), __uint_as_float(c1));
    }
}

// ------------------------------------------------------------------------------------------------
// edge discovery (v4: packed fp16 candidate filter + exact fp32 decision)
//
// 99.4 % of the pair tests that survive the tile statistics find no edge (103 tests, 0.66 edges per box on
// the bench workload), and the kernel is bound by the ALU/FMA issue rate (FMNMX, HMNMX2, HFMA2: 2 warp
// instructions per clock and SM, profiles/r2_ubench_pipes.txt).  So the bulk of the tests runs as a
// CONSERVATIVE filter on packed halves — one HMNMX2/HADD2/HFMA2 handles two columns — in a coordinate frame
// local to the row tile (origin = centre of the tile's bounding box, power-of-two scale that maps its half
// extent to [2^h, 2^(h+1))).  Every conversion is rounded in the direction that can only ADD candidates
// (lower corners down, upper corners up, areas towards zero, thresholds with 2^-6 of slack), boxes too small
// for the half format in that frame are made unconditional candidates, and whatever passes is queued and
// decided 32 at a time in fp32 with torchvision's arithmetic (division-free margin test, exact fma/div when
// ambiguous).  The edge set is therefore exactly the set of pairs with IoU > thr; only the amount of work
// depends on the half-precision stage.  Images the filter cannot serve (NaN/Inf/inverted boxes, thresholds
// outside [0.03, 1e30]) are not handled here at all: the resolve kernel's greedy pass decides them.
// ------------------------------------------------------------------------------------------------
struct Rec16 {   // 32 bytes: a box (rows: the same box in both halves; columns: two neighbouring columns) in the
    uint4 box;   // row tile's local frame: half2 x1 | y1 | x2 | y2, corners rounded outwards
    u32 area;    // half2 area, rounded towards zero, capped at 32768; -60000 = "too small for the format"
    u32 pad[3];
};
constexpr int kSlots = 8;            // sub-tiles per chunk: 64 columns, two per lane
constexpr int kQueue = kTile + 8;    // sub-tiles queued per level-2 step (<= 32) + < 8 left over
constexpr int kCand = 3 * kTile;     // candidate queue: processed when >= 32, one iteration adds <= 64
constexpr int kItems = kSlots * kTile + kSlots;
struct EdgeWarp {
    Rec16 row16[kTile + 1];          // [kTile]: a sentinel that overlaps nothing (pads the item list)
    Rec16 col16[kSlots * 4];         // lane l of the chunk: columns 2*(l&3), +1 of slot l>>2
    u32 tl[kTile];                   // compaction scratch: surviving tiles of the current step
    u32 sub[kQueue];                 // queue of surviving sub-tile ids (J*kSubs+s) ...
    uint2 sbx[kQueue];               // ... their bounding boxes in the local frame: half2 {x1,y1} down | {x2,y2} up ...
    u32 sar[kQueue];                 // ... and half2 {max area up, -(t3 * min area) down}
    unsigned short irow[kItems];     // packed work list of a chunk: byte offset of the item's row record in row16 ...
    unsigned short icol[kItems];     // ... and of its slot's first column record in col16
    u32 cand[kCand];                 // (row << 24) | column position: pairs that passed the half-precision filter
    uint2 ebuf[2 * kTile];           // edges found, flushed to the image's list 33..64 at a time
};

__device__ __forceinline__ u32 pack2(const unsigned short lo, const unsigned short hi) { return (u32)lo | ((u32)hi << 16); }
__device__ __forceinline__ __half2 as_h2(const u32 v) { return *reinterpret_cast<const __half2*>(&v); }
__device__ __forceinline__ u32 as_u32(const __half2 v) { return *reinterpret_cast<const u32*>(&v); }
// float -> half bits with directed rounding (F2F.F16.F32.RM / .RP / .RZ)
__device__ __forceinline__ unsigned short h_dn(const float v) { return __half_as_ushort(__float2half_rd(v)); }
__device__ __forceinline__ unsigned short h_up(const float v) { return __half_as_ushort(__float2half_ru(v)); }
__device__ __forceinline__ unsigned short h_rz(const float v) { return __half_as_ushort(__float2half_rz(v)); }
// both halves of a >= the halves of b (ordered: false on NaN): one HSETP2 + PLOP3
__device__ __forceinline__ bool h2_all_ge(const u32 a, const u32 b) {
    u32 r;
    asm("{\n\t.reg .pred p, q;\n\tsetp.ge.f16x2 p|q, %1, %2;\n\tand.pred p, p, q;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(r) : "r"(a), "r"(b));
    return r != 0u;
}
// either half of a >= 0 (ordered)
__device__ __forceinline__ bool h2_any_ge0(const u32 a) {
    u32 r;
    asm("{\n\t.reg .pred p, q;\n\tsetp.ge.f16x2 p|q, %1, %2;\n\tor.pred p, p, q;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(r) : "r"(a), "r"(0u));
    return r != 0u;
}
// bit 0: low half of a >= 0, bit 1: high half (rare path)
__device__ __forceinline__ u32 h2_ge0_bits(const u32 a) {
    u32 r;
    asm("{\n\t.reg .pred p, q;\n\t.reg .u32 t;\n\tsetp.ge.f16x2 p|q, %1, %2;\n\tselp.u32 %0, 1, 0, p;\n\tselp.u32 t, 2, 0, q;\n\tor.b32 %0, %0, t;\n\t}"
        : "=r"(r) : "r"(a), "r"(0u));
    return r;
}

struct Frame {   // local frame of a row tile
    float S, S2, nox, noy;   // scale (power of two), its square, -origin*S
};
// area of a nice box in the local frame as half bits: lower bound (towards zero), capped; boxes whose area
// would leave the half format's normal range become unconditional candidates (marker -60000)
__device__ __forceinline__ unsigned short area16_down(const float4 q, const float S2) {
    const float a = __fmul_rd(__fmul_rd(__fsub_rd(q.z, q.x), __fsub_rd(q.w, q.y)), S2);
    return a >= 2.44140625e-4f ? h_rz(fminf(a, 32768.0f)) : (unsigned short)0xfb53u;   // 0xfb53 = -60000 (rz)
}

// Append the warp's buffered edges to image b's list.
__device__ __forceinline__ void edge_flush(const GArgs& a, EdgeWarp& w, int b, uint2* __restrict__ edges, int& n_buf) {
    if (n_buf == 0) return;  // warp-uniform
    const int lane = threadIdx.x & 31;
    __syncwarp();
    u32 base = 0;
    if (lane == 0) base = atomicAdd(&a.info[b].n_edges, (u32)n_buf);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (int k = lane; k < n_buf; k += 32) {
        const u64 idx = (u64)base + k;
        if (idx < a.edges_per_img) edges[idx] = w.ebuf[k];
    }
    __syncwarp();
    n_buf = 0;
}

// Exact stage: the queued (row, column) pairs, one per lane.  torchvision's predicate with the division-free
// margin test; ambiguous lanes redo the exact fma/div arithmetic with the higher-scored box as `a`.
__device__ __forceinline__ void cand_process(const GArgs& a, EdgeWarp& w, const int M, const bool class_mode, int b, int I,
                                          int& n_cand, int& n_buf) {
    // (rare path: the pointers are rebuilt here instead of being kept in registers through the hot loops)
    const size_t off = (size_t)b * a.cap;
    const float4* __restrict__ sb = a.sboxes + (size_t)b * a.scap;
    const u32* __restrict__ skey = a.skey + off;
    const u32* __restrict__ sidx = a.pos + off;   // position -> original index (graph_spatial_kernel)
    const u32* __restrict__ scls = a.scls + off;
    uint2* __restrict__ edges = a.edges + (u64)b * a.edges_per_img;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const float kEps = 9.5367431640625e-07f;     // 2^-20
    const float kTiny = 7.888609052210118e-31f;  // 2^-100
    const float thr = a.thr;
    __syncwarp();
    for (int k0 = 0; k0 < n_cand; k0 += 32) {
        const int k = k0 + lane;
        bool fin = false;
        u32 rlo = 0u, rhi = 0u;
        if (k < n_cand) {
            const u32 e = w.cand[k];
            const int rp = I * kTile + (int)(e >> 24);
            const int qp = (int)(e & 0xffffffu);
            if (qp > rp && qp < M && rp < M) {   // each unordered pair once; padding columns of the last sub-tile never
                const float4 r = sb[rp];
                const float4 q = sb[qp];
                // score order without the score sort: descending-score key, ties by the lower original index
                const u32 rk = skey[rp], ck = skey[qp], ri = sidx[rp], ci = sidx[qp];
                const bool row_first = rk < ck || (rk == ck && ri < ci);
                const float qw = q.z - q.x, qh = q.w - q.y, rw = r.z - r.x, rh = r.w - r.y;
                const float qarea = qw * qh, rarea = rw * rh;
                const float left = fmaxf(r.x, q.x), right = fminf(r.z, q.z);
                const float top = fmaxf(r.y, q.y), bottom = fminf(r.w, q.w);
                const float iw = fmaxf(right - left, 0.0f), ih = fmaxf(bottom - top, 0.0f);
                const float inter = iw * ih;
                // den within a few ulp of torchvision's fma form whichever box plays `a`: inside the margin
                const float den0 = (qarea + rarea) - inter;
                const float tt = thr * den0;
                const float d = inter - tt;
                bool pr = d > 0.0f;
                if (!(fabsf(d) > __fmaf_rn(tt, kEps, kTiny))) {
                    const bool row_a = row_first;   // torchvision devIoU with a = the higher-scored box
                    const float sa = row_a ? rarea : qarea;
                    const float bw = row_a ? qw : rw, bh = row_a ? qh : rh;
                    const float den = __fmaf_rn(bw, bh, sa) - inter;
                    pr = (inter / den) > thr;
                }
                fin = pr && (!class_mode || scls[qp] == scls[rp]);
                rlo = row_first ? ri : ci;   // edge = (higher-scored, lower-scored) ORIGINAL indices; the resolve
                rhi = row_first ? ci : ri;   // kernel turns them into score ranks
            }
        }
        const unsigned em = __ballot_sync(0xffffffffu, fin);
        if (em) {
            if (fin) w.ebuf[n_buf + __popc(em & lt_mask)] = make_uint2(rlo, rhi);
            n_buf += __popc(em);
            if (n_buf > kTile) edge_flush(a, w, b, edges, n_buf);  // no room for another 32
        }
    }
    __syncwarp();
    n_cand = 0;
}

// One chunk = up to 8 sub-tiles (8 columns each) that survived the tile-level tests against row tile I.
// Lane l stages columns 2*(l&3), +1 of slot l>>2 as one packed record.  Rows are culled against each slot's
// statistics (one half2 step per dimension pair), the surviving (row, slot) items are packed, and every
// iteration of the pair loop filters 8 items x 8 columns: 64 pairs on 32 lanes whatever the culling pattern.
// FULL: all 8 slots are valid (every chunk but the last one of a row tile).
template <bool FULL>
__device__ __forceinline__ void edge_chunk(const GArgs& a, EdgeWarp& w, const int M, const bool class_mode, int b, int I,
                                           int head, const float4* __restrict__ sb, const Frame& f, const int n_slots,
                                           const u32 r_lo, const u32 r_hi, const u32 r_wh_t, const u32 r_ar,
                                           u32& n_evals, u32& n_cands, int& n_cand, int& n_buf) {
    const int lane = threadIdx.x & 31;
    const int grp = lane >> 2, j = lane & 3;
    const unsigned lt_mask = (1u << lane) - 1u;
    const u32* sub = w.sub + head;  // the chunk's eight sub-tile ids (kNoSub = empty slot)

    // ---- this lane's two columns, in the row tile's frame (positions up to the end of the last sub-tile hold
    //      far-away padding boxes: graph_gather_kernel) ----
    {
        const u32 e = sub[grp];
        uint4 box = make_uint4(0x63d063d0u, 0x63d063d0u, 0xe3d0e3d0u, 0xe3d0e3d0u);   // +1000 | -1000: overlaps nothing
        u32 area = 0x3c003c00u;                                                          // 1.0
        if (FULL || e != kNoSub) {
            const float4* src = sb + (size_t)e * kSub + 2 * j;
            const float4 q0 = src[0], q1 = src[1];
            box.x = pack2(h_dn(__fmaf_rd(q0.x, f.S, f.nox)), h_dn(__fmaf_rd(q1.x, f.S, f.nox)));
            box.y = pack2(h_dn(__fmaf_rd(q0.y, f.S, f.noy)), h_dn(__fmaf_rd(q1.y, f.S, f.noy)));
            box.z = pack2(h_up(__fmaf_ru(q0.z, f.S, f.nox)), h_up(__fmaf_ru(q1.z, f.S, f.nox)));
            box.w = pack2(h_up(__fmaf_ru(q0.w, f.S, f.noy)), h_up(__fmaf_ru(q1.w, f.S, f.noy)));
            area = pack2(area16_down(q0, f.S2), area16_down(q1, f.S2));
        }
        Rec16& c = w.col16[lane];
        c.box = box;
        c.area = area;
    }

    // ---- row culling against the statistics of each slot; pack the surviving (row, slot) items.  Rows beyond
    //      the image carry +-1000 sentinels in r_lo / r_hi / r_wh_t and never pass. ----
    int n_items = 0;
    const unsigned short my_row = (unsigned short)(lane * (int)sizeof(Rec16));
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
        if (FULL || s < n_slots) {   // warp-uniform: the valid slots of a chunk are a prefix
            const uint2 sbx = w.sbx[head + s];
            const __half2 o = __hsub2(__hmin2(as_h2(r_hi), as_h2(sbx.y)), __hmax2(as_h2(r_lo), as_h2(sbx.x)));
            const bool rok = h2_all_ge(as_u32(o), r_wh_t) & h2_all_ge(w.sar[head + s], r_ar);
            const unsigned m = __ballot_sync(0xffffffffu, rok);
            if (rok) {
                const int at = n_items + __popc(m & lt_mask);
                w.irow[at] = my_row;
                w.icol[at] = (unsigned short)(s * 4 * (int)sizeof(Rec16));
            }
            n_items += __popc(m);
        }
    }
    if (lane < kSlots) {   // pad with the sentinel row (against slot 0)
        w.irow[n_items + lane] = (unsigned short)(kTile * (int)sizeof(Rec16));
        w.icol[n_items + lane] = 0;
    }
    n_evals += (u32)n_items;
    __syncwarp();

    // ---- half-precision filter: 8 items x 8 columns per iteration ----
    const __half2 K = as_h2(a.k16x2);
    const __half2 zero = as_h2(0u);
    const char* rowbase = reinterpret_cast<const char*>(w.row16);
    const char* colbase = reinterpret_cast<const char*>(w.col16) + j * sizeof(Rec16);
    const unsigned short* irow = w.irow + grp;
    const unsigned short* icol = w.icol + grp;
    for (int it = 0; it < n_items; it += kSlots) {
        const u32 ro = irow[it], co = icol[it];
        const Rec16* rr = reinterpret_cast<const Rec16*>(rowbase + ro);
        const Rec16* cc = reinterpret_cast<const Rec16*>(colbase + co);
        const uint4 R = rr->box, C = cc->box;
        const __half2 iw = __hmax2(__hsub2(__hmin2(as_h2(R.z), as_h2(C.z)), __hmax2(as_h2(R.x), as_h2(C.x))), zero);
        const __half2 ih = __hsub2(__hmin2(as_h2(R.w), as_h2(C.w)), __hmax2(as_h2(R.y), as_h2(C.y)));
        const __half2 sum = __hadd2(as_h2(rr->area), as_h2(cc->area));
        const u32 x = as_u32(__hfma2(__hmul2(iw, ih), K, __hneg2(sum)));
        if (!__any_sync(0xffffffffu, h2_any_ge0(x))) continue;   // NaN (no overlap at all) is not a candidate
        // queue the candidates: (row, first column of the hit) in one ballot; a lane whose two columns both hit
        // (rare) queues the second one in another round
        const u32 i = ro / (u32)sizeof(Rec16);
        const u32 qp = sub[co / (u32)(4 * sizeof(Rec16))] * kSub + 2 * j;
        // every unordered pair once, and never a box with itself (each box passes the filter against itself: that
        // alone was 1.5 of the 3.2 candidates per edge); padding items never
        const u32 rp = (u32)I * kTile + i;
        u32 hit = ro < (u32)(kTile * sizeof(Rec16)) ? h2_ge0_bits(x) : 0u;
        if (qp <= rp) hit &= 2u;
        if (qp + 1u <= rp) hit = 0u;
        const unsigned m0 = __ballot_sync(0xffffffffu, hit != 0u);
        if (m0 == 0u) continue;   // only self / mirrored pairs of a diagonal tile
        if (hit) w.cand[n_cand + __popc(m0 & lt_mask)] = (i << 24) | (qp + (hit == 2u ? 1u : 0u));
        n_cand += __popc(m0);
        const unsigned m1 = __ballot_sync(0xffffffffu, hit == 3u);
        if (m1) {
            if (hit == 3u) w.cand[n_cand + __popc(m1 & lt_mask)] = (i << 24) | (qp + 1u);
            n_cand += __popc(m1);
        }
        if (n_cand >= kTile) {
            n_cands += (u32)n_cand;
            cand_process(a, w, M, class_mode, b, I, n_cand, n_buf);
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kEdgeThreads, kEdgeCtasPerSM) graph_edge_kernel(const GArgs a) {
    __shared__ __align__(16) EdgeWarp s_w[kEdgeThreads / 32];
    const int lane = threadIdx.x & 31;
    EdgeWarp& w = s_w[threadIdx.x >> 5];
    const unsigned lt_mask = (1u << lane) - 1u;
    // work items = (row tile, image), all images' tile 0 first, up to the largest image's last tile: with half-full
    // images (conf 0.5: 394 of 788 tiles) the tickets beyond it are empty atomic + load round trips (1 % of the kernel)
    int max_tiles = 0;
    for (int i = lane; i < a.B; i += 32) max_tiles = max(max_tiles, a.iflags[i] ? 0 : a.info[i].n_tiles);
    max_tiles = __reduce_max_sync(0xffffffffu, max_tiles);
    const u32 total = (u32)max_tiles * (u32)a.B;

    bool first = true;
    for (;;) {
        // a warp's first item is its own index (4,736 atomics on one word at kernel start otherwise); after that one
        // ticket at a time (taking it one item ahead, to hide the atomic's round trip, measured 6 % SLOWER: a warp then
        // sits on a reserved item while others run dry at the end)
        u32 ticket = blockIdx.x * (kEdgeThreads / 32) + (threadIdx.x >> 5);
        if (!first) {
            if (lane == 0) ticket = atomicAdd(a.ticket, 1u);
            ticket = __shfl_sync(0xffffffffu, ticket, 0);
        }
        first = false;
        if (ticket >= total) break;
        // Row tiles are handed out from the LAST one down: a row tile only looks at column tiles J >= I, but the order
        // is by area bucket and the big-box buckets at the end overlap far more sub-tiles (tools/nms_cull_sim.py:
        // 28 sub-tile pairs per row tile in the first tenth of an image, 60 in the ninth, single tiles up to 160), so
        // ascending order left the heaviest items for the end of the kernel.
        const int I = max_tiles - 1 - (int)(ticket / (u32)a.B), b = (int)(ticket % (u32)a.B);
        const float4* sb = a.sboxes + (size_t)b * a.scap;
        const float4* ts = a.tstat + (size_t)b * a.tcap * 2;
        const float4* ss = a.sstat + (size_t)b * a.tcap * kSubs * 2;
        // the item's first loads are issued together with the image record (all in bounds for any I < tcap): one
        // round trip to L2 per work item instead of three dependent ones
        const float4 ib = ts[I * 2], ia = ts[I * 2 + 1];
        const int rp = I * kTile + lane;
        const float4 rq = rp < a.scap ? sb[rp] : make_float4(0.f, 0.f, 0.f, 0.f);
        const int flagged = a.iflags[b];
        const GImg info = a.info[b];
        if (I >= info.n_tiles || flagged) continue;
        const int M = info.M;
        const float t2 = info.t2;
        const float t3 = a.t3;
        const bool class_mode = info.mode == G_CLASS;

        // ---- the tile's local frame ----
        Frame f;
        {
            const float ext = fmaxf(ib.z - ib.x, ib.w - ib.y) * 0.5f;
            int k = a.hexp - ((int)((__float_as_uint(ext) >> 23) & 255u) - 127);   // ext * 2^k in [2^hexp, 2^(hexp+1))
            k = max(-60, min(60, k));
            f.S = __int_as_float((127 + k) << 23);
            f.S2 = f.S * f.S;
            f.nox = -((ib.x + ib.z) * 0.5f) * f.S;
            f.noy = -((ib.y + ib.w) * 0.5f) * f.S;
        }
        const u32 icmin = __float_as_uint(ia.z), icmax = __float_as_uint(ia.w);

        // ---- this lane's row of tile I in the local frame: packed record for the filter, registers for the row tests ----
        const bool rvalid = rp < M;
        u32 r_lo = 0x63d063d0u, r_hi = 0xe3d0e3d0u, r_wh_t = 0x63d063d0u, r_ar = 0xe3d063d0u;   // +-1000: passes nothing
        __syncwarp();
        {
            Rec16& r16 = w.row16[lane];
            uint4 box = make_uint4(0x63d063d0u, 0x63d063d0u, 0xe3d0e3d0u, 0xe3d0e3d0u);
            u32 area = 0x3c003c00u;
            if (rvalid) {
                const unsigned short x1 = h_dn(__fmaf_rd(rq.x, f.S, f.nox)), y1 = h_dn(__fmaf_rd(rq.y, f.S, f.noy));
                const unsigned short x2 = h_up(__fmaf_ru(rq.z, f.S, f.nox)), y2 = h_up(__fmaf_ru(rq.w, f.S, f.noy));
                const unsigned short ar = area16_down(rq, f.S2);
                const float wd = __fsub_rd(rq.z, rq.x), hd = __fsub_rd(rq.w, rq.y);
                const unsigned short wt = h_rz(__fmul_rd(__fmul_rd(wd, f.S), t3)), ht = h_rz(__fmul_rd(__fmul_rd(hd, f.S), t3));
                const unsigned short at = h_rz(__fmul_rd(__fmul_rd(__fmul_rd(wd, hd), f.S2), t3));
                const unsigned short au = h_up(__fmul_ru(__fmul_ru(__fsub_ru(rq.z, rq.x), __fsub_ru(rq.w, rq.y)), f.S2)) ^ 0x8000u;
                r_lo = pack2(x1, y1);
                r_hi = pack2(x2, y2);
                r_wh_t = pack2(wt, ht);
                r_ar = pack2(at, au);          // slot {amax up, -(t3*amin) down} >= {t3*area down, -(area up)}
                box = make_uint4(pack2(x1, x1), pack2(y1, y1), pack2(x2, x2), pack2(y2, y2));
                area = pack2(ar, ar);
            }
            r16.box = box;
            r16.area = area;
            if (lane == 0) {
                Rec16& sen = w.row16[kTile];
                sen.box = make_uint4(0x63d063d0u, 0x63d063d0u, 0xe3d0e3d0u, 0xe3d0e3d0u);
                sen.area = 0x3c003c00u;
            }
        }
        __syncwarp();
        u32 n_evals = 0, n_cands = 0;
        int n_q = 0;     // sub-tiles waiting in w.sub
        int n_buf = 0;   // edges waiting in w.ebuf
        int n_cand = 0;  // filter survivors waiting in w.cand

        // Tiles are walked chunk by chunk (one chunk at 640^2).  Inside a chunk the tiles are ordered by class in
        // per-class mode, so the walk enters a chunk at the first tile that can hold icmin (32-way probe) and
        // leaves it when every lane is past icmax.
        constexpr int kCT = kChunk / kTile;
        int chunk_end = min(info.n_tiles, (I / kCT + 1) * kCT);
        for (int J0 = I;; J0 += 32) {
            if (J0 >= chunk_end && chunk_end < info.n_tiles) {   // enter the next chunk
                J0 = chunk_end;
                chunk_end = min(info.n_tiles, chunk_end + kCT);
                if (class_mode) {
                    const int step = (chunk_end - J0 + 31) / 32;
                    const int Jp = J0 + lane * step;
                    const bool ge = Jp < chunk_end && __float_as_uint(ts[Jp * 2 + 1].w) >= icmin;   // class max: non-decreasing
                    const unsigned m = __ballot_sync(0xffffffffu, ge);
                    const int first = m ? __ffs(m) - 1 : 32;
                    J0 += max(first - 1, 0) * step;   // everything before the last probe below icmin is below icmin
                }
            }
            const bool tail = J0 >= info.n_tiles;  // one extra round flushes the last partial chunk
            int n_t = 0;
            int n_tail = kSlots;   // valid slots of the chunk about to be processed (< kSlots only in the tail)
            if (tail) {
                if (n_q == 0) break;
                if (lane >= n_q && lane < kSlots) w.sub[lane] = kNoSub;
                n_tail = n_q;
                n_q = kSlots;
                __syncwarp();
            } else {
                // level 1: lane <-> tile J
                const int Jl = J0 + lane;
                bool ok = Jl < chunk_end;
                float4 jb = make_float4(0.f, 0.f, 0.f, 0.f), ja = jb;
                if (ok) { jb = ts[Jl * 2]; ja = ts[Jl * 2 + 1]; }
                if (class_mode) {
                    const bool beyond = !ok || __float_as_uint(ja.z) > icmax;
                    ok = ok && !beyond && __float_as_uint(ja.w) >= icmin;
                    // the chunk's tiles are ordered by class: nothing beyond the last tile that can hold icmax
                    if (__ballot_sync(0xffffffffu, beyond) == 0xffffffffu) {
                        J0 = chunk_end - 32;  // next round: the next chunk, or the tail
                        continue;
                    }
                }
                if (ok) {
                    ok = fminf(ib.z, jb.z) > fmaxf(ib.x, jb.x) && fminf(ib.w, jb.w) > fmaxf(ib.y, jb.y) &&
                         ia.y >= t2 * ja.x && ja.y >= t2 * ia.x;
                }
                const unsigned cand = __ballot_sync(0xffffffffu, ok);
                if (cand == 0u) continue;
                if (ok) w.tl[__popc(cand & lt_mask)] = (u32)Jl;
                __syncwarp();
                n_t = __popc(cand);
            }
            int t0 = 0;
            do {
                if (t0 < n_t) {
                    // level 2: lane <-> (surviving tile, sub-tile), 8 tiles per step
                    const int k = t0 + lane / kSubs;
                    bool sok = k < n_t;
                    u32 e = 0u;
                    float4 sbx = make_float4(0.f, 0.f, 0.f, 0.f), sar = sbx;
                    if (sok) {
                        e = w.tl[k] * kSubs + (u32)(lane & (kSubs - 1));
                        sok = (int)(e * kSub) < M;
                        if (sok) {
                            sbx = ss[e * 2]; sar = ss[e * 2 + 1];
                            sok = fminf(ib.z, sbx.z) > fmaxf(ib.x, sbx.x) && fminf(ib.w, sbx.w) > fmaxf(ib.y, sbx.y) &&
                                  ia.y >= t2 * sar.x && sar.y >= t2 * ia.x;
                        }
                    }
                    const unsigned sm = __ballot_sync(0xffffffffu, sok);
                    if (sok) {
                        const int qi = n_q + __popc(sm & lt_mask);
                        w.sub[qi] = e;
                        // statistics in the local frame, rounded so that the row test can only pass more often
                        w.sbx[qi] = make_uint2(pack2(h_dn(__fmaf_rd(sbx.x, f.S, f.nox)), h_dn(__fmaf_rd(sbx.y, f.S, f.noy))),
                                               pack2(h_up(__fmaf_ru(sbx.z, f.S, f.nox)), h_up(__fmaf_ru(sbx.w, f.S, f.noy))));
                        const unsigned short amax = h_up(__fmul_ru(__fmul_ru(sar.y, 1.0000077f), f.S2));
                        const unsigned short amin = h_rz(__fmul_rd(__fmul_rd(__fmul_rd(sar.x, 0.9999923f), f.S2), t3));
                        w.sar[qi] = pack2(amax, amin ^ 0x8000u);
                    }
                    n_q += __popc(sm);
                    __syncwarp();
                }
                // level 3: chunks of 8 sub-tiles
                int head = 0;
                if (n_tail == kSlots) {
                    for (; n_q - head >= kSlots; head += kSlots)
                        edge_chunk<true>(a, w, M, class_mode, b, I, head, sb, f, kSlots, r_lo, r_hi, r_wh_t, r_ar, n_evals,
                                         n_cands, n_cand, n_buf);
                } else {   // the last, partial chunk of the row tile
                    edge_chunk<false>(a, w, M, class_mode, b, I, 0, sb, f, n_tail, r_lo, r_hi, r_wh_t, r_ar, n_evals, n_cands,
                                      n_cand, n_buf);
                    head = kSlots;
                }
                if (head) {  // move the < 8 leftovers to the front
                    const int rem = n_q - head;
                    u32 v = 0u, va = 0u;
                    uint2 vb = make_uint2(0u, 0u);
                    if (lane < rem) { v = w.sub[head + lane]; vb = w.sbx[head + lane]; va = w.sar[head + lane]; }
                    __syncwarp();
                    if (lane < rem) { w.sub[lane] = v; w.sbx[lane] = vb; w.sar[lane] = va; }
                    __syncwarp();
                    n_q = rem;
                }
                t0 += 32 / kSubs;
            } while (t0 < n_t);
        }
        if (n_cand) {
            n_cands += (u32)n_cand;
            cand_process(a, w, M, class_mode, b, I, n_cand, n_buf);
        }
        edge_flush(a, w, b, a.edges + (u64)b * a.edges_per_img, n_buf);
        if (lane == 0 && n_evals) {
            atomicAdd(&a.info[b].n_evals, (unsigned long long)n_evals);
            atomicAdd(&a.info[b].n_cands, (unsigned long long)n_cands);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Fallback inside the resolve kernel: an image whose suppression graph does not fit the edge list
// (heavily clustered boxes: a trained detector at a low confidence threshold, thousands of copies of
// one box) or whose class ids exceed the sort key is resolved right here, on the device, by the
// blocked greedy algorithm: ranks are walked in blocks of 512; a block is first tested against every
// box kept so far, then against itself through a 512x512 bitmask with a serial scan by one warp.
// Work is M x K pair tests (K = kept boxes), small exactly when the graph is dense.  torchvision's
// exact fma/div arithmetic with the higher-scored box as `a`; no workspace beyond the caller's
// keep row.  The host never sees n_keep = -1 from the graph algorithm any more.
// ------------------------------------------------------------------------------------------------
constexpr int kGreedyT = 512;
struct GreedySmem {
    float4 blk[kGreedyT];        // boxes of the current block of ranks (offset trick applied)
    float4 kt[kGreedyT];         // a tile of already kept boxes
    long long bcls[kGreedyT];
    long long kcls[kGreedyT];
    u32 mask[kGreedyT * (kGreedyT / 32)];
    u32 dead[kGreedyT / 32];
    u32 keptbits[kGreedyT / 32];
    u32 K;
};

__device__ __forceinline__ bool iou_gt_exact(const float4 a, const float4 b, const float thr) {
    // torchvision devIoU (sm_100 SASS): den = fma(bw, bh, Sa) - inter, IEEE division, a = higher score
    const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
    const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
    const float w = fmaxf(right - left, 0.0f), h = fmaxf(bottom - top, 0.0f);
    const float inter = w * h;
    const float sa = (a.z - a.x) * (a.w - a.y);
    const float den = __fmaf_rn(b.z - b.x, b.w - b.y, sa) - inter;
    return (inter / den) > thr;
}

__device__ void greedy_resolve(const GArgs& a, const GImg& info, int b, GreedySmem& g) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = tid & (kGreedyT - 1), h = tid / kGreedyT;       // kResolveThreads == 2 * kGreedyT
    const int M = info.M;
    const size_t off = (size_t)b * a.cap;
    const float4* boxes = a.boxes + off;
    const int64_t* classes = a.classes ? a.classes + off : nullptr;
    const u32* __restrict__ order = a.order + off;
    int64_t* keep = a.keep + off;
    const float thr = a.thr;
    const bool trick = info.mode == G_TRICK, per_class = info.mode == G_CLASS;
    const float far = 3.0e38f;
    constexpr int W = kGreedyT / 32, HALF = kGreedyT / 2;
    if (tid == 0) g.K = 0u;
    __syncthreads();
    for (int r0 = 0; r0 < M; r0 += kGreedyT) {
        const int n = min(kGreedyT, M - r0);
        if (tid < kGreedyT) {
            float4 q = make_float4(far, far, far, far);
            long long c = 0;
            if (tid < n) {
                const u32 idx = order[r0 + tid];
                q = boxes[idx];
                if (classes) c = classes[idx];
                if (trick) {
                    const float o = (float)c * info.s_off;
                    q.x += o; q.y += o; q.z += o; q.w += o;
                }
            }
            g.blk[tid] = q;
            g.bcls[tid] = c;
        }
        if (tid < W) {
            const int rem = n - tid * 32;
            g.dead[tid] = rem >= 32 ? 0u : (rem <= 0 ? 0xffffffffu : ~((1u << rem) - 1u));
        }
        __syncthreads();
        const float4 q = g.blk[i];
        const long long qc = g.bcls[i];
        bool dead = i >= n;
        // phase 1: against everything kept so far (each half of the CTA takes half of every tile)
        const int K = (int)g.K;
        for (int k0 = 0; k0 < K; k0 += kGreedyT) {
            const int nk = min(kGreedyT, K - k0);
            if (tid < nk) {
                const int64_t kidx = keep[k0 + tid];
                float4 kq = boxes[kidx];
                long long c = 0;
                if (classes) c = classes[kidx];
                if (trick) {
                    const float o = (float)c * info.s_off;
                    kq.x += o; kq.y += o; kq.z += o; kq.w += o;
                }
                g.kt[tid] = kq;
                g.kcls[tid] = c;
            }
            __syncthreads();
            if (!dead) {
                const int j1 = min(nk, h * HALF + HALF);
                for (int j = h * HALF; j < j1; ++j) {
                    if ((!per_class || g.kcls[j] == qc) && iou_gt_exact(g.kt[j], q, thr)) { dead = true; break; }
                }
            }
            __syncthreads();
        }
        if (dead && i < n) atomicOr(&g.dead[i >> 5], 1u << (i & 31));
        __syncthreads();
        dead = (g.dead[i >> 5] >> (i & 31)) & 1u;
        // phase 2: row i of the block's bitmask, columns j > i (this half's 8 words)
#pragma unroll 1
        for (int w = 0; w < W / 2; ++w) {
            const int wi = h * (W / 2) + w;
            u32 word = 0u;
            if (!dead && wi * 32 + 31 > i) {
                for (int bit = 0; bit < 32; ++bit) {
                    const int j = wi * 32 + bit;
                    if (j > i && j < n && (!per_class || g.bcls[j] == qc) && iou_gt_exact(q, g.blk[j], thr)) word |= 1u << bit;
                }
            }
            g.mask[i * W + wi] = word;
        }
        __syncthreads();
        if (warp == 0) {  // serial scan in rank order; lanes < W hold the removed words
            u32 rem = lane < W ? g.dead[lane] : 0xffffffffu;
            u32 kb = 0u;
            for (int r = 0; r < n; ++r) {
                const u32 rw = __shfl_sync(0xffffffffu, rem, r >> 5);
                if (!((rw >> (r & 31)) & 1u)) {
                    if (lane == (r >> 5)) kb |= 1u << (r & 31);
                    if (lane < W) rem |= g.mask[r * W + lane];
                }
            }
            if (lane < W) g.keptbits[lane] = kb;
        }
        __syncthreads();
        u32 total = 0, before = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const u32 c = __popc(g.keptbits[w]);
            if (w < (i >> 5)) before += c;
            total += c;
        }
        if (tid < n) {
            const u32 bits = g.keptbits[tid >> 5];
            if ((bits >> (tid & 31)) & 1u)
                keep[g.K + before + __popc(bits & ((1u << (tid & 31)) - 1u))] = (int64_t)order[r0 + tid];
        }
        __syncthreads();
        if (tid == 0) g.K += total;
        __syncthreads();
    }
    if (tid == 0) a.n_keep[b] = (int)g.K;
}

// ------------------------------------------------------------------------------------------------
// fixed-point resolve + emit
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kResolveThreads) graph_resolve_kernel(const GArgs a) {
    extern __shared__ __align__(16) u32 s_bits[];  // U | K | fK | fU, each nw words (or a GreedySmem)
    __shared__ u32 s_scan[32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const GImg info = a.info[b];
    const int M = info.M;
    if (M == 0) {
        if (tid == 0) a.n_keep[b] = 0;
        return;
    }
    if (a.iflags[b] || (u64)info.n_edges > a.edges_per_img) {
        greedy_resolve(a, info, b, *reinterpret_cast<GreedySmem*>(s_bits));
        return;
    }
    const int nw = (M + 31) >> 5;
    const int nw_cap = (a.cap + 31) >> 5;
    u32 *U = s_bits, *K = s_bits + nw_cap, *fK = s_bits + 2 * nw_cap, *fU = s_bits + 3 * nw_cap;
    for (int w = tid; w < nw; w += kResolveThreads) {
        const int rem = M - w * 32;
        U[w] = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
        K[w] = 0u; fK[w] = 0u; fU[w] = 0u;
    }
    uint2* gedges = a.edges + (u64)b * a.edges_per_img;
    const u32 E = info.n_edges;
    // The edge kernel recorded ORIGINAL indices (it does not wait for the score sort): to score ranks, once — and,
    // when the list fits, into shared memory (two 16-bit ranks per word when M <= 65536: 37 K edges per image at
    // 640^2), so that the rounds below never touch global memory.
    u32* sedges = s_bits + 4 * nw_cap;
    const bool pack16 = M <= 65536 && E <= a.resolve_smem_words;
    const bool in_smem = pack16 || 2u * E <= a.resolve_smem_words;
    {
        const u32* __restrict__ rinv = a.rinv + (size_t)b * a.cap;
        // four edges per thread and step: the loads of a step are independent (12 in flight per thread)
        for (u32 e0 = tid; e0 < E; e0 += 4 * kResolveThreads) {
            uint2 sd[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const u32 e = e0 + k * kResolveThreads;
                sd[k] = e < E ? gedges[e] : make_uint2(0u, 0u);
            }
            u32 rx[4], ry[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { rx[k] = rinv[sd[k].x]; ry[k] = rinv[sd[k].y]; }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const u32 e = e0 + k * kResolveThreads;
                if (e < E) {
                    if (pack16) sedges[e] = (rx[k] << 16) | ry[k];
                    else if (in_smem) reinterpret_cast<uint2*>(sedges)[e] = make_uint2(rx[k], ry[k]);
                    else gedges[e] = make_uint2(rx[k], ry[k]);
                }
            }
        }
    }
    __syncthreads();
    // Every thread owns a contiguous segment of the list and compacts it in place: an edge whose destination has
    // been decided, or whose source has been suppressed, can never matter again (odd segment length: the strided
    // shared-memory accesses of a warp fall into 32 different banks).
    const u32 per = ((E + kResolveThreads - 1) / kResolveThreads) | 1u;
    const u32 my_beg = min((u32)tid * per, E);
    u32 my_n = min(per, E - my_beg);
    auto load_edge = [&](u32 e) -> uint2 {
        if (pack16) { const u32 v = sedges[e]; return make_uint2(v >> 16, v & 0xffffu); }
        return in_smem ? reinterpret_cast<const uint2*>(sedges)[e] : gedges[e];
    };
    auto store_edge = [&](u32 e, uint2 v) {
        if (pack16) sedges[e] = (v.x << 16) | v.y;
        else if (in_smem) reinterpret_cast<uint2*>(sedges)[e] = v;
        else gedges[e] = v;
    };
    for (;;) {
        u32 live = 0;
        for (u32 k = 0; k < my_n; ++k) {
            const uint2 sd = load_edge(my_beg + k);  // x = higher-scored (lower rank), y = lower-scored
            const u32 dw = sd.y >> 5, dbit = 1u << (sd.y & 31);
            if (U[dw] & dbit) {
                const u32 sw = sd.x >> 5, sbit = 1u << (sd.x & 31);
                if (K[sw] & sbit) atomicOr(&fK[dw], dbit);           // destination suppressed this round: edge done
                else if (U[sw] & sbit) {
                    atomicOr(&fU[dw], dbit);
                    if (live != k) store_edge(my_beg + live, sd);
                    ++live;
                }
            }
        }
        my_n = live;
        __syncthreads();
        int left = 0;
        for (int w = tid; w < nw; w += kResolveThreads) {
            const u32 u = U[w];
            const u32 sup = u & fK[w];
            const u32 kept = u & ~fK[w] & ~fU[w];
            const u32 nu = u & ~(sup | kept);
            U[w] = nu;
            K[w] |= kept;
            fK[w] = 0u; fU[w] = 0u;
            left |= nu != 0u;
        }
        if (!__syncthreads_or(left)) break;
    }
    // ---- emit kept indices in descending score order (K is indexed by score rank) ----
    // word prefix sums (fK reused), then one thread per rank: coalesced reads of `order`, near-coalesced writes
    const u32* __restrict__ order = a.order + (size_t)b * a.cap;
    int64_t* __restrict__ keep = a.keep + (size_t)b * a.cap;
    const int chunk = (nw + kResolveThreads - 1) / kResolveThreads;
    const int c0 = min(tid * chunk, nw), c1 = min(c0 + chunk, nw);
    u32 sum = 0;
    for (int w = c0; w < c1; ++w) sum += __popc(K[w]);
    u32 inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) s_scan[warp] = inc;
    __syncthreads();
    u32 base = inc - sum, total = 0;
    for (int w = 0; w < 32; ++w) {
        if (w < warp) base += s_scan[w];
        total += s_scan[w];
    }
    for (int w = c0; w < c1; ++w) {
        fK[w] = base;
        base += __popc(K[w]);
    }
    __syncthreads();
    for (int r0 = tid; r0 < M; r0 += 4 * kResolveThreads) {   // four independent loads of `order` per step
        u32 idx[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int r = r0 + k * kResolveThreads;
            idx[k] = r < M ? order[r] : 0u;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int r = r0 + k * kResolveThreads;
            if (r < M) {
                const u32 bits = K[r >> 5];
                if ((bits >> (r & 31)) & 1u) keep[fK[r >> 5] + __popc(bits & ((1u << (r & 31)) - 1u))] = (int64_t)idx[k];
            }
        }
    }
    if (tid == 0) a.n_keep[b] = (int)total;
}

// ---- host side -------------------------------------------------------------------------------
struct GLayout {
    size_t k[2], order, rinv, pos, sboxes, skey, scls, tstat, sstat, info, iflags, ticket, edges, total;
};

static inline size_t g_align(size_t x) { return (x + 255) / 256 * 256; }

static GLayout graph_layout(int B, int cap) {
    GLayout L;
    size_t o = 0;
    const size_t n = (size_t)B * cap;
    const size_t tcap = ((size_t)cap + kTile - 1) / kTile;
    const size_t nk = cap > kS16MaxM ? n : 0;   // sorted runs: only images larger than one shared-memory sort
    for (int i = 0; i < 2; ++i) { L.k[i] = o; o = g_align(o + nk * 4); }
    L.order = o; o = g_align(o + n * 4);
    L.rinv = o; o = g_align(o + n * 4);
    L.pos = o; o = g_align(o + n * 4);
    L.sboxes = o; o = g_align(o + (size_t)B * (((size_t)cap + kSub - 1) / kSub * kSub) * 16);
    L.skey = o; o = g_align(o + n * 4);
    L.scls = o; o = g_align(o + n * 4);
    L.tstat = o; o = g_align(o + (size_t)B * tcap * 32);
    L.sstat = o; o = g_align(o + (size_t)B * tcap * kSubs * 32);
    L.info = o; o = g_align(o + (size_t)B * sizeof(GImg));
    L.iflags = o; o = g_align(o + (size_t)B * sizeof(int));
    L.ticket = o; o = g_align(o + 256);
    L.edges = o;
    L.total = o;
    return L;
}

size_t graph_min_workspace(int B, int cap) { return graph_layout(B, cap).total + (size_t)B * 8; }

// The score sort only feeds the resolve kernel, so it runs on a side stream between two events (fork after the
// caller's previous work, join before the resolve kernel) and overlaps with the spatial kernel and the edge
// discovery; under stream capture the fork/join becomes two parallel branches of the graph.  One side stream and a
// ring of event pairs per device, created on first use (never during a capture: a warm-up call comes first).
struct ForkJoin {
    cudaStream_t side = nullptr;
    static constexpr int kRing = 32;
    cudaEvent_t fork[kRing] = {}, join[kRing] = {};
    std::atomic<unsigned> next{0};
    bool ok = false;
};
static ForkJoin* fork_join_for_device() {
    static std::mutex mu;
    static ForkJoin* per_dev[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    if (!per_dev[dev]) {
        ForkJoin* f = new ForkJoin();
        bool ok = cudaStreamCreateWithFlags(&f->side, cudaStreamNonBlocking) == cudaSuccess;
        for (int i = 0; ok && i < ForkJoin::kRing; ++i)
            ok = cudaEventCreateWithFlags(&f->fork[i], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&f->join[i], cudaEventDisableTiming) == cudaSuccess;
        f->ok = ok;
        if (!ok) cudaGetLastError();
        per_dev[dev] = f;
    }
    return per_dev[dev]->ok ? per_dev[dev] : nullptr;
}

int graph_nms(const float* boxes, const float* scores, const int64_t* classes, const int* counts, int B, int cap,
              double iou_threshold, long long trick_max_numel, int64_t* keep, int* n_keep, void* ws,
              size_t ws_bytes, cudaStream_t st) {
    GLayout L = graph_layout(B, cap);
    YB_CHECK_ARG(ws_bytes >= L.total + (size_t)B * 8, "nms(graph): workspace too small (%zu < %zu)", ws_bytes,
                 L.total + (size_t)B * 8);
    char* w = reinterpret_cast<char*>(ws);
    GArgs a;
    a.boxes = reinterpret_cast<const float4*>(boxes); a.scores = scores; a.classes = classes; a.counts = counts;
    a.B = B; a.cap = cap; a.tcap = (cap + kTile - 1) / kTile; a.scap = (cap + kSub - 1) / kSub * kSub;
    a.n_chunks = (cap + kChunk - 1) / kChunk;
    a.thr = (float)iou_threshold;
    a.thr_fast_ok = (a.thr >= 0.03f && a.thr <= 1e30f) ? 1 : 0;
    a.t3 = a.thr * (1.0f - 0.015625f);
    {
        // K = (1+th)/th with 2^-6 of slack, rounded up into a half; the filter keeps a pair iff K*inter >= area_a + area_b
        const double th = (double)a.thr * (1.0 - 1.0 / 4096.0);
        const double K = a.thr_fast_ok ? (1.0 + th) / th * (1.0 + 1.0 / 64.0) * (1.0 + 1.0 / 512.0) : 1.0;
        const unsigned short kb = __half_as_ushort(__float2half_rn((float)K));   // rn error 2^-11 < the 2^-9 bias above
        a.k16x2 = (u32)kb | ((u32)kb << 16);
        a.hexp = a.thr >= 0.35f ? 5 : (a.thr >= 0.1f ? 4 : 3);   // keeps K * (row area) inside the half range
    }
    a.trick_max_numel = trick_max_numel;
    a.k0 = (u32*)(w + L.k[0]); a.v0 = (u32*)(w + L.k[1]);
    a.n_score_chunks = (cap + kScoreChunk - 1) / kScoreChunk;
    a.order = (u32*)(w + L.order); a.rinv = (u32*)(w + L.rinv); a.pos = (u32*)(w + L.pos);
    a.sboxes = (float4*)(w + L.sboxes); a.skey = (u32*)(w + L.skey); a.scls = (u32*)(w + L.scls);
    a.tstat = (float4*)(w + L.tstat);
    a.sstat = (float4*)(w + L.sstat);
    a.info = (GImg*)(w + L.info);
    a.iflags = (int*)(w + L.iflags);
    a.ticket = (u32*)(w + L.ticket);
    a.ticket_init = (u32)(sm_count() * kEdgeCtasPerSM * (kEdgeThreads / 32));
    a.edges = (uint2*)(w + L.edges);
    a.edges_per_img = (ws_bytes - L.edges) / 8 / (size_t)B;
    a.keep = keep; a.n_keep = n_keep;

    // ---- fork: score order on the side stream ----
    ForkJoin* fj = fork_join_for_device();
    cudaStream_t ss = st;
    int slot = 0;
    if (fj) {
        slot = (int)(fj->next.fetch_add(1u) % ForkJoin::kRing);
        YB_CUDA(cudaEventRecord(fj->fork[slot], st));
        YB_CUDA(cudaStreamWaitEvent(fj->side, fj->fork[slot], 0));
        ss = fj->side;
    }
    {
        const size_t sort_smem = s16_smem_bytes(cap < kScoreChunk ? cap : kScoreChunk);
        YB_CUDA(cudaFuncSetAttribute(graph_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem));
        YB_LAUNCH("graph_score_kernel", ss,
                  (graph_score_kernel<<<dim3(a.n_score_chunks, B), kS16Threads, sort_smem, ss>>>(a)));
        if (a.n_score_chunks > 1)
            YB_LAUNCH("graph_score_merge_kernel", ss,
                      (graph_score_merge_kernel<<<dim3((cap + 255) / 256, B), 256, 0, ss>>>(a)));
    }
    if (fj) YB_CUDA(cudaEventRecord(fj->join[slot], fj->side));

    // ---- main branch: spatial order + gather, edge discovery ----
    if (a.n_chunks > 1) YB_CUDA(cudaMemsetAsync(a.iflags, 0, (size_t)B * sizeof(int), st));
    {
        const size_t smem = spatial_smem_bytes(cap < kChunk ? cap : kChunk);
        YB_CUDA(cudaFuncSetAttribute(graph_spatial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        YB_LAUNCH("graph_spatial_kernel", st, graph_spatial_kernel<<<dim3(a.n_chunks, B), kSpThreads, smem, st>>>(a));
    }
    const dim3 ggrid((a.tcap + kGatherThreads / 32 - 1) / (kGatherThreads / 32), B);
    YB_LAUNCH("graph_gather_kernel", st, graph_gather_kernel<<<ggrid, kGatherThreads, 0, st>>>(a));
    const int ctas = sm_count() * kEdgeCtasPerSM;  // resident CTAs per SM (launch bounds)
    // (a.ticket_init was set before the spatial kernel's launch: it writes the counter)
    YB_LAUNCH("graph_edge_kernel", st, graph_edge_kernel<<<ctas, kEdgeThreads, 0, st>>>(a));

    // ---- join, resolve ----
    if (fj) YB_CUDA(cudaStreamWaitEvent(st, fj->join[slot], 0));
    static_assert(kResolveThreads == 2 * kGreedyT, "greedy_resolve splits the CTA in two halves");
    const size_t bits_bytes = ((size_t)((cap + 31) / 32) * 4 + 2) * 4;
    YB_CHECK_ARG(bits_bytes <= 150 * 1024, "nms(graph): cap too large for the resolve kernel");
    size_t dyn = 160 * 1024;                      // bitmaps + as many edges as fit
    a.resolve_smem_words = (u32)((dyn - bits_bytes) / 4);
    if (dyn < sizeof(GreedySmem)) dyn = sizeof(GreedySmem);
    YB_CUDA(cudaFuncSetAttribute(graph_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    YB_LAUNCH("graph_resolve_kernel", st, graph_resolve_kernel<<<B, kResolveThreads, dyn, st>>>(a));
    return 0;
}

int graph_stats(const void* ws, size_t ws_bytes, int B, int cap, unsigned long long* evals, unsigned long long* edges,
                unsigned long long* cands, cudaStream_t st) {
    GLayout L = graph_layout(B, cap);
    YB_CHECK_ARG(ws && ws_bytes >= L.total, "nms_graph_stats: not a graph workspace");
    std::vector<GImg> h((size_t)B);
    YB_CUDA(cudaMemcpyAsync(h.data(), reinterpret_cast<const char*>(ws) + L.info, (size_t)B * sizeof(GImg),
                            cudaMemcpyDeviceToHost, st));
    YB_CUDA(cudaStreamSynchronize(st));
    unsigned long long ev = 0, ed = 0, ca = 0;
    for (auto& g : h) { ev += g.n_evals * (unsigned long long)kSub; ed += g.n_edges; ca += g.n_cands; }  // items -> pair tests
    if (evals) *evals = ev;
    if (edges) *edges = ed;
    if (cands) *cands = ca;
    return 0;
}

}  // namespace yb
