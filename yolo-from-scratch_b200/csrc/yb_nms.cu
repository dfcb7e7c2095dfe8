// yb_nms.cu — K5: batched, cross-scale global NMS with torchvision's exact CUDA arithmetic.
// Reference: torchvision.ops.batched_nms as called by predict(), train.py:1232-1233
// (torchvision 0.26.0: ops/boxes.py:51-120 and csrc/ops/cuda/nms_kernel.cu).
//
// Three launches for a whole batch of images:
//   nms_sort_kernel  one CTA per image: validity/max-coordinate reduction, stable descending
//                    radix sort by score (LSD, 8-bit digits, per-warp histograms so that no
//                    atomics are needed and stability is by construction), optional stable
//                    partition by class (per-class mode), gather of the sorted (offset) boxes,
//                    segment ends and compact mask-row offsets.
//   nms_mask_kernel  blocked IoU bitmask.  CTA = 64 sorted rows staged in shared memory x all
//                    column tiles of the rows' segment; lane = column, 8 columns per lane held in
//                    registers, one warp ballot per 32 pairs builds the suppression words.
//                    IoU > thr is decided without a division: |inter - thr*den| is compared with
//                    a 2^-21 relative margin and only ambiguous groups are recomputed with the
//                    IEEE division torchvision uses, so the bits are identical.
//   nms_scan_kernel  one CTA per image: greedy scan over 64-box groups (serial resolve of the
//                    diagonal word block in registers, CTA-wide OR of the kept rows), then the
//                    kept set is emitted in descending score order.
// Bound: fp32 SIMT issue (IoU pairs/s); tensor cores do not apply.
#include "yb_common.cuh"
#include "yb_sort.cuh"

namespace yb {

constexpr int kScanThreads = 1024;
constexpr int kMaskThreads = 128;
constexpr int kGroupTiles = 4;  // column tiles (of 64) per warp iteration

enum { MODE_PLAIN = 0, MODE_TRICK = 1, MODE_CLASS = 2 };

struct NmsImg {
    int M;          // candidates
    int mode;
    int exact;      // 1: boxes not provably "nice" -> exact division path
    int overflow;   // mask rows do not fit the workspace
    u32 words;      // total mask words of this image
    float s_off;    // max coordinate + 1 (trick mode)
    int pad[2];
};

struct NmsArgs {
    const float4* boxes;
    const float* scores;
    const int64_t* classes;
    const int* counts;
    int B, cap;
    float thr;
    int thr_fast_ok;
    long long trick_max_numel;
    // workspace
    u32 *k0, *v0, *k1, *v1, *k2, *v2;  // (B,cap) each; v0 ends as rank -> original index
    float4* sboxes;                    // (B,cap) boxes in position order (offset applied)
    u32* pos_rank;                     // (B,cap) position -> score rank (class mode)
    u32* seg_end;                      // (B,cap) position -> end of its segment
    u32* rowoff;                       // (B,cap) position -> first mask word of its row
    NmsImg* info;                      // (B)
    u64* mask;                         // B * mask_words_per_img
    u64 mask_words_per_img;
    int64_t* keep;
    int* n_keep;
};

__global__ void __launch_bounds__(kSortThreads) nms_sort_kernel(const NmsArgs a) {
    __shared__ u32 s_hist[32 * 256];
    __shared__ u32 s_tot[256];
    __shared__ float s_fmax[32];
    __shared__ int s_flag[32];
    __shared__ int s_maxcls[32];
    __shared__ u32 s_scan[32];
    __shared__ u32 s_first[32];
    __shared__ int s_skip;

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t off = (size_t)b * a.cap;
    int M = a.counts ? a.counts[b] : a.cap;
    M = M < 0 ? 0 : (M > a.cap ? a.cap : M);
    NmsImg* info = a.info + b;
    if (M == 0) {
        if (tid == 0) {
            NmsImg z = {0, MODE_PLAIN, 0, 0, 0u, 0.0f, {0, 0}};
            *info = z;
        }
        return;
    }
    const float4* boxes = a.boxes + off;
    const float* scores = a.scores + off;
    const int64_t* classes = a.classes ? a.classes + off : nullptr;
    const int mode = !classes ? MODE_PLAIN
                              : ((long long)M * 4 <= a.trick_max_numel ? MODE_TRICK : MODE_CLASS);

    // ---- reductions: max coordinate (boxes.max(), boxes.py:98), niceness, max class id ----
    float mx = -INFINITY;
    int bad = 0, has_nan = 0, maxcls = 0;
    for (int i = tid; i < M; i += kSortThreads) {
        const float4 q = boxes[i];
        mx = fmaxf(mx, fmaxf(fmaxf(q.x, q.y), fmaxf(q.z, q.w)));
        const bool nan = (q.x != q.x) || (q.y != q.y) || (q.z != q.z) || (q.w != q.w);
        has_nan |= nan;
        const float big = fmaxf(fmaxf(fabsf(q.x), fabsf(q.y)), fmaxf(fabsf(q.z), fabsf(q.w)));
        bad |= nan || !(big <= 1e17f) || !(q.z >= q.x) || !(q.w >= q.y);
        if (classes) {
            const long long c = classes[i];
            bad |= (c < 0 || c >= 65536) ? 2 : 0;
            maxcls = max(maxcls, (int)(c & 0xffff));
        }
    }
    mx = warp_max(mx);
    bad = __reduce_or_sync(0xffffffffu, bad | (has_nan ? 4 : 0));
    maxcls = __reduce_max_sync(0xffffffffu, maxcls);
    if (lane == 0) { s_fmax[warp] = mx; s_flag[warp] = bad; s_maxcls[warp] = maxcls; }
    __syncthreads();
    mx = s_fmax[0]; bad = 0; maxcls = 0;
    for (int w = 0; w < 32; ++w) {
        mx = fmaxf(mx, s_fmax[w]);
        bad |= s_flag[w];
        maxcls = max(maxcls, s_maxcls[w]);
    }
    float s_off = mx + 1.0f;                       // boxes.py:99
    if (bad & 4) s_off = __int_as_float(0x7fc00000);  // torch max propagates NaN
    if (mode == MODE_TRICK && !((float)maxcls * s_off + fabsf(mx) <= 1e17f)) bad |= 1;

    // ---- stable descending sort by score: 4 LSD passes ----
    u32 *k0 = a.k0 + off, *v0 = a.v0 + off, *k1 = a.k1 + off, *v1 = a.v1 + off;
    u32 *k2 = a.k2 + off, *v2 = a.v2 + off;
    for (int i = tid; i < M; i += kSortThreads) { k0[i] = desc_key(scores[i]); v0[i] = (u32)i; }
    __syncthreads();
    {
        u32 *ka = k0, *va = v0, *kb = k1, *vb = v1;
        radix_sort(ka, va, kb, vb, M, 0, 32, s_hist, s_tot, &s_skip);
        if (va != v0) {  // keep the documented location: v0[r] = original index of score rank r
            for (int i = tid; i < M; i += kSortThreads) v0[i] = va[i];
            __syncthreads();
        }
    }

    // ---- per-class mode: stable partition of the ranks by class id ----
    const u32* cls_sorted = nullptr;
    const u32* prank = nullptr;
    if (mode == MODE_CLASS) {
        for (int i = tid; i < M; i += kSortThreads) { k1[i] = (u32)classes[v0[i]] & 0xffffu; v1[i] = (u32)i; }
        __syncthreads();
        u32 *ka = k1, *va = v1, *kb = k2, *vb = v2;
        radix_sort(ka, va, kb, vb, M, 0, maxcls >= 256 ? 16 : 8, s_hist, s_tot, &s_skip);
        cls_sorted = ka; prank = va;
    }

    // ---- gather boxes in position order; coordinate-offset trick (boxes.py:99-101) ----
    float4* sb = a.sboxes + off;
    u32* pos_rank = a.pos_rank + off;
    u32* seg_end = a.seg_end + off;
    for (int p = tid; p < M; p += kSortThreads) {
        const u32 r = prank ? prank[p] : (u32)p;
        const u32 idx = v0[r];
        float4 q = boxes[idx];
        if (mode == MODE_TRICK) {
            const float o = (float)classes[idx] * s_off;
            q.x += o; q.y += o; q.z += o; q.w += o;
        }
        sb[p] = q;
        pos_rank[p] = r;
        if (mode != MODE_CLASS) seg_end[p] = (u32)M;
    }

    // ---- per-class mode: end of each class run ----
    if (mode == MODE_CLASS) {
        const int per = (((M + 31) / 32) + 31) & ~31;
        const int beg = warp * per, end = min(beg + per, M);
        // first run end inside each warp's range
        u32 first = 0xffffffffu;
        for (int i0 = beg; i0 < end && first == 0xffffffffu; i0 += 32) {
            const int p = i0 + lane;
            const bool last = p < end && (p == M - 1 || cls_sorted[p] != cls_sorted[p + 1]);
            const unsigned bal = __ballot_sync(0xffffffffu, last);
            if (bal) first = (u32)(i0 + __ffs(bal));  // end is exclusive: position + 1
        }
        if (lane == 0) s_first[warp] = first;
        __syncthreads();
        u32 carry = (u32)M;
        for (int w = 31; w > warp; --w)
            if (s_first[w] != 0xffffffffu) carry = s_first[w];
        // walk the range right to left
        const int nit = end > beg ? (end - beg + 31) / 32 : 0;
        for (int it = nit - 1; it >= 0; --it) {
            const int p = beg + it * 32 + lane;
            const bool last = p < end && (p == M - 1 || cls_sorted[p] != cls_sorted[p + 1]);
            const unsigned bal = __ballot_sync(0xffffffffu, last);
            const unsigned at_or_after = bal >> lane;
            if (p < end) seg_end[p] = at_or_after ? (u32)(p + __ffs(at_or_after)) : carry;
            if (bal) carry = (u32)(beg + it * 32 + __ffs(bal));
        }
    }
    __syncthreads();

    // ---- compact mask rows: row p holds words [p/64, (seg_end-1)/64], padded to 4 words ----
    u32* rowoff = a.rowoff + off;
    const int chunk = (M + kSortThreads - 1) / kSortThreads;
    const int c0 = min(tid * chunk, M), c1 = min(c0 + chunk, M);
    u32 sum = 0;
    for (int p = c0; p < c1; ++p) {
        const u32 w = ((seg_end[p] - 1) >> 6) - ((u32)p >> 6) + 1;
        sum += (w + 3u) & ~3u;
    }
    u32 inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) s_scan[warp] = inc;
    __syncthreads();
    u32 base = inc - sum;
    u64 total = 0;
    for (int w = 0; w < 32; ++w) {
        if (w < warp) base += s_scan[w];
        total += s_scan[w];
    }
    for (int p = c0; p < c1; ++p) {
        rowoff[p] = base;
        const u32 w = ((seg_end[p] - 1) >> 6) - ((u32)p >> 6) + 1;
        base += (w + 3u) & ~3u;
    }
    if (tid == 0) {
        NmsImg o;
        o.M = M; o.mode = mode;
        o.exact = (bad & 1) || (bad & 4) || !a.thr_fast_ok;
        o.overflow = (total > a.mask_words_per_img) || (bad & 2);
        o.words = (u32)total; o.s_off = s_off; o.pad[0] = o.pad[1] = 0;
        *info = o;
    }
}

// ------------------------------------------------------------------------------------------
// IoU predicate.  a = row box (higher score), b = column box.  torchvision devIoU, fp32:
//   inter = max(0, min(a.z,b.z) - max(a.x,b.x)) * max(0, min(a.w,b.w) - max(a.y,b.y))
//   den   = fma(b.z-b.x, b.w-b.y, (a.z-a.x)*(a.w-a.y)) - inter
//   suppress = inter / den > thr
// ------------------------------------------------------------------------------------------
struct ColBox { float x1, y1, x2, y2, w, h; };

__device__ __forceinline__ void iou_terms(const float4& r, float sa, const ColBox& c, float& inter, float& den) {
    const float left = fmaxf(r.x, c.x1), right = fminf(r.z, c.x2);
    const float top = fmaxf(r.y, c.y1), bottom = fminf(r.w, c.y2);
    const float w = fmaxf(right - left, 0.0f), h = fmaxf(bottom - top, 0.0f);
    inter = w * h;
    den = __fmaf_rn(c.w, c.h, sa) - inter;
}

template <bool EXACT>
__device__ __forceinline__ void mask_group(const float4* s_row, const float* s_area, const ColBox (&col)[kGroupTiles][2],
                                           float thr, int lane, u64 (&words)[2][kGroupTiles], bool& amb) {
    const float kEps = 4.76837158203125e-07f;    // 2^-21
    const float kTiny = 7.888609052210118e-31f;  // 2^-100
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#pragma unroll 2
        for (int ii = 0; ii < 32; ++ii) {
            const int i = half * 32 + ii;
            const float4 r = s_row[i];
            const float sa = s_area[i];
#pragma unroll
            for (int t = 0; t < kGroupTiles; ++t) {
                unsigned bl[2];
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    float inter, den;
                    iou_terms(r, sa, col[t][hh], inter, den);
                    bool pr;
                    if (EXACT) {
                        pr = (inter / den) > thr;
                    } else {
                        const float tt = thr * den;
                        const float d = inter - tt;
                        const float e = __fmaf_rn(tt, kEps, kTiny);
                        pr = d > 0.0f;
                        amb |= !(fabsf(d) > e);
                    }
                    bl[hh] = __ballot_sync(0xffffffffu, pr);
                }
                if (lane == ii) words[half][t] = ((u64)bl[1] << 32) | (u64)bl[0];
            }
        }
    }
}

__global__ void __launch_bounds__(kMaskThreads) nms_mask_kernel(const NmsArgs a) {
    __shared__ float4 s_row[64];
    __shared__ float s_area[64];
    __shared__ u32 s_segend[64];
    __shared__ u32 s_rowoff[64];
    const int b = blockIdx.y, I = blockIdx.x;
    const NmsImg info = a.info[b];
    const int M = info.M;
    if (I * 64 >= M || info.overflow) return;
    const size_t off = (size_t)b * a.cap;
    const float4* sb = a.sboxes + off;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 64) {
        const int p = I * 64 + tid;
        if (p < M) {
            const float4 q = sb[p];
            s_row[tid] = q;
            s_area[tid] = (q.z - q.x) * (q.w - q.y);
            s_segend[tid] = a.seg_end[off + p];
            s_rowoff[tid] = a.rowoff[off + p];
        } else {
            const float far = -3.0e38f;
            s_row[tid] = make_float4(far, far, far, far);
            s_area[tid] = 0.0f;
            s_segend[tid] = 0;
            s_rowoff[tid] = 0;
        }
    }
    __syncthreads();
    const int last_row = min(63, M - 1 - I * 64);
    const int Jend = (int)((s_segend[last_row] - 1) >> 6) + 1;  // exclusive
    const int ngroups = (Jend - I + kGroupTiles - 1) / kGroupTiles;
    u64* mask = a.mask + (u64)b * a.mask_words_per_img;

    for (int g = warp; g < ngroups; g += kMaskThreads / 32) {
        const int J0 = I + g * kGroupTiles;
        ColBox col[kGroupTiles][2];
#pragma unroll
        for (int t = 0; t < kGroupTiles; ++t)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int c = (J0 + t) * 64 + hh * 32 + lane;
                float4 q = make_float4(3.0e38f, 3.0e38f, 3.0e38f, 3.0e38f);
                if (c < M) q = sb[c];
                col[t][hh].x1 = q.x; col[t][hh].y1 = q.y; col[t][hh].x2 = q.z; col[t][hh].y2 = q.w;
                col[t][hh].w = q.z - q.x; col[t][hh].h = q.w - q.y;
            }
        u64 words[2][kGroupTiles];
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int t = 0; t < kGroupTiles; ++t) words[k][t] = 0;
        bool amb = false;
        if (info.exact) {
            mask_group<true>(s_row, s_area, col, a.thr, lane, words, amb);
        } else {
            mask_group<false>(s_row, s_area, col, a.thr, lane, words, amb);
            if (__any_sync(0xffffffffu, amb)) mask_group<true>(s_row, s_area, col, a.thr, lane, words, amb);
        }
        // post masks (diagonal, segment end) and store: lane owns rows lane and lane+32
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int i = k * 32 + lane;
            const int p = I * 64 + i;
            if (p >= M) continue;
            const u32 se = s_segend[i];
            const u32 width = ((((se - 1) >> 6) - (u32)I + 1) + 3u) & ~3u;
            if ((u32)(g * kGroupTiles) >= width) continue;
            u64 w[kGroupTiles];
#pragma unroll
            for (int t = 0; t < kGroupTiles; ++t) {
                const int J = J0 + t;
                u64 x = words[k][t];
                if (J == I) x &= (i == 63) ? 0ull : (~0ull << (i + 1));
                const long long lim = (long long)se - (long long)J * 64;
                if (lim <= 0) x = 0ull;
                else if (lim < 64) x &= (1ull << lim) - 1ull;
                w[t] = x;
            }
            ulonglong2* dst = reinterpret_cast<ulonglong2*>(mask + s_rowoff[i] + (u32)(g * kGroupTiles));
            dst[0] = make_ulonglong2(w[0], w[1]);
            dst[1] = make_ulonglong2(w[2], w[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// greedy scan + emit, one CTA per image
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kScanThreads) nms_scan_kernel(const NmsArgs a) {
    extern __shared__ u64 s_dyn[];  // remv[nwords_cap] | keepbits[nwords_cap] | rankbits (u32)
    __shared__ u64 s_diag[64];
    __shared__ u32 s_roff[64];
    __shared__ u32 s_wid[64];
    __shared__ u64 s_keep;
    __shared__ unsigned char s_klist[64];
    __shared__ u32 s_scan[32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const NmsImg info = a.info[b];
    const int M = info.M;
    if (M == 0) {
        if (tid == 0) a.n_keep[b] = 0;
        return;
    }
    if (info.overflow) {
        if (tid == 0) a.n_keep[b] = -1;
        return;
    }
    const size_t off = (size_t)b * a.cap;
    const int nw = (M + 63) >> 6;
    const int nw_cap = (a.cap + 63) >> 6;
    u64* remv = s_dyn;
    u64* keepw = s_dyn + nw_cap;
    u32* rankbits = reinterpret_cast<u32*>(s_dyn + 2 * nw_cap);
    for (int i = tid; i < nw; i += kScanThreads) { remv[i] = 0ull; keepw[i] = 0ull; }
    for (int i = tid; i < 2 * nw; i += kScanThreads) rankbits[i] = 0u;
    const u64* mask = a.mask + (u64)b * a.mask_words_per_img;
    const u32* rowoff = a.rowoff + off;
    const u32* seg_end = a.seg_end + off;
    __syncthreads();

    for (int G = 0; G < nw; ++G) {
        const int nrows = min(64, M - G * 64);
        if (tid < 64) {
            u64 d = 0ull;
            u32 ro = 0, wd = 0;
            if (tid < nrows) {
                const int p = G * 64 + tid;
                ro = rowoff[p];
                wd = ((seg_end[p] - 1) >> 6) - (u32)G + 1;
                d = mask[ro];
            }
            s_diag[tid] = d; s_roff[tid] = ro; s_wid[tid] = wd;
        }
        __syncthreads();
        if (tid == 0) {
            u64 rem = remv[G];
            if (nrows < 64) rem |= ~0ull << nrows;
            u64 kb = 0ull;
#pragma unroll
            for (int i = 0; i < 64; ++i) {
                const u64 d = s_diag[i];
                if (!((rem >> i) & 1ull)) { kb |= 1ull << i; rem |= d; }
            }
            s_keep = kb;
            keepw[G] = kb;
        }
        __syncthreads();
        // OR the kept rows into the later words.  The CTA is laid out as C row-slots x Wp words
        // so that all 1024 threads have independent loads in flight (4 per thread per step).
        const u64 kb = s_keep;
        const int nk = __popcll(kb);
        if (tid < 64 && ((kb >> tid) & 1ull)) s_klist[__popcll(kb & ((1ull << tid) - 1ull))] = (unsigned char)tid;
        __syncthreads();
        const int W = nw - 1 - G;
        if (W > 0 && nk > 0) {
            const int Wp = (W + 31) & ~31;
            const int stride_k = Wp < kScanThreads ? Wp : kScanThreads;
            const int C = kScanThreads / stride_k;
            const int c = tid / stride_k;
            if (c < C) {
                for (int k = tid - c * stride_k; k < W; k += stride_k) {
                    u64 acc = 0ull;
                    for (int j = c; j < nk; j += 4 * C) {
                        u64 v[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int jj = j + u * C;
                            v[u] = 0ull;
                            if (jj < nk) {
                                const int i = s_klist[jj];
                                if ((u32)(k + 1) < s_wid[i]) v[u] = __ldg(mask + s_roff[i] + k + 1);
                            }
                        }
                        acc |= (v[0] | v[1]) | (v[2] | v[3]);
                    }
                    if (acc) {
                        u32* dst = reinterpret_cast<u32*>(remv + G + 1 + k);
                        if ((u32)acc) atomicOr(dst, (u32)acc);
                        if ((u32)(acc >> 32)) atomicOr(dst + 1, (u32)(acc >> 32));
                    }
                }
            }
        }
        __syncthreads();
    }

    // ---- emit kept indices in descending score order ----
    const u32* pos_rank = a.pos_rank + off;
    const u32* order = a.v0 + off;
    for (int p = tid; p < M; p += kScanThreads) {
        if ((keepw[p >> 6] >> (p & 63)) & 1ull) {
            const u32 r = pos_rank[p];
            atomicOr(&rankbits[r >> 5], 1u << (r & 31));
        }
    }
    __syncthreads();
    const int nw32 = (M + 31) >> 5;
    const int chunk = (nw32 + kScanThreads - 1) / kScanThreads;
    const int c0 = min(tid * chunk, nw32), c1 = min(c0 + chunk, nw32);
    u32 sum = 0;
    for (int w = c0; w < c1; ++w) sum += __popc(rankbits[w]);
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
    int64_t* keep = a.keep + off;
    for (int w = c0; w < c1; ++w) {
        u32 bits = rankbits[w];
        while (bits) {
            const int j = __ffs(bits) - 1;
            bits &= bits - 1u;
            keep[base++] = (int64_t)order[w * 32 + j];
        }
    }
    if (tid == 0) a.n_keep[b] = (int)total;
}

// ---- host side ----------------------------------------------------------------------------
struct NmsLayout {
    size_t k[6], sboxes, pos_rank, seg_end, rowoff, info, mask, total;
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static NmsLayout nms_layout(int B, int cap) {
    NmsLayout L;
    size_t o = 0;
    const size_t n = (size_t)B * cap;
    for (int i = 0; i < 6; ++i) { L.k[i] = o; o = align_up(o + n * 4, 256); }
    L.sboxes = o; o = align_up(o + n * 16, 256);
    L.pos_rank = o; o = align_up(o + n * 4, 256);
    L.seg_end = o; o = align_up(o + n * 4, 256);
    L.rowoff = o; o = align_up(o + n * 4, 256);
    L.info = o; o = align_up(o + (size_t)B * sizeof(NmsImg), 256);
    L.mask = o;
    L.total = o;
    return L;
}

// worst case words of one image: a single segment, rows padded to 4 words
static size_t nms_worst_words(int cap) {
    const size_t nw = ((size_t)cap + 63) / 64;
    // tile I has 64 rows of width round4(nw - I)
    size_t words = 0;
    for (size_t I = 0; I < nw; ++I) words += 64 * (((nw - I) + 3) & ~(size_t)3);
    return words;
}

// yb_nms_graph.cu
size_t graph_min_workspace(int B, int cap);
int graph_nms(const float* boxes, const float* scores, const int64_t* classes, const int* counts, int B, int cap,
              double iou_threshold, long long trick_max_numel, int64_t* keep, int* n_keep, void* ws,
              size_t ws_bytes, cudaStream_t st);

}  // namespace yb

namespace yb {
int graph_stats(const void* ws, size_t ws_bytes, int B, int cap, unsigned long long* evals, unsigned long long* edges,
                unsigned long long* cands, cudaStream_t st);
}
extern "C" int yb_nms_graph_stats(const void* ws, size_t ws_bytes, int B, int cap, unsigned long long* evals_host,
                                  unsigned long long* edges_host, unsigned long long* cands_host, void* stream) {
    return yb::graph_stats(ws, ws_bytes, B, cap, evals_host, edges_host, cands_host, (cudaStream_t)stream);
}

extern "C" size_t yb_nms_graph_workspace_bytes(int B, int cap, int edges_per_box) {
    if (B <= 0 || cap <= 0) return 256;
    if (edges_per_box < 1) edges_per_box = 1;
    return yb::graph_min_workspace(B, cap) + (size_t)B * cap * (size_t)edges_per_box * 8 + 256;
}

extern "C" size_t yb_nms_workspace_bytes(int B, int cap) {
    if (B <= 0 || cap <= 0) return 256;
    yb::NmsLayout L = yb::nms_layout(B, cap);
    size_t dense = L.total + (size_t)B * yb::nms_worst_words(cap) * 8 + 256;
    size_t graph = yb_nms_graph_workspace_bytes(B, cap, 16);
    return dense > graph ? dense : graph;
}

extern "C" size_t yb_nms_min_workspace_bytes(int B, int cap) {
    if (B <= 0 || cap <= 0) return 256;
    size_t dense = yb::nms_layout(B, cap).total + (size_t)B * 32;
    size_t graph = yb::graph_min_workspace(B, cap);
    return dense > graph ? dense : graph;
}

extern "C" int yb_batched_nms(const float* boxes, const float* scores, const int64_t* classes,
                              const int* counts, int B, int cap, double iou_threshold,
                              long long trick_max_numel, int algo, int64_t* keep, int* n_keep, void* ws,
                              size_t ws_bytes, void* stream) {
    using namespace yb;
    YB_CHECK_ARG(B >= 0 && cap >= 0, "nms: bad B/cap");
    if (B == 0) return 0;
    YB_CHECK_ARG(n_keep, "nms: null n_keep");
    cudaStream_t st = (cudaStream_t)stream;
    if (cap == 0) {
        YB_CUDA(cudaMemsetAsync(n_keep, 0, sizeof(int) * B, st));
        return 0;
    }
    YB_CHECK_ARG(boxes && scores && keep && ws, "nms: null pointer");
    YB_CHECK_ARG(aligned16(boxes) && aligned16(ws), "nms: boxes/workspace must be 16-byte aligned");
    YB_CHECK_ARG(B <= 65535, "nms: B too large");
    YB_CHECK_ARG(cap <= 400000, "nms: cap too large");
    YB_CHECK_ARG(algo == YB_NMS_GRAPH || algo == YB_NMS_BITMASK, "nms: unknown algo %d", algo);
    if (algo == YB_NMS_GRAPH)
        return graph_nms(boxes, scores, classes, counts, B, cap, iou_threshold, trick_max_numel, keep, n_keep, ws,
                         ws_bytes, st);
    NmsLayout L = nms_layout(B, cap);
    YB_CHECK_ARG(ws_bytes >= L.total + (size_t)B * 32, "nms: workspace too small (%zu < %zu)", ws_bytes, L.total + (size_t)B * 32);
    char* w = reinterpret_cast<char*>(ws);
    NmsArgs a;
    a.boxes = reinterpret_cast<const float4*>(boxes); a.scores = scores; a.classes = classes; a.counts = counts;
    a.B = B; a.cap = cap;
    a.thr = (float)iou_threshold;
    a.thr_fast_ok = (a.thr >= 0.0f && a.thr <= 1e30f) ? 1 : 0;
    a.trick_max_numel = trick_max_numel;
    a.k0 = (u32*)(w + L.k[0]); a.v0 = (u32*)(w + L.k[1]); a.k1 = (u32*)(w + L.k[2]);
    a.v1 = (u32*)(w + L.k[3]); a.k2 = (u32*)(w + L.k[4]); a.v2 = (u32*)(w + L.k[5]);
    a.sboxes = (float4*)(w + L.sboxes);
    a.pos_rank = (u32*)(w + L.pos_rank); a.seg_end = (u32*)(w + L.seg_end); a.rowoff = (u32*)(w + L.rowoff);
    a.info = (NmsImg*)(w + L.info);
    a.mask = (u64*)(w + L.mask);
    a.mask_words_per_img = ((ws_bytes - L.mask) / 8 / (size_t)B) & ~(size_t)3;
    a.keep = keep; a.n_keep = n_keep;

    YB_LAUNCH("nms_sort_kernel", st, nms_sort_kernel<<<B, kSortThreads, 0, st>>>(a));
    dim3 grid((cap + 63) / 64, B);
    YB_LAUNCH("nms_mask_kernel", st, nms_mask_kernel<<<grid, kMaskThreads, 0, st>>>(a));
    const size_t nw_cap = ((size_t)cap + 63) / 64;
    const size_t dyn = nw_cap * 8 * 2 + nw_cap * 2 * 4;
    YB_CHECK_ARG(dyn <= 200 * 1024, "nms: cap too large for the scan kernel");
    if (dyn > 40 * 1024)
        YB_CUDA(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    YB_LAUNCH("nms_scan_kernel", st, nms_scan_kernel<<<B, kScanThreads, dyn, st>>>(a));
    return 0;
}
