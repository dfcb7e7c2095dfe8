// yb_filter.cu — K4: fused decode + sigmoid + confidence filter + ordered compaction.
// Reference: the per-scale body of predict(), train.py:1152-1222, and the P3->P4->P5
// concatenation :1227-1229, batched over B images.
//
// Bound: HBM.  Algorithmic bytes: rows*min(row_bytes,32) for the objectness column,
// plus M*row_bytes for the rows that pass and 28*M for the candidate list.
// Two launches:
//   filter_count_kernel  one objectness test per row, per-tile pass counts
//   filter_emit_kernel   per-image exclusive prefix over the tile counts (order preserving),
//                        in-tile ballot scan, and for the passing rows decode + class max +
//                        box build.  Dense tiles are staged in shared memory with coalesced
//                        128-bit loads (row stride 5+nc words; conflict-free when odd).
#include "yb_common.cuh"

namespace yb {

constexpr int kFTile = 128;  // rows per tile == threads per CTA

struct FilterScale {
    const float* pred;
    const float* anchors;
    uint32_t rows;        // rows per image at this scale = H*W*A
    uint32_t tile_begin;  // first tile (within an image) of this scale
    float inv_w, inv_h;
    FastDiv d_A, d_W;
};

struct FilterArgs {
    int S, A, nc, cap;
    uint32_t row, tiles_per_img;
    float img, inv_img, conf;
    int stage_ok;              // shared-memory staging available for this row length
    const float* letterbox;    // (B,3) scale, pad_top, pad_left or null
    FilterScale sc[YB_MAX_SCALES];
    int* tile_counts;          // (B, tiles_per_img)
    float4* boxes;
    float* scores;
    int64_t* classes;
    int* counts;
};

__device__ __forceinline__ int tile_scale(const FilterArgs& a, uint32_t tile) {
    int s = 0;
#pragma unroll
    for (int k = 1; k < YB_MAX_SCALES; ++k)
        if (k < a.S && tile >= a.sc[k].tile_begin) s = k;
    return s;
}

__global__ void __launch_bounds__(kFTile) filter_count_kernel(const FilterArgs a) {
    const uint32_t tile = blockIdx.x, b = blockIdx.y;
    const FilterScale& L = a.sc[tile_scale(a, tile)];
    const uint32_t r = (tile - L.tile_begin) * kFTile + threadIdx.x;
    bool pass = false;
    if (r < L.rows) {
        const float x = __ldg(L.pred + ((size_t)b * L.rows + r) * a.row + 4);
        pass = sigmoidf_ref(x) > a.conf;  // :1157,:1166-1167 objectness only
    }
    const int n = __syncthreads_count(pass);
    if (threadIdx.x == 0) a.tile_counts[b * a.tiles_per_img + tile] = n;
}

// first index of the maximum of sigmoid(x[0..nc)) — torch.max(dim=1) semantics (:1189).
// sigmoid is monotone, so the maximum is sigmoid(max logit); an earlier, smaller logit can
// only tie after fp32 rounding if it lies within 1.0 of min(max logit, 14).
template <typename Load>
__device__ __forceinline__ void class_max(int nc, Load ld, float& prob, int& id) {
    float m = ld(0);
    int mi = 0;
    for (int c = 1; c < nc; ++c) {
        const float v = ld(c);
        if (v > m) { m = v; mi = c; }
    }
    prob = sigmoidf_ref(m);
    id = mi;
    const float lo = fminf(m, 14.0f) - 1.0f;
    for (int c = 0; c < mi; ++c) {
        const float v = ld(c);
        if (v > lo && sigmoidf_ref(v) == prob) { id = c; break; }
    }
}

__global__ void __launch_bounds__(kFTile) filter_emit_kernel(const FilterArgs a) {
    extern __shared__ float s_tile[];
    __shared__ int s_red[kFTile / 32];
    __shared__ int s_wbase[kFTile / 32];
    const uint32_t tile = blockIdx.x, b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int* tc = a.tile_counts + b * a.tiles_per_img;

    // exclusive prefix of this tile within its image
    int part = 0;
    for (uint32_t t = threadIdx.x; t < tile; t += kFTile) part += tc[t];
    part = __reduce_add_sync(0xffffffffu, part);
    if (lane == 0) s_red[warp] = part;
    __syncthreads();
    int prefix = 0;
#pragma unroll
    for (int k = 0; k < kFTile / 32; ++k) prefix += s_red[k];
    const int count = tc[tile];
    if (tile == a.tiles_per_img - 1 && threadIdx.x == 0) {
        const int tot = prefix + count;
        a.counts[b] = tot < a.cap ? tot : a.cap;
    }
    if (count == 0) return;

    const int s = tile_scale(a, tile);
    const FilterScale& L = a.sc[s];
    const uint32_t row0 = (tile - L.tile_begin) * kFTile;
    const uint32_t nrows = min((uint32_t)kFTile, L.rows - row0);
    const size_t base = ((size_t)b * L.rows + row0) * a.row;  // float offset of the tile
    const float* g = L.pred + base;

    // dense tiles of long rows: stage through shared memory with coalesced loads
    const bool staged = a.stage_ok && a.row > 8 && count * 4 >= (int)nrows;
    if (staged) {
        const uint32_t nfl = nrows * a.row;
        if ((base & 3) == 0) {
            const float4* g4 = reinterpret_cast<const float4*>(g);
            float4* s4 = reinterpret_cast<float4*>(s_tile);
            for (uint32_t v = threadIdx.x; v < (nfl >> 2); v += kFTile) s4[v] = __ldcs(g4 + v);
            for (uint32_t e = (nfl & ~3u) + threadIdx.x; e < nfl; e += kFTile) s_tile[e] = g[e];
        } else {
            for (uint32_t e = threadIdx.x; e < nfl; e += kFTile) s_tile[e] = g[e];
        }
        __syncthreads();
    }
    const float* x = staged ? (s_tile + threadIdx.x * a.row) : (g + (size_t)threadIdx.x * a.row);

    bool pass = false;
    float sobj = 0.0f;
    if (threadIdx.x < nrows) {
        sobj = sigmoidf_ref(x[4]);
        pass = sobj > a.conf;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, pass);
    if (lane == 0) s_wbase[warp] = __popc(bal);
    __syncthreads();
    int pos = prefix + __popc(bal & ((1u << lane) - 1));
    for (int k = 0; k < warp; ++k) pos += s_wbase[k];
    if (!pass || pos >= a.cap) return;

    // decode (:1154) with the model's img_size
    const uint32_t r = row0 + threadIdx.x;
    uint32_t cell, an, gy, gx;
    L.d_A.divmod(r, cell, an);
    L.d_W.divmod(cell, gy, gx);
    const float aw = __ldg(L.anchors + an * 2), ah = __ldg(L.anchors + an * 2 + 1);
    const float bx = decode_xy(x[0], (float)gx, L.inv_w);
    const float by = decode_xy(x[1], (float)gy, L.inv_h);
    const float bw = decode_wh(x[2], aw, a.inv_img);
    const float bh = decode_wh(x[3], ah, a.inv_img);
    // class probability and id (:1184-1189)
    float cprob;
    int cid;
    class_max(a.nc, [&](int c) { return x[5 + c]; }, cprob, cid);
    // pixels, corners, letterbox reverse (:1192-1213)
    const float xc = bx * a.img, yc = by * a.img, wp = bw * a.img, hp = bh * a.img;
    float x1 = xc - wp * 0.5f, y1 = yc - hp * 0.5f, x2 = xc + wp * 0.5f, y2 = yc + hp * 0.5f;
    if (a.letterbox) {
        const float inv_s = 1.0f / a.letterbox[b * 3 + 0];
        const float pt = a.letterbox[b * 3 + 1], pl = a.letterbox[b * 3 + 2];
        x1 = (x1 - pl) * inv_s; y1 = (y1 - pt) * inv_s;
        x2 = (x2 - pl) * inv_s; y2 = (y2 - pt) * inv_s;
    }
    const size_t o = (size_t)b * a.cap + pos;
    a.boxes[o] = make_float4(x1, y1, x2, y2);
    a.scores[o] = sobj * cprob;  // :1216
    a.classes[o] = (int64_t)cid;
}

static int filter_fill(const yb_heads_desc* d, FilterArgs& a) {
    YB_CHECK_ARG(d, "filter: null descriptor");
    YB_CHECK_ARG(d->S >= 1 && d->S <= YB_MAX_SCALES && d->B >= 0 && d->A > 0 && d->A <= YB_MAX_ANCHORS && d->nc >= 1,
                 "filter: bad S/B/A/nc (nc must be >= 1, train.py:1184-1189)");
    a.S = d->S; a.A = d->A; a.nc = d->nc; a.row = 5 + d->nc;
    a.img = d->img_size; a.inv_img = 1.0f / d->img_size;
    uint32_t tile = 0;
    for (int s = 0; s < d->S; ++s) {
        YB_CHECK_ARG(d->H[s] > 0 && d->W[s] > 0, "filter: bad grid at scale %d", s);
        unsigned long long n = (unsigned long long)d->B * d->H[s] * d->W[s] * d->A * (5 + d->nc);
        YB_CHECK_ARG(n < (1ull << 32), "filter: scale %d too large", s);
        FilterScale& L = a.sc[s];
        L.pred = d->pred[s]; L.anchors = d->anchors[s];
        L.rows = (uint32_t)d->H[s] * d->W[s] * d->A;
        L.tile_begin = tile;
        tile += (L.rows + kFTile - 1) / kFTile;
        L.inv_w = 1.0f / (float)d->W[s]; L.inv_h = 1.0f / (float)d->H[s];
        L.d_A = FastDiv(d->A); L.d_W = FastDiv(d->W[s]);
    }
    a.tiles_per_img = tile;
    return 0;
}

}  // namespace yb

extern "C" size_t yb_filter_workspace_bytes(const yb_heads_desc* d) {
    yb::FilterArgs a;
    if (yb::filter_fill(d, a)) return 0;
    return (size_t)d->B * a.tiles_per_img * sizeof(int) + 16;
}

extern "C" int yb_filter_compact(const yb_heads_desc* d, float conf_thres, const float* letterbox,
                                 float* boxes, float* scores, int64_t* classes, int* counts, int cap,
                                 void* ws, size_t ws_bytes, void* stream) {
    using namespace yb;
    FilterArgs a;
    int rc = filter_fill(d, a);
    if (rc) return rc;
    if (d->B == 0) return 0;
    YB_CHECK_ARG(boxes && scores && classes && counts && ws && cap > 0, "filter: null output");
    YB_CHECK_ARG(aligned16(boxes), "filter: boxes must be 16-byte aligned");
    for (int s = 0; s < d->S; ++s)
        YB_CHECK_ARG(d->pred[s] && d->anchors[s] && aligned16(d->pred[s]), "filter: bad tensor at scale %d", s);
    if (ws_bytes < yb_filter_workspace_bytes(d)) {
        set_error("filter: workspace too small");
        return YB_EWORKSPACE;
    }
    YB_CHECK_ARG(d->B <= 65535, "filter: B too large");
    a.cap = cap; a.conf = conf_thres; a.letterbox = letterbox;
    a.tile_counts = reinterpret_cast<int*>(ws);
    a.boxes = reinterpret_cast<float4*>(boxes); a.scores = scores; a.classes = classes; a.counts = counts;
    const size_t smem = (size_t)kFTile * a.row * sizeof(float);
    a.stage_ok = smem <= 200 * 1024;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(a.tiles_per_img, d->B);
    YB_LAUNCH("filter_count_kernel", st, filter_count_kernel<<<grid, kFTile, 0, st>>>(a));
    const size_t dyn = (a.stage_ok && a.row > 8) ? smem : 0;
    if (dyn > 48 * 1024)
        YB_CUDA(cudaFuncSetAttribute(filter_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    YB_LAUNCH("filter_emit_kernel", st, filter_emit_kernel<<<grid, kFTile, dyn, st>>>(a));
    return 0;
}
