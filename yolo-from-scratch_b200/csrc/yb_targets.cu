// yb_targets.cu — K2: anchor-to-ground-truth target assignment.
// Reference: YOLODataset.compute_anchor_iou train.py:108-131 and the label loop of
// YOLODataset.__getitem__ train.py:147-205.  Integer/index work: results are bit-exact.
//
// Bound: HBM write of the dense zero-filled targets (T bytes per scale, cudaMemsetAsync);
// the assignment itself touches <= max_gt rows per image.  One CTA per image; the reference's
// sequential "first ground truth wins a slot" rule (:193) is resolved in parallel: ground truth
// i writes iff no j < i maps to the same (scale, cell, anchor) slot.
#include "yb_common.cuh"

namespace yb {

__device__ __forceinline__ float shape_iou(float w, float h, float aw, float ah) {
    // train.py:119-130, all fp32
    const float box_area = w * h;
    const float anchor_area = aw * ah;
    const float inter = fminf(w, aw) * fminf(h, ah);
    const float uni = (box_area + anchor_area) - inter;
    return inter / (uni + 1e-16f);
}

__global__ void anchor_iou_kernel(const float* __restrict__ box_wh, const float* __restrict__ anchors,
                                  float* __restrict__ out, int n, int A) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * A) return;
    int b = i / A, k = i - b * A;
    out[i] = shape_iou(box_wh[b * 2], box_wh[b * 2 + 1], anchors[k * 2], anchors[k * 2 + 1]);
}

struct TargetArgs {
    const double* labels;     // (B,max_gt,5)
    const int* n_gt;          // (B)
    const double* letterbox;  // (B,5) orig_w, orig_h, scale, pad_top, pad_left
    const float* anchors;     // (S,A,2)
    float* tgt[YB_MAX_SCALES];
    int G[YB_MAX_SCALES];
    int S, A, nc, max_gt, img;
    int* status;
};

struct GtSlot {
    uint32_t key;  // 0xffffffff = rejected
    float x, y, w, h;
    int cls;
    int pad0, pad1;  // 32 bytes: the label copy behind the slots stays 8-byte aligned
};

static inline size_t assign_smem_bytes(int max_gt, int S, int A) {
    return (size_t)max_gt * (sizeof(GtSlot) + 5 * sizeof(double)) + (size_t)S * A * 2 * sizeof(float) + 16;
}

// python: int(v) truncates toward zero; the index then wraps once if negative (list/tensor
// indexing), and is an IndexError outside [-G, G).
__device__ __forceinline__ bool py_cell(double v, int G, int& cell) {
    if (!(v == v) || fabs(v) > 2.0e9) return false;
    int g = (int)v;
    if (g > G - 1) g = G - 1;
    if (g < 0) g += G;
    if (g < 0) return false;
    cell = g;
    return true;
}

// Phase 1 of both assignment kernels: every ground truth of image b -> its slot key and fp32 row.
// One ground truth per 16-lane group: lanes 0-3 evaluate the four letterbox-adjusted values (fp64
// division chains), every lane one anchor's shape IoU (fp32 IEEE division), then a 4-step shuffle
// reduction picks the winner — the serial version (one thread per ground truth) was a ~1,000
// instruction dependent chain and took 25 us for 50 boxes.
__device__ __forceinline__ bool assign_slots(const TargetArgs& a, int b, int n, GtSlot* s_gt) {
    // labels of this image and the anchors go to shared memory first: one round of global latency
    // instead of one per ground-truth round
    double* s_lab = reinterpret_cast<double*>(s_gt + a.max_gt);
    float* s_anch = reinterpret_cast<float*>(s_lab + (size_t)a.max_gt * 5);
    for (int k = threadIdx.x; k < n * 5; k += blockDim.x) s_lab[k] = a.labels[(size_t)b * a.max_gt * 5 + k];
    for (int k = threadIdx.x; k < a.S * a.A * 2; k += blockDim.x) s_anch[k] = a.anchors[k];
    __syncthreads();
    const double ow = a.letterbox[b * 5 + 0], oh = a.letterbox[b * 5 + 1];
    const double sc = a.letterbox[b * 5 + 2], pt = a.letterbox[b * 5 + 3], pl = a.letterbox[b * 5 + 4];
    const double img = (double)a.img;
    const int sub = threadIdx.x & 15, grp = threadIdx.x >> 4, ngrp = blockDim.x >> 4;
    const unsigned gmask = 0xffffu << (threadIdx.x & 16);
    const int n_anch = a.S * a.A;
    bool bad = false;
    for (int i0 = 0; i0 < n; i0 += ngrp) {
        const int i = i0 + grp;
        const bool live = i < n;
        const double* L = s_lab + (size_t)(live ? i : 0) * 5;
        // :159-162 — python double arithmetic, left to right, no fusion
        double val = 0.0;
        if (sub < 4) {
            const double dim = (sub & 1) ? oh : ow;               // x, w scale with the width; y, h with the height
            const double pad = sub == 0 ? pl : (sub == 1 ? pt : 0.0);
            double t = __dmul_rn(__dmul_rn(L[1 + sub], dim), sc);
            if (sub < 2) t = __dadd_rn(t, pad);
            val = __ddiv_rn(t, img);
        }
        const double xc = __shfl_sync(gmask, val, 0, 16), yc = __shfl_sync(gmask, val, 1, 16);
        const double wd = __shfl_sync(gmask, val, 2, 16), hd = __shfl_sync(gmask, val, 3, 16);
        // :165-167 — pixels, then torch.tensor(...) rounds to fp32
        const float wpx = (float)__dmul_rn(wd, img), hpx = (float)__dmul_rn(hd, img);
        // :170-180 — best anchor over all scales: strict '>' across scales, first argmax within a scale
        // = the first maximum in (scale, anchor) order
        float best = -1.0f;
        int bq = 0x7fffffff;
        for (int q = sub; q < n_anch; q += 16) {
            const float* an = s_anch + q * 2;
            const float iou = shape_iou(wpx, hpx, an[0], an[1]);
            if (iou > best) { best = iou; bq = q; }
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(gmask, best, o, 16);
            const int oq = __shfl_xor_sync(gmask, bq, o, 16);
            if (ob > best || (ob == best && oq < bq)) { best = ob; bq = oq; }
        }
        if (live && sub == 0) {
            if (bq == 0x7fffffff) bq = 0;  // every IoU was NaN / not above -1: the reference keeps scale 0, anchor 0
            const int bs = bq / a.A, ba = bq - bs * a.A;
            // :183-189
            const int G = a.G[bs];
            int gx = 0, gy = 0;
            GtSlot g;
            g.cls = (int)L[0];
            const bool ok = py_cell(__dmul_rn(xc, (double)G), G, gx) && py_cell(__dmul_rn(yc, (double)G), G, gy) &&
                            (a.nc <= 1 || (g.cls >= 0 && g.cls < a.nc));
            if (ok) {
                g.key = ((uint32_t)bs << 28) | ((uint32_t)ba << 24) | ((uint32_t)gy << 12) | (uint32_t)gx;
            } else {
                g.key = 0xffffffffu;
                bad = true;
            }
            g.x = (float)xc; g.y = (float)yc; g.w = (float)wd; g.h = (float)hd;  // :195-197
            s_gt[i] = g;
        }
    }
    return bad;
}

// :193 — the slot is already owned by an earlier ground truth of the same image
__device__ __forceinline__ bool slot_taken(const GtSlot* s_gt, int i, uint32_t key) {
    for (int j = 0; j < i; ++j)
        if (s_gt[j].key == key) return true;
    return false;
}

__global__ void __launch_bounds__(256) build_targets_kernel(const TargetArgs a) {
    extern __shared__ GtSlot s_gt[];
    const int b = blockIdx.x;
    int n = a.n_gt[b];
    n = n < 0 ? 0 : (n > a.max_gt ? a.max_gt : n);
    const bool bad = assign_slots(a, b, n, s_gt);
    __syncthreads();
    const int row = 5 + a.nc;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const GtSlot g = s_gt[i];
        if (g.key == 0xffffffffu || slot_taken(s_gt, i, g.key)) continue;
        const int s = g.key >> 28, an = (g.key >> 24) & 15, gy = (g.key >> 12) & 4095, gx = g.key & 4095;
        const int G = a.G[s];
        float* t = a.tgt[s] + ((((size_t)b * G + gy) * G + gx) * a.A + an) * row;
        t[0] = g.x; t[1] = g.y; t[2] = g.w; t[3] = g.h;
        t[4] = 1.0f;                                   // :199
        if (a.nc == 1) t[5] = 1.0f;                    // :201-202
        else if (a.nc > 1) t[5 + g.cls] = 1.0f;        // :205
    }
    if (bad && a.status) atomicOr(a.status, 1);
}

// Sparse form of the same assignment: instead of dense (G,G,A,5+nc) tensors the
// winners are appended to per-scale positive lists (row index + entry id), their target rows are
// kept as 32-byte entries, and a 1-bit-per-row map marks them for the objectness pass.
__global__ void __launch_bounds__(256) assign_sparse_kernel(const TargetArgs a, const SparseOut o) {
    extern __shared__ GtSlot s_gt[];
    __shared__ int s_cnt[YB_MAX_SCALES], s_base[YB_MAX_SCALES];
    const int b = blockIdx.x;
    int n = a.n_gt[b];
    n = n < 0 ? 0 : (n > a.max_gt ? a.max_gt : n);
    if (threadIdx.x < YB_MAX_SCALES) s_cnt[threadIdx.x] = 0;
    const bool bad = assign_slots(a, b, n, s_gt);
    __syncthreads();
    // winners reserve a slot of their scale's list: counted per CTA in shared memory, ONE global atomic per
    // CTA and scale (3,000 single-thread atomics on three addresses serialised at ~6 ns each: 18 us)
    for (int i0 = 0; i0 < n; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        GtSlot g;
        g.key = 0xffffffffu;
        if (i < n) g = s_gt[i];
        const bool win = g.key != 0xffffffffu && !slot_taken(s_gt, i, g.key);
        const int s = win ? (int)(g.key >> 28) : 0;
        int local = 0;
        if (win) local = atomicAdd(&s_cnt[s], 1);
        __syncthreads();
        if (threadIdx.x < a.S) {
            const int c = s_cnt[threadIdx.x];
            s_base[threadIdx.x] = c ? atomicAdd(o.pos_count + threadIdx.x, c) : 0;
        }
        __syncthreads();
        if (win) {
            const int an = (g.key >> 24) & 15, gy = (g.key >> 12) & 4095, gx = g.key & 4095;
            const int G = a.G[s];
            const uint32_t r = (uint32_t)((((size_t)b * G + gy) * G + gx) * a.A + an);
            const uint32_t e = (uint32_t)b * a.max_gt + i;
            o.entries[e] = SparseEntry{g.x, g.y, g.w, g.h, a.nc == 1 ? 0 : g.cls, 0, 0, 0};  // :201-202
            atomicOr(o.bits + o.bits_begin[s] + (r >> 5), 1u << (r & 31));
            const uint32_t k = o.list_begin[s] + (uint32_t)(s_base[s] + local);
            o.pos_list[k] = r;
            o.pos_ent[k] = e;
        }
        __syncthreads();
        if (threadIdx.x < YB_MAX_SCALES) s_cnt[threadIdx.x] = 0;
        __syncthreads();
    }
    if (bad && a.status) atomicOr(a.status, 1);
}

int launch_assign_sparse(const double* labels, const int* n_gt, const double* letterbox, const float* anchors, int B,
                         int max_gt, int S, const int* G, int A, int nc, int img_size, int* status,
                         const SparseOut& o, cudaStream_t st) {
    TargetArgs a;
    a.labels = labels; a.n_gt = n_gt; a.letterbox = letterbox; a.anchors = anchors;
    a.S = S; a.A = A; a.nc = nc; a.max_gt = max_gt; a.img = img_size; a.status = status;
    for (int s = 0; s < S; ++s) {
        YB_CHECK_ARG(G[s] > 0 && G[s] <= 4096, "assign: bad grid at scale %d", s);
        a.G[s] = G[s];
        a.tgt[s] = nullptr;
    }
    if (status) YB_CUDA(cudaMemsetAsync(status, 0, sizeof(int), st));
    if (max_gt == 0 || B == 0) return 0;
    size_t smem = assign_smem_bytes(max_gt, S, A);
    YB_CHECK_ARG(smem <= 200 * 1024, "assign: max_gt=%d too large", max_gt);
    if (smem > 48 * 1024)
        YB_CUDA(cudaFuncSetAttribute(assign_sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    YB_LAUNCH("assign_sparse_kernel", st, assign_sparse_kernel<<<B, 256, smem, st>>>(a, o));
    return 0;
}

}  // namespace yb

extern "C" int yb_anchor_iou(const float* box_wh, const float* anchors, float* iou_out, int n, int A,
                             void* stream) {
    using namespace yb;
    YB_CHECK_ARG(n >= 0 && A > 0, "anchor_iou: bad shape");
    if (n == 0) return 0;
    YB_CHECK_ARG(box_wh && anchors && iou_out, "anchor_iou: null pointer");
    int tot = n * A;
    cudaStream_t st = (cudaStream_t)stream;
    YB_LAUNCH("anchor_iou_kernel", st, anchor_iou_kernel<<<(tot + 127) / 128, 128, 0, st>>>(box_wh, anchors, iou_out, n, A));
    return 0;
}

extern "C" int yb_build_targets(const double* labels, const int* n_gt, const double* letterbox,
                                const float* anchors, float* const targets_host[YB_MAX_SCALES],
                                int B, int max_gt, int S, const int* G_host, int A, int nc,
                                int img_size, int* status, void* stream) {
    using namespace yb;
    YB_CHECK_ARG(B >= 0 && max_gt >= 0 && S >= 1 && S <= YB_MAX_SCALES && A > 0 && A <= YB_MAX_ANCHORS &&
                     nc >= 0 && img_size > 0 && G_host && targets_host,
                 "build_targets: bad arguments");
    if (B == 0) return 0;
    YB_CHECK_ARG(n_gt && letterbox && anchors && (max_gt == 0 || labels), "build_targets: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    TargetArgs a;
    a.labels = labels; a.n_gt = n_gt; a.letterbox = letterbox; a.anchors = anchors;
    a.S = S; a.A = A; a.nc = nc; a.max_gt = max_gt; a.img = img_size; a.status = status;
    for (int s = 0; s < S; ++s) {
        YB_CHECK_ARG(G_host[s] > 0 && G_host[s] <= 4096 && targets_host[s], "build_targets: bad scale %d", s);
        a.G[s] = G_host[s];
        a.tgt[s] = targets_host[s];
        size_t bytes = (size_t)B * G_host[s] * G_host[s] * A * (5 + nc) * sizeof(float);
        {
            KernelScope ks("targets_memset", st);
            YB_CUDA(cudaMemsetAsync(targets_host[s], 0, bytes, st));  // :141-145 torch.zeros
        }
        count_launch();
    }
    if (status) YB_CUDA(cudaMemsetAsync(status, 0, sizeof(int), st));
    if (max_gt == 0) return 0;
    size_t smem = assign_smem_bytes(max_gt, S, A);
    YB_CHECK_ARG(smem <= 200 * 1024, "build_targets: max_gt=%d too large", max_gt);
    if (smem > 48 * 1024)
        YB_CUDA(cudaFuncSetAttribute(build_targets_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    YB_LAUNCH("build_targets_kernel", st, build_targets_kernel<<<B, 256, smem, st>>>(a));
    return 0;
}
