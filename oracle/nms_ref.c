/*
 * nms_ref.c — ORACLE (test infrastructure, not product code).
 *
 * Plain-C restatement of torchvision 0.26.0's `nms` operator, the third-party dependency the
 * reference calls at train.py:1232-1233 through torchvision.ops.batched_nms.  torchvision is
 * not vendored under /root/reference (README.md:25 just says `pip install torch torchvision`),
 * so the algorithm is restated from its published kernels:
 *   arith = 0  CUDA kernel (csrc/ops/cuda/nms_kernel.cu, devIoU, as compiled for sm_100:
 *              the union's Sa + Sb is contracted to fma(bw, bh, Sa); float compare)
 *   arith = 1  CPU kernel (csrc/ops/cpu/nms_kernel.cpp: areas rounded separately,
 *              float IoU compared with the double threshold)
 * Both: stable descending score order (NaN first, lower index wins ties), greedy suppression
 * with strict '>'.
 * Pinned by tests/test_oracle_cpu.py against the installed operator (arith=1, here) and by
 * tests/test_nms_gpu.py against its CUDA kernel (arith=0, on the GPU box).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may load this.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int score_before(float a, float b) {
    /* torch.sort(descending=True): NaN is the largest value */
    int na = a != a, nb = b != b;
    if (na || nb) return na && !nb;
    return a > b;
}

static void merge_sort(const float* s, int64_t* idx, int64_t* tmp, int64_t lo, int64_t hi) {
    if (hi - lo < 2) return;
    int64_t mid = lo + (hi - lo) / 2;
    merge_sort(s, idx, tmp, lo, mid);
    merge_sort(s, idx, tmp, mid, hi);
    int64_t i = lo, j = mid, k = lo;
    while (i < mid && j < hi) {
        /* stable: take from the right half only if it is strictly before */
        if (score_before(s[idx[j]], s[idx[i]])) tmp[k++] = idx[j++];
        else tmp[k++] = idx[i++];
    }
    while (i < mid) tmp[k++] = idx[i++];
    while (j < hi) tmp[k++] = idx[j++];
    memcpy(idx + lo, tmp + lo, (size_t)(hi - lo) * sizeof(int64_t));
}

static float maxf(float a, float b) { return a > b ? a : b; }
static float minf(float a, float b) { return a < b ? a : b; }

/* returns the number of kept boxes; keep_out[k] = original indices in descending score order */
int64_t yb_oracle_nms(const float* boxes, const float* scores, int64_t n, double thr, int arith,
                      int64_t* keep_out) {
    if (n <= 0) return 0;
    int64_t* order = (int64_t*)malloc((size_t)n * sizeof(int64_t));
    int64_t* tmp = (int64_t*)malloc((size_t)n * sizeof(int64_t));
    unsigned char* dead = (unsigned char*)calloc((size_t)n, 1);
    float* area = (float*)malloc((size_t)n * sizeof(float));
    for (int64_t i = 0; i < n; ++i) {
        order[i] = i;
        const float* b = boxes + 4 * i;
        const float w = b[2] - b[0], h = b[3] - b[1];
        area[i] = w * h;
    }
    merge_sort(scores, order, tmp, 0, n);
    const float thr_f = (float)thr;
    int64_t kept = 0;
    for (int64_t oi = 0; oi < n; ++oi) {
        const int64_t i = order[oi];
        if (dead[i]) continue;
        keep_out[kept++] = i;
        const float* a = boxes + 4 * i;
        for (int64_t oj = oi + 1; oj < n; ++oj) {
            const int64_t j = order[oj];
            if (dead[j]) continue;
            const float* b = boxes + 4 * j;
            const float iw = maxf(minf(a[2], b[2]) - maxf(a[0], b[0]), 0.0f);
            const float ih = maxf(minf(a[3], b[3]) - maxf(a[1], b[1]), 0.0f);
            const float inter = iw * ih;
            int sup;
            if (arith == 0) {
                const float bw = b[2] - b[0], bh = b[3] - b[1];
                const float den0 = fmaf(bw, bh, area[i]);
                const float den = den0 - inter;
                const float iou = inter / den;
                sup = iou > thr_f;
            } else {
                const float uni0 = area[i] + area[j];
                const float uni = uni0 - inter;
                const float iou = inter / uni;
                sup = (double)iou > thr;
            }
            if (sup) dead[j] = 1;
        }
    }
    free(order); free(tmp); free(dead); free(area);
    return kept;
}
