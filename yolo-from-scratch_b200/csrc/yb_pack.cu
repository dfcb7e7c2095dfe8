// yb_pack.cu — device-side detection list: the (x1,y1,x2,y2,conf,class) tuples predict() builds
// one .item() at a time (train.py:1236-1246), packed for a whole batch so that one D2H copy of
// sum(n_keep) rows replaces K*6 host syncs.
#include "yb_common.cuh"

namespace yb {

__global__ void __launch_bounds__(256) pack_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores,
                                                   const int64_t* __restrict__ classes, const int64_t* __restrict__ keep,
                                                   const int* __restrict__ n_keep, int B, int cap,
                                                   float* __restrict__ out, int* __restrict__ offsets) {
    __shared__ int s_base;
    const int b = blockIdx.y;
    if (threadIdx.x == 0) {
        int base = 0;
        for (int i = 0; i < b; ++i) base += max(n_keep[i], 0);
        s_base = base;
        if (blockIdx.x == 0) {
            offsets[b] = base;
            if (b == B - 1) {
                // an image NMS gave up on (n_keep < 0: the bitmask algorithm with too small a workspace) must not
                // read as "no detections": the total becomes -1 and the host raises at its synchronisation point
                bool failed = false;
                for (int i = 0; i < B; ++i) failed |= n_keep[i] < 0;
                offsets[B] = failed ? -1 : base + max(n_keep[b], 0);
            }
        }
    }
    __syncthreads();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_keep[b]) return;
    const size_t src = (size_t)b * cap + (size_t)keep[(size_t)b * cap + k];
    const float4 q = boxes[src];
    float* o = out + (size_t)(s_base + k) * 6;
    o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w;
    o[4] = scores[src];
    o[5] = (float)classes[src];
}

}  // namespace yb

extern "C" int yb_pack_detections(const float* boxes, const float* scores, const int64_t* classes,
                                  const int64_t* keep, const int* n_keep, int B, int cap, float* out,
                                  int* offsets, void* stream) {
    using namespace yb;
    YB_CHECK_ARG(B >= 0 && cap >= 0, "pack: bad B/cap");
    if (B == 0 || cap == 0) return 0;
    YB_CHECK_ARG(boxes && scores && classes && keep && n_keep && out && offsets && aligned16(boxes), "pack: null pointer");
    YB_CHECK_ARG(B <= 65535, "pack: B too large");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((cap + 255) / 256, B);
    YB_LAUNCH("pack_kernel", st,
              pack_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(boxes), scores, classes, keep, n_keep, B,
                                                cap, out, offsets));
    return 0;
}
