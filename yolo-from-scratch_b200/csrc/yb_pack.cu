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
    const int mine = n_keep[b];
    if ((int)(blockIdx.x * blockDim.x) >= mine && !(blockIdx.x == 0)) return;   // nothing to pack in this CTA
    if (threadIdx.x < 32) {   // first row of image b = sum of the earlier images' counts: one warp, coalesced loads
        int base = 0, failed = 0;
        for (int i = threadIdx.x; i < B; i += 32) {
            const int n = n_keep[i];
            failed |= n < 0;
            if (i < b) base += max(n, 0);
        }
        base = __reduce_add_sync(0xffffffffu, base);
        failed = __reduce_or_sync(0xffffffffu, failed);
        if (threadIdx.x == 0) {
            s_base = base;
            if (blockIdx.x == 0) {
                offsets[b] = base;
                // an image NMS gave up on (n_keep < 0: the bitmask algorithm with too small a workspace) must not
                // read as "no detections": the total becomes -1 and the host raises at its synchronisation point
                if (b == B - 1) offsets[B] = failed ? -1 : base + max(mine, 0);
            }
        }
    }
    __syncthreads();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= mine) return;
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
