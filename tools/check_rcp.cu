// Exhaustive check of yb::rcp_normal (yb_common.cuh) against the compiler's IEEE division 1.0f / y over every
// float with 2^-126 <= |y| < 2^126 (both signs), and of the filter's sigmoid form 1/(1+expf(-x)) against
// sigmoidf_ref over every float x whose 1+e^-x is below 2^126.  Prints the number of mismatching bit patterns.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/check_rcp tools/check_rcp.cu
#include <cstdio>
#include "../yolo-from-scratch_b200/csrc/yb_common.cuh"

__global__ void check_rcp(unsigned long long* bad, unsigned int* first) {
    const unsigned long long n = 0xfcull << 23;   // exponents 1..252
    unsigned long long mism = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < 2 * n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned int bits = (unsigned int)(i % n) + (1u << 23) | (i >= n ? 0x80000000u : 0u);
        const float y = __uint_as_float(bits);
        const float a = 1.0f / y, b = yb::rcp_normal(y);
        if (__float_as_uint(a) != __float_as_uint(b)) { ++mism; atomicMin(first, bits & 0x7fffffffu); }
    }
    if (mism) atomicAdd(bad, mism);
}
__global__ void check_sigmoid(unsigned long long* bad, unsigned long long* outside) {
    unsigned long long mism = 0, out = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < (1ull << 32);
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((unsigned int)i);
        if (x != x) continue;
        const float d = 1.0f + expf(-x);
        if (!(d < yb::kRcpNormalMax)) { ++out; continue; }   // the kernel redoes the full division there
        if (__float_as_uint(yb::sigmoidf_ref(x)) != __float_as_uint(yb::rcp_normal(d))) ++mism;
    }
    if (mism) atomicAdd(bad, mism);
    if (out) atomicAdd(outside, out);
}
int main() {
    unsigned long long *d, h[3] = {0, 0, 0};
    unsigned int* f, hf = 0xffffffffu;
    cudaMalloc(&d, sizeof(h)); cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaMalloc(&f, 4); cudaMemcpy(f, &hf, 4, cudaMemcpyHostToDevice);
    check_rcp<<<148 * 8, 256>>>(d, f);
    check_sigmoid<<<148 * 8, 256>>>(d + 1, d + 2);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(&hf, f, 4, cudaMemcpyDeviceToHost);
    printf("check_rcp: %s; rcp_normal vs 1.0f/y over 2 x %llu floats: %llu mismatches (first |y| bits 0x%08x); "
           "sigmoid form over all non-NaN x: %llu mismatches, %llu values outside the fast range\n",
           cudaGetErrorString(e), 0xfcull << 23, h[0], hf, h[1], h[2]);
    return (e != cudaSuccess || h[0] || h[1]) ? 1 : 0;
}
