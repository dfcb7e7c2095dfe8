// Micro-benchmark: issue throughput (warp-instructions per clock per SM) of the instruction families the NMS
// kernels are made of, to decide between fp32 and packed fp16 culling tests.  Design aid, not product code.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_pipes tools/ubench_pipes.cu
#include <cuda_fp16.h>
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048
#define CHAINS 8

template <int OP>
__global__ void __launch_bounds__(1024) k(float* out, unsigned long long* cycles, float seed) {
    float f[CHAINS];
    __half2 h[CHAINS];
    int v[CHAINS];
    const float s2 = seed * 1.0001f;
    const __half2 hs = __floats2half2_rn(seed, s2);
    for (int c = 0; c < CHAINS; ++c) {
        f[c] = seed + c + threadIdx.x;
        h[c] = __floats2half2_rn(f[c], f[c] * 0.5f);
        v[c] = (int)f[c];
    }
    __syncthreads();
    const unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (OP == 0) f[c] = fmaxf(f[c], s2) ;                       // FMNMX
            if (OP == 1) f[c] = __fmaf_rn(f[c], seed, s2);              // FFMA
            if (OP == 2) f[c] = f[c] + s2;                              // FADD
            if (OP == 3) h[c] = __hmax2(h[c], hs);                      // HMNMX2
            if (OP == 4) h[c] = __hfma2(h[c], hs, hs);                  // HFMA2
            if (OP == 5) h[c] = __hadd2(h[c], hs);                      // HADD2
            if (OP == 6) v[c] = max(v[c], (int)it);                     // IMNMX / VIMNMX
            if (OP == 7) v[c] = v[c] * 3 + it;                          // IMAD
            if (OP == 8) v[c] = (v[c] & 0x7e0) ^ (v[c] >> 3);           // LOP3 + SHF
            if (OP == 9) { f[c] = fmaxf(f[c], s2); f[c] = __fmaf_rn(f[c], seed, s2); }   // FMNMX + FFMA pair (dual pipe)
            if (OP == 10) { h[c] = __hmax2(h[c], hs); h[c] = __hfma2(h[c], hs, hs); }    // HMNMX2 + HFMA2
            if (OP == 11) { h[c] = __hmax2(h[c], hs); f[c] = __fmaf_rn(f[c], seed, s2); } // HMNMX2 + FFMA
            if (OP == 12) v[c] = __vimax3_s16x2(v[c], it, (int)threadIdx.x);             // DPX 3-way s16x2 max
            if (OP == 13) v[c] = __viaddmax_s16x2(v[c], it, (int)threadIdx.x);           // DPX add+max s16x2
            if (OP == 14) { v[c] += __popc(__ballot_sync(0xffffffffu, f[c] > s2)); }     // FSETP+VOTE+POPC+IADD
            if (OP == 15) { v[c] += (__hgt2(h[c], hs).x != __half(0)) ? 1 : 0; }         // HSETP2-ish
            if (OP == 16) f[c] = fminf(fmaxf(f[c], s2), seed);                           // FMNMX x2 (maybe FMNMX3)
        }
    }
    const unsigned long long t1 = clock64();
    float acc = 0;
    for (int c = 0; c < CHAINS; ++c) acc += f[c] + __low2float(h[c]) + __high2float(h[c]) + v[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter) {
    int nb = 148;
    float* out; unsigned long long* cyc;
    cudaMalloc(&out, nb * 1024 * sizeof(float));
    cudaMalloc(&cyc, nb * sizeof(unsigned long long));
    k<OP><<<nb, 1024>>>(out, cyc, 1.5f);
    k<OP><<<nb, 1024>>>(out, cyc, 1.5f);
    cudaDeviceSynchronize();
    unsigned long long h[148];
    cudaMemcpy(h, cyc, nb * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < nb; ++i) avg += h[i];
    avg /= nb;
    const double winstr = 32.0 * ITERS * CHAINS * per_iter;   // warp-instructions of the measured kind per SM (32 warps)
    printf("%-28s %7.3f warp-instr/clk/SM  (%d counted per chain step)  err=%s\n", name, winstr / avg, per_iter,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("FMNMX", 1); run<1>("FFMA", 1); run<2>("FADD", 1); run<3>("HMNMX2", 1); run<4>("HFMA2", 1); run<5>("HADD2", 1);
    run<6>("IMNMX", 1); run<7>("IMAD", 1); run<8>("LOP3+SHF", 2); run<9>("FMNMX+FFMA", 2); run<10>("HMNMX2+HFMA2", 2);
    run<11>("HMNMX2+FFMA", 2); run<12>("VIMAX3.S16x2", 1); run<13>("VIADDMAX.S16x2", 1); run<14>("FSETP+VOTE+POPC+IADD", 4);
    run<15>("HSETP2+sel", 2); run<16>("FMNMX,FMNMX", 2);
    return 0;
}
