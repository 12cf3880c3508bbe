// Instruction-throughput probe (GPU box): FFMA vs FHFMA (fma.rn.f32.f16) vs cvt+cvt+FFMA, 8 independent chains per thread.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/bench_fhfma.cu -o /tmp/bench_fhfma && /tmp/bench_fhfma
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ float fhfma(unsigned short a, unsigned short b, float c) {
  float d;
  asm volatile("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c));
  return d;
}
template <int MODE>
__global__ void k(const uint32_t* in, float* out, int iters) {
  uint32_t x = in[threadIdx.x & 31], y = in[32 + (threadIdx.x & 31)];
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MODE == 0) {
        acc[j] = fmaf(__uint_as_float(x), __uint_as_float(y), acc[j]);
      } else if (MODE == 1) {
        acc[j] = fhfma((unsigned short)(x >> (j & 1 ? 16 : 0)), (unsigned short)(y >> (j & 1 ? 16 : 0)), acc[j]);
      } else {
        __half2 hx = *reinterpret_cast<__half2*>(&x), hy = *reinterpret_cast<__half2*>(&y);
        float a = (j & 1) ? __high2float(hx) : __low2float(hx), b = (j & 1) ? __high2float(hy) : __low2float(hy);
        asm volatile("" : "+f"(a), "+f"(b));
        acc[j] = fmaf(a, b, acc[j]);
      }
    }
    x += 0x10001u;
  }
  float s = 0;
  for (int j = 0; j < 8; ++j) s += acc[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  uint32_t* in; float* out;
  cudaMalloc(&in, 256); cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaMemset(in, 0x3c, 256);
  const int iters = 20000;
  for (int mode = 0; mode < 3; ++mode) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 8, 256>>>(in, out, iters);
      if (mode == 1) k<1><<<148 * 8, 256>>>(in, out, iters);
      if (mode == 2) k<2><<<148 * 8, 256>>>(in, out, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = 148.0 * 8 * 256 * iters * 8;
    printf("%s: %.3f ms  %.2f Tfma/s\n", mode == 0 ? "FFMA          " : mode == 1 ? "FHFMA         " : "cvt+cvt+FFMA  ", ms, fma / ms / 1e9);
  }
  return 0;
}
