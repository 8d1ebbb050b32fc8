// Microbenchmark: throughput of legacy mma.sync.m16n8k16 (f16 x f16 -> f32) and of FFMA2 on sm_100a,
// per SM, as a function of resident warps.  Used to decide whether the conv stack should move from
// FFMA to mma.sync (design note in DESIGN.md).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

__global__ void hmma_kernel(float* out, int iters) {
  unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 0x3c003c00u, b1 = 0x3c003c00u;
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void ffma2_kernel(float* out, int iters) {
  float2 acc[16]; for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x, i);
  float2 x = make_float2(1.0001f, 0.9999f), w = make_float2(0.5f, 0.25f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = __ffma2_rn(x, acc[i], w);
  }
  float s = 0; for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void ffma_kernel(float* out, int iters) {
  float acc[32]; for (int i = 0; i < 32; ++i) acc[i] = threadIdx.x + i;
  float x = 1.0001f, w = 0.5f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = fmaf(x, acc[i], w);
  }
  float s = 0; for (int i = 0; i < 32; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float* d; cudaMalloc(&d, 148 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int warps : {4, 8, 12, 16, 32}) {
    for (int which = 0; which < 3; ++which) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (which == 0) hmma_kernel<<<148, warps * 32>>>(d, iters);
        else if (which == 1) ffma2_kernel<<<148, warps * 32>>>(d, iters);
        else ffma_kernel<<<148, warps * 32>>>(d, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double per_sm_per_clk;
      if (which == 0) per_sm_per_clk = double(warps) * iters * 8 * 2048 / (ms * 1e-3 * 1.965e9);
      else per_sm_per_clk = double(warps) * 32 * iters * 32 / (ms * 1e-3 * 1.965e9);
      printf("%s warps/SM=%2d  %.3f ms  %.0f MAC/clk/SM (at 1.965 GHz)%s\n", which == 0 ? "HMMA m16n8k16" : (which == 1 ? "FFMA2        " : "FFMA         "),
             warps, ms, per_sm_per_clk, which == 0 ? "" : "  [peak 128]");
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
