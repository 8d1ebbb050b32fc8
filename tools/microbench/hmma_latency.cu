// Microbenchmark: latency of a dependent mma.sync.m16n8k16 (f32 accumulate) chain and issue interval of independent
// ones, for one warp alone on an SM (clock64), plus the same for the packed-fp32 epilogue chain of the conv stage.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void hmma_chain(long long* out, float* sink, int iters) {
  unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 0x3c003c00u, b1 = 0x3c003c00u;
  float c[CHAINS][4];
  for (int i = 0; i < CHAINS; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  const long long t1 = clock64();
  float s = 0; for (int i = 0; i < CHAINS; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  sink[threadIdx.x] = s;
  if (threadIdx.x == 0) out[0] = t1 - t0;
}

// HMMA -> dependent FFMA2 -> HMMA (conv1 -> act -> conv2 pattern): measures the HMMA result latency seen by a consumer
__global__ void hmma_to_alu(long long* out, float* sink, int iters) {
  unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 0x3c003c00u, b1 = 0x3c003c00u;
  float c[4] = {0, 0, 0, 0};
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    a0 = __float_as_uint(c[0] * 1.0001f);   // one dependent FMUL feeding the next HMMA's A operand
  }
  const long long t1 = clock64();
  sink[threadIdx.x] = c[0] + c[1] + c[2] + c[3];
  if (threadIdx.x == 0) out[0] = t1 - t0;
}

int main() {
  long long* d; float* s; cudaMalloc(&d, 8); cudaMalloc(&s, 4096);
  const int iters = 4096;
  long long h;
#define RUN(K, N) K<<<1, 32>>>(d, s, iters); cudaDeviceSynchronize(); K<<<1, 32>>>(d, s, iters); cudaDeviceSynchronize(); \
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); printf("%-28s %.2f clk per HMMA\n", #K, double(h) / (double(iters) * N));
  RUN(hmma_chain<1>, 1) RUN(hmma_chain<2>, 2) RUN(hmma_chain<3>, 3) RUN(hmma_chain<4>, 4) RUN(hmma_chain<8>, 8)
  RUN(hmma_to_alu, 1)
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
