// Probe: do legacy mma.sync (HMMA) warps and tcgen05.mma (UTCHMMA) share the tensor pipe of an SM?
// One CTA per SM: warp 0 issues back-to-back tcgen05.mma (M = 128, N = 256, K = 16, kind::f16, operands in shared
// memory), warps 1..W run an mma.sync.m16n8k16 loop with 8 independent accumulators.  Each role is timed alone and
// together (clock64 inside the kernel, CTA 0; CUDA events around the launch).  If the combined run takes
// max(T_umma, T_hmma) the two paths are independent; if it takes the sum they serialise on one pipe -- which bounds
// what a fused conv (mma.sync) + FC (tcgen05) kernel can gain from overlap.  A third role (FFMA2 warps) does the same
// for the SIMT epilogue work.
// Build + run on a B200: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/coissue coissue_probe.cu && /tmp/coissue
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity))
    if (++spins > (1u << 26)) __trap();
}
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3ffffu) >> 4);
  d |= uint64_t(1) << 16;
  d |= uint64_t(1024 >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}
__device__ __forceinline__ uint32_t idesc_f16(int n) { return (1u << 4) | (uint32_t(n >> 3) << 17) | (uint32_t(128 >> 4) << 24); }
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p, e;\n.reg .b32 rx;\nelect.sync rx|e, 0xffffffff;\nsetp.ne.b32 p, %4, 0;\n"
               "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("{\n.reg .pred e;\n.reg .b32 rx;\nelect.sync rx|e, 0xffffffff;\n"
               "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(smem_u32(bar)) : "memory");
}

// roles: bit 0 = tcgen05 issuer (warp 0), bit 1 = mma.sync warps, bit 2 = FFMA2 warps.
// warps 1 .. n_hmma run HMMA, the following n_ffma warps run FFMA2.
__global__ void __launch_bounds__(1024, 1) probe(int roles, int n_umma, int hmma_iters, int ffma_iters, int n_hmma, int n_ffma,
                                                 long long* cyc, float* sink) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;              // 128 x 64 fp16 = 16 KB
  uint8_t* sB = smem + 16384;      // 256 x 64 fp16 = 32 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 ones
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = __shfl_sync(0xffffffffu, *slot, 0);
  const long long t0 = clock64();
  if (warp == 0) {
    if (roles & 1) {
      const uint64_t da = umma_desc_sw128(smem_u32(sA)), db = umma_desc_sw128(smem_u32(sB));
      const uint32_t idesc = idesc_f16(256);
      for (int i = 0; i < n_umma; i += 4) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) mma_ss(tmem + uint32_t((i >> 2) & 1) * 256u, da + uint64_t(ks * 2), db + uint64_t(ks * 2), idesc, 1);
      }
      commit(bar);
      mbar_wait(bar, 0);
      if (blockIdx.x == 0 && tid == 0) cyc[0] = clock64() - t0;
    }
  } else if (warp <= n_hmma) {
    if (roles & 2) {
      unsigned a0 = tid, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 0x3c003c00u, b1 = 0x3c003c00u;
      float c[8][4];
      for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
      for (int it = 0; it < hmma_iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                       : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                       : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      }
      float s = 0;
      for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 4; ++j) s += c[i][j];
      sink[blockIdx.x * blockDim.x + tid] = s;
      if (blockIdx.x == 0 && tid == 32) cyc[1] = clock64() - t0;
    }
  } else if (warp <= n_hmma + n_ffma) {
    if (roles & 4) {
      float2 acc[16];
      for (int i = 0; i < 16; ++i) acc[i] = make_float2(tid, i);
      const float2 x = make_float2(1.0001f, 0.9999f), w = make_float2(0.5f, 0.25f);
      for (int it = 0; it < ffma_iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = __ffma2_rn(x, acc[i], w);
      }
      float s = 0;
      for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
      sink[blockIdx.x * blockDim.x + tid] = s;
      if (blockIdx.x == 0 && tid == 32 * (n_hmma + 1)) cyc[2] = clock64() - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
  const int smem_bytes = 49152 + 64 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  long long* dC;
  float* dS;
  cudaMalloc(&dC, 32), cudaMalloc(&dS, 148 * 1024 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  const int n_umma = 16384;      // x 128 clk at peak = 2.1 M clk
  struct Cfg { int n_hmma, n_ffma, hmma_iters, ffma_iters; };
  // per warp: hmma_iters x 8 MMAs; 8 warps x 8.6 clk per MMA per SMSP (2 warps per SMSP) -> choose ~2 M clk alone
  const Cfg cfgs[] = {{8, 0, 14000, 0}, {12, 0, 9400, 0}, {16, 0, 7000, 0}, {8, 8, 14000, 30000}, {0, 8, 0, 60000}, {0, 16, 0, 30000}};
  for (const Cfg& c : cfgs) {
    const int threads = 32 * (1 + c.n_hmma + c.n_ffma);
    printf("--- %d HMMA warps (%d x 8 MMAs each), %d FFMA2 warps (%d x 16 FFMA2 each), %d tcgen05 MMAs N=256\n", c.n_hmma, c.hmma_iters, c.n_ffma,
           c.ffma_iters, n_umma);
    for (int roles : {1, 2, 4, 3, 5, 6, 7}) {
      if ((roles & 2) && !c.n_hmma) continue;
      if ((roles & 4) && !c.n_ffma) continue;
      long long h[3] = {0, 0, 0};
      float ms = 0;
      for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(dC, 0, 32);
        cudaEventRecord(e0);
        probe<<<148, threads, smem_bytes>>>(roles, n_umma, c.hmma_iters, c.ffma_iters, c.n_hmma, c.n_ffma, dC, dS);
        cudaEventRecord(e1);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) {
          printf("CUDA error %s\n", cudaGetErrorString(e));
          return 1;
        }
        cudaEventElapsedTime(&ms, e0, e1);
      }
      cudaMemcpy(h, dC, 24, cudaMemcpyDeviceToHost);
      printf("roles %s%s%s: kernel %.3f ms | cycles umma %lld (%.1f clk/MMA)  hmma %lld (%.2f clk/MMA/SMSP-equivalent: %.0f MAC/clk/SM)  ffma2 %lld (%.1f FFMA/clk/SM)\n",
             (roles & 1) ? "U" : "-", (roles & 2) ? "H" : "-", (roles & 4) ? "F" : "-", ms, h[0], h[0] ? double(h[0]) / n_umma : 0.0, h[1],
             h[1] ? double(h[1]) / (double(c.hmma_iters) * 8 * c.n_hmma / 4) : 0.0,
             h[1] ? double(c.n_hmma) * c.hmma_iters * 8 * 2048 / double(h[1]) : 0.0, h[2],
             h[2] ? double(c.n_ffma) * 32 * c.ffma_iters * 32 / double(h[2]) : 0.0);
    }
  }
  return 0;
}
