// Probe of the tcgen05.mma forms a tensor-core conv stage would use (one CTA, M = 128, kind::f16, small N):
//   SS  A and B from shared memory (128-byte swizzle, K-major)           -- the form fc_fused.cu uses
//   TS  A from tensor memory (written with tcgen05.st), B from shared memory
// For each (mode, N) it checks D = A * B^T against the host and times back-to-back MMAs (cycles per K = 16 MMA), which
// answers: is a skinny-N MMA bound by the shared-memory read of its A operand, and does A-in-TMEM remove that?
// Build + run on a B200:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_probe umma_probe.cu && /tmp/umma_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

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
    if (++spins > (1u << 24)) __trap();
}
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3ffffu) >> 4);
  d |= uint64_t(1) << 16;
  d |= uint64_t(1024 >> 4) << 32;   // stride between 8-row groups
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;           // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint32_t idesc_f16(int n) { return (1u << 4) | (uint32_t(n >> 3) << 17) | (uint32_t(128 >> 4) << 24); }

// Issued by ALL lanes of a converged warp with warp-uniform operands; elect.sync picks the one lane that really issues.
// (Issuing from inside `if (lane == 0)` makes ptxas wrap every MMA in a divergence "waterfall" loop -- ELECT / R2UR.BROADCAST /
// BRA.U.ANY -- that costs ~110 cycles per MMA: harmless under a 128-cycle N = 256 MMA, fatal for skinny ones.)
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p, e;\n.reg .b32 rx;\nelect.sync rx|e, 0xffffffff;\nsetp.ne.b32 p, %4, 0;\n"
               "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p, e;\n.reg .b32 rx;\nelect.sync rx|e, 0xffffffff;\nsetp.ne.b32 p, %4, 0;\n"
               "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("{\n.reg .pred e;\n.reg .b32 rx;\nelect.sync rx|e, 0xffffffff;\n"
               "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int kK = 64;   // one swizzle row of fp16

// element (row, k) of a K-major [rows][64] fp16 tile in the 128-byte-swizzled layout
__device__ __forceinline__ uint32_t sw128_off(int row, int k) {
  return uint32_t((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + (k & 7) * 2);
}

// mode 0 = SS, 1 = TS.  reps > 1: timing (D is then reps * A B^T).
__global__ void __launch_bounds__(128, 1) probe(const __half* A, const __half* B, float* D, int N, int mode, int reps, long long* cycles,
                                                int nacc /* accumulators used round-robin (timing only) */) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // 16 KB
  uint8_t* sB = smem + 16384;         // N * 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * kK; i += 128) *reinterpret_cast<__half*>(sA + sw128_off(i / kK, i % kK)) = A[i];
  for (int i = tid; i < N * kK; i += 128) *reinterpret_cast<__half*>(sB + sw128_off(i / kK, i % kK)) = B[i];
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
  const uint32_t lane_addr = uint32_t(warp * 32) << 16;
  constexpr uint32_t kACol = 256;     // A operand lives at columns 256 .. 287 (K = 64 fp16 = 32 columns)
  if (mode == 1) {
    // row tid of A -> TMEM lane tid: column j holds (A[row][2j], A[row][2j+1]), low half = even k
    uint32_t r[32];
    for (int j = 0; j < 32; ++j) {
      const __half2 h = __halves2half2(A[tid * kK + 2 * j], A[tid * kK + 2 * j + 1]);
      r[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(tmem + lane_addr + kACol),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (warp == 0) {   // the whole warp, converged
    const uint64_t da = umma_desc_sw128(smem_u32(sA)), db = umma_desc_sw128(smem_u32(sB));
    const uint32_t idesc = idesc_f16(N);
    const long long t0 = clock64();
    int acc = 0;         // nacc = 1: one accumulator, dependent MMAs; > 1: rotation over independent ones
    if (mode == 0) {
      for (int rep = 0; rep < reps; ++rep)
#pragma unroll
        for (int ks = 0; ks < kK / 16; ++ks) {
          mma_ss(tmem + uint32_t(acc * N), da + uint64_t(ks * 2), db + uint64_t(ks * 2), idesc, (rep | ks) != 0);
          acc = (acc + 1 == nacc) ? 0 : acc + 1;
        }
    } else {
      for (int rep = 0; rep < reps; ++rep)
#pragma unroll
        for (int ks = 0; ks < kK / 16; ++ks) {
          mma_ts(tmem + uint32_t(acc * N), tmem + kACol + ks * 8, db + uint64_t(ks * 2), idesc, (rep | ks) != 0);
          acc = (acc + 1 == nacc) ? 0 : acc + 1;
        }
    }
    commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (cycles && tid == 0) *cycles = t1 - t0;
  }
  __syncthreads();
  mbar_wait(bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(tmem + lane_addr + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) D[tid * N + c0 + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
  const int smem_bytes = 16384 + 32768 + 64 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  std::vector<__half> hA(128 * kK), hB(256 * kK);
  srand(7);
  for (auto& v : hA) v = __float2half(float(rand() % 9 - 4));
  for (auto& v : hB) v = __float2half(float(rand() % 5 - 2));
  __half *dA, *dB;
  float* dD;
  long long* dC;
  cudaMalloc(&dA, hA.size() * 2), cudaMalloc(&dB, hB.size() * 2), cudaMalloc(&dD, 128 * 256 * 4), cudaMalloc(&dC, 8);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  const int Ns[] = {16, 32, 48, 64, 128, 256};
  for (int mode = 0; mode < 2; ++mode)
    for (int N : Ns) {
      std::vector<float> hD(128 * N);
      probe<<<1, 128, smem_bytes>>>(dA, dB, dD, N, mode, 1, dC, 1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("%s N=%d: CUDA error %s\n", mode ? "TS" : "SS", N, cudaGetErrorString(e));
        return 1;
      }
      cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          double s = 0;
          for (int k = 0; k < kK; ++k) s += double(__half2float(hA[m * kK + k])) * double(__half2float(hB[n * kK + k]));
          const double d = fabs(s - hD[m * N + n]);
          if (d > maxerr) maxerr = d;
        }
      const int reps = 512;
      printf("%s M=128 N=%3d K=16: max|err| = %g; cycles per MMA (%d back to back) with 1/2/4/8 accumulators in rotation:", mode ? "TS" : "SS", N,
             maxerr, reps * 4);
      for (int nacc = 1; nacc <= 8 && nacc * N <= 256; nacc *= 2) {
        long long cyc = 0;
        probe<<<1, 128, smem_bytes>>>(dA, dB, dD, N, mode, reps, dC, nacc);
        cudaDeviceSynchronize();
        cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
        printf("  %.1f", double(cyc) / (reps * 4));
      }
      printf("\n");
    }
  return 0;
}
