// Stage FC1 on tcgen05: D[n][448] = leaky(A[n][2688] * W[2688][448] + b), three heads side by side.
//
// Precision.  A single fp16/bf16/tf32 pass misses the reference by ~3e-4 in probability (SURVEY.md
// section 7.3-A), so both operands are split into fp16 hi + lo parts (after power-of-two scaling so the
// lo parts stay normal) and three MMAs accumulate hi*hi + hi*lo + lo*hi in fp32 in TMEM; the dropped
// lo*lo term is 2^-22 relative.
//
// Tiling.  One output tile = 128 CTUs (UMMA M = 128, one TMEM lane per CTU) x 224 columns (half of the
// 448; UMMA N = 224).  K is streamed in 64-element slices: per slice TMA brings A_hi, A_lo (128 x 64)
// and B_hi, B_lo (224 x 64) into 128-byte-swizzled shared memory (88 KB per stage, 2 stages) and one
// elected thread issues 4 (K = 16 steps) x 3 (passes) tcgen05.mma into a 224-column fp32 accumulator.
// Two accumulators (2 x 256 TMEM columns) let the epilogue of tile i (tcgen05.ld -> scale, bias, leaky
// -> 128-byte row stores) overlap the MMAs of tile i+1.  Persistent CTAs, static tile striding.
// Warp roles: 0 = TMA producer, 1 = TMEM owner + MMA issuer, 2..5 = epilogue (TMEM lane quarter w % 4).
#include <cstring>

#include "fc1_tc.h"
#include "kernels.h"
#include "ptx_sm100.cuh"

namespace ethcnn {
namespace {

constexpr int kBM = 128, kBN = 224, kBK = 64;
constexpr int kStages = 2;
constexpr int kABytes = kBM * kBK * 2;   // 16384
constexpr int kBBytes = kBN * kBK * 2;   // 28672
constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;  // 90112
constexpr int kKSteps = kFeat / kBK;     // 42
constexpr int kAccCols = 256;            // TMEM columns reserved per accumulator
constexpr int kThreads = 192;
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*alignment slack*/ + 256 /*barriers*/;

static_assert(kFeat % kBK == 0, "K must be a multiple of the slice");
static_assert(kBN % 16 == 0 && kBN <= 256, "invalid UMMA N for M = 128");

// K-major operand tile in 128-byte-swizzled shared memory: rows of 64 fp16 (128 B), 8-row groups of
// 1024 B (stride byte offset), start address in 16-byte units, descriptor version 1 (sm_100),
// layout type 2 = SWIZZLE_128B.  The leading byte offset is unused for swizzled K-major tiles.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3ffffu) >> 4);       // start address  [0,14)
  d |= uint64_t(1) << 16;                           // leading byte offset (ignored) [16,30)
  d |= uint64_t(1024 >> 4) << 32;                   // stride byte offset [32,46)
  d |= uint64_t(1) << 46;                           // version [46,48)
  d |= uint64_t(2) << 61;                           // layout type [61,64)
  return d;
}

// Instruction descriptor, kind::f16: D = fp32 (bits 4-5 = 1), A = B = fp16 (format 0), both K-major,
// N >> 3 at bits 17-22, M >> 4 at bits 24-28.
constexpr uint32_t kInstrDesc = (1u << 4) | (uint32_t(kBN >> 3) << 17) | (uint32_t(kBM >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(kInstrDesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float leaky(float v) { return fmaxf(0.2f * v, v); }

__global__ void __launch_bounds__(kThreads, 1)
fc1_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
              const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
              const float* __restrict__ b1, float unscale, float* __restrict__ out, int n_ctus, int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full = bars;                   // [kStages]  TMA -> MMA
  uint64_t* empty = bars + kStages;        // [kStages]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * kStages; // [2]        MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;      // [2]        epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a_hi), prefetch_tmap(&map_a_lo), prefetch_tmap(&map_b_hi), prefetch_tmap(&map_b_lo);
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], 1);
    for (int a = 0; a < 2; ++a) mbar_init(&acc_full[a], 1), mbar_init(&acc_empty[a], 4);
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int m0 = (t >> 1) * kBM, n0 = (t & 1) * kBN;
        for (int ks = 0; ks < kKSteps; ++ks, ++it) {
          const int s = it % kStages;
          mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1);
          uint8_t* st = smem + s * kStageBytes;
          mbar_arrive_expect_tx(&full[s], kStageBytes);
          tma_load_2d(st, &map_a_hi, &full[s], ks * kBK, m0);
          tma_load_2d(st + kABytes, &map_a_lo, &full[s], ks * kBK, m0);
          tma_load_2d(st + 2 * kABytes, &map_b_hi, &full[s], ks * kBK, n0);
          tma_load_2d(st + 2 * kABytes + kBBytes, &map_b_lo, &full[s], ks * kBK, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    int it = 0, tile_i = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tile_i) {
      const int acc = tile_i & 1;
      mbar_wait(&acc_empty[acc], ((tile_i >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kAccCols;
      for (int ks = 0; ks < kKSteps; ++ks, ++it) {
        const int s = it % kStages;
        mbar_wait(&full[s], (it / kStages) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_hi = smem_u32(smem + s * kStageBytes);
          const uint64_t da_hi = umma_desc_sw128(a_hi), da_lo = umma_desc_sw128(a_hi + kABytes);
          const uint64_t db_hi = umma_desc_sw128(a_hi + 2 * kABytes), db_lo = umma_desc_sw128(a_hi + 2 * kABytes + kBBytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t adv = uint64_t(k * 32 >> 4);  // 16 fp16 = 32 bytes along K inside the swizzle atom
            umma_f16(d_tmem, da_hi + adv, db_hi + adv, (ks | k) != 0);
            umma_f16(d_tmem, da_hi + adv, db_lo + adv, 1);
            umma_f16(d_tmem, da_lo + adv, db_hi + adv, 1);
          }
          umma_commit(&empty[s]);                               // frees the smem stage when these MMAs retire
          if (ks == kKSteps - 1) umma_commit(&acc_full[acc]);   // accumulator complete
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------ epilogue ------------------------------
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int tile_i = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tile_i) {
      const int acc = tile_i & 1;
      const int m0 = (t >> 1) * kBM, n0 = (t & 1) * kBN;
      mbar_wait(&acc_full[acc], (tile_i >> 1) & 1);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      float* orow = out + size_t(row) * kFc1 + n0;
      const uint32_t taddr = tmem_base + acc * kAccCols + (uint32_t(q * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < kBN; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + c0, r);
        if (row < n_ctus) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(b1 + n0 + c0 + j));
            float4 o;
            o.x = leaky(fmaf(__uint_as_float(r[j]), unscale, b.x));
            o.y = leaky(fmaf(__uint_as_float(r[j + 1]), unscale, b.y));
            o.z = leaky(fmaf(__uint_as_float(r[j + 2]), unscale, b.z));
            o.w = leaky(fmaf(__uint_as_float(r[j + 3]), unscale, b.w));
            *reinterpret_cast<float4*>(orow + c0 + j) = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    cudaDriverEntryPointQueryResult qres;
    void* f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  return fn;
}

// [rows][2688] fp16 row-major -> 2-D map (inner = K), box 64 x box_rows, 128-byte swizzle.
bool make_kmajor_map(CUtensorMap* map, const __half* base, uint64_t rows, uint32_t box_rows, const char** err) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    *err = "cuTensorMapEncodeTiled not available from the driver";
    return false;
  }
  cuuint64_t dims[2] = {cuuint64_t(kFeat), rows};
  cuuint64_t strides[1] = {cuuint64_t(kFeat) * 2};
  cuuint32_t box[2] = {cuuint32_t(kBK), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    *err = "cuTensorMapEncodeTiled failed";
    return false;
  }
  return true;
}

}  // namespace

bool fc1_tc_prepare_weights(const __half* w_hi, const __half* w_lo, Fc1TcWeights* out, const char** err) {
  out->valid = make_kmajor_map(&out->map_hi, w_hi, kFc1, kBN, err) && make_kmajor_map(&out->map_lo, w_lo, kFc1, kBN, err);
  return out->valid;
}

cudaError_t fc1_tc_configure() {
  return cudaFuncSetAttribute(fc1_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
}

cudaError_t launch_fc1_tc(const __half* feat_hi, const __half* feat_lo, const Fc1TcWeights& w, const float* b1, float unscale,
                          float* fc1_out, int n_ctus, int sm_count, cudaStream_t stream) {
  if (n_ctus <= 0) return cudaSuccess;
  if (!w.valid) return cudaErrorInvalidValue;
  const int m_tiles = (n_ctus + kBM - 1) / kBM;
  const int n_tiles = 2 * m_tiles;
  CUtensorMap map_a_hi, map_a_lo;
  const char* err = nullptr;
  if (!make_kmajor_map(&map_a_hi, feat_hi, uint64_t(m_tiles) * kBM, kBM, &err) ||
      !make_kmajor_map(&map_a_lo, feat_lo, uint64_t(m_tiles) * kBM, kBM, &err))
    return cudaErrorInvalidValue;
  const int grid = n_tiles < sm_count ? n_tiles : sm_count;
  fc1_tc_kernel<<<grid, kThreads, kSmemBytes, stream>>>(map_a_hi, map_a_lo, w.map_hi, w.map_lo, b1, unscale, fc1_out, n_ctus,
                                                       n_tiles);
  return cudaGetLastError();
}

}  // namespace ethcnn
