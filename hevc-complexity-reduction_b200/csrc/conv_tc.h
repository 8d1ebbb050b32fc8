// Stage CONV on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a; see conv_tc.cu.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace ethcnn {

// Per-branch weight image as it sits in shared memory (byte offsets).  Every filter is a K-major fp16 tile in the
// 128-byte-swizzled UMMA layout, rows = output channels, 64 K values per row:
//   element (n, k) at (n / 8) * 1024 + (n % 8) * 128 + (((k / 8) ^ (n % 8)) * 16) + (k % 8) * 2
// hi = fp16(w * 2^e), lo = fp16(w * 2^e - hi), the exponents of model.cpp (conv_exp).
constexpr int kTcW1Hi = 0;          // [16][64], K = tap (ky * 4 + kx) < 16 used
constexpr int kTcW1Lo = 2048;
constexpr int kTcW2Hi = 4096;       // [32][64], rows < 24 used, K = patch * 16 + ci
constexpr int kTcW2Lo = 8192;
constexpr int kTcW3Hi = 12288;      // 2 tiles [32][64]: K = region * 24 + ci = 64 * tile + k, K < 96 used
constexpr int kTcW3Lo = 20480;
constexpr int kTcTab = 28672;       // floats: b1 * 2^e_c1 [16] | conv1 tap sums as stored [16] | b2 * 2^feat_exp [24] | b3 * 2^feat_exp [32]
constexpr int kTcBranchBytes = 29696;   // 29 KB, keeps the next block 1024-byte aligned
constexpr int kTcBlobBytes = 3 * kTcBranchBytes;   // branch order S, M, L

struct ConvTcLaunch {
  const uint8_t* blob;       // device [kTcBlobBytes]
  __half* feat_hi;           // device [rows][2688]
  __half* feat_lo;
  int n_ctus;                // CTUs in this launch
  int ctu_begin;             // global index (frame-major raster) of the first CTU
  int ctus_per_row, ctus_per_frame;
  // per branch (S, M, L): epilogue factors, see conv_tc.cu
  float u1x8[3];             // conv1: activation * 2^e_c1 per unit of sum((s - centre) * w_stored)
  float u1[3];               // mean term: (1024 pool^2 - W / 32) * u1 multiplies the tap sums
  float u2[3];               // conv2: feature * 2^feat_exp per unit of the accumulator
  float u3[3];
  int phase_mask;            // bit b: run branch b (7 = all; anything else is for timing experiments only)
};

cudaError_t conv_tc_configure();
// Needs the TMA tile loader (16-byte aligned luma base / pitch / frame stride); the caller falls back to the mma.sync
// kernel (conv_stage.cu) otherwise.
cudaError_t launch_conv_tc(const CUtensorMap& tmap, const ConvTcLaunch& p, int sm_count, cudaStream_t stream);

}  // namespace ethcnn
