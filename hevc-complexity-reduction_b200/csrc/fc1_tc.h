// Stage FC1 on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
// See fc1_tc.cu for the design; net_CNN.py:156,166,178 for the arithmetic.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace ethcnn {

struct Fc1TcWeights {
  CUtensorMap map_hi;  // [448][2688] fp16, K-major, box 64 x 224, 128-byte swizzle
  CUtensorMap map_lo;
  bool valid = false;
};

// Builds the TMA descriptors over the packed FC1 weights (device pointers).
bool fc1_tc_prepare_weights(const __half* w_hi, const __half* w_lo, Fc1TcWeights* out, const char** err);

cudaError_t fc1_tc_configure();

// fc1_out[n][448] = leaky(unscale * (Ahi*Bhi + Ahi*Blo + Alo*Bhi) + b1).  feat_hi/lo must be allocations
// whose row count is padded to a multiple of 128 (rows >= n_ctus are read but never stored).
cudaError_t launch_fc1_tc(const __half* feat_hi, const __half* feat_lo, const Fc1TcWeights& w, const float* b1, float unscale,
                          float* fc1_out, int n_ctus, int sm_count, cudaStream_t stream);

}  // namespace ethcnn
