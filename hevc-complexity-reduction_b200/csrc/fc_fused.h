// Stages FC1 + FC2 + FC3 fused in one tcgen05 kernel (sm_100a); see fc_fused.cu.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace ethcnn {

struct FusedWeights {
  struct Maps {
    CUtensorMap w1_hi_t0, w1_lo_t0;  // [448][2688] fp16 K-major, box K x 192 (heads 64 + 32: rows 0..191)
    CUtensorMap w1_hi_t1, w1_lo_t1;  // same tensor, box K x 256 (head 16: rows 192..447)
    CUtensorMap w2_hi[3], w2_lo[3];  // per head [n2][n1] fp16 K-major, box K x n2
  };
  Maps maps[2];                      // [0]: one CTA per tile; [1]: CTA pairs, each CTA stages half of the box rows
  bool valid = false;
};

struct FusedParams {
  float b2eff[336];  // b2 + q * (qp row of W2): [48 | 96 | 192]
  float b3eff[21];   // b3 + q * (qp row of W3): [1 | 4 | 16]
  float unscale1;    // 2^-(feat_exp + w_exp)
  float a1_scale;    // 2^a1_exp
  float unscale2;    // 2^-(a1_exp + w2_exp)
  float t1, t2;
  const float* b1;   // [448]
  const float* w3;   // device [48*1 | 96*4 | 192*16]
  float* prob;       // [total][21] or nullptr
  float* fc1_out;    // [n][448] or nullptr (LDP tap)
  unsigned* flags;   // gate flags or nullptr
  int n_ctus, ctu_begin, ctus_per_frame, chunks_per_frame;
};

bool fc_fused_prepare_weights(const __half* w1_hi, const __half* w1_lo, const __half* const w2_hi[3],
                              const __half* const w2_lo[3], FusedWeights* out, const char** err);
cudaError_t fc_fused_configure();
// ctas_per_tile: 1 = a CTA per 128 CTUs; 2 = a CTA pair (tcgen05 cta_group::2) per 256 CTUs, half the weight traffic.
cudaError_t launch_fc_fused(const __half* feat_hi, const __half* feat_lo, const FusedWeights& w, const FusedParams& p,
                            int ctas_per_tile, int sm_count, cudaStream_t stream);

}  // namespace ethcnn
