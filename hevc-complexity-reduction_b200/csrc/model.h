// Host-side packing of one ETH-CNN checkpoint into the layouts the kernels consume.
// Tensor names / roles: SURVEY.md section 8(c) table; creation order in net_CNN.py:126-141
// (L = Variable..Variable_5, M = Variable_6..11, S = Variable_12..17), FC tensors named
// h_fc1__{64,32,16}__{w,b}, h_fc2__..., y_conv_flat__....
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "tf_bundle.h"

namespace ethcnn {

struct PackedModel {
  std::vector<float> conv;        // [3 branches S,M,L][kConvBranchFloats], see kernels.h
  std::vector<uint8_t> conv_tc;   // [3 branches S,M,L][kTcBranchBytes]: the same filters as swizzled UMMA tiles, see conv_tc.h
  std::vector<float> w1;          // [2688][448] fp32, heads side by side (64 | 128 | 256)
  std::vector<float> b1;          // [448]
  std::vector<uint16_t> w1_hi;    // [448][2688] fp16 bits of w1 * 2^w_exp (K-major, for tcgen05)
  std::vector<uint16_t> w1_lo;    // [448][2688] fp16 bits of the residual
  int conv_exp[3][4] = {};        // per branch: e1w, e_c1, e2w, e3w (power-of-two scales of the conv operands)
  int feat_exp = 0;               // features are stored as value * 2^feat_exp (split into fp16 hi + lo)
  int w_exp = 0;
  float feat_bound = 0.f;         // rigorous bound on |feature| for inputs |x| <= input_bound
  // heads: index 0,1,2 = 64,32,16
  std::vector<float> w2[3], w2q[3], b2[3], w3[3], w3q[3], b3[3];
  // FC2 for the fused tensor-core kernel: per head [n2][n1] fp16 bits of w2^T * 2^w2_exp (K-major) and residual
  std::vector<uint16_t> w2_hi[3], w2_lo[3];
  int a1_exp = 0;                 // FC1 activations are re-split as a1 * 2^a1_exp
  int w2_exp = 0;
  float a1_bound = 0.f;           // rigorous bound on |a1|
};

// input_bound: max |x| after scaling and mean removal (1.0 for AI, 10.0 for LDP).
bool pack_model(const std::map<std::string, BundleTensor>& t, float input_bound, PackedModel* out, std::string* err);

uint16_t f32_to_f16_bits(float f);   // round-to-nearest-even, IEEE binary16 incl. subnormals
float f16_bits_to_f32(uint16_t h);

}  // namespace ethcnn
