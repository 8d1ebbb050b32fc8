// Device-side stages of the ETH-CNN forward pass (sm_100a) and their host launchers.
//   stage CONV  : 64x64 luma tiles (TMA) -> mean removal -> three conv branches -> 2688 features,
//                 stored as scaled fp16 hi/lo pairs                      (net_CNN.py:105-150)
//   stage FC1   : [n,2688] x [2688,448] + bias + leaky -> [n,448] fp32    (net_CNN.py:156,166,178)
//                 tcgen05 (3-pass hi/lo split fp16, fp32 accumulate in TMEM) or SIMT fp32
//   stage HEADS : FC2 + FC3 + sigmoid, raw probabilities + gate flags      (net_CNN.py:158-185)
//   stage GATE  : per <=1024-CTU sub-batch gates                           (net_CNN.py:175,187)
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace ethcnn {

constexpr int kCtu = 64;
constexpr int kFeat = 2688;        // net_CNN.py:27
constexpr int kFc1 = 448;          // 64 + 128 + 256
constexpr int kProbs = 21;
constexpr int kSubBatch = 1024;    // video_to_cu_depth.py:64

// feature offsets inside the 2688-vector (net_CNN.py:150 concat order)
constexpr int kOffC3S = 0, kOffC3M = 512, kOffC3L = 640, kOffC2S = 672, kOffC2M = 2208, kOffC2L = 2592;

// conv weight block of one branch as laid out in shared memory (32-bit words).  Filters are stored as
// mma.sync.m16n8k16 B fragments (fp16 hi and lo halves of w * 2^e): fragment (k-step j, n-tile nt) is
// 32 lanes x 2 registers; lane (g = lane / 4, d = lane % 4) holds {W[16j+2d][8nt+g], W[16j+2d+1][8nt+g]}
// and {W[16j+2d+8][8nt+g], W[16j+2d+9][8nt+g]} with K in the natural TF order (ky, kx, ci).
constexpr int kHdrOff = 0;       // floats: [0] 32 * 2^-e1w, [1] 2^-(e_c1 + e2w), [2] 2^-(feat_exp + e3w), [3] 2^e_c1
constexpr int kB1Off = 16;       // [16] conv1 bias
constexpr int kB2Off = 32;       // [24] conv2 bias
constexpr int kB3Off = 56;       // [32] conv3 bias
constexpr int kF1HiOff = 96;     // conv1:  1 k-step  x 2 n-tiles x 64 words
constexpr int kF1LoOff = 224;
constexpr int kF2HiOff = 352;    // conv2:  4 k-steps x 3 n-tiles
constexpr int kF2LoOff = 1120;
constexpr int kF3HiOff = 1888;   // conv3:  6 k-steps x 4 n-tiles
constexpr int kF3LoOff = 3424;
constexpr int kW1SumOff = 4960;  // [16] floats: per conv1 output channel, the sum over the 16 taps of (hi + lo) as stored above
constexpr int kConvBranchFloats = 4976;
constexpr int kConvFloats = 3 * kConvBranchFloats;  // branch order S, M, L

// conv-stage tiling
constexpr int kGroupCtus = 16;    // CTUs per shared-memory tile group
constexpr int kGroupTasks = 21;   // 16 S + 4 M + 1 L warp tasks per group (two regions per lane)
constexpr int kConvStages = 2;
#ifndef ETHCNN_CONV_COMPUTE_WARPS
#define ETHCNN_CONV_COMPUTE_WARPS 11
#endif
constexpr int kConvComputeWarps = ETHCNN_CONV_COMPUTE_WARPS;  // + 1 producer warp; 12 warps -> 168 registers, 16 warps -> 128
constexpr int kConvThreads = 32 * (1 + kConvComputeWarps);

struct ConvLaunch {
  const float* convw;        // device [kConvFloats]
  __half* feat_hi;           // device [n_ctus][kFeat], value * feat_scale rounded to fp16
  __half* feat_lo;           // device [n_ctus][kFeat], residual of the above
  int n_ctus;                // CTUs in this launch
  int dump_row;              // a feature row nobody reads (stores of a tail group's absent CTUs go there)
  int ctu_begin;             // global index (frame-major raster) of the first CTU of this launch
  int ctus_per_row;
  int ctus_per_frame;
  float cst[3];              // S, M, L: input_scale / (256 * pool^2)
  float feat_scale;          // power of two
  unsigned* clear_flags;     // gate flags to zero before anything downstream sets them (first chunk of a call), or nullptr
  int n_clear_flags;
  // plain-load tile loader (used when the TMA preconditions do not hold)
  const uint8_t* luma;
  size_t pitch, frame_stride;
  int width, height;
};

struct HeadWeights {         // one head (64 / 32 / 16); all device pointers
  const float* w2;           // [n1][n2]
  const float* w2q;          // [n2]   the qp row of the FC2 matrix
  const float* b2;           // [n2]
  const float* w3;           // [n2][n3]
  const float* w3q;          // [n3]
  const float* b3;           // [n3]
};

struct HeadsLaunch {
  const float* fc1;          // [n_ctus][fc1_stride] input activations (FC1 output, or the LSTM h for the LDP step)
  int fc1_stride;            // floats between rows (448, or 896 when reading h out of the LSTM state rows)
  HeadWeights head[3];
  float q;                   // scaled qp
  float* prob;               // [total][21], row of CTU i at prob + (ctu_begin + i) * 21
  unsigned* flags;           // [n_frames * chunks_per_frame], bit0: any y64 > t1, bit1: any y32 > t2
  float t1, t2;
  int n_ctus, ctu_begin, ctus_per_frame, chunks_per_frame;
};

// CUtensorMap over the luma planes: dims (width, height, n_frames), box 64 x 64 x 1, zero OOB fill.
bool make_luma_tensor_map(CUtensorMap* map, const uint8_t* d_y, int width, int height, int n_frames, size_t pitch,
                          size_t frame_stride, const char** err);

cudaError_t launch_conv_features(const CUtensorMap* tmap /* nullptr = plain loads */, const ConvLaunch& p, int sm_count,
                                 cudaStream_t stream);
cudaError_t conv_features_configure();  // one-off cudaFuncSetAttribute calls

cudaError_t launch_fc1_simt(const __half* feat_hi, const __half* feat_lo, float inv_feat_scale, const float* w1 /*[2688][448]*/,
                            const float* b1, float* fc1_out, int n_ctus, cudaStream_t stream);

cudaError_t heads_configure();
cudaError_t launch_heads(const HeadsLaunch& p, cudaStream_t stream);

cudaError_t launch_gate(float* prob, const unsigned* flags, float t2, long long n_total, int ctus_per_frame,
                        int chunks_per_frame, cudaStream_t stream);

// Gates applied while copying the raw probabilities from a staging buffer to their destination with coalesced stores
// (the destination may be peer memory over NVLink); flags == nullptr copies without gates.
cudaError_t launch_gate_export(const float* src, float* dst, const unsigned* flags, float t2, long long n_total,
                               int ctus_per_frame, int chunks_per_frame, cudaStream_t stream);

// Gates + HM's threshold rule on the device: float rows to dst (in place when dst == src) and one 64-bit word of 21 two-bit
// decisions per CTU to map (dense_stages.cu: gate_map_kernel).  thr6 = up, down per depth.  flags == nullptr: no gates.
cudaError_t launch_gate_map(const float* src, float* dst, const unsigned* flags, float t2, long long n_total, int ctus_per_frame,
                            int chunks_per_frame, unsigned long long* map, const float thr6[6], cudaStream_t stream);

// One LSTM step for the three heads (lstm_stage.cu).  state rows are [c(448) | h(448)]; z_scratch is [rows][1792].
cudaError_t launch_lstm_step(const float* fc1, const float* state_in, float* state_out, float* z_scratch,
                             const float* const kernel[3], const float* const bias[3], int rows, cudaStream_t stream);

cudaError_t launch_decisions(const float* prob, unsigned char* dec, long long n_values, const float* thr6_dev,
                             cudaStream_t stream);

}  // namespace ethcnn
