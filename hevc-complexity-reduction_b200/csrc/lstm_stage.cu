// One step of the three ETH-LSTM cells of the LDP predictor (sm_100a, fp32 SIMT).
//   HM-16.5_Test_LDP/bin/net_CNN_LSTM_one_step.py:201-264: tf.contrib.rnn.LSTMCell(n, forget_bias=1, cell_clip=5),
//   one layer, one time step per encoded frame:
//     z = [x, h_prev] K + b;  (i, j, f, o) = split(z, 4)
//     c = sigmoid(f + 1) c_prev + sigmoid(i) tanh(j), clipped to +-5;  h = sigmoid(o) tanh(c)
//   x = the head's slice of the 448-vector (FC1 activations), state rows = [c(448) | h(448)], heads 64|128|256.
// This is the per-frame latency path (one frame = a few hundred CTUs per call, 0.76 MFLOP each), so it is a
// plain tiled fp32 GEMM + an elementwise gate kernel; the FC2 / FC3 heads reuse heads_kernel with h as input.
#include "kernels.h"

namespace ethcnn {
namespace {

constexpr int kTB = 64, kTK = 16;

// z[r][zoff + col] = sum_k A[r][k] * K[k][col] + bias[col],  A[r] = [x (n) | h_prev (n)],  col < 4n
__global__ void __launch_bounds__(256) lstm_gemm_kernel(const float* __restrict__ x, int x_stride, const float* __restrict__ hprev,
                                                        int h_stride, const float* __restrict__ kernel, const float* __restrict__ bias,
                                                        float* __restrict__ z, int z_stride, int n_units, int rows) {
  __shared__ float As[kTK][kTB + 4];
  __shared__ float Bs[kTK][kTB];
  const int row0 = blockIdx.x * kTB, col0 = blockIdx.y * kTB;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int n4 = 4 * n_units, kdim = 2 * n_units;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int ar = threadIdx.x >> 2, ak = (threadIdx.x & 3) * 4;
  const int bk = threadIdx.x >> 4, bc = (threadIdx.x & 15) * 4;
  for (int k0 = 0; k0 < kdim; k0 += kTK) {
    const int r = row0 + ar;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < rows) {
      const int k = k0 + ak;  // n_units is a multiple of 16, so a K tile never straddles x | h_prev
      a = (k < n_units) ? *reinterpret_cast<const float4*>(x + size_t(r) * x_stride + k)
                        : *reinterpret_cast<const float4*>(hprev + size_t(r) * h_stride + (k - n_units));
    }
    As[ak][ar] = a.x, As[ak + 1][ar] = a.y, As[ak + 2][ar] = a.z, As[ak + 3][ar] = a.w;
    *reinterpret_cast<float4*>(&Bs[bk][bc]) = *reinterpret_cast<const float4*>(kernel + size_t(k0 + bk) * n4 + col0 + bc);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kTK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float4 b = *reinterpret_cast<const float4*>(bias + col0 + tx * 4);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row0 + ty * 4 + i;
    if (r < rows)
      *reinterpret_cast<float4*>(z + size_t(r) * z_stride + col0 + tx * 4) =
          make_float4(acc[i][0] + b.x, acc[i][1] + b.y, acc[i][2] + b.z, acc[i][3] + b.w);
  }
}

__device__ __forceinline__ float sigm(float v) { return 1.0f / (1.0f + expf(-v)); }

// gates for all three heads: thread = (row, unit u of 448)
__global__ void lstm_gate_kernel(const float* __restrict__ z, const float* __restrict__ state_in, float* __restrict__ state_out,
                                 int rows) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * kFc1) return;
  const int r = idx / kFc1, u = idx - r * kFc1;
  const int off = u < 64 ? 0 : (u < 192 ? 64 : 192), n = u < 64 ? 64 : (u < 192 ? 128 : 256);
  const float* zr = z + size_t(r) * (4 * kFc1) + 4 * off;   // the head's 4n gate pre-activations
  const int lu = u - off;
  const float gi = zr[lu], gj = zr[n + lu], gf = zr[2 * n + lu], go = zr[3 * n + lu];
  const float c_prev = state_in[size_t(r) * (2 * kFc1) + u];
  float c = sigm(gf + 1.0f) * c_prev + sigm(gi) * tanhf(gj);
  c = fminf(fmaxf(c, -5.0f), 5.0f);
  state_out[size_t(r) * (2 * kFc1) + u] = c;
  state_out[size_t(r) * (2 * kFc1) + kFc1 + u] = sigm(go) * tanhf(c);
}

}  // namespace

cudaError_t launch_lstm_step(const float* fc1, const float* state_in, float* state_out, float* z_scratch,
                             const float* const kernel[3], const float* const bias[3], int rows, cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  const int n[3] = {64, 128, 256}, off[3] = {0, 64, 192};
  for (int h = 0; h < 3; ++h) {
    dim3 grid((rows + kTB - 1) / kTB, 4 * n[h] / kTB);
    lstm_gemm_kernel<<<grid, 256, 0, stream>>>(fc1 + off[h], kFc1, state_in + kFc1 + off[h], 2 * kFc1, kernel[h], bias[h],
                                               z_scratch + 4 * off[h], 4 * kFc1, n[h], rows);
  }
  const int total = rows * kFc1;
  lstm_gate_kernel<<<(total + 255) / 256, 256, 0, stream>>>(z_scratch, state_in, state_out, rows);
  return cudaGetLastError();
}

}  // namespace ethcnn
