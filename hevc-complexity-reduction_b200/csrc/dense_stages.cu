// Stages FC1 (SIMT fp32 variant), HEADS, GATE and the HM decision quantiser (sm_100a).
//   FC1   : net_CNN.py:156,166,178   a1 = leaky(f W1 + b1), the three heads side by side (448 columns)
//   HEADS : net_CNN.py:158-163,168-173,180-185   a2 = leaky([a1,q] W2 + b2); y = sigmoid([a2,q] W3 + b3)
//   GATE  : net_CNN.py:175,187 over each <=1024-CTU sub-batch of one frame (video_to_cu_depth.py:64-70)
//   decisions: TLibEncoder/TEncCu.cpp:448-462
#include <cstdlib>

#include "kernels.h"
#include "ptx_sm100.cuh"

namespace ethcnn {
namespace {

__device__ __forceinline__ float leaky(float v) { return fmaxf(0.2f * v, v); }

// ------------------------------------------------------------------------------------------------
// FC1, SIMT fp32: C[n][448] = leaky(A[n][2688] * W[2688][448] + b), A rebuilt from its hi/lo halves.
// 64 x 64 output tile per CTA, 256 threads, 4 x 4 outputs per thread, K tile 16.
constexpr int kSB = 64, kSK = 16;

__global__ void __launch_bounds__(256) fc1_simt_kernel(const __half* __restrict__ fhi, const __half* __restrict__ flo,
                                                       float inv_scale, const float* __restrict__ w1,
                                                       const float* __restrict__ b1, float* __restrict__ out, int n) {
  __shared__ float As[kSK][kSB + 4];
  __shared__ float Bs[kSK][kSB];
  const int row0 = blockIdx.x * kSB, col0 = blockIdx.y * kSB;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // loader mapping: A tile 64 rows x 16 k -> thread loads 4 consecutive k of one row (8 bytes of hi, of lo)
  const int ar = threadIdx.x >> 2, ak = (threadIdx.x & 3) * 4;
  // B tile 16 k x 64 cols -> thread loads 4 consecutive cols of one k row
  const int bk = threadIdx.x >> 4, bc = (threadIdx.x & 15) * 4;

  for (int k0 = 0; k0 < kFeat; k0 += kSK) {
    {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      const int r = row0 + ar;
      if (r < n) {
        const uint2 h = *reinterpret_cast<const uint2*>(fhi + size_t(r) * kFeat + k0 + ak);
        const uint2 l = *reinterpret_cast<const uint2*>(flo + size_t(r) * kFeat + k0 + ak);
        const __half2 h0 = *reinterpret_cast<const __half2*>(&h.x), h1 = *reinterpret_cast<const __half2*>(&h.y);
        const __half2 l0 = *reinterpret_cast<const __half2*>(&l.x), l1 = *reinterpret_cast<const __half2*>(&l.y);
        v[0] = (__low2float(h0) + __low2float(l0)) * inv_scale;
        v[1] = (__high2float(h0) + __high2float(l0)) * inv_scale;
        v[2] = (__low2float(h1) + __low2float(l1)) * inv_scale;
        v[3] = (__high2float(h1) + __high2float(l1)) * inv_scale;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) As[ak + i][ar] = v[i];
      const float4 w = *reinterpret_cast<const float4*>(w1 + size_t(k0 + bk) * kFc1 + col0 + bc);
      *reinterpret_cast<float4*>(&Bs[bk][bc]) = w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float4 bias = *reinterpret_cast<const float4*>(b1 + col0 + tx * 4);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row0 + ty * 4 + i;
    if (r < n) {
      float4 o;
      o.x = leaky(acc[i][0] + bias.x), o.y = leaky(acc[i][1] + bias.y);
      o.z = leaky(acc[i][2] + bias.z), o.w = leaky(acc[i][3] + bias.w);
      *reinterpret_cast<float4*>(out + size_t(r) * kFc1 + col0 + tx * 4) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// HEADS: one CTA = 32 CTUs of one head.  FC2 runs on 4-column x 8-CTU register tiles (weights streamed
// once through registers, activations broadcast from shared memory), then the 32 x N3 FC3 dot products
// are spread over the block.
constexpr int kHT = 32;  // CTUs per CTA

template <int N1, int N2, int N3, int OUT_OFF, int COL_OFF>
__device__ __forceinline__ void heads_body(const HeadsLaunch& p, const HeadWeights& w, float* smem) {
  const int NT = blockDim.x;       // every thread of the CTA takes part in the loads and barriers
  float* a1s = smem;               // [N1][32]
  float* a2s = smem + N1 * kHT;    // [N2][33]
  const int n0 = blockIdx.x * kHT;
  const int tid = threadIdx.x;
  for (int i = tid; i < (N1 / 4) * kHT; i += NT) {
    const int c = i % kHT, k4 = i / kHT;  // CTU fastest: conflict-free transposed store
    const int n = n0 + c;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < p.n_ctus) v = *reinterpret_cast<const float4*>(p.fc1 + size_t(n) * p.fc1_stride + COL_OFF + 4 * k4);
    a1s[(4 * k4 + 0) * kHT + c] = v.x;
    a1s[(4 * k4 + 1) * kHT + c] = v.y;
    a1s[(4 * k4 + 2) * kHT + c] = v.z;
    a1s[(4 * k4 + 3) * kHT + c] = v.w;
  }
  __syncthreads();
  // FC2: thread = 4 output columns x 8 CTUs (register tile): per k one 128-bit weight load (coalesced
  // across threads) and two broadcast 128-bit activation loads feed 32 FMAs.
  constexpr int CG = N2 / 4;                      // column groups
  if (tid < CG * 4) {
    const int cg = tid % CG, rg = tid / CG;
    float acc[4][8];
    {
      const float4 bq = __ldg(reinterpret_cast<const float4*>(w.w2q) + cg), bb = __ldg(reinterpret_cast<const float4*>(w.b2) + cg);
      const float init[4] = {fmaf(p.q, bq.x, bb.x), fmaf(p.q, bq.y, bb.y), fmaf(p.q, bq.z, bb.z), fmaf(p.q, bq.w, bb.w)};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = init[i];
    }
    const float4* wp = reinterpret_cast<const float4*>(w.w2) + cg;
    // 16 independent weight loads in flight per thread: the FC2 matrices live in L2, not L1
#pragma unroll 16
    for (int k = 0; k < N1; ++k) {
      const float4 wv = __ldg(wp + k * CG);
      const float4 a0 = *reinterpret_cast<const float4*>(a1s + k * kHT + 8 * rg);
      const float4 a1 = *reinterpret_cast<const float4*>(a1s + k * kHT + 8 * rg + 4);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float wq[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[j], wq[i], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) a2s[(4 * cg + i) * (kHT + 1) + 8 * rg + j] = leaky(acc[i][j]);
  }
  __syncthreads();
  for (int i = tid; i < kHT * N3; i += NT) {
    const int o = i / kHT, c = i - o * kHT;
    const int n = n0 + c;
    float acc = fmaf(p.q, w.w3q[o], w.b3[o]);
    for (int j = 0; j < N2; ++j) acc = fmaf(a2s[j * (kHT + 1) + c], __ldg(w.w3 + j * N3 + o), acc);
    const float y = 1.0f / (1.0f + expf(-acc));
    if (n < p.n_ctus) {
      const int gn = p.ctu_begin + n;
      p.prob[size_t(gn) * kProbs + OUT_OFF + o] = y;
      if (p.flags != nullptr && N3 <= 4) {
        const float thr = (N3 == 1) ? p.t1 : p.t2;
        if (y > thr) {
          const unsigned bit = (N3 == 1) ? 1u : 2u;
          const int f = gn / p.ctus_per_frame, r = gn - f * p.ctus_per_frame;
          unsigned* fl = p.flags + f * p.chunks_per_frame + r / kSubBatch;
          if (!(*reinterpret_cast<volatile unsigned*>(fl) & bit)) atomicOr(fl, bit);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(192) heads_kernel(const HeadsLaunch p) {
  extern __shared__ __align__(16) float hsm[];
  // blockIdx.y selects the head (warp-uniform); threads beyond a head's FC2 width only help with loads.
  if (blockIdx.y == 0) {
    heads_body<64, 48, 1, 0, 0>(p, p.head[0], hsm);
  } else if (blockIdx.y == 1) {
    heads_body<128, 96, 4, 1, 64>(p, p.head[1], hsm);
  } else {
    heads_body<256, 192, 16, 5, 192>(p, p.head[2], hsm);
  }
}

// ------------------------------------------------------------------------------------------------
// GATE (AI deployment): per sub-batch  g1 = any(y64 > t1);  y32 := g1 ? y32 : 0;
//                                      g2 = any(y32_gated > t2);  y16 := g2 ? y16 : 0.
// When g1 is false the gated y32 is all zeros, so g2 = (0 > t2).
__global__ void gate_kernel(float* __restrict__ prob, const unsigned* __restrict__ flags, float t2, int ctus_per_frame,
                            int chunks_per_frame) {
  pdl_wait();   // launched as a programmatic dependent of the FC kernel: its rows and flags are complete from here on
  // one block per (frame, sub-batch); almost always both gates are open and the block has nothing to do
  const int f = blockIdx.x / chunks_per_frame, ch = blockIdx.x - f * chunks_per_frame;
  const unsigned fl = flags[blockIdx.x];
  const bool g1 = (fl & 1u) != 0;
  const bool g2 = g1 ? ((fl & 2u) != 0) : (0.0f > t2);
  if (g1 && g2) return;
  const int r0 = ch * kSubBatch, rows = min(kSubBatch, ctus_per_frame - r0);
  float* base = prob + (size_t(f) * ctus_per_frame + r0) * kProbs;
  for (int i = threadIdx.x; i < rows * 20; i += blockDim.x) {
    const int n = i / 20, slot = 1 + i - n * 20;
    if (!((slot < 5) ? g1 : g2)) base[n * kProbs + slot] = 0.0f;
  }
}

// GATE + EXPORT: the same rule, reading the raw probabilities from a library-owned staging buffer and writing the final
// rows to `dst` with consecutive threads on consecutive floats (full 128-byte store transactions per warp).  Used when
// dst is PEER memory (another GPU's gather buffer mapped over NVLink, ethcnn_peer_buffer_open): the dense kernel's own
// 4-byte row-strided stores would cross the link as one small packet each.  flags == nullptr: no gates (LDP CNN).
__global__ void gate_export_kernel(const float* __restrict__ src, float* __restrict__ dst, const unsigned* __restrict__ flags,
                                   float t2, long long n_total, int ctus_per_frame, int chunks_per_frame) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;  // one thread per float
  if (i >= n_total * kProbs) return;
  float v = src[i];
  if (flags != nullptr) {
    const long long n = i / kProbs;
    const int slot = int(i - n * kProbs);
    if (slot > 0) {
      const long long f = n / ctus_per_frame;
      const int r = int(n - f * ctus_per_frame);
      const unsigned fl = flags[f * chunks_per_frame + r / kSubBatch];
      const bool g1 = (fl & 1u) != 0;
      const bool g2 = g1 ? ((fl & 2u) != 0) : (0.0f > t2);
      if (!((slot < 5) ? g1 : g2)) v = 0.0f;
    }
  }
  dst[i] = v;
}

// HM's use of one probability (TLibEncoder/TEncCu.cpp:448-457): "> up" = split only (2), "<= down" = no split (0),
// otherwise both candidates are checked (1).
__device__ __forceinline__ unsigned decide(float p, float up, float down) { return (p > up) ? 2u : ((p <= down) ? 0u : 1u); }

__global__ void decisions_kernel(const float* __restrict__ prob, unsigned char* __restrict__ dec, long long n_values,
                                 const float* __restrict__ thr6) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_values) return;
  const int slot = int(i % kProbs);
  const int lvl = slot == 0 ? 0 : (slot < 5 ? 1 : 2);
  dec[i] = (unsigned char)decide(prob[i], thr6[2 * lvl], thr6[2 * lvl + 1]);
}

// GATE + DECISION MAP: the gates as above, the finished float rows written to dst (== src: in place,
// only the gated slots are touched; != src: every float, consecutive threads on consecutive floats, as gate_export_kernel)
// and, next to them, HM's threshold rule applied on the device: one 64-bit word per CTU holding 21 two-bit decisions,
// bits [2k, 2k+1] = decide(row[k]) with k the position in the cu_depth.dat row (0: 64x64, 1..4: 32x32, 5..20: 16x16,
// TEncCu.cpp:434-447), bits 42..63 zero.  A block handles 256 CTUs through shared memory (row stride 21 words: conflict-free).
struct Thr6 {
  float v[6];   // up, down per depth (the AI order of Thr_info.txt, TEncCu.cpp:250)
};
constexpr int kMapTile = 256;
__global__ void __launch_bounds__(kMapTile) gate_map_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                            const unsigned* __restrict__ flags, float t2, long long n_total,
                                                            int ctus_per_frame, int chunks_per_frame,
                                                            unsigned long long* __restrict__ map, const Thr6 thr) {
  __shared__ float rows[kMapTile * kProbs];
  const long long n0 = blockIdx.x * (long long)kMapTile;
  const int n_here = int(min((long long)kMapTile, n_total - n0));
  for (int i = threadIdx.x; i < n_here * kProbs; i += kMapTile) {
    const int c = i / kProbs, slot = i - c * kProbs;
    float v = src[n0 * kProbs + i];
    bool zeroed = false;
    if (flags != nullptr && slot > 0) {
      const long long n = n0 + c;
      const long long f = n / ctus_per_frame;
      const int r = int(n - f * ctus_per_frame);
      const unsigned fl = flags[f * chunks_per_frame + r / kSubBatch];
      const bool g1 = (fl & 1u) != 0;
      const bool g2 = g1 ? ((fl & 2u) != 0) : (0.0f > t2);
      if (!((slot < 5) ? g1 : g2)) v = 0.0f, zeroed = true;
    }
    if (dst != src || zeroed) dst[n0 * kProbs + i] = v;
    rows[i] = v;
  }
  __syncthreads();
  if (int(threadIdx.x) < n_here) {
    const float* r = rows + threadIdx.x * kProbs;
    unsigned long long m = decide(r[0], thr.v[0], thr.v[1]);
#pragma unroll
    for (int k = 1; k < 5; ++k) m |= (unsigned long long)decide(r[k], thr.v[2], thr.v[3]) << (2 * k);
#pragma unroll
    for (int k = 5; k < kProbs; ++k) m |= (unsigned long long)decide(r[k], thr.v[4], thr.v[5]) << (2 * k);
    map[n0 + threadIdx.x] = m;
  }
}

}  // namespace

cudaError_t launch_fc1_simt(const __half* feat_hi, const __half* feat_lo, float inv_feat_scale, const float* w1,
                            const float* b1, float* fc1_out, int n_ctus, cudaStream_t stream) {
  if (n_ctus <= 0) return cudaSuccess;
  dim3 grid((n_ctus + kSB - 1) / kSB, kFc1 / kSB);
  fc1_simt_kernel<<<grid, 256, 0, stream>>>(feat_hi, feat_lo, inv_feat_scale, w1, b1, fc1_out, n_ctus);
  return cudaGetLastError();
}

constexpr int kHeadsSmem = (256 * kHT + 192 * (kHT + 1)) * 4;  // 58112 B

cudaError_t heads_configure() {  // function attributes are per device: call once on every device
  return cudaFuncSetAttribute(heads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHeadsSmem);
}

cudaError_t launch_heads(const HeadsLaunch& p, cudaStream_t stream) {
  if (p.n_ctus <= 0) return cudaSuccess;
  dim3 grid((p.n_ctus + kHT - 1) / kHT, 3);
  heads_kernel<<<grid, 192, kHeadsSmem, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_gate(float* prob, const unsigned* flags, float t2, long long n_total, int ctus_per_frame,
                        int chunks_per_frame, cudaStream_t stream) {
  if (n_total <= 0) return cudaSuccess;
  const long long n_frames = n_total / ctus_per_frame;   // a call always covers whole frames
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(unsigned(n_frames * chunks_per_frame)), cfg.blockDim = dim3(256), cfg.dynamicSmemBytes = 0, cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = getenv("ETHCNN_NO_PDL") ? 0 : 1;   // measurement switch
  return cudaLaunchKernelEx(&cfg, gate_kernel, prob, flags, t2, ctus_per_frame, chunks_per_frame);
}

cudaError_t launch_gate_export(const float* src, float* dst, const unsigned* flags, float t2, long long n_total,
                               int ctus_per_frame, int chunks_per_frame, cudaStream_t stream) {
  if (n_total <= 0) return cudaSuccess;
  const long long work = n_total * kProbs;
  gate_export_kernel<<<unsigned((work + 255) / 256), 256, 0, stream>>>(src, dst, flags, t2, n_total, ctus_per_frame, chunks_per_frame);
  return cudaGetLastError();
}

cudaError_t launch_gate_map(const float* src, float* dst, const unsigned* flags, float t2, long long n_total, int ctus_per_frame,
                            int chunks_per_frame, unsigned long long* map, const float thr6[6], cudaStream_t stream) {
  if (n_total <= 0) return cudaSuccess;
  Thr6 t;
  for (int i = 0; i < 6; ++i) t.v[i] = thr6[i];
  gate_map_kernel<<<unsigned((n_total + kMapTile - 1) / kMapTile), kMapTile, 0, stream>>>(src, dst, flags, t2, n_total, ctus_per_frame,
                                                                                        chunks_per_frame, map, t);
  return cudaGetLastError();
}

cudaError_t launch_decisions(const float* prob, unsigned char* dec, long long n_values, const float* thr6_dev,
                             cudaStream_t stream) {
  if (n_values <= 0) return cudaSuccess;
  decisions_kernel<<<unsigned((n_values + 255) / 256), 256, 0, stream>>>(prob, dec, n_values, thr6_dev);
  return cudaGetLastError();
}

}  // namespace ethcnn
