// Stage CONV: luma tiles -> 2688 conv features per CTU (sm_100a).
//
// Follows net_CNN.py:105-150 (AI) / ETH-CNN_Training_LDP/net_CTU64.py:102-175 (LDP):
//   x = X * scale; L = meanremove16(avgpool4(x)); M = meanremove16(avgpool2(x)); S = meanremove16(x)
//   per branch: conv 4x4/s4 1->16, conv 2x2/s2 16->24, conv 2x2/s2 24->32, leaky(0.2) after each
//   features = [c3_S | c3_M | c3_L | c2_S | c2_M | c2_L], each NHWC-flattened.
//
// All three convolutions are non-overlapping, i.e. small GEMMs, and they run on the tensor cores through
// warp-level mma.sync.m16n8k16 with the accumulator fragment of one layer re-used, in registers, as the
// A fragment of the next (no shared-memory or shuffle traffic between layers):
//   * after pooling all branches look alike: a "region" = 8x8 pooled samples = 2x2 conv1 patches = one
//     conv2 position; a "quad" = 2x2 regions = one 16x16 mean-removal window = one conv3 position.
//     S: pool 1, 16 quads per CTU; M: pool 2, 4 quads per CTU; L: pool 4, 1 quad per CTU.
//   * a warp owns 16 quads (two sets A/B of 8, one per lane group g = lane / 4); the 4 lanes d = lane % 4
//     of a group split every K dimension the way the A fragment wants it:
//       conv1  rows = (region pair T, patch p), K = 16 taps: lane d supplies taps (ky = d/2 (+2), kx = 2(d%2)+{0,1})
//       conv2  rows = regions, K = 64 = (patch, channel): the conv1 C fragment (channels 2d,2d+1,8+2d,9+2d) IS the A fragment
//       conv3  rows = quads (set A rows 0-7, set B rows 8-15), K = 96 = (region, channel): the conv2 C fragments
//              of the four regions, in natural order, ARE the A fragments.
//     Weights sit in shared memory pre-arranged as B fragments (one conflict-free 64-bit load per lane).
//   * precision: every operand is split into fp16 hi + lo (after exact power-of-two scaling) and three MMAs
//     (hi*hi, hi*lo, lo*hi) accumulate in fp32 -- 2^-22 relative, like the FC stages.  The mean removal is
//     exact: with s = pooled integer sum and W = integer window sum, (256 s - W) / 32 is split exactly.
// A warp task is one CTU (S), four CTUs (M) or sixteen CTUs (L); a group of 16 CTUs is 16 + 4 + 1 = 21
// warp tasks of identical MMA count (312 mma.sync each).  Round-1 history: an FFMA version of this stage
// reached 39 % of the fp32 peak and was bound by shared-memory wavefronts for the broadcast weights
// (profiles/r01c_conv_v2.md); mma.sync does the same MACs 7.4x faster per SM (tools/microbench/hmma_rate.cu).
//
// A persistent CTA (one per SM) keeps the 58 KB of weight fragments resident in shared memory and streams
// 16-CTU tile groups through a 2-deep TMA ring (3-D tensor map over (x, y, frame), 64x64x1 box; the zero
// fill of out-of-bounds rows/columns IS the reference's zero padding, video_to_cu_depth.py:54-57).
// One producer warp issues TMA, eleven compute warps take warp tasks round-robin.
#include <cstring>

#include "kernels.h"
#include "ptx_sm100.cuh"

namespace ethcnn {
namespace {

__device__ __forceinline__ float leaky(float v) { return fmaxf(0.2f * v, v); }  // Maximum(alpha*x, x), alpha = 0.2

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint2 lds_u64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds_u128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) { return __uint_as_float(lds_u32(addr)); }
__device__ __forceinline__ float2 lds_f32x2(uint32_t addr) {
  const uint2 v = lds_u64(addr);
  return make_float2(__uint_as_float(v.x), __uint_as_float(v.y));
}

// D(16x8, fp32) += A(16x16, fp16) * B(16x8, fp16); fragment layouts as in the PTX ISA.
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint2 b) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}
// split-precision product: d += a_hi b_hi + a_hi b_lo + a_lo b_hi
__device__ __forceinline__ void mma3(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], const uint2 bh, const uint2 bl) {
  mma16816(d, ah, bh);
  mma16816(d, ah, bl);
  mma16816(d, al, bh);
}

// (s0, s1) -> packed fp16 hi pair and lo pair with hi + lo == s to ~22 bits
__device__ __forceinline__ void split2(float s0, float s1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(s0, s1);
  const __half2 l = __floats2half2_rn(s0 - __low2float(h), s1 - __high2float(h));
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// Integer sum of a quad's pixel block ((16 pool) x (16 pool) pixels at shared address blk, row pitch 64):
// lane d adds pooled rows d, d+4, d+8, d+12; the caller reduces over the quad.
__device__ __forceinline__ uint32_t block_rows_sum(int pool, uint32_t blk, int d) {
  uint32_t s = 0;
  const int chunks = pool;  // 16-byte chunks per pixel row
#pragma unroll 1
  for (int r = d; r < 16; r += 4) {
#pragma unroll 1
    for (int a = 0; a < pool; ++a) {
      const uint32_t row = blk + (pool * r + a) * kCtu;
#pragma unroll 1
      for (int k = 0; k < chunks; ++k) {
        const uint4 v = lds_u128(row + 16 * k);
        s = __dp4a(v.x, 0x01010101u, s);
        s = __dp4a(v.y, 0x01010101u, s);
        s = __dp4a(v.z, 0x01010101u, s);
        s = __dp4a(v.w, 0x01010101u, s);
      }
    }
  }
  return s;
}

// conv1 A operand of region pair T of a quad: for patch p, register reg = rx + 2*ki holds the pooled pair
// (y = 8T + 4(p/2) + d/2 + 2ki, x = 8rx + 4(p%2) + 2(d%2) + {0,1}) as exact fp16 hi/lo of (256 s - W) / 32.
__device__ __forceinline__ void cvt_pair(int s0, int s1, float wneg, uint32_t& hi, uint32_t& lo) {
  // (256 s - W) / 32 = 8 s - W / 32, exact in fp32 (|.| < 2^16 with 5 fractional bits); wneg = -W / 32
  const float f0 = fmaf(__int2float_rn(s0), 8.0f, wneg), f1 = fmaf(__int2float_rn(s1), 8.0f, wneg);
  const __half2 h = __floats2half2_rn(f0, f1);
  const __half2 l = __floats2half2_rn(f0 - __low2float(h), f1 - __high2float(h));
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

template <int P>
__device__ __forceinline__ void load_x_p(uint32_t blk, int T, int d, float wsum, uint32_t (&xh)[16], uint32_t (&xl)[16]) {
  const int ky0 = d >> 1, xs = 2 * (d & 1);
#pragma unroll
  for (int ph = 0; ph < 2; ++ph)
#pragma unroll
    for (int ki = 0; ki < 2; ++ki) {
      const int y = 8 * T + 4 * ph + ky0 + 2 * ki;
      if (P == 1) {
        const uint4 row = lds_u128(blk + y * kCtu);
        const uint32_t w[4] = {row.x, row.y, row.z, row.w};
#pragma unroll
        for (int pl = 0; pl < 2; ++pl)
#pragma unroll
          for (int rx = 0; rx < 2; ++rx) {
            const uint32_t hw = (w[2 * rx + pl] >> (8 * xs)) & 0xffffu;
            cvt_pair(int(hw & 0xffu), int(hw >> 8), wsum, xh[4 * (2 * ph + pl) + rx + 2 * ki], xl[4 * (2 * ph + pl) + rx + 2 * ki]);
          }
      } else {
#pragma unroll
        for (int pl = 0; pl < 2; ++pl)
#pragma unroll
          for (int rx = 0; rx < 2; ++rx) {
            const int x0 = 8 * rx + 4 * pl + xs;
            uint32_t s0 = 0, s1 = 0;
            if (P == 2) {
              const uint32_t a = lds_u32(blk + (2 * y) * kCtu + 2 * x0), b = lds_u32(blk + (2 * y + 1) * kCtu + 2 * x0);
              s0 = __dp4a(b, 0x00000101u, __dp4a(a, 0x00000101u, 0u));
              s1 = __dp4a(b, 0x01010000u, __dp4a(a, 0x01010000u, 0u));
            } else {
#pragma unroll
              for (int a = 0; a < 4; ++a) {
                const uint2 v = lds_u64(blk + (4 * y + a) * kCtu + 4 * x0);
                s0 = __dp4a(v.x, 0x01010101u, s0);
                s1 = __dp4a(v.y, 0x01010101u, s1);
              }
            }
            cvt_pair(int(s0), int(s1), wsum, xh[4 * (2 * ph + pl) + rx + 2 * ki], xl[4 * (2 * ph + pl) + rx + 2 * ki]);
          }
      }
    }
}

__device__ __forceinline__ void load_x(int pool, uint32_t blk, int T, int d, float wsum, uint32_t (&xh)[16], uint32_t (&xl)[16]) {
  if (pool == 1) {
    load_x_p<1>(blk, T, d, wsum, xh, xl);
  } else if (pool == 2) {
    load_x_p<2>(blk, T, d, wsum, xh, xl);
  } else {
    load_x_p<4>(blk, T, d, wsum, xh, xl);
  }
}

struct QuadSet {        // what a lane needs to know about its quad in set A or B
  uint32_t blk;         // shared-memory address of the quad's pixel block inside its CTU tile
  __half* hi;           // feature rows of the quad's CTU (global memory)
  __half* lo;
  int c2_off;           // offset of the 24 conv2 features of region 0 of the quad
  int c3_off;           // offset of the quad's 32 conv3 features
  bool valid;           // CTU exists (tail groups run with masked stores)
};

// One warp task: the conv stack for 16 quads (8 per set).
//   pool   1 / 2 / 4 (warp-uniform)    wb  shared-memory byte address of the branch's weight block
//   g24    feature distance between vertically adjacent regions (regions per row * 24)
__device__ __noinline__ void warp_task(const int pool, const uint32_t wb, const float cst, const float fscale, const int lane,
                                       const QuadSet qa, const QuadSet qb, const int g24) {
  const int d = lane & 3;
  // leaky(v) * 2^e == leaky(v * 2^e): the power-of-two operand scales are folded into the unscale factors and biases
  const float sc1 = lds_f32(wb + 4 * (kHdrOff + 3));
  const float u1 = lds_f32(wb + 4 * kHdrOff) * cst * sc1;
  const float u2 = lds_f32(wb + 4 * (kHdrOff + 1)) * fscale, u3 = lds_f32(wb + 4 * (kHdrOff + 2)) * fscale;
  // this lane's output channels: conv1 {2d, 2d+1, 8+2d, 9+2d}; conv2 / conv3 {8 nt + 2d, +1}
  float2 b1lo = lds_f32x2(wb + 4 * (kB1Off + 2 * d)), b1hi = lds_f32x2(wb + 4 * (kB1Off + 8 + 2 * d));
  b1lo.x *= sc1, b1lo.y *= sc1, b1hi.x *= sc1, b1hi.y *= sc1;
  uint2 f1h[2], f1l[2];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    f1h[nt] = lds_u64(wb + 4 * kF1HiOff + (nt * 32 + lane) * 8);
    f1l[nt] = lds_u64(wb + 4 * kF1LoOff + (nt * 32 + lane) * 8);
  }
  float d3[4][4];   // conv3 accumulators: rows 0-7 = quads of set A, rows 8-15 = quads of set B
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) d3[nt][0] = d3[nt][1] = d3[nt][2] = d3[nt][3] = 0.f;
  uint32_t keep_h[6], keep_l[6];   // set A's conv2 outputs of the current region pair, waiting for set B's
#pragma unroll
  for (int i = 0; i < 6; ++i) keep_h[i] = keep_l[i] = 0u;
  float wneg_a = 0.f, wneg_b = 0.f;

  // (region pair T, set st) = (it / 2, it % 2); rolled so that the loop body stays inside the instruction cache
#pragma unroll 1
  for (int it = 0; it < 4; ++it) {
    const int T = it >> 1, st = it & 1;
    const uint32_t blk = st ? qb.blk : qa.blk;
    __half* const hi = st ? qb.hi : qa.hi;
    __half* const lo = st ? qb.lo : qa.lo;
    const int c2_off = st ? qb.c2_off : qa.c2_off;
    const bool valid = st ? qb.valid : qa.valid;
    if (T == 0) {
      // integer sum over the quad's 16x16 pooled window (256 pool^2 pixels), kept as -W/32 (exact)
      uint32_t ws = block_rows_sum(pool, blk, d);
      ws += __shfl_xor_sync(0xffffffffu, ws, 1);
      ws += __shfl_xor_sync(0xffffffffu, ws, 2);
      const float w = -__uint2float_rn(ws) * 0.03125f;
      if (st) wneg_b = w; else wneg_a = w;
    }
    uint32_t xh[16], xl[16];
    load_x(pool, blk, T, d, st ? wneg_b : wneg_a, xh, xl);
    float d2[3][4];
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) d2[nt][0] = d2[nt][1] = d2[nt][2] = d2[nt][3] = 0.f;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      // conv1 for patch p of regions 2T (row g) and 2T+1 (row g+8)
      const uint32_t a1h[4] = {xh[4 * p], xh[4 * p + 1], xh[4 * p + 2], xh[4 * p + 3]};
      const uint32_t a1l[4] = {xl[4 * p], xl[4 * p + 1], xl[4 * p + 2], xl[4 * p + 3]};
      float d1[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        d1[nt][0] = d1[nt][1] = d1[nt][2] = d1[nt][3] = 0.f;
        mma3(d1[nt], a1h, a1l, f1h[nt], f1l[nt]);
      }
      // bias + leaky + re-split: the C fragments become the conv2 A fragment of k-step p
      uint32_t a2h[4], a2l[4];
      split2(leaky(fmaf(d1[0][0], u1, b1lo.x)), leaky(fmaf(d1[0][1], u1, b1lo.y)), a2h[0], a2l[0]);
      split2(leaky(fmaf(d1[0][2], u1, b1lo.x)), leaky(fmaf(d1[0][3], u1, b1lo.y)), a2h[1], a2l[1]);
      split2(leaky(fmaf(d1[1][0], u1, b1hi.x)), leaky(fmaf(d1[1][1], u1, b1hi.y)), a2h[2], a2l[2]);
      split2(leaky(fmaf(d1[1][2], u1, b1hi.x)), leaky(fmaf(d1[1][3], u1, b1hi.y)), a2h[3], a2l[3]);
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) {
        const uint2 wh = lds_u64(wb + 4 * kF2HiOff + ((p * 3 + nt) * 32 + lane) * 8);
        const uint2 wl = lds_u64(wb + 4 * kF2LoOff + ((p * 3 + nt) * 32 + lane) * 8);
        mma3(d2[nt], a2h, a2l, wh, wl);
      }
    }
    // conv2 outputs of regions 2T (c0, c1) and 2T+1 (c2, c3): features (already scaled by 2^feat_exp, hi/lo)
    uint32_t cur_h[6], cur_l[6];   // pair index 3 (r - 2T) + nt
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
      float2 b2 = lds_f32x2(wb + 4 * (kB2Off + 8 * nt + 2 * d));
      b2.x *= fscale, b2.y *= fscale;
      split2(leaky(fmaf(d2[nt][0], u2, b2.x)), leaky(fmaf(d2[nt][1], u2, b2.y)), cur_h[nt], cur_l[nt]);
      split2(leaky(fmaf(d2[nt][2], u2, b2.x)), leaky(fmaf(d2[nt][3], u2, b2.y)), cur_h[3 + nt], cur_l[3 + nt]);
      if (valid) {
        const int o = c2_off + T * g24 + 8 * nt + 2 * d;
        *reinterpret_cast<uint32_t*>(hi + o) = cur_h[nt];
        *reinterpret_cast<uint32_t*>(lo + o) = cur_l[nt];
        *reinterpret_cast<uint32_t*>(hi + o + 24) = cur_h[3 + nt];
        *reinterpret_cast<uint32_t*>(lo + o + 24) = cur_l[3 + nt];
      }
    }
    if (st == 0) {
#pragma unroll
      for (int i = 0; i < 6; ++i) keep_h[i] = cur_h[i], keep_l[i] = cur_l[i];
    } else {
      // conv3 k-steps 3T .. 3T+2: K = 96 in natural (region, channel) order, regions 2T and 2T+1 of both sets
#pragma unroll
      for (int jj = 0; jj < 3; ++jj) {
        const uint32_t a3h[4] = {keep_h[2 * jj], cur_h[2 * jj], keep_h[2 * jj + 1], cur_h[2 * jj + 1]};
        const uint32_t a3l[4] = {keep_l[2 * jj], cur_l[2 * jj], keep_l[2 * jj + 1], cur_l[2 * jj + 1]};
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const uint32_t fo = (((3 * T + jj) * 4 + nt) * 32 + lane) * 8;
          const uint2 wh = lds_u64(wb + 4 * kF3HiOff + fo);
          const uint2 wl = lds_u64(wb + 4 * kF3LoOff + fo);
          mma3(d3[nt], a3h, a3l, wh, wl);
        }
      }
    }
  }

#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    float2 b3 = lds_f32x2(wb + 4 * (kB3Off + 8 * nt + 2 * d));
    b3.x *= fscale, b3.y *= fscale;
    uint32_t h0, l0, h1, l1;
    split2(leaky(fmaf(d3[nt][0], u3, b3.x)), leaky(fmaf(d3[nt][1], u3, b3.y)), h0, l0);
    split2(leaky(fmaf(d3[nt][2], u3, b3.x)), leaky(fmaf(d3[nt][3], u3, b3.y)), h1, l1);
    if (qa.valid) {
      *reinterpret_cast<uint32_t*>(qa.hi + qa.c3_off + 8 * nt + 2 * d) = h0;
      *reinterpret_cast<uint32_t*>(qa.lo + qa.c3_off + 8 * nt + 2 * d) = l0;
    }
    if (qb.valid) {
      *reinterpret_cast<uint32_t*>(qb.hi + qb.c3_off + 8 * nt + 2 * d) = h1;
      *reinterpret_cast<uint32_t*>(qb.lo + qb.c3_off + 8 * nt + 2 * d) = l1;
    }
  }
}

constexpr int kTileBytes = kCtu * kCtu;                       // 4096
constexpr int kStageBytes = kGroupCtus * kTileBytes;          // 65536
constexpr int kWeightBytes = ((kConvFloats * 4 + 127) / 128) * 128;
constexpr int kConvSmemBytes = kWeightBytes + kConvStages * kStageBytes + 2 * kConvStages * 8 + 128;

template <bool kTma>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_features_kernel(const __grid_constant__ CUtensorMap tmap, const ConvLaunch p) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* wsm = reinterpret_cast<float*>(smem);
  uint8_t* tiles = smem + kWeightBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + kConvStages * kStageBytes);
  uint64_t* empty = full + kConvStages;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const float4* src = reinterpret_cast<const float4*>(p.convw);
    float4* dst = reinterpret_cast<float4*>(wsm);
    for (int i = threadIdx.x; i < kConvFloats / 4; i += kConvThreads) dst[i] = src[i];
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kConvStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kGroupTasks);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const int n_groups = (p.n_ctus + kGroupCtus - 1) / kGroupCtus;

  if (warp == 0) {
    // ---------------- producer: stream tile groups into the ring ----------------
    for (int j = 0;; ++j) {
      const int g = blockIdx.x + j * gridDim.x;
      if (g >= n_groups) break;
      const int stage = j % kConvStages;
      const uint32_t parity = (j / kConvStages) & 1;
      mbar_wait(&empty[stage], parity ^ 1);
      const int nv = min(kGroupCtus, p.n_ctus - g * kGroupCtus);
      uint8_t* dst = tiles + stage * kStageBytes;
      if (kTma) {
        if (lane == 0) mbar_arrive_expect_tx(&full[stage], nv * kTileBytes);
        __syncwarp();
        if (lane < nv) {
          const int n = p.ctu_begin + g * kGroupCtus + lane;
          const int f = n / p.ctus_per_frame, r = n - f * p.ctus_per_frame;
          const int cy = r / p.ctus_per_row, cx = r - cy * p.ctus_per_row;
          tma_load_3d(dst + lane * kTileBytes, &tmap, &full[stage], cx * kCtu, cy * kCtu, f);
        }
      } else {
        for (int c = 0; c < nv; ++c) {
          const int n = p.ctu_begin + g * kGroupCtus + c;
          const int f = n / p.ctus_per_frame, r = n - f * p.ctus_per_frame;
          const int cy = r / p.ctus_per_row, cx = r - cy * p.ctus_per_row;
          const uint8_t* src = p.luma + size_t(f) * p.frame_stride + size_t(cy) * kCtu * p.pitch + size_t(cx) * kCtu;
          const int rows = min(kCtu, p.height - cy * kCtu), cols = min(kCtu, p.width - cx * kCtu);
          uint8_t* t = dst + c * kTileBytes;
          for (int i = lane; i < kTileBytes; i += 32) {
            const int y = i >> 6, x = i & 63;
            t[i] = (y < rows && x < cols) ? src[size_t(y) * p.pitch + x] : uint8_t(0);  // zero padding
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[stage]);
      }
    }
  } else {
    // ---------------- compute warps: warp tasks round-robin ----------------
    const int cw = warp - 1;
    const uint32_t wsm_addr = smem_u32(wsm);
    const int g = lane >> 2;
    for (int t = cw;; t += kConvComputeWarps) {
      const int j = t / kGroupTasks, task = t - j * kGroupTasks;
      const int grp = blockIdx.x + j * gridDim.x;
      if (grp >= n_groups) break;
      const int stage = j % kConvStages;
      const uint32_t parity = (j / kConvStages) & 1;
      mbar_wait(&full[stage], parity);
      const uint32_t tile0 = smem_u32(tiles + stage * kStageBytes);
      const int ctu0 = grp * kGroupCtus;  // index inside this launch
      // lane group g owns quad (qy, qx) of CTU ca (set A) and of CTU cb / quad row qy + 2 (set B)
      int pool, br, ca, cb, qy_a, qy_b, qx, rg, qg, c2_base, c3_base;
      if (task < 16) {            // S: one CTU, 16 quads: set A = upper half, set B = lower half
        pool = 1, br = 0, ca = cb = task, qy_a = g >> 2, qy_b = qy_a + 2, qx = g & 3, rg = 8, qg = 4;
        c2_base = kOffC2S, c3_base = kOffC3S;
      } else if (task < 20) {     // M: four CTUs, 4 quads each: sets A / B in CTUs two apart
        pool = 2, br = 1, ca = 4 * (task - 16) + (g >> 2), cb = ca + 2, qy_a = qy_b = (g & 3) >> 1, qx = g & 1, rg = 4, qg = 2;
        c2_base = kOffC2M, c3_base = kOffC3M;
      } else {                    // L: sixteen CTUs, one quad each: sets A / B in CTUs eight apart
        pool = 4, br = 2, ca = g, cb = g + 8, qy_a = qy_b = 0, qx = 0, rg = 2, qg = 1;
        c2_base = kOffC2L, c3_base = kOffC3L;
      }
      const int bpx = 16 * pool;  // quad block edge in pixels
      QuadSet qa, qb;
      qa.valid = (ctu0 + ca) < p.n_ctus, qb.valid = (ctu0 + cb) < p.n_ctus;
      const size_t row_a = size_t(qa.valid ? ctu0 + ca : 0) * kFeat, row_b = size_t(qb.valid ? ctu0 + cb : 0) * kFeat;
      qa.blk = tile0 + ca * kTileBytes + (bpx * qy_a) * kCtu + bpx * qx;
      qb.blk = tile0 + cb * kTileBytes + (bpx * qy_b) * kCtu + bpx * qx;
      qa.hi = p.feat_hi + row_a, qa.lo = p.feat_lo + row_a;
      qb.hi = p.feat_hi + row_b, qb.lo = p.feat_lo + row_b;
      qa.c2_off = c2_base + ((2 * qy_a) * rg + 2 * qx) * 24, qb.c2_off = c2_base + ((2 * qy_b) * rg + 2 * qx) * 24;
      qa.c3_off = c3_base + (qy_a * qg + qx) * 32, qb.c3_off = c3_base + (qy_b * qg + qx) * 32;
      warp_task(pool, wsm_addr + 4 * br * kConvBranchFloats, p.cst[br], p.feat_scale, lane, qa, qb, rg * 24);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
    }
  }
}

}  // namespace

cudaError_t conv_features_configure() {
  cudaError_t e = cudaFuncSetAttribute(conv_features_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kConvSmemBytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(conv_features_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kConvSmemBytes);
}

cudaError_t launch_conv_features(const CUtensorMap* tmap, const ConvLaunch& p, int sm_count, cudaStream_t stream) {
  if (p.n_ctus <= 0) return cudaSuccess;
  const int n_groups = (p.n_ctus + kGroupCtus - 1) / kGroupCtus;
  const int grid = n_groups < sm_count ? n_groups : sm_count;
  if (tmap) {
    conv_features_kernel<true><<<grid, kConvThreads, kConvSmemBytes, stream>>>(*tmap, p);
  } else {
    CUtensorMap dummy;
    memset(&dummy, 0, sizeof(dummy));
    conv_features_kernel<false><<<grid, kConvThreads, kConvSmemBytes, stream>>>(dummy, p);
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// Tensor map over the luma planes.  cuTensorMapEncodeTiled is fetched through the runtime so that the
// library does not link libcuda directly.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

void* get_encode_tiled() {
  static void* fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    cudaDriverEntryPointQueryResult qres;
    void* f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = f;
  }
  return fn;
}

bool make_luma_tensor_map(CUtensorMap* map, const uint8_t* d_y, int width, int height, int n_frames, size_t pitch,
                          size_t frame_stride, const char** err) {
  EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(get_encode_tiled());
  if (!encode) {
    *err = "cuTensorMapEncodeTiled not available from the driver";
    return false;
  }
  if ((reinterpret_cast<uintptr_t>(d_y) & 15) || (pitch & 15) || (frame_stride & 15)) {
    *err = "luma base/pitch/frame stride not 16-byte aligned";
    return false;
  }
  cuuint64_t dims[3] = {cuuint64_t(width), cuuint64_t(height), cuuint64_t(n_frames)};
  cuuint64_t strides[2] = {cuuint64_t(pitch), cuuint64_t(frame_stride)};
  cuuint32_t box[3] = {kCtu, kCtu, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(d_y), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    *err = "cuTensorMapEncodeTiled failed";
    return false;
  }
  return true;
}

}  // namespace ethcnn
