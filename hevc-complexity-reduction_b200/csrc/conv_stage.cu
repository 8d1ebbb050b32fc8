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
//   * precision: conv1's A operand is the centred integer pooled sum s - 128 pool^2 (|.| <= 2048: exact in fp16), its
//     filters are fp16 hi + lo (two MMAs); the mean removal is a per-window constant and moves into the bias:
//     conv(s - mean) = conv(s - centre) + (centre - mean) * sum(taps).  Conv2 / conv3 operands are split into fp16
//     hi + lo (after exact power-of-two scaling) and three MMAs (hi*hi, hi*lo, lo*hi) accumulate in fp32: 2^-22
//     relative, like the FC stages.
// A warp task is one CTU (S), four CTUs (M) or sixteen CTUs (L); a group of 16 CTUs is 16 + 4 + 1 = 21
// warp tasks of identical MMA count (280 mma.sync each).  Round-1 history: an FFMA version of this stage
// reached 39 % of the fp32 peak and was bound by shared-memory wavefronts for the broadcast weights
// (profiles/r01c_conv_v2.md); mma.sync does the same MACs 7.4x faster per SM (tools/microbench/hmma_rate.cu).
//
// A persistent CTA (one per SM) keeps the 58 KB of weight fragments resident in shared memory and streams
// 16-CTU tile groups through a 2-deep TMA ring (3-D tensor map over (x, y, frame), 64x64x1 box; the zero
// fill of out-of-bounds rows/columns IS the reference's zero padding, video_to_cu_depth.py:54-57).
// One producer warp issues TMA and tabulates the 16x16-pixel block sums of the group (the mean-removal windows of
// the three branches are 1, 4 and 16 of those blocks); eleven compute warps take warp tasks round-robin.
#include <cstdlib>
#include <cstring>

#include "kernels.h"
#include "ptx_sm100.cuh"

namespace ethcnn {
namespace {


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
#ifdef ETHCNN_EXP_NO_HMMA   // measurement only: the SIMT skeleton without the tensor instructions
  d[0] += __uint_as_float(a[0] ^ b.x), d[1] += __uint_as_float(a[1]), d[2] += __uint_as_float(a[2] ^ b.y), d[3] += __uint_as_float(a[3]);
  return;
#endif
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

// One accumulator pair -> leaky(d * u + b) -> packed fp16 hi pair and lo pair with hi + lo == value to ~22 bits.
__device__ __forceinline__ void act_split(float d0, float d1, float2 u, float2 b, uint32_t& hi, uint32_t& lo) {
#ifdef ETHCNN_EXP_NO_ACT    // measurement only: the MMA skeleton without the fragment epilogues
  hi = __float_as_uint(d0 + u.x), lo = __float_as_uint(d1 + b.x);
  return;
#endif
  const float2 t = fma2(make_float2(d0, d1), u, b);
  const float2 m = mul2(t, make_float2(0.2f, 0.2f));
  const float2 r = make_float2(fmaxf(m.x, t.x), fmaxf(m.y, t.y));   // Maximum(alpha*x, x)
  const __half2 h = __floats2half2_rn(r.x, r.y);
  const float2 l = sub2(r, make_float2(__low2float(h), __high2float(h)));
  const __half2 lh = __floats2half2_rn(l.x, l.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&lh);
}

// two fp16 features -> global memory
__device__ __forceinline__ void stg32(__half* p, uint32_t v) {
#ifdef ETHCNN_EXP_NO_STG    // measurement only: (almost) no feature stores
  if (v != 0x7fff7fffu) return;
#endif
  *reinterpret_cast<uint32_t*>(p) = v;
}

__device__ __forceinline__ uint32_t centred_half2(uint32_t biased_bits, float centre) {
  // biased_bits: two fp16 bit patterns 0x6400 + n = 1024 + n (n < 1024); subtract 1024 + centre: exact
  const __half2 c = __floats2half2_rn(centre, centre);
  const __half2 v = __hsub2(*reinterpret_cast<const __half2*>(&biased_bits), c);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// conv1 A operand of region pair T of a quad: for patch p = 2 ph + pl, register reg = rx + 2 ki holds the pooled pair
// (y = 8T + 4 ph + d/2 + 2 ki, x = 8 rx + 4 pl + 2 (d%2) + {0,1}) as s - 128 pool^2, s = integer sum of the pool x pool
// pixels: an integer of magnitude <= 2048, i.e. EXACT in fp16 -- conv1 needs no lo part for its A operand, and the mean
// removal (a per-window constant) moves into the bias: conv(s - mean) = conv(s - centre) + (centre - mean) * sum(taps).
// xs[c] = v[c ^ m] for m in 0..3 (two conditional swaps, eight selects)
__device__ __forceinline__ void unpermute4(const uint32_t (&v)[4], int m, uint32_t (&xs)[4]) {
  const bool b0 = m & 1, b1 = m & 2;
  const uint32_t w0 = b0 ? v[1] : v[0], w1 = b0 ? v[0] : v[1], w2 = b0 ? v[3] : v[2], w3 = b0 ? v[2] : v[3];
  xs[0] = b1 ? w2 : w0, xs[1] = b1 ? w3 : w1, xs[2] = b1 ? w0 : w2, xs[3] = b1 ? w1 : w3;
}

template <int P>
__device__ __forceinline__ void load_x_p(uint32_t blk, int T, int lane, uint32_t (&xh)[16]) {
  const int d = lane & 3, g = lane >> 2;
  const int ky0 = d >> 1, e = d & 1;
  // Pooled branches: the eight lane groups of a warp read the same offsets of DIFFERENT tiles / quads, i.e. the same banks
  // (tile stride 4096 B; TMA wants 128-byte aligned tiles, so the tiles cannot be skewed): 16 wavefronts per LDS.64 in the L task,
  // 4 per LDS.32 in the M tasks (profiles/r02h_conv.md: 37 % of all shared-memory wavefronts were conflicts, 70 % of them from the
  // L task).  Each group therefore walks its four (pl, rx) elements of a row pair in its own order (index c ^ m) and, for the 4x4
  // pooling, its four pixel rows starting at its own parity -- sums are order-independent -- and the elements are put back
  // in place with selects: M conflict-free, L 4 wavefronts instead of 16.
  const int m = P == 2 ? ((g >> 1) & 3) : (g & 3);
  const int apar = (g >> 2) & 1;
#pragma unroll
  for (int ph = 0; ph < 2; ++ph)
#pragma unroll
    for (int ki = 0; ki < 2; ++ki) {
      const int y = 8 * T + 4 * ph + ky0 + 2 * ki;
      if (P == 1) {
        const uint4 row = lds_u128(blk + y * kCtu);
        const uint32_t w[4] = {row.x, row.y, row.z, row.w};
        const uint32_t sel = e ? 0x4342u : 0x4140u;   // bytes (2e, 2e+1) -> low bytes of the two halves, 0x64 above: 1024 + pixel
#pragma unroll
        for (int pl = 0; pl < 2; ++pl)
#pragma unroll
          for (int rx = 0; rx < 2; ++rx)
            xh[4 * (2 * ph + pl) + rx + 2 * ki] = centred_half2(__byte_perm(w[2 * rx + pl], 0x64646464u, sel), 1024.f + 128.f);
      } else {
        uint32_t v[4], xs[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int cc = c ^ m;                               // this group's element (pl, rx) = (cc / 2, cc % 2) at step c
          const int x0 = 8 * (cc & 1) + 4 * (cc >> 1) + 2 * e;
          if (P == 2) {
            // the two pixel rows in lane-dependent order (the sum is symmetric): lanes d/2 = 0, 1 hit different banks
            const uint32_t a = lds_u32(blk + (2 * y + ky0) * kCtu + 2 * x0), b = lds_u32(blk + (2 * y + (ky0 ^ 1)) * kCtu + 2 * x0);
            const uint32_t s0 = __dp4a(b, 0x00000101u, __dp4a(a, 0x00000101u, 0x6400u));   // 0x6400 + s: fp16 bits of 1024 + s
            const uint32_t s1 = __dp4a(b, 0x01010000u, __dp4a(a, 0x01010000u, 0x6400u));
            v[c] = centred_half2(__byte_perm(s0, s1, 0x5410u), 1024.f + 512.f);
          } else {
            uint32_t s0 = 0u - 2048u, s1 = 0u - 2048u;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
              const uint2 q = lds_u64(blk + (4 * y + (a ^ apar)) * kCtu + 4 * x0);
              s0 = __dp4a(q.x, 0x01010101u, s0);
              s1 = __dp4a(q.y, 0x01010101u, s1);
            }
            const __half2 h = __floats2half2_rn(__int2float_rn(int(s0)), __int2float_rn(int(s1)));
            v[c] = *reinterpret_cast<const uint32_t*>(&h);
          }
        }
        unpermute4(v, m, xs);
#pragma unroll
        for (int c = 0; c < 4; ++c) xh[4 * (2 * ph + (c >> 1)) + (c & 1) + 2 * ki] = xs[c];
      }
    }
}

__device__ __forceinline__ void load_x(int pool, uint32_t blk, int T, int lane, uint32_t (&xh)[16]) {
  if (pool == 1) {
    load_x_p<1>(blk, T, lane, xh);
  } else if (pool == 2) {
    load_x_p<2>(blk, T, lane, xh);
  } else {
    load_x_p<4>(blk, T, lane, xh);
  }
}

#ifdef ETHCNN_EXP_TIMING   // measurement only: cycles per phase of the warp tasks (per-warp counters in shared memory, folded at exit)
__device__ unsigned long long g_conv_phase[3][8];
__shared__ unsigned int s_conv_phase[16][24];
#define PH_BEGIN() long long ph_t = clock64()
#define PH_MARK(br, k)                                                                     \
  do {                                                                                     \
    const long long ph_n = clock64();                                                      \
    if ((threadIdx.x & 31) == 0) s_conv_phase[threadIdx.x >> 5][(br) * 8 + (k)] += (unsigned int)(ph_n - ph_t); \
    ph_t = ph_n;                                                                           \
  } while (0)
#else
#define PH_BEGIN()
#define PH_MARK(br, k)
#endif

#ifdef ETHCNN_EXP_UMMA_LOAD   // measurement only: the producer warp issues dummy tcgen05 MMAs (what a fused FC1 would put on
                              // the tensor pipe) while the compute warps run the unchanged conv stage
__device__ __forceinline__ void exp_umma(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
  asm volatile(
      "{\n.reg .pred p, e;\n.reg .b32 rx;\nelect.sync rx|e, 0xffffffff;\nsetp.ne.b32 p, 1, 0;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ uint64_t exp_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3ffffu) >> 4);
  d |= uint64_t(1) << 16;
  d |= uint64_t(1024 >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}
#endif

struct QuadSet {        // what a lane needs to know about its quad in set A or B
  uint32_t blk;         // shared-memory address of the quad's pixel block inside its CTU tile
  uint32_t wsum;        // shared-memory address of this lane's share of the quad's window sum (block-sum table)
  __half* hi;           // feature rows of the quad's CTU (global memory)
  __half* lo;
  int c2_off;           // offset of the 24 conv2 features of region 0 of the quad
  int c3_off;           // offset of the quad's 32 conv3 features
};

// One warp task: the conv stack for 16 quads (8 per set).
//   pool   1 / 2 / 4 (warp-uniform)    wb  shared-memory byte address of the branch's weight block
//   g24    feature distance between vertically adjacent regions (regions per row * 24)
__device__ __forceinline__ void warp_task(const int pool, const uint32_t wb, const float cst, const float fscale, const int lane,
                                       const QuadSet qa, const QuadSet qb, const int g24) {
  const int d = lane & 3;
  const int phb = pool == 1 ? 0 : (pool == 2 ? 1 : 2);
  (void)phb;
  PH_BEGIN();
  // leaky(v) * 2^e == leaky(v * 2^e): the power-of-two operand scales are folded into the unscale factors and biases
  const float sc1 = lds_f32(wb + 4 * (kHdrOff + 3));
  const float u1 = lds_f32(wb + 4 * kHdrOff) * cst * sc1;     // per unit of sum((256 s - W) / 32 * B)
  const float2 u1x8 = make_float2(8.f * u1, 8.f * u1);         // per unit of sum((s - centre) * B)
  const float u2s = lds_f32(wb + 4 * (kHdrOff + 1)) * fscale, u3s = lds_f32(wb + 4 * (kHdrOff + 2)) * fscale;
  const float2 u2 = make_float2(u2s, u2s), u3 = make_float2(u3s, u3s), fs2 = make_float2(fscale, fscale);
  // this lane's output channels: conv1 {2d, 2d+1, 8+2d, 9+2d}; conv2 / conv3 {8 nt + 2d, +1}
  const float2 b1lo = mul2(lds_f32x2(wb + 4 * (kB1Off + 2 * d)), make_float2(sc1, sc1));
  const float2 b1hi = mul2(lds_f32x2(wb + 4 * (kB1Off + 8 + 2 * d)), make_float2(sc1, sc1));
  const float2 t1lo = lds_f32x2(wb + 4 * (kW1SumOff + 2 * d)), t1hi = lds_f32x2(wb + 4 * (kW1SumOff + 8 + 2 * d));
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
  // conv1 bias + mean term of the quad, sets A / B.  W = integer sum over the quad's 16x16 pooled window (256 pool^2 pixels)
  // from the 16x16-pixel block sums the producer warp tabulated; the mean term (centre - W/256) * 256/32 = 8 centre - W/32 is
  // exact in fp32.  Done for both sets up front, where its latency (LDS -> shuffles -> I2F -> FFMA) hides behind the first
  // pixel loads instead of sitting at the head of two iterations of the rolled loop below.
  float2 ba_lo, ba_hi, bb_lo, bb_hi;
  {
    uint32_t wsa, wsb;
    if (pool == 4) {
      const uint4 va = lds_u128(qa.wsum), vb = lds_u128(qb.wsum);
      wsa = va.x + va.y + va.z + va.w, wsb = vb.x + vb.y + vb.z + vb.w;
    } else {
      wsa = lds_u32(qa.wsum), wsb = lds_u32(qb.wsum);
    }
    if (pool != 1) {
      wsa += __shfl_xor_sync(0xffffffffu, wsa, 1), wsb += __shfl_xor_sync(0xffffffffu, wsb, 1);
      wsa += __shfl_xor_sync(0xffffffffu, wsa, 2), wsb += __shfl_xor_sync(0xffffffffu, wsb, 2);
    }
    const float centre = float(1024 * pool * pool);
    const float mua = fmaf(__uint2float_rn(wsa), -0.03125f, centre) * u1, mub = fmaf(__uint2float_rn(wsb), -0.03125f, centre) * u1;
    ba_lo = fma2(make_float2(mua, mua), t1lo, b1lo), ba_hi = fma2(make_float2(mua, mua), t1hi, b1hi);
    bb_lo = fma2(make_float2(mub, mub), t1lo, b1lo), bb_hi = fma2(make_float2(mub, mub), t1hi, b1hi);
  }

  // (region pair T, set st) = (it / 2, it % 2).  Unrolled by two: the pixel loads and conv1 MMAs of the second half overlap the
  // conv2 epilogue / stores of the first (0.186 -> 0.178 ms per 25 500 CTUs); fully unrolled the body no longer fits the
  // instruction cache (0.190 ms) -- profiles/r02c_conv_variants.md.
#ifndef ETHCNN_CONV_IT_UNROLL
#define ETHCNN_CONV_IT_UNROLL 2
#endif
  constexpr int kItUnroll = ETHCNN_CONV_IT_UNROLL;
#pragma unroll kItUnroll
  for (int it = 0; it < 4; ++it) {
    const int T = it >> 1, st = it & 1;
    const uint32_t blk = st ? qb.blk : qa.blk;
    __half* const hi = st ? qb.hi : qa.hi;
    __half* const lo = st ? qb.lo : qa.lo;
    const int c2_off = st ? qb.c2_off : qa.c2_off;
    const float2 be_lo = st ? bb_lo : ba_lo, be_hi = st ? bb_hi : ba_hi;
    PH_MARK(phb, 0);   // task set-up / window sums
    uint32_t xh[16];
    load_x(pool, blk, T, lane, xh);
    PH_MARK(phb, 1);   // pixel loads + conversion
    float d2[3][4];
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) d2[nt][0] = d2[nt][1] = d2[nt][2] = d2[nt][3] = 0.f;
#ifdef ETHCNN_CONV_SPLIT_ACC   // correction passes (hi*lo, lo*hi) into their own accumulators: dependent chains of 4 + 8 instead of 12
    float e2[3][4];
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) e2[nt][0] = e2[nt][1] = e2[nt][2] = e2[nt][3] = 0.f;
#endif
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      // conv1 for patch p of regions 2T (row g) and 2T+1 (row g+8): the A operand is exact, two passes (B hi, B lo)
      const uint32_t a1[4] = {xh[4 * p], xh[4 * p + 1], xh[4 * p + 2], xh[4 * p + 3]};
      float d1[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        d1[nt][0] = d1[nt][1] = d1[nt][2] = d1[nt][3] = 0.f;
        mma16816(d1[nt], a1, f1h[nt]);
        mma16816(d1[nt], a1, f1l[nt]);
      }
      // bias + leaky + re-split: the C fragments become the conv2 A fragment of k-step p
      uint32_t a2h[4], a2l[4];
      act_split(d1[0][0], d1[0][1], u1x8, be_lo, a2h[0], a2l[0]);
      act_split(d1[0][2], d1[0][3], u1x8, be_lo, a2h[1], a2l[1]);
      act_split(d1[1][0], d1[1][1], u1x8, be_hi, a2h[2], a2l[2]);
      act_split(d1[1][2], d1[1][3], u1x8, be_hi, a2h[3], a2l[3]);
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) {
        const uint2 wh = lds_u64(wb + 4 * kF2HiOff + ((p * 3 + nt) * 32 + lane) * 8);
        const uint2 wl = lds_u64(wb + 4 * kF2LoOff + ((p * 3 + nt) * 32 + lane) * 8);
#ifdef ETHCNN_CONV_SPLIT_ACC
        mma16816(d2[nt], a2h, wh);
        mma16816(e2[nt], a2h, wl);
        mma16816(e2[nt], a2l, wh);
#else
        mma3(d2[nt], a2h, a2l, wh, wl);
#endif
      }
    }
#ifdef ETHCNN_CONV_SPLIT_ACC
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) d2[nt][i] += e2[nt][i];
#endif
    PH_MARK(phb, 2);   // conv1 + epilogue + conv2 MMAs
    // conv2 outputs of regions 2T (c0, c1) and 2T+1 (c2, c3): features (already scaled by 2^feat_exp, hi/lo)
    uint32_t cur_h[6], cur_l[6];   // pair index 3 (r - 2T) + nt
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
      const float2 b2 = mul2(lds_f32x2(wb + 4 * (kB2Off + 8 * nt + 2 * d)), fs2);
      act_split(d2[nt][0], d2[nt][1], u2, b2, cur_h[nt], cur_l[nt]);
      act_split(d2[nt][2], d2[nt][3], u2, b2, cur_h[3 + nt], cur_l[3 + nt]);
      const int o = c2_off + T * g24 + 8 * nt + 2 * d;
      stg32(hi + o, cur_h[nt]);
      stg32(lo + o, cur_l[nt]);
      stg32(hi + o + 24, cur_h[3 + nt]);
      stg32(lo + o + 24, cur_l[3 + nt]);
    }
    PH_MARK(phb, 3);   // conv2 epilogue + feature stores
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
      PH_MARK(phb, 4);   // conv3 MMAs
    }
  }

#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const float2 b3 = mul2(lds_f32x2(wb + 4 * (kB3Off + 8 * nt + 2 * d)), fs2);
    uint32_t h0, l0, h1, l1;
    act_split(d3[nt][0], d3[nt][1], u3, b3, h0, l0);
    act_split(d3[nt][2], d3[nt][3], u3, b3, h1, l1);
    stg32(qa.hi + qa.c3_off + 8 * nt + 2 * d, h0);
    stg32(qa.lo + qa.c3_off + 8 * nt + 2 * d, l0);
    stg32(qb.hi + qb.c3_off + 8 * nt + 2 * d, h1);
    stg32(qb.lo + qb.c3_off + 8 * nt + 2 * d, l1);
  }
  PH_MARK(phb, 5);     // conv3 epilogue + stores
}

constexpr int kTileBytes = kCtu * kCtu;                       // 4096
constexpr int kTileStride = kTileBytes;                       // (TMA needs 128-byte aligned destinations: no bank skew between tiles)
constexpr int kStageBytes = kGroupCtus * kTileStride;         // 65536
constexpr int kWeightBytes = ((kConvFloats * 4 + 127) / 128) * 128;
constexpr int kSumWords = kGroupCtus * 16;                   // 16x16-pixel block sums of a group: [ctu][4 qy + qx]
constexpr int kConvSmemBytes = kWeightBytes + kConvStages * (kStageBytes + kSumWords * 4) + 3 * kConvStages * 8 + 128;

template <bool kTma>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_features_kernel(const __grid_constant__ CUtensorMap tmap, const ConvLaunch p) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* wsm = reinterpret_cast<float*>(smem);
  uint8_t* tiles = smem + kWeightBytes;
  uint32_t* sums = reinterpret_cast<uint32_t*>(tiles + kConvStages * kStageBytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(sums + kConvStages * kSumWords);   // tile bytes landed
  uint64_t* empty = full + kConvStages;                                           // all warp tasks of the group done
  uint64_t* ready = empty + kConvStages;                                          // block sums of the group tabulated
  unsigned int* next_task = reinterpret_cast<unsigned int*>(ready + kConvStages);   // warp-task dispenser of this CTA

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef ETHCNN_EXP_TIMING
  if (lane < 24) s_conv_phase[warp][lane] = 0;
#endif
  pdl_launch_dependents();   // the FC kernel's CTAs may take over each SM as soon as this CTA leaves it (their prologue overlaps our tail)
  {
    const float4* src = reinterpret_cast<const float4*>(p.convw);
    float4* dst = reinterpret_cast<float4*>(wsm);
    // all of a thread's loads in flight before its first store: a rolled load -> store loop pays the L2 latency ten times in a row
    constexpr int kCopyIters = (kConvFloats / 4 + kConvThreads - 1) / kConvThreads;
    float4 v[kCopyIters];
#pragma unroll
    for (int k = 0; k < kCopyIters; ++k)
      if (int(threadIdx.x) + k * kConvThreads < kConvFloats / 4) v[k] = src[threadIdx.x + k * kConvThreads];
#pragma unroll
    for (int k = 0; k < kCopyIters; ++k)
      if (int(threadIdx.x) + k * kConvThreads < kConvFloats / 4) dst[threadIdx.x + k * kConvThreads] = v[k];
  }
#ifdef ETHCNN_EXP_UMMA_LOAD
  __shared__ uint32_t exp_tmem_slot;
  __shared__ uint64_t exp_bar;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&exp_tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
#endif
  if (threadIdx.x == 0) {
#ifdef ETHCNN_EXP_UMMA_LOAD
    mbar_init(&exp_bar, 1);
#endif
    for (int s = 0; s < kConvStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kGroupTasks);
      mbar_init(&ready[s], 1);
    }
    *next_task = 0;
    mbar_fence_init();
  }
  // Launched as a programmatic dependent of whatever kernel precedes it in the stream (the FC kernel of the previous feature
  // chunk, the gate kernel of the previous call): the weight copy and barrier set-up above ran while that kernel drained.
  // Everything below may conflict with it (the flags its gate reads, the feature rows its FC reads, luma a caller's kernel
  // wrote), so it starts here.
  pdl_wait();
  // the gate flags of the call are cleared here (the FC kernel, next in the stream, is the first to set them): one launch
  // less per call than a separate memset
  if (p.clear_flags != nullptr)
    for (int i = blockIdx.x * kConvThreads + threadIdx.x; i < p.n_clear_flags; i += gridDim.x * kConvThreads) p.clear_flags[i] = 0u;
  __syncthreads();

  const int n_groups = (p.n_ctus + kGroupCtus - 1) / kGroupCtus;

  if (warp == 0) {
    // ---------------- producer: stream tile groups into the ring, tabulate their 16x16 block sums ----------------
    for (int j = 0;; ++j) {
      const int g = blockIdx.x + j * gridDim.x;
      if (g >= n_groups) break;
      const int stage = j % kConvStages;
      const uint32_t parity = (j / kConvStages) & 1;
      mbar_wait_relaxed(&empty[stage], parity ^ 1);   // a whole group period away: do not compete for issue slots
      const int nv = min(kGroupCtus, p.n_ctus - g * kGroupCtus);
      uint8_t* dst = tiles + stage * kStageBytes;
      if (kTma) {
        if (lane == 0) mbar_arrive_expect_tx(&full[stage], nv * kTileBytes);
        __syncwarp();
        if (lane < nv) {
          const int n = p.ctu_begin + g * kGroupCtus + lane;
          const int f = n / p.ctus_per_frame, r = n - f * p.ctus_per_frame;
          const int cy = r / p.ctus_per_row, cx = r - cy * p.ctus_per_row;
          tma_load_3d(dst + lane * kTileStride, &tmap, &full[stage], cx * kCtu, cy * kCtu, f);
        }
      } else {
        for (int c = 0; c < nv; ++c) {
          const int n = p.ctu_begin + g * kGroupCtus + c;
          const int f = n / p.ctus_per_frame, r = n - f * p.ctus_per_frame;
          const int cy = r / p.ctus_per_row, cx = r - cy * p.ctus_per_row;
          const uint8_t* src = p.luma + size_t(f) * p.frame_stride + size_t(cy) * kCtu * p.pitch + size_t(cx) * kCtu;
          const int rows = min(kCtu, p.height - cy * kCtu), cols = min(kCtu, p.width - cx * kCtu);
          uint8_t* t = dst + c * kTileStride;
          for (int i = lane; i < kTileBytes; i += 32) {
            const int y = i >> 6, x = i & 63;
            t[i] = (y < rows && x < cols) ? src[size_t(y) * p.pitch + x] : uint8_t(0);  // zero padding
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[stage]);
      }
      // Integer sums of the 16 blocks of 16x16 pixels of every tile: the mean-removal windows of the three branches
      // are 1, 4 and 16 of these blocks.  Lane l reads the 16-byte chunk l%4 of row 8 jj + l/4: 512 contiguous
      // bytes per instruction, conflict-free; the compute warps are busy with the previous group meanwhile.
      mbar_wait(&full[stage], parity);
      const uint32_t t0 = smem_u32(dst) + (lane >> 2) * kCtu + (lane & 3) * 16;
      uint32_t* sm = sums + stage * kSumWords;
      for (int c = 0; c < nv; ++c) {
        uint32_t acc[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const uint4 v = lds_u128(t0 + c * kTileStride + jj * 8 * kCtu);
          acc[jj >> 1] = __dp4a(v.w, 0x01010101u, __dp4a(v.z, 0x01010101u, __dp4a(v.y, 0x01010101u, __dp4a(v.x, 0x01010101u, acc[jj >> 1]))));
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], 4);
          acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], 8);
          acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], 16);
        }
        if (lane < 4) {
#pragma unroll
          for (int q = 0; q < 4; ++q) sm[c * 16 + q * 4 + lane] = acc[q];
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&ready[stage]);
#ifdef ETHCNN_EXP_UMMA_LOAD
      {   // ETHCNN_EXP_UMMA_LOAD dummy MMAs (M = 128, N = 224, K = 16: 112 cycles each) per 16-CTU group, in bursts of 6 like the
          // K = 16 stages of a fused FC1 (930 tensor cycles per CTU = 133 such MMAs per group)
        const uint32_t tm = exp_tmem_slot;
        const uint64_t da = exp_desc(smem_u32(tiles)), db = exp_desc(smem_u32(tiles) + 16384);
        const uint32_t idesc = (1u << 4) | (uint32_t(224 >> 3) << 17) | (uint32_t(128 >> 4) << 24);
        for (int b = 0; b < ETHCNN_EXP_UMMA_LOAD; b += 6) {
#pragma unroll
          for (int k = 0; k < 6; ++k) exp_umma(tm, da + uint64_t(2 * (k & 3)), db + uint64_t(2 * (k & 3)), idesc);
          __nanosleep(ETHCNN_EXP_UMMA_SLEEP);
        }
      }
#endif
    }
#ifdef ETHCNN_EXP_UMMA_LOAD
    asm volatile("{\n.reg .pred e;\n.reg .b32 rx;\nelect.sync rx|e, 0xffffffff;\n"
                 "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(smem_u32(&exp_bar)) : "memory");
    mbar_wait(&exp_bar, 0);
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(exp_tmem_slot), "r"(512));
#endif
  } else {
    // ---------------- compute warps: warp tasks from a dispenser ----------------
    // Tasks are handed out in order from a shared counter instead of round-robin: M / L tasks cost 1.2 - 1.6x an S task, and
    // with a static assignment the warps that keep drawing them set the pace while the others wait on the 2-deep ring
    // (9.6 % of the compute warps' time, profiles/r02a_conv_ablation.md).
    const uint32_t wsm_addr = smem_u32(wsm);
    const int g = lane >> 2;
    for (;;) {
      int t = 0;
      if (lane == 0) t = int(atomicAdd(next_task, 1u));
      t = __shfl_sync(0xffffffffu, t, 0);
      // the four M and the L task of a group take longest (pooling on the fly): they go first, the 16 S tasks fill up
      const int j = t / kGroupTasks, task = (t - j * kGroupTasks + 16) % kGroupTasks;
      const int grp = blockIdx.x + j * gridDim.x;
      if (grp >= n_groups) break;
      const int stage = j % kConvStages;
      const uint32_t parity = (j / kConvStages) & 1;
#ifdef ETHCNN_EXP_TIMING
      const long long w0 = clock64();
#endif
      mbar_wait(&full[stage], parity);
      mbar_wait(&ready[stage], parity);
#ifdef ETHCNN_EXP_TIMING
      if (lane == 0) s_conv_phase[warp][6] += (unsigned int)(clock64() - w0);
#endif
      const uint32_t tile0 = smem_u32(tiles + stage * kStageBytes);
      const uint32_t sum0 = smem_u32(sums + stage * kSumWords);
      const int d = lane & 3;
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
      // CTUs past the end of a tail group are computed on stale tile bytes and stored to the dump row
      const size_t row_a = size_t((ctu0 + ca) < p.n_ctus ? ctu0 + ca : p.dump_row) * kFeat;
      const size_t row_b = size_t((ctu0 + cb) < p.n_ctus ? ctu0 + cb : p.dump_row) * kFeat;
      qa.blk = tile0 + ca * kTileStride + (bpx * qy_a) * kCtu + bpx * qx;
      qb.blk = tile0 + cb * kTileStride + (bpx * qy_b) * kCtu + bpx * qx;
      // this lane's share of the window sum: S = the quad's own block; M = block (2 qy + d/2, 2 qx + d%2) of the quad's
      // 2x2 blocks; L = block row d (four blocks, one 16-byte load) of the CTU's 4x4
      const int sa = pool == 1 ? 4 * qy_a + qx : (pool == 2 ? (2 * qy_a + (d >> 1)) * 4 + 2 * qx + (d & 1) : 4 * d);
      const int sb = pool == 1 ? 4 * qy_b + qx : (pool == 2 ? (2 * qy_b + (d >> 1)) * 4 + 2 * qx + (d & 1) : 4 * d);
      qa.wsum = sum0 + 4 * (ca * 16 + sa), qb.wsum = sum0 + 4 * (cb * 16 + sb);
      qa.hi = p.feat_hi + row_a, qa.lo = p.feat_lo + row_a;
      qb.hi = p.feat_hi + row_b, qb.lo = p.feat_lo + row_b;
      qa.c2_off = c2_base + ((2 * qy_a) * rg + 2 * qx) * 24, qb.c2_off = c2_base + ((2 * qy_b) * rg + 2 * qx) * 24;
      qa.c3_off = c3_base + (qy_a * qg + qx) * 32, qb.c3_off = c3_base + (qy_b * qg + qx) * 32;
#ifdef ETHCNN_EXPERIMENT_S_ONLY   // measurement only (wrong results): time the S tasks / the M and L tasks alone
      if (task < 16)
#endif
#ifdef ETHCNN_EXPERIMENT_ML_ONLY
      if (task >= 16)
#endif
      warp_task(pool, wsm_addr + 4 * br * kConvBranchFloats, p.cst[br], p.feat_scale, lane, qa, qb, rg * 24);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
    }
#ifdef ETHCNN_EXP_TIMING
    __syncwarp();
    if (lane < 24 && s_conv_phase[warp][lane]) atomicAdd(&g_conv_phase[lane >> 3][lane & 7], (unsigned long long)s_conv_phase[warp][lane]);
#endif
  }
}

}  // namespace

#ifdef ETHCNN_EXP_TIMING
}  // namespace ethcnn
extern "C" int ethcnn_debug_conv_phases(unsigned long long* out24, int reset) {
  unsigned long long z[24] = {};
  if (cudaMemcpyFromSymbol(out24, ethcnn::g_conv_phase, sizeof(z)) != cudaSuccess) return -1;
  if (reset && cudaMemcpyToSymbol(ethcnn::g_conv_phase, z, sizeof(z)) != cudaSuccess) return -1;
  return 0;
}
namespace ethcnn {
#endif

cudaError_t conv_features_configure() {
  cudaError_t e = cudaFuncSetAttribute(conv_features_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kConvSmemBytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(conv_features_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kConvSmemBytes);
}

cudaError_t launch_conv_features(const CUtensorMap* tmap, const ConvLaunch& p, int sm_count, cudaStream_t stream) {
  if (p.n_ctus <= 0) return cudaSuccess;
  const int n_groups = (p.n_ctus + kGroupCtus - 1) / kGroupCtus;
  const int grid = n_groups < sm_count ? n_groups : sm_count;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid), cfg.blockDim = dim3(kConvThreads), cfg.dynamicSmemBytes = kConvSmemBytes, cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = getenv("ETHCNN_NO_PDL") ? 0 : 1;   // measurement switch
  if (tmap) return cudaLaunchKernelEx(&cfg, conv_features_kernel<true>, *tmap, p);
  CUtensorMap dummy;
  memset(&dummy, 0, sizeof(dummy));
  return cudaLaunchKernelEx(&cfg, conv_features_kernel<false>, dummy, p);
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
