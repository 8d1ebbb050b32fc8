// Stage CONV: luma tiles -> 2688 conv features per CTU (sm_100a).
//
// Follows net_CNN.py:105-150 (AI) / ETH-CNN_Training_LDP/net_CTU64.py:102-175 (LDP):
//   x = X * scale; L = meanremove16(avgpool4(x)); M = meanremove16(avgpool2(x)); S = meanremove16(x)
//   per branch: conv 4x4/s4 1->16, conv 2x2/s2 16->24, conv 2x2/s2 24->32, leaky(0.2) after each
//   features = [c3_S | c3_M | c3_L | c2_S | c2_M | c2_L], each NHWC-flattened.
//
// Work decomposition.  All three branches have the same shape once the input is pooled: a "region" is an
// 8x8 block of (pooled) samples = 2x2 conv1 patches = one conv2 output position, and 2x2 neighbouring
// regions (a "quad", held by 4 adjacent lanes) share one mean-removal window (16x16 pooled samples) and
// one conv3 output position.
//   S: pool 1, region  8x8  px, 64 regions per CTU      M: pool 2, 16x16 px, 16 per CTU
//   L: pool 4, region 32x32 px,  4 regions per CTU
// Every lane owns TWO regions (A and B) that use the same filter taps, so each weight fetched from shared
// memory feeds two packed FMAs (fma.rn.f32x2: two output channels at once) -- a broadcast 128-bit LDS
// costs two shared-memory wavefronts, and at one region per lane the kernel was bound by that pipe and by
// instruction fetch (round-1 ncu: 52 % LSU wavefronts, 46 % I-cache misses).  A warp task is therefore
// one CTU (S), four CTUs (M) or sixteen CTUs (L); a group of 16 CTUs is exactly 16 + 4 + 1 = 21 warp
// tasks of identical cost.  One code body serves the three branches (the pool factor only changes the
// small integer-sum loaders), conv3 is rolled over its four 8-channel output groups, and the whole hot
// loop stays inside the 32 KB instruction cache.
//
// Mean removal is exact: with s = pooled integer sum and W = integer sum of the whole window,
//   pooled - mean = (256*s - W) / (256*pool^2), one rounding when multiplied by scale/(256*pool^2).
//
// A persistent CTA (one per SM) keeps the 58 KB of conv weights resident in shared memory and streams
// 16-CTU tile groups through a 2-deep TMA ring (3-D tensor map over (x, y, frame), 64x64x1 box; the zero
// fill of out-of-bounds rows/columns IS the reference's zero padding, video_to_cu_depth.py:54-57).
// One producer warp issues TMA, eleven compute warps take warp tasks round-robin.
#include <cstring>

#include "kernels.h"
#include "ptx_sm100.cuh"

namespace ethcnn {
namespace {

// 128-bit shared-memory load of four consecutive weights as two channel pairs.  `asm volatile` on
// purpose: the conv1 filter is invariant across the patch loop and the compiler would otherwise hoist
// all 64 loads (256 registers).
__device__ __forceinline__ void lds_pairs(uint32_t addr, float2& p0, float2& p1) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(p0.x), "=f"(p0.y), "=f"(p1.x), "=f"(p1.y) : "r"(addr));
}

__device__ __forceinline__ float leaky(float v) { return fmaxf(0.2f * v, v); }  // Maximum(alpha*x, x), alpha = 0.2
__device__ __forceinline__ float2 leaky2(float2 v) { return make_float2(leaky(v.x), leaky(v.y)); }
__device__ __forceinline__ float2 splat(float v) { return make_float2(v, v); }

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// (v.x, v.y) * scale -> packed fp16 hi pair and lo pair with hi + lo == v * scale to ~22 bits
__device__ __forceinline__ void split_pair(float2 v, float scale, uint32_t& hi, uint32_t& lo) {
  const float sx = v.x * scale, sy = v.y * scale;
  const __half2 h = __floats2half2_rn(sx, sy);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = pack_half2(sx - __low2float(h), sy - __high2float(h));
}

// Explicit shared-memory loads of pixels (32-bit shared addresses keep the compiler from emitting generic LD).
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

// Integer sum of a whole region ((8*pool) x (8*pool) pixels at shared address reg0, row pitch 64).
__device__ __forceinline__ uint32_t region_sum(int pool, uint32_t reg0) {
  uint32_t s = 0;
  if (pool == 1) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const uint2 v = lds_u64(reg0 + r * kCtu);
      s = __dp4a(v.x, 0x01010101u, s);
      s = __dp4a(v.y, 0x01010101u, s);
    }
  } else {
    const int rows = 8 * pool, chunks = pool >> 1;  // 16 rows x 16 B or 32 rows x 32 B
#pragma unroll 4
    for (int r = 0; r < rows; ++r) {
      for (int k = 0; k < chunks; ++k) {
        const uint4 v = lds_u128(reg0 + r * kCtu + 16 * k);
        s = __dp4a(v.x, 0x01010101u, s);
        s = __dp4a(v.y, 0x01010101u, s);
        s = __dp4a(v.z, 0x01010101u, s);
        s = __dp4a(v.w, 0x01010101u, s);
      }
    }
  }
  return s;
}

// Pooled integer sums of one conv1 patch: 4x4 pooled samples = (4P) x (4P) pixels at p0; s[ky*4 + kx].
template <int P>
__device__ __forceinline__ void load_patch_p(uint32_t p0, int (&s)[16]) {
#pragma unroll
  for (int ky = 0; ky < 4; ++ky) {
    if (P == 1) {
      const uint32_t w = lds_u32(p0 + ky * kCtu);
      s[ky * 4 + 0] = int(w & 0xffu);
      s[ky * 4 + 1] = int((w >> 8) & 0xffu);
      s[ky * 4 + 2] = int((w >> 16) & 0xffu);
      s[ky * 4 + 3] = int(w >> 24);
    } else if (P == 2) {
      const uint2 a = lds_u64(p0 + (2 * ky) * kCtu);
      const uint2 b = lds_u64(p0 + (2 * ky + 1) * kCtu);
      s[ky * 4 + 0] = int(__dp4a(b.x, 0x00000101u, __dp4a(a.x, 0x00000101u, 0u)));
      s[ky * 4 + 1] = int(__dp4a(b.x, 0x01010000u, __dp4a(a.x, 0x01010000u, 0u)));
      s[ky * 4 + 2] = int(__dp4a(b.y, 0x00000101u, __dp4a(a.y, 0x00000101u, 0u)));
      s[ky * 4 + 3] = int(__dp4a(b.y, 0x01010000u, __dp4a(a.y, 0x01010000u, 0u)));
    } else {
      uint32_t t0 = 0, t1 = 0, t2 = 0, t3 = 0;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const uint4 u = lds_u128(p0 + (4 * ky + a) * kCtu);
        t0 = __dp4a(u.x, 0x01010101u, t0);
        t1 = __dp4a(u.y, 0x01010101u, t1);
        t2 = __dp4a(u.z, 0x01010101u, t2);
        t3 = __dp4a(u.w, 0x01010101u, t3);
      }
      s[ky * 4 + 0] = int(t0), s[ky * 4 + 1] = int(t1), s[ky * 4 + 2] = int(t2), s[ky * 4 + 3] = int(t3);
    }
  }
}

__device__ __forceinline__ void load_patch(int pool, uint32_t p0, int (&s)[16]) {
  if (pool == 1) {
    load_patch_p<1>(p0, s);
  } else if (pool == 2) {
    load_patch_p<2>(p0, s);
  } else {
    load_patch_p<4>(p0, s);
  }
}

struct Region {
  uint32_t px;         // shared-memory address of the region origin inside its CTU tile
  __half* hi;          // feature rows of the region's CTU (global memory)
  __half* lo;
  bool valid;          // CTU exists (tail groups run with masked stores)
};

// One warp task: every lane runs the conv stack for its two regions A and B.
//   pool    1 / 2 / 4 (warp-uniform)       wb  shared-memory byte address of the branch's weight block
//   d       position of the lane inside its quad: conv3 tap (ky = d >> 1, kx = d & 1)
//   c2_off  offset of a region's 24 conv2 features, c3_off offset of its quad's 32 conv3 features
__device__ __noinline__ void warp_task(const int pool, const uint32_t wb, const float cst, const float fscale, const int d,
                                       const Region ra, const Region rb, const int c2_off_a, const int c2_off_b,
                                       const int c3_off_a, const int c3_off_b) {
  // ---- mean-removal window sums (integer, over the quad's 16x16 pooled samples)
  uint32_t sa = region_sum(pool, ra.px), sb = region_sum(pool, rb.px);
  sa += __shfl_xor_sync(0xffffffffu, sa, 1);
  sb += __shfl_xor_sync(0xffffffffu, sb, 1);
  sa += __shfl_xor_sync(0xffffffffu, sa, 2);
  sb += __shfl_xor_sync(0xffffffffu, sb, 2);
  const int wsum_a = int(sa), wsum_b = int(sb);

  // ---- conv1 (4x4/s4, 1->16) feeding conv2 (2x2/s2, 16->24) patch by patch
  float2 acc_a[12], acc_b[12];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    lds_pairs(wb + 4 * (kB2Off + 4 * i), acc_a[2 * i], acc_a[2 * i + 1]);
    acc_b[2 * i] = acc_a[2 * i], acc_b[2 * i + 1] = acc_a[2 * i + 1];
  }
  const int patch_step = 4 * pool;
#pragma unroll 1
  for (int patch = 0; patch < 4; ++patch) {
    const int poff = ((patch >> 1) * kCtu + (patch & 1)) * patch_step;
    int ps_a[16], ps_b[16];
    load_patch(pool, ra.px + poff, ps_a);
    load_patch(pool, rb.px + poff, ps_b);
    float2 a1_a[8], a1_b[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      lds_pairs(wb + 4 * (kB1Off + 4 * i), a1_a[2 * i], a1_a[2 * i + 1]);
      a1_b[2 * i] = a1_a[2 * i], a1_b[2 * i + 1] = a1_a[2 * i + 1];
    }
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      const float2 xa = splat(__int2float_rn(ps_a[t] * 256 - wsum_a) * cst);
      const float2 xb = splat(__int2float_rn(ps_b[t] * 256 - wsum_b) * cst);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 w0, w1;
        lds_pairs(wb + 4 * (kW1Off + t * 16 + 4 * i), w0, w1);
        a1_a[2 * i] = __ffma2_rn(xa, w0, a1_a[2 * i]);
        a1_a[2 * i + 1] = __ffma2_rn(xa, w1, a1_a[2 * i + 1]);
        a1_b[2 * i] = __ffma2_rn(xb, w0, a1_b[2 * i]);
        a1_b[2 * i + 1] = __ffma2_rn(xb, w1, a1_b[2 * i + 1]);
      }
    }
    const uint32_t w2 = wb + 4 * (kW2Off + patch * 16 * 24);
#pragma unroll
    for (int ci = 0; ci < 16; ++ci) {
      const float2 ca = splat(leaky((ci & 1) ? a1_a[ci >> 1].y : a1_a[ci >> 1].x));
      const float2 cb = splat(leaky((ci & 1) ? a1_b[ci >> 1].y : a1_b[ci >> 1].x));
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        float2 w0, w1;
        lds_pairs(w2 + 4 * (ci * 24 + 4 * i), w0, w1);
        acc_a[2 * i] = __ffma2_rn(ca, w0, acc_a[2 * i]);
        acc_a[2 * i + 1] = __ffma2_rn(ca, w1, acc_a[2 * i + 1]);
        acc_b[2 * i] = __ffma2_rn(cb, w0, acc_b[2 * i]);
        acc_b[2 * i + 1] = __ffma2_rn(cb, w1, acc_b[2 * i + 1]);
      }
    }
  }

  // ---- conv2 outputs: 24 features per region, stored as fp16 hi/lo
#pragma unroll
  for (int i = 0; i < 12; ++i) acc_a[i] = leaky2(acc_a[i]), acc_b[i] = leaky2(acc_b[i]);
  {
    uint32_t hw[12], lw[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) split_pair(acc_a[i], fscale, hw[i], lw[i]);
    if (ra.valid) {
      uint4* ph = reinterpret_cast<uint4*>(ra.hi + c2_off_a);
      uint4* pl = reinterpret_cast<uint4*>(ra.lo + c2_off_a);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        ph[i] = make_uint4(hw[4 * i], hw[4 * i + 1], hw[4 * i + 2], hw[4 * i + 3]);
        pl[i] = make_uint4(lw[4 * i], lw[4 * i + 1], lw[4 * i + 2], lw[4 * i + 3]);
      }
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) split_pair(acc_b[i], fscale, hw[i], lw[i]);
    if (rb.valid) {
      uint4* ph = reinterpret_cast<uint4*>(rb.hi + c2_off_b);
      uint4* pl = reinterpret_cast<uint4*>(rb.lo + c2_off_b);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        ph[i] = make_uint4(hw[4 * i], hw[4 * i + 1], hw[4 * i + 2], hw[4 * i + 3]);
        pl[i] = make_uint4(lw[4 * i], lw[4 * i + 1], lw[4 * i + 2], lw[4 * i + 3]);
      }
    }
  }

  // ---- conv3 (2x2/s2, 24->32): lane d contributes tap d; rolled over the four 8-channel output groups,
  // each summed across the quad with two xor-shuffles; lane d keeps group d.
  float2 res_a[4], res_b[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) res_a[i] = res_b[i] = make_float2(0.f, 0.f);
  const uint32_t w3 = wb + 4 * (kW3Off + d * kW3Stride);
#pragma unroll 1
  for (int og = 0; og < 4; ++og) {
    float2 pa[4], pb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) pa[i] = pb[i] = make_float2(0.f, 0.f);
    const uint32_t wg = w3 + 4 * (og * 24 * 8);
#pragma unroll
    for (int ci = 0; ci < 24; ++ci) {
      const float2 ca = splat((ci & 1) ? acc_a[ci >> 1].y : acc_a[ci >> 1].x);
      const float2 cb = splat((ci & 1) ? acc_b[ci >> 1].y : acc_b[ci >> 1].x);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float2 w0, w1;
        lds_pairs(wg + 4 * (ci * 8 + 4 * i), w0, w1);
        pa[2 * i] = __ffma2_rn(ca, w0, pa[2 * i]);
        pa[2 * i + 1] = __ffma2_rn(ca, w1, pa[2 * i + 1]);
        pb[2 * i] = __ffma2_rn(cb, w0, pb[2 * i]);
        pb[2 * i + 1] = __ffma2_rn(cb, w1, pb[2 * i + 1]);
      }
    }
    const bool mine = (og == d);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 va = pa[i], vb = pb[i];
      va.x += __shfl_xor_sync(0xffffffffu, va.x, 1), va.y += __shfl_xor_sync(0xffffffffu, va.y, 1);
      vb.x += __shfl_xor_sync(0xffffffffu, vb.x, 1), vb.y += __shfl_xor_sync(0xffffffffu, vb.y, 1);
      va.x += __shfl_xor_sync(0xffffffffu, va.x, 2), va.y += __shfl_xor_sync(0xffffffffu, va.y, 2);
      vb.x += __shfl_xor_sync(0xffffffffu, vb.x, 2), vb.y += __shfl_xor_sync(0xffffffffu, vb.y, 2);
      if (mine) res_a[i] = va, res_b[i] = vb;
    }
  }
  {
    float2 b[4];
    lds_pairs(wb + 4 * (kB3Off + 8 * d), b[0], b[1]);
    lds_pairs(wb + 4 * (kB3Off + 8 * d + 4), b[2], b[3]);
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      split_pair(leaky2(make_float2(res_a[i].x + b[i].x, res_a[i].y + b[i].y)), fscale, hw[i], lw[i]);
    if (ra.valid) {
      *reinterpret_cast<uint4*>(ra.hi + c3_off_a + 8 * d) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
      *reinterpret_cast<uint4*>(ra.lo + c3_off_a + 8 * d) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      split_pair(leaky2(make_float2(res_b[i].x + b[i].x, res_b[i].y + b[i].y)), fscale, hw[i], lw[i]);
    if (rb.valid) {
      *reinterpret_cast<uint4*>(rb.hi + c3_off_b + 8 * d) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
      *reinterpret_cast<uint4*>(rb.lo + c3_off_b + 8 * d) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
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
    const int d = lane & 3;
    for (int t = cw;; t += kConvComputeWarps) {
      const int j = t / kGroupTasks, task = t - j * kGroupTasks;
      const int g = blockIdx.x + j * gridDim.x;
      if (g >= n_groups) break;
      const int stage = j % kConvStages;
      const uint32_t parity = (j / kConvStages) & 1;
      mbar_wait(&full[stage], parity);
      const uint32_t tile0 = smem_u32(tiles + stage * kStageBytes);
      const int ctu0 = g * kGroupCtus;  // index inside this launch
      int pool, br, ca, cb, ry_a, rx_a, ry_b, rx_b, c2a, c2b, c3a, c3b;
      if (task < 16) {            // S: one CTU; regions A / B = upper / lower half
        const int q = lane >> 2, qy = q >> 2, qx = q & 3;
        pool = 1, br = 0, ca = cb = task;
        ry_a = 2 * qy + (d >> 1), ry_b = ry_a + 4, rx_a = rx_b = 2 * qx + (d & 1);
        c2a = kOffC2S + (ry_a * 8 + rx_a) * 24, c2b = kOffC2S + (ry_b * 8 + rx_b) * 24;
        c3a = kOffC3S + (qy * 4 + qx) * 32, c3b = kOffC3S + ((qy + 2) * 4 + qx) * 32;
      } else if (task < 20) {     // M: four CTUs; regions A / B in CTUs two apart
        const int l16 = lane & 15, q = l16 >> 2, qy = q >> 1, qx = q & 1;
        pool = 2, br = 1, ca = 4 * (task - 16) + (lane >> 4), cb = ca + 2;
        ry_a = ry_b = 2 * qy + (d >> 1), rx_a = rx_b = 2 * qx + (d & 1);
        c2a = c2b = kOffC2M + (ry_a * 4 + rx_a) * 24;
        c3a = c3b = kOffC3M + (qy * 2 + qx) * 32;
      } else {                    // L: sixteen CTUs; regions A / B in CTUs eight apart
        pool = 4, br = 2, ca = lane >> 2, cb = ca + 8;
        ry_a = ry_b = d >> 1, rx_a = rx_b = d & 1;
        c2a = c2b = kOffC2L + (ry_a * 2 + rx_a) * 24;
        c3a = c3b = kOffC3L;
      }
      const int rpx = 8 * pool;   // region edge in pixels
      Region ra, rb;
      ra.valid = (ctu0 + ca) < p.n_ctus, rb.valid = (ctu0 + cb) < p.n_ctus;
      const size_t row_a = size_t(ra.valid ? ctu0 + ca : 0) * kFeat, row_b = size_t(rb.valid ? ctu0 + cb : 0) * kFeat;
      ra.px = tile0 + ca * kTileBytes + (rpx * ry_a) * kCtu + rpx * rx_a;
      rb.px = tile0 + cb * kTileBytes + (rpx * ry_b) * kCtu + rpx * rx_b;
      ra.hi = p.feat_hi + row_a, ra.lo = p.feat_lo + row_a;
      rb.hi = p.feat_hi + row_b, rb.lo = p.feat_lo + row_b;
      warp_task(pool, wsm_addr + 4 * br * kConvBranchFloats, p.cst[br], p.feat_scale, d, ra, rb, c2a, c2b, c3a, c3b);
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
