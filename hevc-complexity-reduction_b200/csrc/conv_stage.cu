// Stage CONV: luma tiles -> 2688 conv features per CTU (sm_100a).
//
// Follows net_CNN.py:105-150 (AI) / ETH-CNN_Training_LDP/net_CTU64.py:102-175 (LDP):
//   x = X * scale; L = meanremove16(avgpool4(x)); M = meanremove16(avgpool2(x)); S = meanremove16(x)
//   per branch: conv 4x4/s4 1->16, conv 2x2/s2 16->24, conv 2x2/s2 24->32, leaky(0.2) after each
//   features = [c3_S | c3_M | c3_L | c2_S | c2_M | c2_L], each NHWC-flattened.
//
// Work decomposition.  All three branches have the same shape once the input is pooled: a lane owns an
// 8x8 block of (pooled) samples = 2x2 conv1 patches = one conv2 output position, and 2x2 neighbouring
// lanes (a "quad") share one mean-removal window (16x16 pooled samples) and one conv3 output position.
//   S: pool 1, lane region  8x8  px, 64 lanes per CTU -> a warp task covers half a CTU
//   M: pool 2, lane region 16x16 px, 16 lanes per CTU -> a warp task covers 2 CTUs
//   L: pool 4, lane region 32x32 px,  4 lanes per CTU -> a warp task covers 8 CTUs
// so a group of 8 CTUs is exactly 16 + 4 + 1 = 21 warp tasks of identical cost.  Every lane of a warp
// uses the same filter taps, so weights are broadcast 128-bit shared-memory loads and activations live
// in registers; the quad exchanges data only through warp shuffles (window sums, conv3 reduce-scatter).
//
// Mean removal is exact: with s = pooled integer sum and W = integer sum of the whole window,
//   pooled - mean = (256*s - W) / (256*pool^2), one rounding when multiplied by scale/(256*pool^2).
//
// A persistent CTA (one per SM) keeps the 58 KB of conv weights resident in shared memory and streams
// 8-CTU tile groups through a 4-deep TMA ring (3-D tensor map over (x, y, frame), 64x64x1 box; the
// zero fill of out-of-bounds rows/columns IS the reference's zero padding, video_to_cu_depth.py:54-57).
// One producer warp issues TMA, eleven compute warps take warp tasks round-robin.
#include "kernels.h"
#include "ptx_sm100.cuh"

namespace ethcnn {
namespace {

// 128-bit shared-memory load of four consecutive weights.  `asm volatile` on purpose: the conv1 filter is
// invariant across the patch loop and the compiler would otherwise hoist all 64 loads (256 registers).
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

__device__ __forceinline__ float leaky(float v) { return fmaxf(0.2f * v, v); }  // Maximum(alpha*x, x), alpha = 0.2

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// value * scale -> (hi, lo) fp16 pair with hi + lo == value*scale to ~22 bits
__device__ __forceinline__ void split_hi_lo(float v, float scale, float& hi, float& lo) {
  float s = v * scale;
  hi = __half2float(__float2half_rn(s));
  lo = s - hi;
}

// Integer sum of the lane's whole region ((8P) x (8P) pixels at reg0, row pitch 64).
template <int P>
__device__ __forceinline__ uint32_t region_sum(const uint8_t* __restrict__ reg0) {
  uint32_t s = 0;
#pragma unroll
  for (int r = 0; r < 8 * P; ++r) {
    const uint8_t* row = reg0 + r * kCtu;
    if (P == 1) {
      uint2 v = *reinterpret_cast<const uint2*>(row);
      s = __dp4a(v.x, 0x01010101u, s);
      s = __dp4a(v.y, 0x01010101u, s);
    } else {
#pragma unroll
      for (int k = 0; k < P / 2; ++k) {
        uint4 v = *reinterpret_cast<const uint4*>(row + 16 * k);
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
__device__ __forceinline__ void load_patch(const uint8_t* __restrict__ p0, uint32_t (&s)[16]) {
#pragma unroll
  for (int ky = 0; ky < 4; ++ky) {
    if (P == 1) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(p0 + ky * kCtu);
      s[ky * 4 + 0] = w & 0xffu;
      s[ky * 4 + 1] = (w >> 8) & 0xffu;
      s[ky * 4 + 2] = (w >> 16) & 0xffu;
      s[ky * 4 + 3] = w >> 24;
    } else if (P == 2) {
      const uint2 a = *reinterpret_cast<const uint2*>(p0 + (2 * ky) * kCtu);
      const uint2 b = *reinterpret_cast<const uint2*>(p0 + (2 * ky + 1) * kCtu);
      s[ky * 4 + 0] = __dp4a(b.x, 0x00000101u, __dp4a(a.x, 0x00000101u, 0u));
      s[ky * 4 + 1] = __dp4a(b.x, 0x01010000u, __dp4a(a.x, 0x01010000u, 0u));
      s[ky * 4 + 2] = __dp4a(b.y, 0x00000101u, __dp4a(a.y, 0x00000101u, 0u));
      s[ky * 4 + 3] = __dp4a(b.y, 0x01010000u, __dp4a(a.y, 0x01010000u, 0u));
    } else {
      uint32_t t0 = 0, t1 = 0, t2 = 0, t3 = 0;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const uint4 u = *reinterpret_cast<const uint4*>(p0 + (4 * ky + a) * kCtu);
        t0 = __dp4a(u.x, 0x01010101u, t0);
        t1 = __dp4a(u.y, 0x01010101u, t1);
        t2 = __dp4a(u.z, 0x01010101u, t2);
        t3 = __dp4a(u.w, 0x01010101u, t3);
      }
      s[ky * 4 + 0] = t0, s[ky * 4 + 1] = t1, s[ky * 4 + 2] = t2, s[ky * 4 + 3] = t3;
    }
  }
}

// The per-lane program shared by the three branches (see file header).
//   reg0    shared-memory address of the lane's region origin inside its CTU tile
//   wb      shared-memory weight block of the branch
//   d       position of the lane inside its quad: conv3 tap (ky = d >> 1, kx = d & 1)
//   hi_row / lo_row  feature rows of the lane's CTU (global memory)
//   c2_off  offset of the lane's 24 conv2 features, c3_off offset of its quad's 32 conv3 features
template <int P>
__device__ __forceinline__ void lane_program(const uint8_t* __restrict__ reg0, const uint32_t wb, float cst,
                                             float fscale, int d, __half* __restrict__ hi_row,
                                             __half* __restrict__ lo_row, int c2_off, int c3_off, bool valid) {
  uint32_t rsum = region_sum<P>(reg0);
  rsum += __shfl_xor_sync(0xffffffffu, rsum, 1);
  rsum += __shfl_xor_sync(0xffffffffu, rsum, 2);
  const int wsum = static_cast<int>(rsum);  // integer sum over the 16x16 pooled window (256*P*P pixels)

  float acc2[24];
  {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      float4 v = lds128(wb + 4 * (kB2Off + 4 * i));
      acc2[4 * i] = v.x, acc2[4 * i + 1] = v.y, acc2[4 * i + 2] = v.z, acc2[4 * i + 3] = v.w;
    }
  }
#pragma unroll 1
  for (int patch = 0; patch < 4; ++patch) {
    const int py = patch >> 1, px = patch & 1;
    uint32_t ps[16];
    load_patch<P>(reg0 + (4 * P * py) * kCtu + 4 * P * px, ps);
    float a1[16];
    {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 v = lds128(wb + 4 * (kB1Off + 4 * i));
        a1[4 * i] = v.x, a1[4 * i + 1] = v.y, a1[4 * i + 2] = v.z, a1[4 * i + 3] = v.w;
      }
    }
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      const float x = __int2float_rn(static_cast<int>(ps[t]) * 256 - wsum) * cst;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 v = lds128(wb + 4 * (kW1Off + t * 16 + 4 * i));
        a1[4 * i] = fmaf(x, v.x, a1[4 * i]);
        a1[4 * i + 1] = fmaf(x, v.y, a1[4 * i + 1]);
        a1[4 * i + 2] = fmaf(x, v.z, a1[4 * i + 2]);
        a1[4 * i + 3] = fmaf(x, v.w, a1[4 * i + 3]);
      }
    }
#pragma unroll
    for (int ci = 0; ci < 16; ++ci) {
      const float c = leaky(a1[ci]);
      const uint32_t w = wb + 4 * (kW2Off + (patch * 16 + ci) * 24);
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        float4 v = lds128(w + 16 * i);
        acc2[4 * i] = fmaf(c, v.x, acc2[4 * i]);
        acc2[4 * i + 1] = fmaf(c, v.y, acc2[4 * i + 1]);
        acc2[4 * i + 2] = fmaf(c, v.z, acc2[4 * i + 2]);
        acc2[4 * i + 3] = fmaf(c, v.w, acc2[4 * i + 3]);
      }
    }
  }

  // conv2 output of this lane: 24 features, stored as hi/lo fp16
#pragma unroll
  for (int i = 0; i < 24; ++i) acc2[i] = leaky(acc2[i]);
  if (valid) {
    uint32_t hw[12], lw[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      float h0, l0, h1, l1;
      split_hi_lo(acc2[2 * i], fscale, h0, l0);
      split_hi_lo(acc2[2 * i + 1], fscale, h1, l1);
      hw[i] = pack_half2(h0, h1);
      lw[i] = pack_half2(l0, l1);
    }
    uint4* ph = reinterpret_cast<uint4*>(hi_row + c2_off);
    uint4* pl = reinterpret_cast<uint4*>(lo_row + c2_off);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      ph[i] = make_uint4(hw[4 * i], hw[4 * i + 1], hw[4 * i + 2], hw[4 * i + 3]);
      pl[i] = make_uint4(lw[4 * i], lw[4 * i + 1], lw[4 * i + 2], lw[4 * i + 3]);
    }
  }

  // conv3: this lane contributes tap d (24 inputs) to all 32 channels of the quad's output position
  float part[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) part[i] = 0.f;
  {
    const uint32_t w3 = wb + 4 * (kW3Off + d * kW3Stride);
#pragma unroll
    for (int ci = 0; ci < 24; ++ci) {
      const float c = acc2[ci];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 v = lds128(w3 + 4 * (ci * 32 + 4 * i));
        part[4 * i] = fmaf(c, v.x, part[4 * i]);
        part[4 * i + 1] = fmaf(c, v.y, part[4 * i + 1]);
        part[4 * i + 2] = fmaf(c, v.z, part[4 * i + 2]);
        part[4 * i + 3] = fmaf(c, v.w, part[4 * i + 3]);
      }
    }
  }
  // reduce-scatter over the quad: lane d ends with channels [8d, 8d+8)
  float r16[16];
  const bool up2 = (d & 2) != 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float send = up2 ? part[i] : part[16 + i];
    const float keep = up2 ? part[16 + i] : part[i];
    r16[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  float r8[8];
  const bool up1 = (d & 1) != 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float send = up1 ? r16[i] : r16[8 + i];
    const float keep = up1 ? r16[8 + i] : r16[i];
    r8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
  }
  {
    const float4 u = lds128(wb + 4 * (kB3Off + 8 * d)), v = lds128(wb + 4 * (kB3Off + 8 * d + 4));
    r8[0] = leaky(r8[0] + u.x), r8[1] = leaky(r8[1] + u.y), r8[2] = leaky(r8[2] + u.z), r8[3] = leaky(r8[3] + u.w);
    r8[4] = leaky(r8[4] + v.x), r8[5] = leaky(r8[5] + v.y), r8[6] = leaky(r8[6] + v.z), r8[7] = leaky(r8[7] + v.w);
  }
  if (valid) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float h0, l0, h1, l1;
      split_hi_lo(r8[2 * i], fscale, h0, l0);
      split_hi_lo(r8[2 * i + 1], fscale, h1, l1);
      hw[i] = pack_half2(h0, h1);
      lw[i] = pack_half2(l0, l1);
    }
    *reinterpret_cast<uint4*>(hi_row + c3_off + 8 * d) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(lo_row + c3_off + 8 * d) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  }
}

constexpr int kTileBytes = kCtu * kCtu;                       // 4096
constexpr int kStageBytes = kGroupCtus * kTileBytes;          // 32768
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
    for (int t = cw;; t += kConvComputeWarps) {
      const int j = t / kGroupTasks, task = t - j * kGroupTasks;
      const int g = blockIdx.x + j * gridDim.x;
      if (g >= n_groups) break;
      const int stage = j % kConvStages;
      const uint32_t parity = (j / kConvStages) & 1;
      mbar_wait(&full[stage], parity);
      const uint8_t* tile0 = tiles + stage * kStageBytes;
      const int ctu0 = g * kGroupCtus;  // index inside this launch
      if (task < 16) {
        const int c = task >> 1, half = task & 1;
        const int q = lane >> 2, d = lane & 3;
        const int qy = half * 2 + (q >> 2), qx = q & 3;
        const int ry = 2 * qy + (d >> 1), rx = 2 * qx + (d & 1);
        const int n = ctu0 + c;
        const bool valid = n < p.n_ctus;
        const size_t row = size_t(valid ? n : 0) * kFeat;
        lane_program<1>(tile0 + c * kTileBytes + (8 * ry) * kCtu + 8 * rx, wsm_addr, p.cst[0], p.feat_scale, d, p.feat_hi + row,
                        p.feat_lo + row, kOffC2S + (ry * 8 + rx) * 24, kOffC3S + (qy * 4 + qx) * 32, valid);
      } else if (task < 20) {
        const int c = 2 * (task - 16) + (lane >> 4);
        const int l16 = lane & 15, q = l16 >> 2, d = l16 & 3;
        const int qy = q >> 1, qx = q & 1;
        const int ry = 2 * qy + (d >> 1), rx = 2 * qx + (d & 1);
        const int n = ctu0 + c;
        const bool valid = n < p.n_ctus;
        const size_t row = size_t(valid ? n : 0) * kFeat;
        lane_program<2>(tile0 + c * kTileBytes + (16 * ry) * kCtu + 16 * rx, wsm_addr + 4 * kConvBranchFloats, p.cst[1], p.feat_scale,
                        d, p.feat_hi + row, p.feat_lo + row, kOffC2M + (ry * 4 + rx) * 24, kOffC3M + (qy * 2 + qx) * 32,
                        valid);
      } else {
        const int c = lane >> 2, d = lane & 3;
        const int ry = d >> 1, rx = d & 1;
        const int n = ctu0 + c;
        const bool valid = n < p.n_ctus;
        const size_t row = size_t(valid ? n : 0) * kFeat;
        lane_program<4>(tile0 + c * kTileBytes + (32 * ry) * kCtu + 32 * rx, wsm_addr + 8 * kConvBranchFloats, p.cst[2],
                        p.feat_scale, d, p.feat_hi + row, p.feat_lo + row, kOffC2L + (ry * 2 + rx) * 24, kOffC3L, valid);
      }
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
