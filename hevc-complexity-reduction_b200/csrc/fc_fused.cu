// Stages FC1 + FC2 + FC3 + sigmoid in ONE tcgen05 kernel (sm_100a).
//   net_CNN.py:156-185:  a1 = leaky(f W1 + b1);  a2 = leaky([a1, q] W2 + b2);  y = sigmoid([a2, q] W3 + b3)
//
// Both dense contractions run on the 5th-generation tensor cores with fp32 accumulation in TMEM.  A single
// fp16 pass misses the reference by ~3e-4 in probability, so every operand is
// scaled by a power of two and split into fp16 hi + lo parts and three MMAs (hi*hi, hi*lo, lo*hi) are
// accumulated; the dropped lo*lo term is 2^-22 relative.  The qp input (last row of W2 / W3) and the
// biases are folded into b2eff / b3eff on the host.
//
// Tiles.  M = 128 CTUs (one TMEM lane per CTU).  N is split on head boundaries so FC2 stays inside a CTA:
//   type 1: head 16        FC1 N = 256 (cols 192..447), FC2 K = 256 (4 slices), N = 192, FC3 192 -> 16
//   type 0: heads 64 + 32  FC1 N = 192 (cols 0..191),   FC2 64 -> 48 (1 slice) and 128 -> 96 (2 slices)
// The static tile order pairs the two tiles of an M block on neighbouring CTAs (tile_of).
//
// One shared-memory ring (2 stages x 96 KB of 128-byte-swizzled K = 64 slices; -DETHCNN_FC_BK=32 builds 4 stages
// x 48 KB with 64-byte swizzle instead) carries BOTH contractions:
//   FC1 stage: A_hi, A_lo (features, TMA) + B_hi, B_lo (W1 slice, TMA)
//   FC2 stage: A_hi, A_lo = the tile's own FC1 activations, written by the epilogue warps straight from
//              TMEM (bias + leaky + re-split) into the swizzled layout + B_hi, B_lo (W2 slice, TMA)
// so a1 never leaves the SM.  TMEM: accumulator 1 (FC1, 256 columns) and accumulator 2 (FC2, <= 192).
// Warp roles: 0 = TMA producer, 1 = TMEM owner + MMA issuer (one elected lane), 2..5 = epilogue
// (TMEM lane quarter = warp % 4): epi1 = acc1 -> a1 slices, epi2 = acc2 -> a2 -> FC3 (FFMA) -> sigmoid ->
// 84-byte probability rows + gate flags.  epi2 of tile i overlaps the FC1 MMAs of tile i+1.
#include <cstring>

#include "fc_fused.h"
#include <cstdlib>

#include "kernels.h"
#include "ptx_sm100.cuh"

namespace ethcnn {
namespace {

#ifndef ETHCNN_FC_BK
#define ETHCNN_FC_BK 64
#endif
constexpr int kBM = 128, kBK = ETHCNN_FC_BK;                 // K slice = one swizzle row: 64 fp16 (128 B) or 32 fp16 (64 B)
static_assert(kBK == 64 || kBK == 32, "K slice must be one 128-byte or one 64-byte swizzle row");
constexpr int kRowBytes = kBK * 2;
constexpr int kABytes = kBM * kRowBytes;                     // 16384 | 8192
constexpr int kKSteps = kFeat / kBK;                         // 42 | 84
constexpr int kAcc2Col = 256;                                // TMEM column of accumulator 2
#ifndef ETHCNN_FC_EPI_WARPS
#define ETHCNN_FC_EPI_WARPS 8
#endif
constexpr int kEpiWarps = ETHCNN_FC_EPI_WARPS;                // 4: one warp per TMEM lane quarter; 8: two, splitting the columns of epi1
static_assert(kEpiWarps == 4 || kEpiWarps == 8, "epilogue warps come in sets of four (one per TMEM lane quarter)");
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kW3Floats = 48 * 1 + 96 * 4 + 192 * 16;        // 3504
constexpr int kTableFloats = kW3Floats + 336 + 21 + kFc1 + 3; // w3 | b2eff | b3eff | b1 (padded to 16 B)
// kCtas = 1: one CTA per 128-CTU tile.  kCtas = 2: a CTA PAIR (cluster of two, tcgen05 cta_group::2) per 256-CTU tile;
// each CTA stages its own 128 feature rows and HALF of the weight rows, so the weight traffic per CTU halves.
template <int kCtas>
struct Geo {
  static constexpr int kBRowsMax = 256 / kCtas;                                   // weight rows a CTA stages per slice
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBRowsMax * kRowBytes;     // K = 64: 96 KB | 64 KB
  static constexpr int kStages = 192 * 1024 / kStageBytes;                        // K = 64: 2 | 3
  static constexpr int kSmemBytes = kStages * kStageBytes + kTableFloats * 4 + 256 + 1024;
  static constexpr int kA2Bars = kStages > 256 / kBK ? kStages : 256 / kBK;      // a2_full: per ring stage, or per a1 slice of a tile
  static_assert((2 * kStages + kA2Bars + 5) * 8 <= 256, "barrier block overflows its shared-memory slot");
};

// K-major operand tile in swizzled shared memory: rows of kBK fp16 (one swizzle row), 8-row groups back to back.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3ffffu) >> 4);  // start address  [0,14)
  d |= uint64_t(1) << 16;                      // leading byte offset (ignored for swizzled K-major)
  d |= uint64_t((8 * kRowBytes) >> 4) << 32;   // stride byte offset: one 8-row group
  d |= uint64_t(1) << 46;                      // descriptor version (sm_100)
  d |= uint64_t(kBK == 64 ? 2 : 4) << 61;      // SWIZZLE_128B | SWIZZLE_64B
  return d;
}
// kind::f16 instruction descriptor: D fp32, A = B = fp16, K-major both, M = 128, N as given.
template <int kCtas>
__device__ __forceinline__ uint32_t idesc_f16(int n) {
  return (1u << 4) | (uint32_t(n >> 3) << 17) | (uint32_t((kBM * kCtas) >> 4) << 24);
}
// The MMA and commit wrappers are called by ALL lanes of the converged issuer warp with warp-uniform operands; elect.sync
// picks the one lane that really issues.  (Issued from inside `if (lane == 0)`, ptxas wraps every tcgen05.mma in a divergence
// "waterfall" loop -- ELECT / R2UR.BROADCAST / BRA.U.ANY -- that costs ~110 cycles per MMA (tools/microbench/umma_probe.cu),
// about as much as the MMA itself.)
template <int kCtas>
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if (kCtas == 1)
    asm volatile(
        "{\n"
        ".reg .pred p, e;\n"
        ".reg .b32 rx;\n"
        "elect.sync rx|e, 0xffffffff;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n"
        ".reg .pred p, e;\n"
        ".reg .b32 rx;\n"
        "elect.sync rx|e, 0xffffffff;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// The same with the A operand in TENSOR MEMORY (lane = row, every 32-bit column two consecutive K elements, low half first).
template <int kCtas>
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if (kCtas == 1)
    asm volatile(
        "{\n.reg .pred p, e;\n.reg .b32 rx;\nelect.sync rx|e, 0xffffffff;\nsetp.ne.b32 p, %4, 0;\n"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n.reg .pred p, e;\n.reg .b32 rx;\nelect.sync rx|e, 0xffffffff;\nsetp.ne.b32 p, %4, 0;\n"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// FC2's A operand: 1 = the epilogue warps write a1 (fp16 hi / lo) back into the tensor-memory columns they just drained and the
// FC2 MMAs read it from there (TS mode): no shared-memory stores, no generic->async proxy fence (~1 000 cycles per slice on the
// tensor pipe's critical path), no ring stage to wait for.  0 = through the shared-memory ring (round 1).
#ifndef ETHCNN_FC_A2_TMEM
#define ETHCNN_FC_A2_TMEM (ETHCNN_FC_BK == 64)
#endif
constexpr bool kA2Tmem = ETHCNN_FC_A2_TMEM;

// Arrive on `bar` when all MMAs issued so far by the elected lane have retired; for a pair, on that barrier in BOTH CTAs.
// (elect.sync returns the same lane every time for a full mask, so the commit tracks the MMAs above.)
template <int kCtas>
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  if (kCtas == 1)
    asm volatile(
        "{\n.reg .pred e;\n.reg .b32 rx;\nelect.sync rx|e, 0xffffffff;\n"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(smem_u32(bar))
        : "memory");
  else
    asm volatile(
        "{\n.reg .pred e;\n.reg .b32 rx;\nelect.sync rx|e, 0xffffffff;\n"
        "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n}\n" ::"r"(smem_u32(bar)),
        "h"(uint16_t(3))
        : "memory");
}
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float leaky(float v) { return fmaxf(0.2f * v, v); }
__device__ __forceinline__ float sigmoidf(float v) { return 1.0f / (1.0f + expf(-v)); }

#ifdef ETHCNN_EXP_FC_TIMING   // measurement only: where the MMA issuer (leader CTA) spends its cycles
__device__ unsigned long long g_fc_issuer[8];
__device__ unsigned long long g_fc_epi[8];
#define FCT_BEGIN() long long fct_t = clock64()
#define FCT_MARK(k)                                                              \
  do {                                                                           \
    const long long fct_n = clock64();                                           \
    if (lane == 0) fct_acc[k] += (unsigned long long)(fct_n - fct_t);            \
    fct_t = fct_n;                                                               \
  } while (0)
#else
#define FCT_BEGIN()
#define FCT_MARK(k)
#endif

struct TileInfo {
  int type;      // 1 = head 16, 0 = heads 64 + 32
  int m0;        // first CTU row of the tile
  int n0;        // first FC1 column
  int n1;        // FC1 N
  int nslices;   // FC2 K slices (= n1 / 64)
};
// Static tile order: CTAs 2j and 2j+1 work on the SAME 128 CTUs at the same time, one on the head-16 columns and
// one on the head-64/32 columns, and swap roles every iteration (type-1 tiles cost 4/3 of type-0 tiles).  The
// feature slices of an M tile are therefore requested twice within a few microseconds and come from HBM once.
// (A "unit" is a CTA, or a CTA pair; an M block is the 128 or 256 CTUs a unit covers.)
__device__ __forceinline__ int tile_of(int unit, int iter, int n_units, int m_blocks) {
  const int m = (unit >> 1) + (n_units >> 1) * iter;
  return m < m_blocks ? 2 * m + ((unit ^ iter) & 1) : -1;
}
__device__ __forceinline__ TileInfo decode_tile(int t, int rows_per_block, int row_ofs) {
  TileInfo ti;
  ti.type = t & 1;
  ti.m0 = (t >> 1) * rows_per_block + row_ofs;
  ti.n0 = ti.type ? 192 : 0;
  ti.n1 = ti.type ? 256 : 192;
  ti.nslices = ti.n1 / kBK;
  return ti;
}
// FC2 slice j of a tile: which head's W2, which K offset inside it, N2, accumulator-2 column, first slice of the head?
__device__ __forceinline__ void fc2_slice(int type, int j, int& head, int& kofs, int& n2, int& acc_col, int& first) {
  if (type) {
    head = 2, kofs = j * kBK, n2 = 192, acc_col = 0, first = (j == 0);
  } else if (j < 64 / kBK) {
    head = 0, kofs = j * kBK, n2 = 48, acc_col = 0, first = (j == 0);
  } else {
    head = 1, kofs = (j - 64 / kBK) * kBK, n2 = 96, acc_col = 48, first = (j == 64 / kBK);
  }
}

#define FC_FUSED_PARAMS                                                                                          \
  const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,                    \
      const __grid_constant__ CUtensorMap w1_hi_t0, const __grid_constant__ CUtensorMap w1_lo_t0,                \
      const __grid_constant__ CUtensorMap w1_hi_t1, const __grid_constant__ CUtensorMap w1_lo_t1,                \
      const __grid_constant__ CUtensorMap w2_hi_0, const __grid_constant__ CUtensorMap w2_lo_0,                  \
      const __grid_constant__ CUtensorMap w2_hi_1, const __grid_constant__ CUtensorMap w2_lo_1,                  \
      const __grid_constant__ CUtensorMap w2_hi_2, const __grid_constant__ CUtensorMap w2_lo_2,                  \
      const __grid_constant__ FusedParams p, const int m_blocks
#define FC_FUSED_ARGS \
  map_a_hi, map_a_lo, w1_hi_t0, w1_lo_t0, w1_hi_t1, w1_lo_t1, w2_hi_0, w2_lo_0, w2_hi_1, w2_lo_1, w2_hi_2, w2_lo_2, p, m_blocks

// Barriers.  In a pair, the MMAs are issued by the leader CTA (cluster rank 0) only, so everything the issuer waits
// for -- full (TMA bytes of BOTH CTAs), a2_full, acc1_empty, acc2_empty (epilogue warps of both CTAs) -- lives in the
// leader's shared memory and the peer arrives remotely; what the issuer signals -- empty, acc1_full, acc2_full -- is
// committed to the same barrier in both CTAs (multicast commit).
template <int kCtas>
__device__ __forceinline__ void fc_fused_body(const CUtensorMap& map_a_hi, const CUtensorMap& map_a_lo, const CUtensorMap& w1_hi_t0,
                                              const CUtensorMap& w1_lo_t0, const CUtensorMap& w1_hi_t1, const CUtensorMap& w1_lo_t1,
                                              const CUtensorMap& w2_hi_0, const CUtensorMap& w2_lo_0, const CUtensorMap& w2_hi_1,
                                              const CUtensorMap& w2_lo_1, const CUtensorMap& w2_hi_2, const CUtensorMap& w2_lo_2,
                                              const FusedParams& p, const int m_blocks) {
  constexpr int kStages = Geo<kCtas>::kStages, kStageBytes = Geo<kCtas>::kStageBytes;
  const uint32_t rank = kCtas == 2 ? cluster_ctarank() : 0u;   // 0 = leader
  const int unit = blockIdx.x / kCtas, n_units = gridDim.x / kCtas;
  const int row_ofs = int(rank) * kBM;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* tab = reinterpret_cast<float*>(smem + kStages * kStageBytes);
  float* w3s = tab;                    // [3504]
  float* b2s = tab + kW3Floats;        // [336]
  float* b3s = b2s + 336;              // [21] (+3 pad)
  float* b1s = b3s + 24;               // [448]
  uint64_t* bars = reinterpret_cast<uint64_t*>(tab + kTableFloats);
  uint64_t* full = bars;                        // [kStages]  TMA bytes landed
  uint64_t* empty = bars + kStages;             // [kStages]  MMAs that read the stage have retired
  // a2_full: the epilogue warps have written the FC2 A operand.  Through shared memory it is indexed by ring stage (the epilogue
  // waits for the stage's `empty` first, so it can never be two phases ahead of the issuer).  Through tensor memory nothing holds
  // the epilogue back within a tile -- with 4 slices on 3 stages it could complete the same stage's barrier TWICE before the
  // issuer looked once (seen as a launch failure from the deadlock trap, with the QP 20~25 weights only) -- so there it is indexed
  // by slice, one phase per tile.
  uint64_t* a2_full = bars + 2 * kStages;       // [kA2Bars]
  uint64_t* acc1_full = a2_full + Geo<kCtas>::kA2Bars;     // FC1 accumulator complete
  uint64_t* acc1_empty = acc1_full + 1;         // epilogue has drained accumulator 1
  uint64_t* acc2_full = acc1_full + 2;
  uint64_t* acc2_empty = acc1_full + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc1_full + 4);
  // cluster addresses of the leader's copies (identical shared-memory layout in both CTAs)
  auto at_leader = [&](uint64_t* bar) -> uint32_t { return kCtas == 2 ? mapa_shared(smem_u32(bar), 0) : smem_u32(bar); };
  auto arrive_at_leader = [&](uint64_t* bar) {
#ifdef ETHCNN_FC_CLUSTER_RELEASE   // the round-1 form: every hand-off with a cluster-scope release
    if (kCtas == 2) mbar_arrive_remote(mapa_shared(smem_u32(bar), 0));
#else
    if (kCtas == 2) mbar_arrive_remote_relaxed_scope(mapa_shared(smem_u32(bar), 0));
#endif
    else mbar_arrive(bar);
  };
  auto wait_shared = [&](uint64_t* bar, uint32_t parity) {   // a leader barrier both CTAs arrive on
    if (kCtas == 2) mbar_wait_cluster(bar, parity);
    else mbar_wait(bar, parity);
  };
  auto load_tile = [&](void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
    if (kCtas == 2) tma_load_2d_pair(dst, map, at_leader(bar), x, y);
    else tma_load_2d(dst, map, bar, x, y);
  };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  pdl_launch_dependents();   // the gate kernel may be scheduled behind us; it waits for our completion itself
  {
    // Tables from global memory: every load of the thread is in flight before its first store (a rolled scalar copy loop pays
    // the L2 latency once per iteration, ~5 us of prologue that the CTAs starting behind the last conv CTAs cannot hide).
    constexpr int kW3Vec = kW3Floats / 4, kW3Iters = (kW3Vec + kThreads - 1) / kThreads;
    static_assert(kW3Floats % 4 == 0 && kFc1 % 4 == 0 && kFc1 / 4 <= kThreads, "vector copy of the tables");
    const float4* w3g = reinterpret_cast<const float4*>(p.w3);
    float4 wv[kW3Iters];
#pragma unroll
    for (int k = 0; k < kW3Iters; ++k)
      if (int(threadIdx.x) + k * kThreads < kW3Vec) wv[k] = w3g[threadIdx.x + k * kThreads];
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x < kFc1 / 4) bv = reinterpret_cast<const float4*>(p.b1)[threadIdx.x];
#pragma unroll
    for (int k = 0; k < kW3Iters; ++k)
      if (int(threadIdx.x) + k * kThreads < kW3Vec) reinterpret_cast<float4*>(w3s)[threadIdx.x + k * kThreads] = wv[k];
    // b1 pre-scaled by 2^a1_exp: leaky(acc u + b) 2^e == leaky(acc (u 2^e) + b 2^e) exactly, one multiply less per activation
    if (threadIdx.x < kFc1 / 4)
      reinterpret_cast<float4*>(b1s)[threadIdx.x] =
          make_float4(bv.x * p.a1_scale, bv.y * p.a1_scale, bv.z * p.a1_scale, bv.w * p.a1_scale);
  }
  for (int i = threadIdx.x; i < 336; i += kThreads) b2s[i] = p.b2eff[i];
  if (threadIdx.x < 21) b3s[threadIdx.x] = p.b3eff[threadIdx.x];

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a_hi), prefetch_tmap(&map_a_lo);
    prefetch_tmap(&w1_hi_t0), prefetch_tmap(&w1_lo_t0), prefetch_tmap(&w1_hi_t1), prefetch_tmap(&w1_lo_t1);
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 2), mbar_init(&empty[s], 1);
    for (int s = 0; s < Geo<kCtas>::kA2Bars; ++s) mbar_init(&a2_full[s], kEpiWarps * kCtas);
    mbar_init(acc1_full, 1), mbar_init(acc1_empty, kEpiWarps * kCtas), mbar_init(acc2_full, 1), mbar_init(acc2_empty, 4 * kCtas);
    mbar_fence_init();
  }
  if (warp == 1) {   // the same warp of both CTAs of a pair allocates collectively
    if (kCtas == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kCtas == 2) cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // provably warp-uniform: feeds the MMA's uniform registers
  // Everything above (tables, barriers, tensor memory) touched only weights and this CTA's own state: when launched as a
  // programmatic dependent of the conv kernel it ran while that kernel was finishing.  The features are read below.
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------ TMA producer ------------------------------------------------
    // Issuing a TMA costs the issuing thread a few hundred cycles, so the work is spread over three lanes that walk
    // the ring in lockstep (all wait on the same empty barrier): lane 0 loads the A tiles (features), lane 1 the B
    // tiles (weights), lane 2 runs an L2 prefetch of the feature slices a few stages ahead (they come from HBM: a
    // chunk of features is larger than L2).  full[s] therefore expects two arrivals, each with its own byte count.
#ifndef ETHCNN_FC_L2_PREFETCH
#define ETHCNN_FC_L2_PREFETCH (384 / ETHCNN_FC_BK)
#endif
    if (lane < 3) {
      int it = 0;
      for (int iter = 0, t; (t = tile_of(unit, iter, n_units, m_blocks)) >= 0; ++iter) {
        const TileInfo ti = decode_tile(t, kBM * kCtas, row_ofs);
        const CUtensorMap* wh = ti.type ? &w1_hi_t1 : &w1_hi_t0;
        const CUtensorMap* wl = ti.type ? &w1_lo_t1 : &w1_lo_t0;
        const int nb = ti.n1 / kCtas;                 // weight rows this CTA stages
        const int nrow = ti.n0 + int(rank) * nb;
        if (lane == 2) {
          for (int ks = 0; ks < ETHCNN_FC_L2_PREFETCH && ks < kKSteps; ++ks) {
            tma_prefetch_l2_2d(&map_a_hi, ks * kBK, ti.m0);
            tma_prefetch_l2_2d(&map_a_lo, ks * kBK, ti.m0);
          }
        }
        for (int ks = 0; ks < kKSteps; ++ks, ++it) {
          const int s = it % kStages;
          mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1);
          uint8_t* st = smem + s * kStageBytes;
          if (lane == 0) {
#ifdef ETHCNN_EXP_FC_SKIP_ALO   // measurement only (wrong results): a quarter less operand traffic -- is the kernel bound by its L2 feed?
            if (rank == 0) mbar_arrive_expect_tx(&full[s], kCtas * kABytes);
            load_tile(st, &map_a_hi, &full[s], ks * kBK, ti.m0);
#else
            if (rank == 0) mbar_arrive_expect_tx(&full[s], kCtas * 2 * kABytes);
            load_tile(st, &map_a_hi, &full[s], ks * kBK, ti.m0);
            load_tile(st + kABytes, &map_a_lo, &full[s], ks * kBK, ti.m0);
#endif
          } else if (lane == 1) {
            if (rank == 0) mbar_arrive_expect_tx(&full[s], kCtas * 2 * nb * kRowBytes);
            load_tile(st + 2 * kABytes, wh, &full[s], ks * kBK, nrow);
            load_tile(st + 2 * kABytes + nb * kRowBytes, wl, &full[s], ks * kBK, nrow);
          } else if (ks + ETHCNN_FC_L2_PREFETCH < kKSteps) {
            tma_prefetch_l2_2d(&map_a_hi, (ks + ETHCNN_FC_L2_PREFETCH) * kBK, ti.m0);
            tma_prefetch_l2_2d(&map_a_lo, (ks + ETHCNN_FC_L2_PREFETCH) * kBK, ti.m0);
          }
        }
        for (int j = 0; j < ti.nslices; ++j, ++it) {  // FC2 stages: only the W2 slice comes through TMA
          const int s = it % kStages;
          mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1);
          uint8_t* st = smem + s * kStageBytes;
          int head, kofs, n2, acc_col, first;
          fc2_slice(ti.type, j, head, kofs, n2, acc_col, first);
          const int nb2 = n2 / kCtas;
          if (lane == 0) {
            if (rank == 0) mbar_arrive(&full[s]);
          } else if (lane == 1) {
            const CUtensorMap* bh = head == 0 ? &w2_hi_0 : (head == 1 ? &w2_hi_1 : &w2_hi_2);
            const CUtensorMap* bl = head == 0 ? &w2_lo_0 : (head == 1 ? &w2_lo_1 : &w2_lo_2);
            if (rank == 0) mbar_arrive_expect_tx(&full[s], kCtas * 2 * nb2 * kRowBytes);
            load_tile(st + 2 * kABytes, bh, &full[s], kofs, int(rank) * nb2);
            load_tile(st + 2 * kABytes + nb2 * kRowBytes, bl, &full[s], kofs, int(rank) * nb2);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (the leader CTA of a pair only) -------------------------
    int it = 0, tile_i = 0, a2_cnt[Geo<kCtas>::kA2Bars] = {};
#ifdef ETHCNN_EXP_FC_TIMING
    unsigned long long fct_acc[8] = {};
#endif
    FCT_BEGIN();
    for (int t; rank == 0 && (t = tile_of(unit, tile_i, n_units, m_blocks)) >= 0; ++tile_i) {
      const TileInfo ti = decode_tile(t, kBM * kCtas, row_ofs);
      const int nb = ti.n1 / kCtas;
      wait_shared(acc1_empty, (tile_i & 1) ^ 1);
      FCT_MARK(0);   // waiting for accumulator 1 to be drained
      tc_fence_after();
      const uint32_t idesc1 = idesc_f16<kCtas>(ti.n1);
      for (int ks = 0; ks < kKSteps; ++ks, ++it) {
        const int s = it % kStages;
        mbar_wait(&full[s], (it / kStages) & 1);
        FCT_MARK(1);   // FC1: waiting for the operands of the stage
        tc_fence_after();
        __syncwarp();
        {   // all lanes, converged; one elected lane issues (see umma_f16)
          const uint32_t base = smem_u32(smem + s * kStageBytes);
          const uint64_t da_hi = umma_desc(base), da_lo = umma_desc(base + kABytes);
          const uint64_t db_hi = umma_desc(base + 2 * kABytes), db_lo = umma_desc(base + 2 * kABytes + nb * kRowBytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t adv = uint64_t(k * 32 >> 4);  // 16 fp16 = 32 bytes along K inside the swizzle atom
            umma_f16<kCtas>(tmem_base, da_hi + adv, db_hi + adv, idesc1, (ks | k) != 0);
            umma_f16<kCtas>(tmem_base, da_hi + adv, db_lo + adv, idesc1, 1);
            umma_f16<kCtas>(tmem_base, da_lo + adv, db_hi + adv, idesc1, 1);
          }
          umma_commit<kCtas>(&empty[s]);
          if (ks == kKSteps - 1) umma_commit<kCtas>(acc1_full);
        }
        __syncwarp();
        FCT_MARK(2);   // FC1: issuing
      }
      // FC2: A operand = this tile's a1 slices written by the epilogue warps into the ring
      wait_shared(acc2_empty, (tile_i & 1) ^ 1);
      FCT_MARK(3);   // waiting for accumulator 2 of the previous tile to be drained
      tc_fence_after();
      for (int j = 0; j < ti.nslices; ++j, ++it) {
        const int s = it % kStages;
        int head, kofs, n2, acc_col, first;
        fc2_slice(ti.type, j, head, kofs, n2, acc_col, first);
        mbar_wait(&full[s], (it / kStages) & 1);
        const int a2i = kA2Tmem ? j : s;   // a slice index is not used by every tile (type 0 has one slice less): count uses
        wait_shared(&a2_full[a2i], a2_cnt[a2i] & 1);
        ++a2_cnt[a2i];
        FCT_MARK(4);   // FC2: waiting for W2 and the a1 slice from the epilogue warps
        tc_fence_after();
        __syncwarp();
        {
          const uint32_t base = smem_u32(smem + s * kStageBytes);
          const uint64_t da_hi = umma_desc(base), da_lo = umma_desc(base + kABytes);
          const uint64_t db_hi = umma_desc(base + 2 * kABytes), db_lo = umma_desc(base + 2 * kABytes + (n2 / kCtas) * kRowBytes);
          const uint32_t idesc2 = idesc_f16<kCtas>(n2);
          const uint32_t d2 = tmem_base + kAcc2Col + acc_col;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t adv = uint64_t(k * 32 >> 4);
            if (kA2Tmem) {
              // a1 of K elements [64 j + 16 k, +16): hi pairs in 8 columns of the 32-column block the values came from, lo 16 further
              const uint32_t a_hi = tmem_base + uint32_t(j * kBK + (k >> 1) * 32 + (k & 1) * 8), a_lo = a_hi + 16;
              umma_f16_ts<kCtas>(d2, a_hi, db_hi + adv, idesc2, (first && k == 0) ? 0u : 1u);
              umma_f16_ts<kCtas>(d2, a_hi, db_lo + adv, idesc2, 1);
              umma_f16_ts<kCtas>(d2, a_lo, db_hi + adv, idesc2, 1);
            } else {
              umma_f16<kCtas>(d2, da_hi + adv, db_hi + adv, idesc2, (first && k == 0) ? 0u : 1u);
              umma_f16<kCtas>(d2, da_hi + adv, db_lo + adv, idesc2, 1);
              umma_f16<kCtas>(d2, da_lo + adv, db_hi + adv, idesc2, 1);
            }
          }
          umma_commit<kCtas>(&empty[s]);
          if (j == ti.nslices - 1) umma_commit<kCtas>(acc2_full);
        }
        __syncwarp();
        FCT_MARK(5);   // FC2: issuing
      }
    }
#ifdef ETHCNN_EXP_FC_TIMING
    if (lane == 0 && rank == 0)
      for (int k = 0; k < 6; ++k) atomicAdd(&g_fc_issuer[k], fct_acc[k]);
    if (lane == 0 && rank == 0) atomicAdd(&g_fc_issuer[6], 1ull);
#endif
  } else {
    // ------------------------------------------------ epilogue warps ------------------------------------------------
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    // with eight epilogue warps the two warps of a quarter split the columns of every a1 slice (the drain of accumulator 1 sits
    // between FC1(i) and FC2(i) on the tensor pipe's critical path); accumulator 2 is drained by the first set only
    const int cset = (warp - 2) >> 2;       // 0 | 1
    const int row_l = q * 32 + lane;        // row inside the tile = TMEM lane
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    int it = 0, tile_i = 0;
#ifdef ETHCNN_EXP_FC_TIMING
    unsigned long long fct_acc[8] = {};
#endif
    FCT_BEGIN();
    for (int t; (t = tile_of(unit, tile_i, n_units, m_blocks)) >= 0; ++tile_i) {
      const TileInfo ti = decode_tile(t, kBM * kCtas, row_ofs);
      const int row = ti.m0 + row_l;
      const bool live = row < p.n_ctus;
      it += kKSteps;
      // ---- epi1: accumulator 1 -> a1 = leaky(acc * unscale1 + b1) -> fp16 hi/lo slices in the ring (FC2 A operand)
      mbar_wait(acc1_full, tile_i & 1);
      FCT_MARK(0);   // epilogue: waiting for accumulator 1 (includes epi2 of the previous tile)
      tc_fence_after();
      for (int j = 0; j < ti.nslices; ++j, ++it) {
        const int s = it % kStages;
        if (!kA2Tmem) mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1);   // the MMAs that last read this stage have retired
        FCT_MARK(1);   // epi1: waiting for the ring stage
        uint8_t* st = smem + s * kStageBytes;
        uint8_t* row_hi = st + row_l * kRowBytes;
#pragma unroll 1
        for (int half = (kEpiWarps == 8 ? cset : 0); half < kBK / 32; half += (kEpiWarps == 8 ? 2 : 1)) {
          uint32_t r[32];
          const int c0 = j * kBK + half * 32;
          tmem_ld_x32(tmem_base + lane_addr + c0, r);
          // a1 * 2^a1_exp = leaky(acc * (unscale1 2^a1_exp) + b1 2^a1_exp) on packed fp32 pairs (FFMA2 / FMUL2), then the exact
          // fp16 hi / lo split: 4.5 instructions per activation (the drain sits on the tensor pipe's critical path)
          const float us = p.unscale1 * p.a1_scale;
          float2 a[16];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const float4 b = *reinterpret_cast<const float4*>(b1s + ti.n0 + c0 + 2 * i);
            const float2 t0 = fma2(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), make_float2(us, us), make_float2(b.x, b.y));
            const float2 t1 = fma2(make_float2(__uint_as_float(r[2 * i + 2]), __uint_as_float(r[2 * i + 3])), make_float2(us, us), make_float2(b.z, b.w));
            const float2 m0 = mul2(t0, make_float2(0.2f, 0.2f)), m1 = mul2(t1, make_float2(0.2f, 0.2f));
            a[i] = make_float2(fmaxf(m0.x, t0.x), fmaxf(m0.y, t0.y));
            a[i + 1] = make_float2(fmaxf(m1.x, t1.x), fmaxf(m1.y, t1.y));
          }
          if (p.fc1_out != nullptr && live) {
            const float inv = 1.0f / p.a1_scale;   // a power of two: exact
            float4* o = reinterpret_cast<float4*>(p.fc1_out + size_t(row) * kFc1 + ti.n0 + c0);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = make_float4(a[2 * i].x * inv, a[2 * i].y * inv, a[2 * i + 1].x * inv, a[2 * i + 1].y * inv);
          }
          if (kA2Tmem) {
            // back into the 32 columns just read: [c0, c0 + 16) the hi pairs, [c0 + 16, c0 + 32) the lo pairs (TS-mode A operand)
            uint32_t hw[16], lw[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float2 v = a[e];
              const __half2 h = __floats2half2_rn(v.x, v.y);
              const float2 lf = sub2(v, make_float2(__low2float(h), __high2float(h)));
              const __half2 l = __floats2half2_rn(lf.x, lf.y);
              hw[e] = *reinterpret_cast<const uint32_t*>(&h);
              lw[e] = *reinterpret_cast<const uint32_t*>(&l);
            }
            tmem_st_x16(tmem_base + lane_addr + c0, hw);
            tmem_st_x16(tmem_base + lane_addr + c0 + 16, lw);
          } else {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {     // 16-byte chunks of 8 fp16; swizzle: chunk ^= row % 8 (128 B) | (row / 2) % 4 (64 B)
              uint32_t hw[4], lw[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 v = a[4 * cc + e];
                const __half2 h = __floats2half2_rn(v.x, v.y);
                const float2 lf = sub2(v, make_float2(__low2float(h), __high2float(h)));
                const __half2 l = __floats2half2_rn(lf.x, lf.y);
                hw[e] = *reinterpret_cast<const uint32_t*>(&h);
                lw[e] = *reinterpret_cast<const uint32_t*>(&l);
              }
              const int chunk = kBK == 64 ? ((half * 4 + cc) ^ (row_l & 7)) * 16 : (cc ^ ((row_l >> 1) & 3)) * 16;
              *reinterpret_cast<uint4*>(row_hi + chunk) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
              *reinterpret_cast<uint4*>(row_hi + kABytes + chunk) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
          }
        }
        FCT_MARK(2);   // epi1: tcgen05.ld + math + store
        if (kA2Tmem) {
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
        } else {
          asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy writes -> visible to the MMA (async proxy)
        }
        __syncwarp();
        if (lane == 0) arrive_at_leader(&a2_full[kA2Tmem ? j : s]);
        FCT_MARK(3);   // epi1: fence + arrive
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_at_leader(acc1_empty);

      // ---- epi2: accumulator 2 -> a2 = leaky(acc * unscale2 + b2eff) -> FC3 -> sigmoid -> probabilities
      if (cset != 0) continue;   // second column set: back to the next tile's accumulator 1
      mbar_wait(acc2_full, tile_i & 1);
      tc_fence_after();
      const uint32_t t2addr = tmem_base + kAcc2Col + lane_addr;
      const int gn = p.ctu_begin + row;
      float* prow = (p.prob != nullptr && live) ? p.prob + size_t(gn) * kProbs : nullptr;
      unsigned flag_bits = 0;
      if (ti.type) {
        float y[16];
#pragma unroll
        for (int o = 0; o < 16; ++o) y[o] = b3s[5 + o];
        const float* w3 = w3s + 48 + 96 * 4;
#pragma unroll 1
        for (int c0 = 0; c0 < 192; c0 += 32) {
          uint32_t r[32];
          tmem_ld_x32(t2addr + c0, r);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float a2 = leaky(fmaf(__uint_as_float(r[i]), p.unscale2, b2s[144 + c0 + i]));
            const float4* w = reinterpret_cast<const float4*>(w3 + (c0 + i) * 16);
#pragma unroll
            for (int o4 = 0; o4 < 4; ++o4) {
              const float4 v = w[o4];
              y[4 * o4] = fmaf(a2, v.x, y[4 * o4]);
              y[4 * o4 + 1] = fmaf(a2, v.y, y[4 * o4 + 1]);
              y[4 * o4 + 2] = fmaf(a2, v.z, y[4 * o4 + 2]);
              y[4 * o4 + 3] = fmaf(a2, v.w, y[4 * o4 + 3]);
            }
          }
        }
        if (prow) {
#pragma unroll
          for (int o = 0; o < 16; ++o) prow[5 + o] = sigmoidf(y[o]);
        }
      } else {
        uint32_t r[32];
        float y64 = b3s[0];
        tmem_ld_x32(t2addr, r);
#pragma unroll
        for (int i = 0; i < 32; ++i) y64 = fmaf(leaky(fmaf(__uint_as_float(r[i]), p.unscale2, b2s[i])), w3s[i], y64);
        tmem_ld_x16(t2addr + 32, r);
#pragma unroll
        for (int i = 0; i < 16; ++i) y64 = fmaf(leaky(fmaf(__uint_as_float(r[i]), p.unscale2, b2s[32 + i])), w3s[32 + i], y64);
        float y32[4] = {b3s[1], b3s[2], b3s[3], b3s[4]};
#pragma unroll 1
        for (int c0 = 0; c0 < 96; c0 += 32) {
          tmem_ld_x32(t2addr + 48 + c0, r);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float a2 = leaky(fmaf(__uint_as_float(r[i]), p.unscale2, b2s[48 + c0 + i]));
            const float4 v = *reinterpret_cast<const float4*>(w3s + 48 + (c0 + i) * 4);
            y32[0] = fmaf(a2, v.x, y32[0]), y32[1] = fmaf(a2, v.y, y32[1]);
            y32[2] = fmaf(a2, v.z, y32[2]), y32[3] = fmaf(a2, v.w, y32[3]);
          }
        }
        const float p64 = sigmoidf(y64);
        float p32[4];
#pragma unroll
        for (int o = 0; o < 4; ++o) p32[o] = sigmoidf(y32[o]);
        if (prow) {
          prow[0] = p64;
#pragma unroll
          for (int o = 0; o < 4; ++o) prow[1 + o] = p32[o];
        }
        if (p64 > p.t1) flag_bits |= 1u;
        if (p32[0] > p.t2 || p32[1] > p.t2 || p32[2] > p.t2 || p32[3] > p.t2) flag_bits |= 2u;
      }
      if (p.flags != nullptr && live && flag_bits) {
        const int f = gn / p.ctus_per_frame, rr = gn - f * p.ctus_per_frame;
        unsigned* fl = p.flags + f * p.chunks_per_frame + rr / kSubBatch;
        if ((*reinterpret_cast<volatile unsigned*>(fl) & flag_bits) != flag_bits) atomicOr(fl, flag_bits);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_at_leader(acc2_empty);
      FCT_MARK(4);   // epi2
    }
#ifdef ETHCNN_EXP_FC_TIMING
    if (lane == 0 && rank == 0 && warp == 2) {
      for (int k = 0; k < 5; ++k) atomicAdd(&g_fc_epi[k], fct_acc[k]);
      atomicAdd(&g_fc_epi[6], 1ull);
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (kCtas == 2) cluster_sync_all();   // neither CTA leaves (or frees TMEM) while its peer still reads its shared memory / arrives on its barriers
  if (warp == 1) {
    tc_fence_after();
    if (kCtas == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

__global__ void __launch_bounds__(kThreads, 1) fc_fused_kernel(FC_FUSED_PARAMS) { fc_fused_body<1>(FC_FUSED_ARGS); }
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) fc_fused_pair_kernel(FC_FUSED_PARAMS) {
  fc_fused_body<2>(FC_FUSED_ARGS);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    cudaDriverEntryPointQueryResult qres;
    void* f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  return fn;
}

// [rows][k_len] fp16 row-major (K contiguous) -> 2-D map, box kBK x box_rows, swizzle span = one box row.
bool make_kmajor_map(CUtensorMap* map, const __half* base, uint64_t k_len, uint64_t rows, uint32_t box_rows, const char** err) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    *err = "cuTensorMapEncodeTiled not available from the driver";
    return false;
  }
  cuuint64_t dims[2] = {k_len, rows};
  cuuint64_t strides[1] = {k_len * 2};
  cuuint32_t box[2] = {cuuint32_t(kBK), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, kBK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    *err = "cuTensorMapEncodeTiled failed";
    return false;
  }
  return true;
}

}  // namespace

#ifdef ETHCNN_EXP_FC_TIMING
}  // namespace ethcnn
extern "C" int ethcnn_debug_fc_issuer(unsigned long long* out8, int reset) {
  unsigned long long z[8] = {};
  if (cudaMemcpyFromSymbol(out8, ethcnn::g_fc_issuer, sizeof(z)) != cudaSuccess) return -1;
  if (reset && cudaMemcpyToSymbol(ethcnn::g_fc_issuer, z, sizeof(z)) != cudaSuccess) return -1;
  return 0;
}
extern "C" int ethcnn_debug_fc_epi(unsigned long long* out8, int reset) {
  unsigned long long z[8] = {};
  if (cudaMemcpyFromSymbol(out8, ethcnn::g_fc_epi, sizeof(z)) != cudaSuccess) return -1;
  if (reset && cudaMemcpyToSymbol(ethcnn::g_fc_epi, z, sizeof(z)) != cudaSuccess) return -1;
  return 0;
}
namespace ethcnn {
#endif

bool fc_fused_prepare_weights(const __half* w1_hi, const __half* w1_lo, const __half* const w2_hi[3],
                              const __half* const w2_lo[3], FusedWeights* out, const char** err) {
  const int n1[3] = {64, 128, 256}, n2[3] = {48, 96, 192};
  bool ok = true;
  for (int ctas = 1; ok && ctas <= 2; ++ctas) {   // box rows: what ONE CTA stages (all of the N rows, or half of them in a pair)
    FusedWeights::Maps& m = out->maps[ctas - 1];
    ok = make_kmajor_map(&m.w1_hi_t0, w1_hi, kFeat, kFc1, 192 / ctas, err) && make_kmajor_map(&m.w1_lo_t0, w1_lo, kFeat, kFc1, 192 / ctas, err) &&
         make_kmajor_map(&m.w1_hi_t1, w1_hi, kFeat, kFc1, 256 / ctas, err) && make_kmajor_map(&m.w1_lo_t1, w1_lo, kFeat, kFc1, 256 / ctas, err);
    for (int h = 0; ok && h < 3; ++h)
      ok = make_kmajor_map(&m.w2_hi[h], w2_hi[h], n1[h], n2[h], n2[h] / ctas, err) &&
           make_kmajor_map(&m.w2_lo[h], w2_lo[h], n1[h], n2[h], n2[h] / ctas, err);
  }
  out->valid = ok;
  return ok;
}

cudaError_t fc_fused_configure() {
  cudaError_t e = cudaFuncSetAttribute(fc_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<1>::kSmemBytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(fc_fused_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<2>::kSmemBytes);
}

cudaError_t launch_fc_fused(const __half* feat_hi, const __half* feat_lo, const FusedWeights& w, const FusedParams& p,
                            int ctas_per_tile, int sm_count, cudaStream_t stream) {
  if (p.n_ctus <= 0) return cudaSuccess;
  if (!w.valid || (ctas_per_tile != 1 && ctas_per_tile != 2)) return cudaErrorInvalidValue;
  const int m_tiles = (p.n_ctus + kBM - 1) / kBM;            // rows the feature buffers hold (multiple of 128)
  const int m_blocks = (m_tiles + ctas_per_tile - 1) / ctas_per_tile;
  CUtensorMap map_a_hi, map_a_lo;
  const char* err = nullptr;
  // a pair's second half may lie past the last row: TMA zero-fills it and the epilogue skips those rows
  if (!make_kmajor_map(&map_a_hi, feat_hi, kFeat, uint64_t(m_tiles) * kBM, kBM, &err) ||
      !make_kmajor_map(&map_a_lo, feat_lo, kFeat, uint64_t(m_tiles) * kBM, kBM, &err))
    return cudaErrorInvalidValue;
  // units (CTAs or CTA pairs) work in twos on one M block (tile_of), so their number is even
  const int n_tiles = 2 * m_blocks;
  int units = (sm_count / ctas_per_tile) & ~1;
  if (units < 2) units = 2;
  if (units > n_tiles) units = n_tiles;
  const FusedWeights::Maps& m = w.maps[ctas_per_tile - 1];
  if (ctas_per_tile == 1)
    fc_fused_kernel<<<units, kThreads, Geo<1>::kSmemBytes, stream>>>(map_a_hi, map_a_lo, m.w1_hi_t0, m.w1_lo_t0, m.w1_hi_t1, m.w1_lo_t1,
                                                                    m.w2_hi[0], m.w2_lo[0], m.w2_hi[1], m.w2_lo[1], m.w2_hi[2], m.w2_lo[2],
                                                                    p, m_blocks);
  else {
    // programmatic dependent launch: the CTAs start (prologue only, see pdl_wait in the kernel) while the conv kernel ahead of
    // us in the stream drains
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * units), cfg.blockDim = dim3(kThreads), cfg.dynamicSmemBytes = Geo<2>::kSmemBytes, cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr, cfg.numAttrs = getenv("ETHCNN_NO_PDL") ? 0 : 1;   // measurement switch
    return cudaLaunchKernelEx(&cfg, fc_fused_pair_kernel, map_a_hi, map_a_lo, m.w1_hi_t0, m.w1_lo_t0, m.w1_hi_t1, m.w1_lo_t1, m.w2_hi[0],
                              m.w2_lo[0], m.w2_hi[1], m.w2_lo[1], m.w2_hi[2], m.w2_lo[2], p, m_blocks);
  }
  return cudaGetLastError();
}

}  // namespace ethcnn
