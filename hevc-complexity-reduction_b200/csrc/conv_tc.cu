// Stage CONV on the 5th-generation tensor cores: luma tiles -> 2688 conv features per CTU (sm_100a, tcgen05 + TMEM).
//
// Same arithmetic as conv_stage.cu (net_CNN.py:105-150 / ETH-CNN_Training_LDP/net_CTU64.py:102-175): per branch
// (S: pool 1, M: pool 2, L: pool 4) conv 4x4/s4 1->16, conv 2x2/s2 16->24, conv 2x2/s2 24->32, leaky(0.2) after each,
// mean removal over 16x16 pooled windows folded into the conv1 bias, operands as exact power-of-two scaled fp16
// hi + lo pairs (conv1's input is an exact integer and needs no lo part), fp32 accumulation.
//
// Shape of the work.  After pooling every branch looks alike: a QUAD = 16x16 pooled samples = one mean-removal window = one
// conv3 position = 2x2 REGIONS (8x8 samples, one conv2 position each) = 4x4 PATCHES (4x4 samples, one conv1 position each).
// A TASK is 128 quads of one branch (8 / 32 / 128 CTUs for S / M / L) and every GEMM of a task has its 128 quads as the M
// dimension, one TMEM lane per quad, so that the output of a layer is, lane for lane, the A operand of the next:
//   conv1  for region r, patch p:  D1[r][p] (128 x 16)  = X[r][p] (128 x 16 taps, shared memory)  * W1          SS, 2 passes
//   conv2  for region r:           D2[r]    (128 x 24)  = sum_p A2[r][p] (128 x 16 ch, TMEM)      * W2[p]       TS, 3 passes
//   conv3:                         D3       (128 x 32)  = sum_r A3[r]    (128 x 24 ch, TMEM)      * W3[r]       TS, 3 passes
// The epilogue warps read an accumulator with tcgen05.ld, apply bias + leaky, split into fp16 hi/lo and write the next A
// operand back into tensor memory with tcgen05.st: activations never touch shared memory.  Measured on a B200
// (tools/microbench/umma_probe.cu): a K = 16 MMA with M = 128 costs max(40, N / 2) cycles with A in shared memory (the
// 4 KB A read) and max(23, N / 2) with A in tensor memory -- skinny N is only affordable with A in TMEM.
//
// TMEM map (512 columns): acc1[g] 64 x 2 | A2[g] (hi 32 | lo 32) x 2 | acc2[r] 32 x 4 | A3 hi 48 | A3 lo 48 | acc3 32,
// g = r % 2 = the epilogue group that owns the unit (task, r).
// Warp roles (448 threads): 0-3 converters (raw tile bytes -> centred pooled fp16 X tiles in the 128-byte-swizzled UMMA
// layout + the window sums), 4-11 two epilogue groups of four warps (TMEM lane quarter = warp % 4), 12 TMA producer,
// 13 MMA issuer (the whole warp runs converged, elect.sync picks the issuing lane).
// Tasks are sorted by branch; a CTA takes every gridDim.x-th task of the global list and reloads the 29 KB weight image
// when the branch changes (twice per launch).
#include <cstring>

#include "conv_tc.h"
#include "kernels.h"
#include "ptx_sm100.cuh"

namespace ethcnn {
namespace {

constexpr int kConvWarps = 4, kEpiWarps = 8;
constexpr int kWarpTma = kConvWarps + kEpiWarps, kWarpMma = kWarpTma + 1;
constexpr int kTcThreads = 32 * (kWarpMma + 1);                      // 448
constexpr int kBatchCtus = 4, kRawSlots = 4, kRawSlotBytes = kBatchCtus * kCtu * kCtu;
constexpr int kXTileBytes = 128 * 128;                               // [128 quads][64 fp16], one region r
constexpr int kXBytes = 4 * kXTileBytes;                             // one task

// shared memory (offsets from a 1024-byte aligned base)
constexpr int kSmW = 0;
constexpr int kSmX = kTcBranchBytes;                                 // 2 tasks
constexpr int kSmRaw = kSmX + 2 * kXBytes;
constexpr int kSmQsum = kSmRaw + kRawSlots * kRawSlotBytes;          // [2][128] u32
constexpr int kSmBars = kSmQsum + 2 * 128 * 4;
constexpr int kNumBars = 24;
constexpr int kSmSlot = kSmBars + kNumBars * 8;
constexpr int kTcSmemBytes = kSmSlot + 16 + 1024;
static_assert(kSmX % 1024 == 0 && kSmRaw % 1024 == 0, "UMMA tiles need 1024-byte alignment");
static_assert(kTcSmemBytes <= 227 * 1024, "shared memory budget");

// barriers
enum {
  kBarRawFull = 0,     // [4]  TMA bytes landed
  kBarRawEmpty = 4,    // [4]  converters done with the slot (4 arrivals)
  kBarXFull = 8,       // [2]  X tiles + window sums of a task written (4 arrivals)
  kBarXFree = 10,      // [2]  conv1 MMAs of the task retired (commit)
  kBarAcc1Full = 12,   // [2]  per group: conv1 of a unit retired (commit)
  kBarAcc1Empty = 14,  // [2]  per group: accumulator 1 drained (4 arrivals)
  kBarA2Full = 16,     // [2]  per group: A2 written (4 arrivals)
  kBarC2Done = 18,     // [4]  per region: conv2 retired = acc2[r] full and A2[r % 2] free (commit)
  kBarA3Full = 22,     // all four regions' A3 slices written (16 arrivals)
  kBarAcc3Full = 23,   // conv3 retired (commit)
};
static_assert(kRawSlots == 4, "barrier table is laid out for four raw slots");

// tensor memory columns
constexpr uint32_t kColAcc1 = 0, kColA2 = 128, kColAcc2 = 256, kColA3Hi = 384, kColA3Lo = 432, kColAcc3 = 480;

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
__device__ __forceinline__ void sts_u128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// K-major operand tile in 128-byte-swizzled shared memory: rows of 64 fp16, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr & 0x3ffffu) >> 4);
  d |= uint64_t(1) << 16;
  d |= uint64_t(1024 >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}
__device__ __forceinline__ constexpr uint32_t idesc_f16(int n) { return (1u << 4) | (uint32_t(n >> 3) << 17) | (uint32_t(128 >> 4) << 24); }

// Called by all lanes of the converged MMA warp with warp-uniform operands (see fc_fused.cu umma_f16).
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p, e;\n.reg .b32 rx;\nelect.sync rx|e, 0xffffffff;\nsetp.ne.b32 p, %4, 0;\n"
               "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p, e;\n.reg .b32 rx;\nelect.sync rx|e, 0xffffffff;\nsetp.ne.b32 p, %4, 0;\n"
               "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("{\n.reg .pred e;\n.reg .b32 rx;\nelect.sync rx|e, 0xffffffff;\n"
               "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(smem_u32(bar)) : "memory");
}

// tcgen05.ld / st of N consecutive 32-bit columns of the warp's 32 lanes (shape 32x32b)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// packed fp32 pair arithmetic (FFMA2 / FMUL2 / FADD2)
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 r;
  asm("{\n.reg .b64 ra, rb, rc, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\nmov.b64 rc, {%6, %7};\n"
      "fma.rn.f32x2 rd, ra, rb, rc;\nmov.b64 {%0, %1}, rd;\n}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 r;
  asm("{\n.reg .b64 ra, rb, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\nmul.rn.f32x2 rd, ra, rb;\nmov.b64 {%0, %1}, rd;\n}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  float2 r;
  asm("{\n.reg .b64 ra, rb, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\nsub.rn.f32x2 rd, ra, rb;\nmov.b64 {%0, %1}, rd;\n}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
// accumulator pair -> leaky(d * u + b) -> packed fp16 hi pair and lo pair (hi + lo == value to ~22 bits)
__device__ __forceinline__ void act_split(uint32_t d0, uint32_t d1, float2 u, float2 b, uint32_t& hi, uint32_t& lo) {
  const float2 t = fma2(make_float2(__uint_as_float(d0), __uint_as_float(d1)), u, b);
  const float2 m = mul2(t, make_float2(0.2f, 0.2f));
  const float2 r = make_float2(fmaxf(m.x, t.x), fmaxf(m.y, t.y));   // Maximum(alpha * x, x)
  const __half2 h = __floats2half2_rn(r.x, r.y);
  const float2 l = sub2(r, make_float2(__low2float(h), __high2float(h)));
  const __half2 lh = __floats2half2_rn(l.x, l.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&lh);
}

__device__ __forceinline__ uint32_t centred_half2(uint32_t biased_bits, float centre) {
  // two fp16 bit patterns 0x6400 + n = 1024 + n (n < 1024); subtract 1024 + centre: exact
  const __half2 c = __floats2half2_rn(centre, centre);
  const __half2 v = __hsub2(*reinterpret_cast<const __half2*>(&biased_bits), c);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// One 16-byte chunk of an X tile: 2 pooled rows x 4 pooled columns, as centred integer sums s - 128 P^2 in fp16 (exact),
// K order (row, column).  addr = the chunk's first raw pixel in the 64-byte-pitch tile.  sum = its raw pixel sum.
template <int P>
__device__ __forceinline__ uint4 convert_chunk(uint32_t addr, uint32_t& sum) {
  uint4 o;
  if (P == 1) {
    const uint32_t w0 = lds_u32(addr), w1 = lds_u32(addr + kCtu);
    sum = __dp4a(w0, 0x01010101u, __dp4a(w1, 0x01010101u, 0u));
    o.x = centred_half2(__byte_perm(w0, 0x64646464u, 0x4140u), 1024.f + 128.f);
    o.y = centred_half2(__byte_perm(w0, 0x64646464u, 0x4342u), 1024.f + 128.f);
    o.z = centred_half2(__byte_perm(w1, 0x64646464u, 0x4140u), 1024.f + 128.f);
    o.w = centred_half2(__byte_perm(w1, 0x64646464u, 0x4342u), 1024.f + 128.f);
  } else if (P == 2) {
    uint32_t out[4];
    sum = 0;
#pragma unroll
    for (int pr = 0; pr < 2; ++pr) {
      const uint2 a = lds_u64(addr + (2 * pr) * kCtu), b = lds_u64(addr + (2 * pr + 1) * kCtu);
      sum = __dp4a(a.x, 0x01010101u, __dp4a(a.y, 0x01010101u, __dp4a(b.x, 0x01010101u, __dp4a(b.y, 0x01010101u, sum))));
      // 0x6400 + s is the fp16 bit pattern of 1024 + s for s < 1024 (s <= 4 * 255)
      const uint32_t s0 = __dp4a(b.x, 0x00000101u, __dp4a(a.x, 0x00000101u, 0x6400u));
      const uint32_t s1 = __dp4a(b.x, 0x01010000u, __dp4a(a.x, 0x01010000u, 0x6400u));
      const uint32_t s2 = __dp4a(b.y, 0x00000101u, __dp4a(a.y, 0x00000101u, 0x6400u));
      const uint32_t s3 = __dp4a(b.y, 0x01010000u, __dp4a(a.y, 0x01010000u, 0x6400u));
      out[2 * pr] = centred_half2(__byte_perm(s0, s1, 0x5410u), 1024.f + 512.f);
      out[2 * pr + 1] = centred_half2(__byte_perm(s2, s3, 0x5410u), 1024.f + 512.f);
    }
    o = make_uint4(out[0], out[1], out[2], out[3]);
  } else {
    uint32_t out[4];
    sum = 0;
#pragma unroll
    for (int pr = 0; pr < 2; ++pr) {
      uint32_t s[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const uint4 q = lds_u128(addr + (4 * pr + a) * kCtu);
        s[0] = __dp4a(q.x, 0x01010101u, s[0]);
        s[1] = __dp4a(q.y, 0x01010101u, s[1]);
        s[2] = __dp4a(q.z, 0x01010101u, s[2]);
        s[3] = __dp4a(q.w, 0x01010101u, s[3]);
      }
      sum += s[0] + s[1] + s[2] + s[3];
      const __half2 h0 = __floats2half2_rn(__int2float_rn(int(s[0]) - 2048), __int2float_rn(int(s[1]) - 2048));
      const __half2 h1 = __floats2half2_rn(__int2float_rn(int(s[2]) - 2048), __int2float_rn(int(s[3]) - 2048));
      out[2 * pr] = *reinterpret_cast<const uint32_t*>(&h0);
      out[2 * pr + 1] = *reinterpret_cast<const uint32_t*>(&h1);
    }
    o = make_uint4(out[0], out[1], out[2], out[3]);
  }
  return o;
}

struct Geometry {   // of one branch
  int pool;         // 1, 2, 4
  int ctus;         // CTUs per task: 8, 32, 128
  int batches;      // raw batches (of 4 CTUs) per task
  int quads;        // quads per raw batch: 64, 16, 4
  int off2, off3;   // feature offsets of conv2 / conv3 of the branch
  int rg, qg;       // regions per row, quads per row inside a CTU
};
__device__ __forceinline__ Geometry geometry(int ph) {
  Geometry g;
  g.pool = 1 << ph;
  g.ctus = 8 << (2 * ph);
  g.batches = g.ctus / kBatchCtus;
  g.quads = 64 >> (2 * ph);
  g.off2 = ph == 0 ? kOffC2S : (ph == 1 ? kOffC2M : kOffC2L);
  g.off3 = ph == 0 ? kOffC3S : (ph == 1 ? kOffC3M : kOffC3L);
  g.rg = 8 >> ph;
  g.qg = 4 >> ph;
  return g;
}
// quad q of a task -> CTU inside the task and quad position inside that CTU
__device__ __forceinline__ void quad_pos(int ph, int q, int& ctu, int& qy, int& qx) {
  if (ph == 0) ctu = q >> 4, qy = (q >> 2) & 3, qx = q & 3;
  else if (ph == 1) ctu = q >> 2, qy = (q >> 1) & 1, qx = q & 1;
  else ctu = q, qy = 0, qx = 0;
}

// One raw batch (4 CTUs) of a task: a warp per quad, a lane per 16-byte chunk of the quad's X rows (region r = lane / 8,
// patch = (lane / 2) % 4, half = lane % 2).  Each warp takes U consecutive quads per iteration so that their shared-memory
// loads are in flight together; the window sum is one REDUX.
template <int PH>
__device__ __forceinline__ void convert_batch(uint32_t rawbase, uint32_t xlane, uint32_t lane_raw, int q0, int warp, int lane, uint32_t* qs) {
  constexpr int P = 1 << PH, NQ = 64 >> (2 * PH), U = NQ >= 16 ? 4 : 1;
  const int pch = (lane >> 1) & 3, h = lane & 1;
  for (int qi = warp * U; qi < NQ; qi += kConvWarps * U) {
    uint4 v[U];
    uint32_t sm[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int q = q0 + qi + u;
      int ctu, qy, qx;
      quad_pos(PH, q, ctu, qy, qx);
      v[u] = convert_chunk<P>(rawbase + (ctu & (kBatchCtus - 1)) * kCtu * kCtu + (16 * P * qy) * kCtu + 16 * P * qx + lane_raw, sm[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int q = q0 + qi + u;
      sts_u128(xlane + (q >> 3) * 1024 + (q & 7) * 128 + (((2 * pch + h) ^ (q & 7)) << 4), v[u]);
      const uint32_t tot = __reduce_add_sync(0xffffffffu, sm[u]);
      if (lane == 0) qs[q] = tot;
    }
  }
}

__global__ void __launch_bounds__(kTcThreads, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap tmap, const ConvTcLaunch p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kSmSlot);
  uint32_t* qsum = reinterpret_cast<uint32_t*>(smem + kSmQsum);
  const float* tab = reinterpret_cast<const float*>(smem + kSmW + kTcTab);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grid = gridDim.x;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kRawSlots; ++i) mbar_init(&bars[kBarRawFull + i], 1), mbar_init(&bars[kBarRawEmpty + i], kConvWarps);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[kBarXFull + i], kConvWarps), mbar_init(&bars[kBarXFree + i], 1);
      mbar_init(&bars[kBarAcc1Full + i], 1), mbar_init(&bars[kBarAcc1Empty + i], 4), mbar_init(&bars[kBarA2Full + i], 4);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&bars[kBarC2Done + i], 1);
    mbar_init(&bars[kBarA3Full], 2 * kEpiWarps), mbar_init(&bars[kBarAcc3Full], 1);
    mbar_fence_init();
  }
  if (warp == kWarpMma) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp == kWarpTma && lane == 0) prefetch_tmap(&tmap);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  int k0 = 0;    // tasks this CTA has finished (the same count in every warp): barrier phases are derived from it
  int c0 = 0;    // raw batches this CTA has consumed
  int base = 0;  // global index of the phase's first task
  for (int ph = 0; ph < 3; ++ph) {
    const Geometry geo = geometry(ph);
    const int n_tasks = (p.n_ctus + geo.ctus - 1) / geo.ctus;
    const int i0 = ((int(blockIdx.x) - base) % grid + grid) % grid;
    const int nt = (i0 < n_tasks && ((p.phase_mask >> ph) & 1)) ? (n_tasks - 1 - i0) / grid + 1 : 0;
    base += n_tasks;
    // the branch's weight image (everything that used the previous one has retired: see the barrier at the end)
    {
      const uint4* src = reinterpret_cast<const uint4*>(p.blob + size_t(ph) * kTcBranchBytes);
      uint4* dst = reinterpret_cast<uint4*>(smem + kSmW);
      for (int i = threadIdx.x; i < kTcBranchBytes / 16; i += kTcThreads) dst[i] = src[i];
      asm volatile("fence.proxy.async;" ::: "memory");   // read by the MMAs through the async proxy
    }
    __syncthreads();

    if (warp == kWarpTma) {
      // ------------------------------------------------ TMA producer ------------------------------------------------
      for (int j = 0; j < nt; ++j) {
        const int first = (i0 + j * grid) * geo.ctus;
        for (int bi = 0; bi < geo.batches; ++bi) {
          const int c = c0 + j * geo.batches + bi, slot = c % kRawSlots;
          if (c >= kRawSlots) mbar_wait_relaxed(&bars[kBarRawEmpty + slot], ((c / kRawSlots) - 1) & 1);
          const int bfirst = first + bi * kBatchCtus;
          const int nv = max(0, min(kBatchCtus, p.n_ctus - bfirst));
          if (lane == 0) {
            if (nv > 0) mbar_arrive_expect_tx(&bars[kBarRawFull + slot], nv * kCtu * kCtu);
            else mbar_arrive(&bars[kBarRawFull + slot]);
          }
          __syncwarp();
          if (lane < nv) {
            const int n = p.ctu_begin + bfirst + lane;
            const int f = n / p.ctus_per_frame, r = n - f * p.ctus_per_frame;
            const int cy = r / p.ctus_per_row, cx = r - cy * p.ctus_per_row;
            tma_load_3d(smem + kSmRaw + slot * kRawSlotBytes + lane * kCtu * kCtu, &tmap, &bars[kBarRawFull + slot], cx * kCtu, cy * kCtu, f);
          }
          __syncwarp();
        }
      }
    } else if (warp < kConvWarps) {
      // ------------------------------------------------ converters ------------------------------------------------
      // a warp per quad, a lane per 16-byte chunk: region r = lane / 8, patch p = (lane / 2) % 4, half h = lane % 2
      const int r = lane >> 3, pch = (lane >> 1) & 3, h = lane & 1;
      const int y0 = 8 * (r >> 1) + 4 * (pch >> 1) + 2 * h, x0 = 8 * (r & 1) + 4 * (pch & 1);   // pooled coordinates inside the quad
      const uint32_t lane_raw = uint32_t(geo.pool * y0 * kCtu + geo.pool * x0);
      for (int j = 0; j < nt; ++j) {
        const int k = k0 + j, xb = k & 1;
        if (k >= 2) mbar_wait(&bars[kBarXFree + xb], ((k >> 1) - 1) & 1);
        const uint32_t xbase = smem_u32(smem + kSmX + xb * kXBytes) + r * kXTileBytes;
        for (int bi = 0; bi < geo.batches; ++bi) {
          const int c = c0 + j * geo.batches + bi, slot = c % kRawSlots;
          mbar_wait(&bars[kBarRawFull + slot], (c / kRawSlots) & 1);
          const uint32_t rawbase = smem_u32(smem + kSmRaw + slot * kRawSlotBytes);
          if (ph == 0) convert_batch<0>(rawbase, xbase, lane_raw, bi * geo.quads, warp, lane, qsum + xb * 128);
          else if (ph == 1) convert_batch<1>(rawbase, xbase, lane_raw, bi * geo.quads, warp, lane, qsum + xb * 128);
          else convert_batch<2>(rawbase, xbase, lane_raw, bi * geo.quads, warp, lane, qsum + xb * 128);
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[kBarRawEmpty + slot]);
        }
        asm volatile("fence.proxy.async;" ::: "memory");   // the X tiles are read by the MMAs (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[kBarXFull + xb]);
      }
    } else if (warp == kWarpMma) {
      // ------------------------------------------------ MMA issuer (whole warp, converged) ------------------------------------------------
      const uint32_t wbase = smem_u32(smem + kSmW);
      const uint64_t w1h = umma_desc(wbase + kTcW1Hi), w1l = umma_desc(wbase + kTcW1Lo);
      const uint64_t w2h = umma_desc(wbase + kTcW2Hi), w2l = umma_desc(wbase + kTcW2Lo);
      const uint64_t w3h = umma_desc(wbase + kTcW3Hi), w3l = umma_desc(wbase + kTcW3Lo);
      // conv1 of unit (k, r): X tile r of the task (shared memory) x W1, patches p = K slices of 32 bytes
      auto conv1 = [&](int k, int r) {
        const int g = r & 1, n = 2 * k + (r >> 1), xb = k & 1;
        if (r == 0) mbar_wait(&bars[kBarXFull + xb], (k >> 1) & 1);
        if (n > 0) mbar_wait(&bars[kBarAcc1Empty + g], (n - 1) & 1);
        tc_fence_after();
        __syncwarp();
        const uint64_t xa = umma_desc(smem_u32(smem + kSmX + xb * kXBytes + r * kXTileBytes));
#pragma unroll
        for (int pch = 0; pch < 4; ++pch) {
          const uint32_t d = tmem + kColAcc1 + 64 * g + 16 * pch;
          mma_ss(d, xa + uint64_t(2 * pch), w1h, idesc_f16(16), 0);
          mma_ss(d, xa + uint64_t(2 * pch), w1l, idesc_f16(16), 1);
        }
        mma_commit(&bars[kBarAcc1Full + g]);
        if (r == 3) mma_commit(&bars[kBarXFree + xb]);
        __syncwarp();
      };
      auto conv2 = [&](int k, int r) {
        const int g = r & 1, n = 2 * k + (r >> 1);
        mbar_wait(&bars[kBarA2Full + g], n & 1);
        tc_fence_after();
        __syncwarp();
        const uint32_t d = tmem + kColAcc2 + 32 * r, ah = tmem + kColA2 + 64 * g, al = ah + 32;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          mma_ts(d, ah + 8 * ks, w2h + uint64_t(2 * ks), idesc_f16(32), ks != 0);
          mma_ts(d, ah + 8 * ks, w2l + uint64_t(2 * ks), idesc_f16(32), 1);
          mma_ts(d, al + 8 * ks, w2h + uint64_t(2 * ks), idesc_f16(32), 1);
        }
        mma_commit(&bars[kBarC2Done + r]);
        __syncwarp();
      };
      auto conv3 = [&](int k) {
        mbar_wait(&bars[kBarA3Full], k & 1);
        tc_fence_after();
        __syncwarp();
        const uint32_t d = tmem + kColAcc3;
#pragma unroll
        for (int s = 0; s < 6; ++s) {
          const uint64_t adv = uint64_t((s >> 2) * (4096 >> 4) + 2 * (s & 3));   // tile s / 4, K slice s % 4
          mma_ts(d, tmem + kColA3Hi + 8 * s, w3h + adv, idesc_f16(32), s != 0);
          mma_ts(d, tmem + kColA3Hi + 8 * s, w3l + adv, idesc_f16(32), 1);
          mma_ts(d, tmem + kColA3Lo + 8 * s, w3h + adv, idesc_f16(32), 1);
        }
        mma_commit(&bars[kBarAcc3Full]);
        __syncwarp();
      };
      if (nt > 0) conv1(k0, 0), conv1(k0, 1);
      for (int j = 0; j < nt; ++j) {
        const int k = k0 + j;
        conv2(k, 0), conv1(k, 2), conv2(k, 1), conv1(k, 3), conv2(k, 2), conv2(k, 3);
        if (j + 1 < nt) conv1(k + 1, 0), conv1(k + 1, 1);
        conv3(k);
      }
    } else {
      // ------------------------------------------------ epilogue groups ------------------------------------------------
      const int e = warp - kConvWarps, g = e >> 2, quarter = warp & 3;
      const int q = quarter * 32 + lane;                   // this lane's quad = TMEM lane
      const uint32_t lane_addr = tmem + (uint32_t(quarter * 32) << 16);
      int ctu_l, qy, qx;
      quad_pos(ph, q, ctu_l, qy, qx);
      const float2 u1x8 = make_float2(p.u1x8[ph], p.u1x8[ph]), u2 = make_float2(p.u2[ph], p.u2[ph]), u3 = make_float2(p.u3[ph], p.u3[ph]);
      const float u1 = p.u1[ph], mu_c = float(1024 * geo.pool * geo.pool);
      float2 be[8];      // conv1 bias + mean term of this lane's quad, per channel pair; valid for the current task of the group
      __half* row_hi = nullptr;
      __half* row_lo = nullptr;

      // unit (k, r): accumulator 1 -> a1 = leaky(acc * u + bias + mean term) -> fp16 hi/lo -> A2[g] in tensor memory
      auto ep1 = [&](int j, int r) {
        const int k = k0 + j, n = 2 * k + (r >> 1);
        if (r < 2) {   // first unit of the task for this group: the task's rows and mean terms
          mbar_wait(&bars[kBarXFull + (k & 1)], (k >> 1) & 1);   // window sums written
          const float mu = fmaf(__uint2float_rn(qsum[(k & 1) * 128 + q]), -0.03125f, mu_c) * u1;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            be[i] = make_float2(fmaf(mu, tab[16 + 2 * i], tab[2 * i]), fmaf(mu, tab[16 + 2 * i + 1], tab[2 * i + 1]));
          const int row = (i0 + j * grid) * geo.ctus + ctu_l;
          const bool live = row < p.n_ctus;
          row_hi = live ? p.feat_hi + size_t(row) * kFeat : nullptr;
          row_lo = live ? p.feat_lo + size_t(row) * kFeat : nullptr;
        }
        mbar_wait(&bars[kBarAcc1Full + g], n & 1);
        // A2[g] is free once conv2 of the group's previous unit has retired
        if (r >= 2) mbar_wait(&bars[kBarC2Done + r - 2], k & 1);
        else if (k > 0) mbar_wait(&bars[kBarC2Done + r + 2], (k - 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          uint32_t a[32], hw[16], lw[16];
          tmem_ld32(lane_addr + kColAcc1 + 64 * g + 32 * half, a);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) act_split(a[2 * i], a[2 * i + 1], u1x8, be[i & 7], hw[i], lw[i]);
          tmem_st16(lane_addr + kColA2 + 64 * g + 16 * half, hw);
          tmem_st16(lane_addr + kColA2 + 64 * g + 32 + 16 * half, lw);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[kBarAcc1Empty + g]), mbar_arrive(&bars[kBarA2Full + g]);
      };
      // unit (k, r): accumulator 2 -> conv2 features (global memory) and the A3 slice of region r
      auto ep2 = [&](int j, int r) {
        const int k = k0 + j;
        mbar_wait(&bars[kBarC2Done + r], k & 1);
        tc_fence_after();
        uint32_t a[24], hw[12], lw[12];
        tmem_ld16(lane_addr + kColAcc2 + 32 * r, a);
        tmem_ld8(lane_addr + kColAcc2 + 32 * r + 16, a + 16);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 12; ++i)
          act_split(a[2 * i], a[2 * i + 1], u2, make_float2(tab[32 + 2 * i], tab[32 + 2 * i + 1]), hw[i], lw[i]);
        tmem_st8(lane_addr + kColA3Hi + 12 * r, hw), tmem_st4(lane_addr + kColA3Hi + 12 * r + 8, hw + 8);
        tmem_st8(lane_addr + kColA3Lo + 12 * r, lw), tmem_st4(lane_addr + kColA3Lo + 12 * r + 8, lw + 8);
        if (row_hi != nullptr) {
          const int o = geo.off2 + ((2 * qy + (r >> 1)) * geo.rg + 2 * qx + (r & 1)) * 24;
          uint4* dh = reinterpret_cast<uint4*>(row_hi + o);
          uint4* dl = reinterpret_cast<uint4*>(row_lo + o);
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            dh[i] = make_uint4(hw[4 * i], hw[4 * i + 1], hw[4 * i + 2], hw[4 * i + 3]);
            dl[i] = make_uint4(lw[4 * i], lw[4 * i + 1], lw[4 * i + 2], lw[4 * i + 3]);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[kBarA3Full]);
      };
      // task k: accumulator 3, columns 16 g .. 16 g + 15 -> conv3 features.
      // It runs after ep1 of the next task has started (overlapping conv3), when row_hi / row_lo already point at task
      // k + 1: the rows of task k are kept in prev_hi / prev_lo.
      __half* prev_hi = nullptr;
      __half* prev_lo = nullptr;
      auto ep3_store = [&](int j) {
        const int k = k0 + j;
        mbar_wait(&bars[kBarAcc3Full], k & 1);
        tc_fence_after();
        uint32_t a[16], hw[8], lw[8];
        tmem_ld16(lane_addr + kColAcc3 + 16 * g, a);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i)
          act_split(a[2 * i], a[2 * i + 1], u3, make_float2(tab[56 + 16 * g + 2 * i], tab[56 + 16 * g + 2 * i + 1]), hw[i], lw[i]);
        if (prev_hi != nullptr) {
          const int o = geo.off3 + (qy * geo.qg + qx) * 32 + 16 * g;
          uint4* dh = reinterpret_cast<uint4*>(prev_hi + o);
          uint4* dl = reinterpret_cast<uint4*>(prev_lo + o);
          dh[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]), dh[1] = make_uint4(hw[4], hw[5], hw[6], hw[7]);
          dl[0] = make_uint4(lw[0], lw[1], lw[2], lw[3]), dl[1] = make_uint4(lw[4], lw[5], lw[6], lw[7]);
        }
        tc_fence_before();
      };
      for (int j = 0; j < nt; ++j) {
        if (j == 0) ep1(j, g);
        ep1(j, g + 2);
        ep2(j, g), ep2(j, g + 2);
        prev_hi = row_hi, prev_lo = row_lo;
        if (j + 1 < nt) ep1(j + 1, g);   // overlaps conv3 of task j
        ep3_store(j);
      }
    }
    k0 += nt;
    c0 += nt * geo.batches;
    tc_fence_before();
    __syncthreads();   // every role is done with the branch: its last MMAs have retired (the epilogue waited for them)
    tc_fence_after();
  }

  if (warp == kWarpMma) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

}  // namespace

cudaError_t conv_tc_configure() {
  return cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes);
}

cudaError_t launch_conv_tc(const CUtensorMap& tmap, const ConvTcLaunch& p, int sm_count, cudaStream_t stream) {
  if (p.n_ctus <= 0) return cudaSuccess;
  const int n_tasks = (p.n_ctus + 7) / 8 + (p.n_ctus + 31) / 32 + (p.n_ctus + 127) / 128;
  const int grid = n_tasks < sm_count ? n_tasks : sm_count;
  conv_tc_kernel<<<grid, kTcThreads, kTcSmemBytes, stream>>>(tmap, p);
  return cudaGetLastError();
}

}  // namespace ethcnn
