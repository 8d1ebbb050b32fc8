#include "model.h"

#include <cmath>
#include <cstring>

#include "conv_tc.h"
#include "kernels.h"

namespace ethcnn {

uint16_t f32_to_f16_bits(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  const uint32_t em = x & 0x7fffffffu;
  if (em >= 0x7f800000u) return uint16_t(sign | 0x7c00u | ((em > 0x7f800000u) ? 0x200u : 0));  // inf / nan
  if (em >= 0x477ff000u) return uint16_t(sign | 0x7c00u);                                      // rounds to >= 65520 -> inf
  if (em < 0x33000001u) return uint16_t(sign);                                                 // < 2^-25 -> 0
  int exp = int(em >> 23) - 127;
  uint32_t mant = (em & 0x7fffffu) | 0x800000u;
  int shift;
  uint32_t base;
  if (exp >= -14) {  // normal half
    shift = 13;
    base = uint32_t(exp + 15) << 10;
    mant &= 0x7fffffu;
  } else {  // subnormal half: value = mant * 2^(exp-23), unit 2^-24
    shift = 13 + (-14 - exp);
    base = 0;
  }
  uint32_t q = mant >> shift;
  const uint32_t rem = mant & ((1u << shift) - 1), halfway = 1u << (shift - 1);
  if (rem > halfway || (rem == halfway && (q & 1))) ++q;
  return uint16_t(sign | (base + q));  // a carry out of the mantissa correctly bumps the exponent
}

float f16_bits_to_f32(uint16_t h) {
  const uint32_t sign = uint32_t(h & 0x8000u) << 16;
  const uint32_t exp = (h >> 10) & 0x1f, mant = h & 0x3ffu;
  float v;
  if (exp == 0) {
    v = std::ldexp(float(mant), -24);
  } else if (exp == 31) {
    v = mant ? NAN : INFINITY;
  } else {
    v = std::ldexp(float(mant | 0x400u), int(exp) - 25);
  }
  uint32_t bits;
  memcpy(&bits, &v, 4);
  bits |= sign;
  memcpy(&v, &bits, 4);
  return v;
}

namespace {

std::string var_name(int i) { return i == 0 ? "Variable" : "Variable_" + std::to_string(i); }

const BundleTensor* find(const std::map<std::string, BundleTensor>& t, const std::string& name,
                         std::initializer_list<int64_t> shape, std::string* err) {
  auto it = t.find(name);
  if (it == t.end()) {
    *err = "checkpoint lacks tensor " + name;
    return nullptr;
  }
  if (it->second.shape != std::vector<int64_t>(shape)) {
    *err = "tensor " + name + " has an unexpected shape";
    return nullptr;
  }
  return &it->second;
}

}  // namespace

bool pack_model(const std::map<std::string, BundleTensor>& t, float input_bound, PackedModel* out, std::string* err) {
  // power-of-two scales so that fp16 hi parts stay below 2^15 and lo parts stay normal for all but tiny values
  auto pick_exp = [](float bound) {
    int e = int(std::floor(std::log2(32768.0 / double(bound))));
    if (e > 14) e = 14;
    if (e < -14) e = -14;
    return e;
  };
  out->conv.assign(kConvFloats, 0.f);
  out->conv_tc.assign(kTcBlobBytes, 0);
  const int branch_base[3] = {12, 6, 0};  // S, M, L -> first Variable index (net_CNN.py:126-141)
  float feat_bound = 0.f;
  const BundleTensor* bt[3][6] = {};
  for (int br = 0; br < 3; ++br) {
    const int v = branch_base[br];
    const BundleTensor* w1 = find(t, var_name(v), {4, 4, 1, 16}, err);
    const BundleTensor* b1 = find(t, var_name(v + 1), {16}, err);
    const BundleTensor* w2 = find(t, var_name(v + 2), {2, 2, 16, 24}, err);
    const BundleTensor* b2 = find(t, var_name(v + 3), {24}, err);
    const BundleTensor* w3 = find(t, var_name(v + 4), {2, 2, 24, 32}, err);
    const BundleTensor* b3 = find(t, var_name(v + 5), {32}, err);
    if (!w1 || !b1 || !w2 || !b2 || !w3 || !b3) return false;
    bt[br][0] = w1, bt[br][1] = b1, bt[br][2] = w2, bt[br][3] = b2, bt[br][4] = w3, bt[br][5] = b3;
    float* dst = out->conv.data() + br * kConvBranchFloats;
    memcpy(dst + kB1Off, b1->data.data(), 16 * 4);
    memcpy(dst + kB2Off, b2->data.data(), 24 * 4);
    memcpy(dst + kB3Off, b3->data.data(), 32 * 4);
    // rigorous magnitude bounds (leaky never increases |.|)
    double B1[16], B2[24], B3[32];
    for (int co = 0; co < 16; ++co) {
      double s = 0;
      for (int k = 0; k < 16; ++k) s += std::fabs(w1->data[k * 16 + co]);
      B1[co] = s * input_bound + std::fabs(b1->data[co]);
    }
    for (int co = 0; co < 24; ++co) {
      double s = 0;
      for (int k = 0; k < 64; ++k) s += std::fabs(w2->data[k * 24 + co]) * B1[k % 16];
      B2[co] = s + std::fabs(b2->data[co]);
      if (B2[co] > feat_bound) feat_bound = float(B2[co]);
    }
    for (int co = 0; co < 32; ++co) {
      double s = 0;
      for (int k = 0; k < 96; ++k) s += std::fabs(w3->data[k * 32 + co]) * B2[k % 24];
      B3[co] = s + std::fabs(b3->data[co]);
      if (B3[co] > feat_bound) feat_bound = float(B3[co]);
    }
    // filters as mma.sync B fragments, fp16 hi/lo of w * 2^e (K in TF order: w1 [16 taps][16], w2 [64][24], w3 [96][32])
    auto maxabs = [](const std::vector<float>& v) {
      float m = 0.f;
      for (float x : v) m = std::fmax(m, std::fabs(x));
      return m > 0.f ? m : 1.f;
    };
    double b1max = 0;
    for (int co = 0; co < 16; ++co) b1max = std::fmax(b1max, B1[co]);
    const int e1w = pick_exp(maxabs(w1->data)), e2w = pick_exp(maxabs(w2->data)), e3w = pick_exp(maxabs(w3->data));
    const int ec1 = pick_exp(float(b1max));
    out->conv_exp[br][0] = e1w, out->conv_exp[br][1] = ec1, out->conv_exp[br][2] = e2w, out->conv_exp[br][3] = e3w;
    auto pack_frags = [&](const std::vector<float>& W, int n_out, int ksteps, int ntiles, int e, int off_hi, int off_lo) {
      uint16_t* hi = reinterpret_cast<uint16_t*>(dst + off_hi);
      uint16_t* lo = reinterpret_cast<uint16_t*>(dst + off_lo);
      const float sc = std::ldexp(1.0f, e);
      for (int j = 0; j < ksteps; ++j)
        for (int nt = 0; nt < ntiles; ++nt)
          for (int lane = 0; lane < 32; ++lane) {
            const int g = lane >> 2, d = lane & 3;
            for (int reg = 0; reg < 2; ++reg)
              for (int el = 0; el < 2; ++el) {
                const float v = W[size_t(16 * j + 2 * d + 8 * reg + el) * n_out + 8 * nt + g] * sc;
                const uint16_t h = f32_to_f16_bits(v);
                const size_t idx = ((size_t(j) * ntiles + nt) * 32 + lane) * 4 + reg * 2 + el;
                hi[idx] = h;
                lo[idx] = f32_to_f16_bits(v - f16_bits_to_f32(h));
              }
          }
    };
    pack_frags(w1->data, 16, 1, 2, e1w, kF1HiOff, kF1LoOff);
    // per output channel: sum over the taps of the conv1 filter AS STORED (hi + lo); the kernel feeds conv1 with
    // centred integer samples and adds (centre - window mean) * this sum to the bias instead of subtracting the mean
    for (int co = 0; co < 16; ++co) {
      double s = 0;
      const float sc = std::ldexp(1.0f, e1w);
      for (int k = 0; k < 16; ++k) {
        const float v = w1->data[size_t(k) * 16 + co] * sc;
        const float h = f16_bits_to_f32(f32_to_f16_bits(v));
        s += double(h) + double(f16_bits_to_f32(f32_to_f16_bits(v - h)));
      }
      dst[kW1SumOff + co] = float(s);
    }
    pack_frags(w2->data, 24, 4, 3, e2w, kF2HiOff, kF2LoOff);
    pack_frags(w3->data, 32, 6, 4, e3w, kF3HiOff, kF3LoOff);
  }
  out->feat_bound = feat_bound;

  static const char* hname[3] = {"64", "32", "16"};
  const int n1[3] = {64, 128, 256}, n2[3] = {48, 96, 192}, n3[3] = {1, 4, 16};
  const int col_off[3] = {0, 64, 192};
  out->w1.assign(size_t(kFeat) * kFc1, 0.f);
  out->b1.assign(kFc1, 0.f);
  float wmax = 0.f;
  for (int h = 0; h < 3; ++h) {
    const std::string hs = hname[h];
    const BundleTensor* w1 = find(t, "h_fc1__" + hs + "__w", {kFeat, n1[h]}, err);
    const BundleTensor* b1 = find(t, "h_fc1__" + hs + "__b", {n1[h]}, err);
    const BundleTensor* w2 = find(t, "h_fc2__" + hs + "__w", {n1[h] + 1, n2[h]}, err);
    const BundleTensor* b2 = find(t, "h_fc2__" + hs + "__b", {n2[h]}, err);
    const BundleTensor* w3 = find(t, "y_conv_flat__" + hs + "__w", {n2[h] + 1, n3[h]}, err);
    const BundleTensor* b3 = find(t, "y_conv_flat__" + hs + "__b", {n3[h]}, err);
    if (!w1 || !b1 || !w2 || !b2 || !w3 || !b3) return false;
    for (int k = 0; k < kFeat; ++k)
      for (int j = 0; j < n1[h]; ++j) {
        const float w = w1->data[size_t(k) * n1[h] + j];
        out->w1[size_t(k) * kFc1 + col_off[h] + j] = w;
        if (std::fabs(w) > wmax) wmax = std::fabs(w);
      }
    memcpy(out->b1.data() + col_off[h], b1->data.data(), n1[h] * 4);
    // the qp input is the LAST row of the FC2 / FC3 matrices (tf.concat([h, qp], axis=1), net_CNN.py:158,161)
    out->w2[h].assign(w2->data.begin(), w2->data.begin() + size_t(n1[h]) * n2[h]);
    out->w2q[h].assign(w2->data.begin() + size_t(n1[h]) * n2[h], w2->data.end());
    out->b2[h] = b2->data;
    out->w3[h].assign(w3->data.begin(), w3->data.begin() + size_t(n2[h]) * n3[h]);
    out->w3q[h].assign(w3->data.begin() + size_t(n2[h]) * n3[h], w3->data.end());
    out->b3[h] = b3->data;
  }
  for (float v : out->w1)
    if (!std::isfinite(v)) {
      *err = "non-finite FC1 weight";
      return false;
    }
  if (!std::isfinite(feat_bound) || feat_bound <= 0.f) {
    *err = "non-finite conv weights";
    return false;
  }

  out->feat_exp = pick_exp(feat_bound);
  for (int br = 0; br < 3; ++br) {
    float* hdr = out->conv.data() + br * kConvBranchFloats + kHdrOff;
    const int* e = out->conv_exp[br];
    hdr[0] = std::ldexp(32.0f, -e[0]);                    // conv1: A = (256 s - W) / 32, B = w1 * 2^e1w (times cst at run time)
    hdr[1] = std::ldexp(1.0f, -(e[1] + e[2]));            // conv2: A = c1 * 2^e_c1, B = w2 * 2^e2w
    hdr[2] = std::ldexp(1.0f, -(out->feat_exp + e[3]));   // conv3: A = c2 * 2^feat_exp, B = w3 * 2^e3w
    hdr[3] = std::ldexp(1.0f, e[1]);
  }
  // the same filters (same exponents, same hi/lo values) as 128-byte-swizzled K-major UMMA tiles for conv_tc.cu
  for (int br = 0; br < 3; ++br) {
    uint8_t* blob = out->conv_tc.data() + size_t(br) * kTcBranchBytes;
    const int* e = out->conv_exp[br];
    auto put = [&](int off_hi, int off_lo, int n, int k, float v) {   // element (row n, K index k) of a [..][64] tile
      const size_t o = size_t(n >> 3) * 1024 + size_t(n & 7) * 128 + size_t(((k >> 3) ^ (n & 7)) << 4) + size_t(k & 7) * 2;
      const uint16_t h = f32_to_f16_bits(v);
      const uint16_t l = f32_to_f16_bits(v - f16_bits_to_f32(h));
      memcpy(blob + off_hi + o, &h, 2);
      memcpy(blob + off_lo + o, &l, 2);
    };
    const float s1 = std::ldexp(1.0f, e[0]), s2 = std::ldexp(1.0f, e[2]), s3 = std::ldexp(1.0f, e[3]);
    for (int k = 0; k < 16; ++k)
      for (int n = 0; n < 16; ++n) put(kTcW1Hi, kTcW1Lo, n, k, bt[br][0]->data[size_t(k) * 16 + n] * s1);
    for (int k = 0; k < 64; ++k)
      for (int n = 0; n < 24; ++n) put(kTcW2Hi, kTcW2Lo, n, k, bt[br][2]->data[size_t(k) * 24 + n] * s2);
    for (int k = 0; k < 96; ++k)
      for (int n = 0; n < 32; ++n) put(kTcW3Hi + (k >> 6) * 4096, kTcW3Lo + (k >> 6) * 4096, n, k & 63, bt[br][4]->data[size_t(k) * 32 + n] * s3);
    float* tab = reinterpret_cast<float*>(blob + kTcTab);
    const float sc1 = std::ldexp(1.0f, e[1]), fs = std::ldexp(1.0f, out->feat_exp);
    const float* v4 = out->conv.data() + br * kConvBranchFloats;
    for (int c = 0; c < 16; ++c) tab[c] = bt[br][1]->data[c] * sc1, tab[16 + c] = v4[kW1SumOff + c];
    for (int c = 0; c < 24; ++c) tab[32 + c] = bt[br][3]->data[c] * fs;
    for (int c = 0; c < 32; ++c) tab[56 + c] = bt[br][5]->data[c] * fs;
  }
  out->w_exp = pick_exp(wmax > 0.f ? wmax : 1.f);
  const float ws = std::ldexp(1.0f, out->w_exp);
  out->w1_hi.resize(size_t(kFc1) * kFeat);
  out->w1_lo.resize(size_t(kFc1) * kFeat);
  for (int k = 0; k < kFeat; ++k)
    for (int j = 0; j < kFc1; ++j) {
      const float s = out->w1[size_t(k) * kFc1 + j] * ws;
      const uint16_t hi = f32_to_f16_bits(s);
      const float lo = s - f16_bits_to_f32(hi);
      out->w1_hi[size_t(j) * kFeat + k] = hi;
      out->w1_lo[size_t(j) * kFeat + k] = f32_to_f16_bits(lo);
    }

  // FC2 operands for the fused kernel
  double a1_bound = 0;
  for (int j = 0; j < kFc1; ++j) {
    double s = 0;
    for (int k = 0; k < kFeat; ++k) s += std::fabs(out->w1[size_t(k) * kFc1 + j]);
    s = s * feat_bound + std::fabs(out->b1[j]);
    if (s > a1_bound) a1_bound = s;
  }
  out->a1_bound = float(a1_bound);
  out->a1_exp = pick_exp(float(a1_bound));
  float w2max = 0.f;
  for (int h = 0; h < 3; ++h)
    for (float v : out->w2[h]) {
      if (!std::isfinite(v)) {
        *err = "non-finite FC2 weight";
        return false;
      }
      if (std::fabs(v) > w2max) w2max = std::fabs(v);
    }
  out->w2_exp = pick_exp(w2max > 0.f ? w2max : 1.f);
  const float w2s = std::ldexp(1.0f, out->w2_exp);
  for (int h = 0; h < 3; ++h) {
    out->w2_hi[h].resize(size_t(n1[h]) * n2[h]);
    out->w2_lo[h].resize(size_t(n1[h]) * n2[h]);
    for (int k = 0; k < n1[h]; ++k)
      for (int j = 0; j < n2[h]; ++j) {
        const float s = out->w2[h][size_t(k) * n2[h] + j] * w2s;
        const uint16_t hi = f32_to_f16_bits(s);
        out->w2_hi[h][size_t(j) * n1[h] + k] = hi;
        out->w2_lo[h][size_t(j) * n1[h] + k] = f32_to_f16_bits(s - f16_bits_to_f32(hi));
      }
  }
  return true;
}

}  // namespace ethcnn
