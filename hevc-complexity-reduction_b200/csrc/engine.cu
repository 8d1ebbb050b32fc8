// Host engine + C ABI of libethcnn_b200.so (see include/ethcnn.h for the contract and the reference
// interfaces each entry point replaces).
//
// Data flow for one call (per device):
//   host luma --H2D (copy stream, slabs of whole frames, Y only)--> device slab [frames][H][pitch16]
//   CONV kernel (mma.sync, TMA tiles) -> features fp16 hi/lo [chunk][2688]   (scratch, chunk = 148 * 256 = 37 888 CTUs)
//   FC kernel (tcgen05, CTA pairs: FC1 + FC2 + FC3 + sigmoid) -> raw probabilities + gate flags
//   GATE kernel over the slab (in place; or gate + export to a staged / peer destination; or gate + 2-bit decision map)
//   D2H of 84 B/CTU (+ 8 B/CTU of decision map) into the caller's buffer (its frame range = the "gather")
// Slabs (~16 MB of luma) rotate through three buffers so copies overlap kernels.  With n_gpus > 1 one host thread per device
// runs the same pipeline on a contiguous frame range (video_to_cu_depth.py:88 loop, sharded).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ethcnn.h"
#include "conv_tc.h"
#include "fc1_tc.h"
#include "fc_fused.h"
#include "kernels.h"
#include "model.h"
#include "tf_bundle.h"

namespace ethcnn {
namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

// ETHCNN_TRACE=1: wall-clock milestones of a call on stderr (where does a one-shot CLI run spend its time?)
void trace(const char* what) {
  static const bool on = getenv("ETHCNN_TRACE") != nullptr;
  if (!on) return;
  static const auto t0 = std::chrono::steady_clock::now();
  static auto last = t0;
  const auto now = std::chrono::steady_clock::now();
  fprintf(stderr, "[ethcnn %8.1f ms  +%7.1f] %s\n", std::chrono::duration<double, std::milli>(now - t0).count(),
          std::chrono::duration<double, std::milli>(now - last).count(), what);
  last = now;
}

#define CUDA_TRY(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t e__ = (expr);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
      return fail(ETHCNN_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));                   \
  } while (0)

struct DeviceModel {
  float* conv = nullptr;
  uint8_t* conv_tc = nullptr;  // the conv filters as swizzled UMMA tiles (conv_tc.h)
  float conv_hdr[3][4] = {};   // per branch: 32 * 2^-e1w, 2^-(e_c1 + e2w), 2^-(feat_exp + e3w), 2^e_c1 (kernels.h kHdrOff)
  float* w1 = nullptr;
  float* b1 = nullptr;
  __half* w1_hi = nullptr;
  __half* w1_lo = nullptr;
  float* heads = nullptr;  // one allocation holding w2/w2q/b2/w3/w3q/b3 of the three heads
  HeadWeights hw[3];
  int feat_exp = 0, w_exp = 0;
  Fc1TcWeights tc;         // tensor maps over w1_hi / w1_lo
  // fused FC1 + FC2 + FC3 kernel
  __half* w2_hi[3] = {nullptr, nullptr, nullptr};
  __half* w2_lo[3] = {nullptr, nullptr, nullptr};
  float* w3_packed = nullptr;  // [48*1 | 96*4 | 192*16]
  FusedWeights fused;
  int a1_exp = 0, w2_exp = 0;
  std::vector<float> h_b2, h_w2q, h_b3, h_w3q;  // host copies: b2eff / b3eff depend on the call's qp
  std::string index_bytes;  // the checkpoint's .index file as loaded (holds the crc32c of every tensor): identity of the weights
};

// One ETH-LSTM checkpoint (HM-16.5_Test_LDP/bin/model_LDP_200000_qp*.dat, 18 tensors) on the device.
struct LstmDeviceModel {
  float* blob = nullptr;            // one allocation
  const float* kernel[3] = {};      // [2n][4n]
  const float* bias[3] = {};        // [4n]
  HeadWeights hw[3];                // w2 = first n rows of fc2 [n+5][n2], w3 = first n2 rows of fc3 [n2+5][n3]; qp rows unused
  std::vector<float> h_w2e, h_b2;   // host: the 5 extra-feature rows of fc2 (per head, [5][n2]) and its bias, heads side by side
  std::vector<float> h_w3e, h_b3;   // same for fc3
};

struct ProfEvent {
  int stage;
  cudaEvent_t a, b;
};

struct DeviceCtx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t s_compute = nullptr, s_h2d = nullptr, s_d2h = nullptr;
  std::map<std::string, DeviceModel> models;
  // scratch for the device path
  size_t chunk_ctus = 0;
  __half* feat_hi = nullptr;
  __half* feat_lo = nullptr;
  float* fc1 = nullptr;
  unsigned* flags = nullptr;
  size_t flags_cap = 0;
  float* gprob = nullptr;         // staged raw probabilities (ETHCNN_OPT_STAGED_OUTPUT): the gate kernel exports them
  size_t gprob_floats = 0;
  std::vector<void*> peer_owned;  // gather buffers this device exported (ethcnn_peer_buffer_create)
  std::vector<void*> peer_opened; // peer buffers mapped into this process (ethcnn_peer_buffer_open)
  cudaEvent_t ev_last = nullptr;  // end of the previous device call (scratch reuse across streams)
  cudaStream_t last_stream = nullptr;
  // staging for the host path
  static constexpr int kSlabs = 3;
  uint8_t* d_slab[kSlabs] = {};
  float* d_prob[kSlabs] = {};
  uint8_t* h_stage[kSlabs] = {};
  float* h_prob[kSlabs] = {};
  unsigned long long* d_map[kSlabs] = {};   // packed decision maps of a slab (ethcnn_predict_luma_map)
  unsigned long long* h_map[kSlabs] = {};
  size_t slab_map_words = 0, hmap_words = 0;
  size_t slab_bytes = 0, slab_prob_floats = 0, stage_bytes = 0, hprob_floats = 0;
  cudaEvent_t ev_h2d[kSlabs] = {}, ev_comp[kSlabs] = {}, ev_d2h[kSlabs] = {};
  // profiling
  bool profiling = false;
  std::vector<ProfEvent> prof;
  std::vector<cudaEvent_t> ev_pool;  // timing events are recycled: creating them per launch costs more than the kernels
  double prof_ms[ETHCNN_N_STAGES] = {0, 0, 0, 0};
  int64_t prof_launches[ETHCNN_N_STAGES] = {0, 0, 0, 0};
  int last_used_tma = 0;
  int last_feat_exp = 0;
  // LDP one-step LSTM path
  std::map<std::string, LstmDeviceModel> lstm_models;
  float* d_state_in = nullptr;
  float* d_state_out = nullptr;
  float* d_z = nullptr;
  float* d_beff = nullptr;          // [336 + 21 (+3 pad) + 336 zeros] folded biases of the call + a zero row
  float* h_beff = nullptr;          // pinned staging for d_beff
  size_t lstm_rows = 0;
};

}  // namespace
}  // namespace ethcnn

using namespace ethcnn;

struct ethcnn_handle {
  int mode = ETHCNN_MODE_AI;
  std::string model_dir;
  float t1 = 0.5f, t2 = 0.5f;
  bool have_thr = false;
  float thr6[6] = {0.5f, 0.5f, 0.5f, 0.5f, 0.5f, 0.5f};   // up, down per depth as HM uses them (TEncCu.cpp:250, 448-457)
  bool have_thr6 = false;
  std::vector<std::unique_ptr<DeviceCtx>> devs;
  std::mutex mu;
  std::atomic<int64_t> launches{0};
  // 0 = SIMT fp32 FC1 + heads kernel, 1 = tcgen05 FC1 + heads kernel, 2 = fused tcgen05 FC1+FC2+FC3 (a CTA per tile),
  // 3 = the fused kernel on CTA pairs (cta_group::2)
  int fc1_path = 3;
  size_t chunk_ctus = 148 * 256;   // 74 CTA pairs x 2 column tiles x 256 CTUs: four fused-FC tiles per pair, 16 conv groups per CTA
  int conv_path = 0;               // 0 = mma.sync conv kernel (conv_stage.cu), 1 = tcgen05 conv kernel (conv_tc.cu)
  bool staged_output = false;      // device path: dense kernel -> local staging -> gate kernel exports with coalesced stores
};

namespace ethcnn {
namespace {

// ----------------------------------------------------------------------------------------------
// Thr_info.txt: first line split on single spaces, tokens [1] and [3] (net_CNN.py:38-45).
int read_thresholds(const std::string& path, float* t1, float* t2, float* all6 = nullptr, bool* have6 = nullptr) {
  FILE* f = fopen(path.c_str(), "r");
  if (!f) return fail(ETHCNN_E_IO, "cannot open " + path);
  char line[4096];
  if (!fgets(line, sizeof(line), f)) {
    fclose(f);
    return fail(ETHCNN_E_FORMAT, path + ": empty");
  }
  fclose(f);
  std::vector<std::string> tok;  // python str.split(' '): consecutive spaces give empty tokens
  std::string cur;
  for (const char* p = line; *p; ++p) {
    if (*p == ' ') {
      tok.push_back(cur);
      cur.clear();
    } else {
      cur.push_back(*p);
    }
  }
  tok.push_back(cur);
  if (tok.size() < 4) return fail(ETHCNN_E_FORMAT, path + ": fewer than 4 tokens");
  auto to_float = [](const std::string& s, float* out) {
    char* end = nullptr;
    const double v = strtod(s.c_str(), &end);
    if (end == s.c_str()) return false;
    while (*end == '\n' || *end == '\r' || *end == '\t') ++end;  // python float() strips whitespace
    if (*end) return false;
    *out = float(v);
    return true;
  };
  if (!to_float(tok[1], t1) || !to_float(tok[3], t2)) return fail(ETHCNN_E_FORMAT, path + ": tokens [1]/[3] are not numbers");
  if (all6 && have6) {   // HM itself reads six numbers (fscanf, TEncCu.cpp:250): needed only for the decision map
    *have6 = tok.size() >= 6;
    for (int i = 0; i < 6 && *have6; ++i) *have6 = to_float(tok[i], &all6[i]);
  }
  return ETHCNN_OK;
}

// Model selection by QP range (video_to_cu_depth.py:126-133) / the single LDP CNN checkpoint
// (resi_to_cu_depth_LDP.py:158-159).
std::string model_prefix(int mode, int qp) {
  if (mode == ETHCNN_MODE_LDP) return "model_LDP_2000000_qp22~37.dat";
  if (qp < 25) return "model_2000000_qp20~25.dat";
  if (qp < 30) return "model_2000000_qp25~30.dat";
  if (qp < 35) return "model_2000000_qp30~35.dat";
  return "model_2000000_qp35~40.dat";
}

float scaled_qp(int mode, int qp) {
  if (mode == ETHCNN_MODE_LDP) return (float(qp) / 51.0f) * 0.18f;  // net_CTU64.py:103
  return float(qp) * float(1.0 / 51.0);                             // net_CNN.py:106 (scalar_mul(1/51.0, qp))
}

template <class T>
int upload(T** dst, const void* src, size_t bytes) {
  CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(dst), bytes));
  CUDA_TRY(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
  return ETHCNN_OK;
}

void free_model(DeviceModel& m) {
  cudaFree(m.conv), cudaFree(m.conv_tc), cudaFree(m.w1), cudaFree(m.b1), cudaFree(m.w1_hi), cudaFree(m.w1_lo), cudaFree(m.heads);
  for (int k = 0; k < 3; ++k) cudaFree(m.w2_hi[k]), cudaFree(m.w2_lo[k]);
  cudaFree(m.w3_packed);
  m = DeviceModel();
}

bool slurp(const std::string& path, std::string* out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  out->clear();
  char buf[4096];
  size_t n;
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) out->append(buf, n);
  fclose(f);
  return true;
}

int get_model(ethcnn_handle* h, DeviceCtx& c, int qp, DeviceModel** out) {
  const std::string prefix = model_prefix(h->mode, qp);
  auto it = c.models.find(prefix);
  if (it != c.models.end()) {
    *out = &it->second;
    return ETHCNN_OK;
  }
  std::map<std::string, BundleTensor> tensors;
  std::string err;
  const std::string path = h->model_dir.empty() ? prefix : h->model_dir + "/" + prefix;
  trace("get_model: start");
  if (!read_tf_bundle(path, &tensors, &err)) {
    const bool io = err.find("cannot open") != std::string::npos;
    return fail(io ? ETHCNN_E_IO : ETHCNN_E_FORMAT, err);
  }
  PackedModel pm;
  trace("get_model: checkpoint read");
  if (!pack_model(tensors, h->mode == ETHCNN_MODE_LDP ? 10.0f : 1.0f, &pm, &err)) return fail(ETHCNN_E_FORMAT, path + ": " + err);
  trace("get_model: packed");
  DeviceModel m;
  int rc;
  slurp(path + ".index", &m.index_bytes);
  if ((rc = upload(&m.conv, pm.conv.data(), pm.conv.size() * 4))) return rc;
  if ((rc = upload(&m.conv_tc, pm.conv_tc.data(), pm.conv_tc.size()))) return rc;
  for (int br = 0; br < 3; ++br)
    for (int i = 0; i < 4; ++i) m.conv_hdr[br][i] = pm.conv[size_t(br) * kConvBranchFloats + kHdrOff + i];
  if ((rc = upload(&m.w1, pm.w1.data(), pm.w1.size() * 4))) return rc;
  if ((rc = upload(&m.b1, pm.b1.data(), pm.b1.size() * 4))) return rc;
  if ((rc = upload(&m.w1_hi, pm.w1_hi.data(), pm.w1_hi.size() * 2))) return rc;
  if ((rc = upload(&m.w1_lo, pm.w1_lo.data(), pm.w1_lo.size() * 2))) return rc;
  std::vector<float> hb;
  size_t off[3][6];
  for (int k = 0; k < 3; ++k) {
    const std::vector<float>* parts[6] = {&pm.w2[k], &pm.w2q[k], &pm.b2[k], &pm.w3[k], &pm.w3q[k], &pm.b3[k]};
    for (int j = 0; j < 6; ++j) {
      while (hb.size() % 4) hb.push_back(0.f);
      off[k][j] = hb.size();
      hb.insert(hb.end(), parts[j]->begin(), parts[j]->end());
    }
  }
  if ((rc = upload(&m.heads, hb.data(), hb.size() * 4))) return rc;
  for (int k = 0; k < 3; ++k) {
    m.hw[k].w2 = m.heads + off[k][0], m.hw[k].w2q = m.heads + off[k][1], m.hw[k].b2 = m.heads + off[k][2];
    m.hw[k].w3 = m.heads + off[k][3], m.hw[k].w3q = m.heads + off[k][4], m.hw[k].b3 = m.heads + off[k][5];
  }
  m.feat_exp = pm.feat_exp;
  m.w_exp = pm.w_exp;
  const char* terr = nullptr;
  if (!fc1_tc_prepare_weights(m.w1_hi, m.w1_lo, &m.tc, &terr)) return fail(ETHCNN_E_CUDA, std::string("FC1 weight tensor map: ") + terr);
  std::vector<float> w3p;
  for (int k = 0; k < 3; ++k) {
    if ((rc = upload(&m.w2_hi[k], pm.w2_hi[k].data(), pm.w2_hi[k].size() * 2))) return rc;
    if ((rc = upload(&m.w2_lo[k], pm.w2_lo[k].data(), pm.w2_lo[k].size() * 2))) return rc;
    w3p.insert(w3p.end(), pm.w3[k].begin(), pm.w3[k].end());
    m.h_b2.insert(m.h_b2.end(), pm.b2[k].begin(), pm.b2[k].end());
    m.h_w2q.insert(m.h_w2q.end(), pm.w2q[k].begin(), pm.w2q[k].end());
    m.h_b3.insert(m.h_b3.end(), pm.b3[k].begin(), pm.b3[k].end());
    m.h_w3q.insert(m.h_w3q.end(), pm.w3q[k].begin(), pm.w3q[k].end());
  }
  if ((rc = upload(&m.w3_packed, w3p.data(), w3p.size() * 4))) return rc;
  m.a1_exp = pm.a1_exp;
  m.w2_exp = pm.w2_exp;
  if (!fc_fused_prepare_weights(m.w1_hi, m.w1_lo, m.w2_hi, m.w2_lo, &m.fused, &terr))
    return fail(ETHCNN_E_CUDA, std::string("fused FC weight tensor maps: ") + terr);
  // cudaMemcpy from pageable memory may return before the DMA has landed, and the kernels run on non-blocking streams
  CUDA_TRY(cudaDeviceSynchronize());
  trace("get_model: uploaded");
  auto ins = c.models.emplace(prefix, m);
  *out = &ins.first->second;
  return ETHCNN_OK;
}

int ensure_scratch(ethcnn_handle* h, DeviceCtx& c, size_t flags_needed) {
  if (c.chunk_ctus != h->chunk_ctus || !c.feat_hi) {
    trace("ensure_scratch: start");
    cudaFree(c.feat_hi), cudaFree(c.feat_lo), cudaFree(c.fc1);
    c.feat_hi = c.feat_lo = nullptr, c.fc1 = nullptr;
    c.chunk_ctus = h->chunk_ctus;
    // rows are padded to a multiple of 128 so the FC1 tile loads never leave the allocation
    const size_t rows = (c.chunk_ctus + 127) / 128 * 128;
    // + 1: the dump row of the conv stage (ConvLaunch::dump_row)
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c.feat_hi), (rows + 1) * kFeat * 2));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c.feat_lo), (rows + 1) * kFeat * 2));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c.fc1), rows * kFc1 * 4));
    CUDA_TRY(cudaMemset(c.feat_hi, 0, rows * kFeat * 2));
    CUDA_TRY(cudaMemset(c.feat_lo, 0, rows * kFeat * 2));
    // The fills run on the legacy stream, which does NOT order against the non-blocking streams the kernels use:
    // without this the zero fill of a fresh handle's buffers could land on top of the first slab's features.
    CUDA_TRY(cudaDeviceSynchronize());
    trace("ensure_scratch: feature buffers allocated + zeroed");
  }
  if (flags_needed > c.flags_cap) {
    cudaFree(c.flags);
    c.flags = nullptr;
    c.flags_cap = flags_needed * 2 + 64;
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c.flags), c.flags_cap * sizeof(unsigned)));
  }
  return ETHCNN_OK;
}

struct StageTimer {
  DeviceCtx& c;
  cudaStream_t s;
  int stage;
  ProfEvent ev{};
  bool on;
  StageTimer(DeviceCtx& ctx, cudaStream_t stream, int st) : c(ctx), s(stream), stage(st), on(ctx.profiling) {
    if (on) {
      ev.stage = st;
      ev.a = take(), ev.b = take();
      cudaEventRecord(ev.a, s);
    }
  }
  ~StageTimer() {
    if (on) {
      cudaEventRecord(ev.b, s);
      c.prof.push_back(ev);
    }
  }
  cudaEvent_t take() {
    if (c.ev_pool.empty()) fold_finished(c);
    cudaEvent_t e = nullptr;
    if (!c.ev_pool.empty()) {
      e = c.ev_pool.back();
      c.ev_pool.pop_back();
    } else {
      cudaEventCreate(&e);
    }
    return e;
  }
  // Fold event pairs that have already completed into the per-stage totals and recycle their events, so a
  // long timed region needs only as many events as are in flight (cudaEventCreate costs ~0.1 ms).
  static void fold_finished(DeviceCtx& c) {
    size_t done = 0;
    while (done < c.prof.size() && cudaEventQuery(c.prof[done].b) == cudaSuccess) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, c.prof[done].a, c.prof[done].b) == cudaSuccess) {
        c.prof_ms[c.prof[done].stage] += t;
        c.prof_launches[c.prof[done].stage] += 1;
      }
      c.ev_pool.push_back(c.prof[done].a), c.ev_pool.push_back(c.prof[done].b);
      ++done;
    }
    cudaGetLastError();  // cudaErrorNotReady from the query is expected
    c.prof.erase(c.prof.begin(), c.prof.begin() + done);
  }
};

// The device-resident forward pass: everything is enqueued on `stream`, nothing is synchronised.
// fc1_out != nullptr selects the LDP FC1 tap (conv + FC1 only, written to fc1_out [n][448]).
struct LdpStep {             // extra inputs of the LDP CNN + one-step LSTM evaluation
  const LstmDeviceModel* lm;
  const float* d_state_in;   // [rows][896]
  float* d_state_out;        // [rows][896]
  float* d_z;                // [rows][1792]
  const float* d_b2eff;      // [336]   b2 + efs rows folded
  const float* d_b3eff;      // [21]
  const float* d_zero;       // [>= 192] zeros (the qp rows are folded already)
};

int run_device(ethcnn_handle* h, DeviceCtx& c, const uint8_t* d_y, int width, int height, size_t pitch, size_t frame_stride,
               int n_frames, int qp, float* d_out, float* fc1_out, cudaStream_t stream, const LdpStep* ldp = nullptr,
               unsigned long long* d_map = nullptr) {
  if (width <= 0 || height <= 0 || n_frames < 0 || pitch < size_t(width)) return fail(ETHCNN_E_ARG, "bad frame geometry");
  if (n_frames == 0) return ETHCNN_OK;
  if (d_map && !h->have_thr6)
    return fail(ETHCNN_E_FORMAT, "the decision map needs the six thresholds of Thr_info.txt (or ethcnn_set_decision_thresholds)");
  if (d_map && (fc1_out || !d_out)) return fail(ETHCNN_E_ARG, "the decision map goes with the probability rows");
  CUDA_TRY(cudaSetDevice(c.device));
  DeviceModel* m = nullptr;
  int rc = get_model(h, c, qp, &m);
  if (rc) return rc;
  const int ctu_cols = (width + kCtu - 1) / kCtu, ctu_rows = (height + kCtu - 1) / kCtu;
  const int ctus_per_frame = ctu_cols * ctu_rows;
  const int chunks_per_frame = (ctus_per_frame + kSubBatch - 1) / kSubBatch;
  const long long total = (long long)n_frames * ctus_per_frame;
  if (total > 0x7fffffffLL / 32) return fail(ETHCNN_E_ARG, "too many CTUs in one call; split the sequence");
  if ((rc = ensure_scratch(h, c, size_t(n_frames) * chunks_per_frame))) return rc;
  // scratch reuse across streams: order this call behind the previous one unless it is on the same stream anyway (then the
  // wait would only stand between the previous call's gate kernel and this call's programmatically dependent conv kernel)
  if (c.ev_last && stream != c.last_stream) CUDA_TRY(cudaStreamWaitEvent(stream, c.ev_last, 0));
  c.last_stream = stream;

  CUtensorMap tmap;
  const char* terr = nullptr;
  const bool use_tma = make_luma_tensor_map(&tmap, d_y, width, height, n_frames, pitch, frame_stride, &terr);
  c.last_used_tma = use_tma ? 1 : 0;
  const bool gated = ((h->mode == ETHCNN_MODE_AI) && fc1_out == nullptr) || ldp != nullptr;
  // staged output: the dense kernels write their row-strided 4-byte stores to a local buffer and the gate kernel
  // copies the finished rows to d_out (peer memory over NVLink in the multi-GPU gather) with coalesced stores
  const bool staged = h->staged_output && fc1_out == nullptr && ldp == nullptr && d_out != nullptr;
  float* const final_out = d_out;
  if (staged) {
    if (size_t(total) * kProbs > c.gprob_floats) {
      cudaFree(c.gprob);
      c.gprob = nullptr, c.gprob_floats = 0;
      CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c.gprob), size_t(total) * kProbs * 4));
      c.gprob_floats = size_t(total) * kProbs;
    }
    d_out = c.gprob;
  }
  // the gate flags are cleared by the first conv launch of the call (mma.sync stage) or by a memset (tcgen05 conv stage)
  const bool conv_clears_flags = gated && !(h->conv_path == 1 && use_tma);
  if (gated && !conv_clears_flags) CUDA_TRY(cudaMemsetAsync(c.flags, 0, size_t(n_frames) * chunks_per_frame * sizeof(unsigned), stream));

  const float in_scale = (h->mode == ETHCNN_MODE_LDP) ? (10.0f / 255.0f) : (1.0f / 255.0f);
  for (long long begin = 0; begin < total; begin += (long long)c.chunk_ctus) {
    const int n = int(std::min<long long>(c.chunk_ctus, total - begin));
    ConvLaunch cl{};
    cl.convw = m->conv;
    cl.feat_hi = c.feat_hi, cl.feat_lo = c.feat_lo;
    cl.n_ctus = n, cl.ctu_begin = int(begin), cl.ctus_per_row = ctu_cols, cl.ctus_per_frame = ctus_per_frame;
    cl.dump_row = int((c.chunk_ctus + 127) / 128 * 128);
    cl.cst[0] = in_scale / 256.0f, cl.cst[1] = in_scale / 1024.0f, cl.cst[2] = in_scale / 4096.0f;
    cl.feat_scale = std::ldexp(1.0f, m->feat_exp);
    c.last_feat_exp = m->feat_exp;
    cl.luma = d_y, cl.pitch = pitch, cl.frame_stride = frame_stride, cl.width = width, cl.height = height;
    cl.clear_flags = (conv_clears_flags && begin == 0) ? c.flags : nullptr;
    cl.n_clear_flags = n_frames * chunks_per_frame;
    if (h->conv_path == 1 && use_tma) {   // tensor-core conv stage; needs the TMA tile loader
      ConvTcLaunch tl{};
      tl.blob = m->conv_tc;
      tl.feat_hi = c.feat_hi, tl.feat_lo = c.feat_lo;
      tl.n_ctus = n, tl.ctu_begin = int(begin), tl.ctus_per_row = ctu_cols, tl.ctus_per_frame = ctus_per_frame;
      for (int br = 0; br < 3; ++br) {
        const float* hd = m->conv_hdr[br];
        tl.u1[br] = hd[0] * cl.cst[br] * hd[3];
        tl.u1x8[br] = 8.f * tl.u1[br];
        tl.u2[br] = hd[1] * cl.feat_scale;
        tl.u3[br] = hd[2] * cl.feat_scale;
      }
      tl.phase_mask = 7;
      if (const char* e = getenv("ETHCNN_TC_PHASES")) tl.phase_mask = atoi(e);   // timing experiments (wrong results)
      StageTimer t(c, stream, ETHCNN_STAGE_CONV);
      CUDA_TRY(launch_conv_tc(tmap, tl, c.sm_count, stream));
      ++h->launches;
    } else {
      StageTimer t(c, stream, ETHCNN_STAGE_CONV);
      CUDA_TRY(launch_conv_features(use_tma ? &tmap : nullptr, cl, c.sm_count, stream));
      ++h->launches;
    }
    float* fc1_dst = fc1_out ? fc1_out + size_t(begin) * kFc1 : c.fc1;
    if (ldp != nullptr) {
      // LDP deployment: FC1 (the 448-vector) -> one LSTM step -> FC2 / FC3 on h with the five extra features folded
      {
        StageTimer t(c, stream, ETHCNN_STAGE_FC1);
        CUDA_TRY(launch_fc1_tc(c.feat_hi, c.feat_lo, m->tc, m->b1, std::ldexp(1.0f, -(m->feat_exp + m->w_exp)), c.fc1, n, c.sm_count,
                               stream));
        ++h->launches;
      }
      StageTimer t(c, stream, ETHCNN_STAGE_HEADS);
      CUDA_TRY(launch_lstm_step(c.fc1, ldp->d_state_in + size_t(begin) * 2 * kFc1, ldp->d_state_out + size_t(begin) * 2 * kFc1, ldp->d_z,
                                ldp->lm->kernel, ldp->lm->bias, n, stream));
      h->launches += 4;
      HeadsLaunch hl{};
      hl.fc1 = ldp->d_state_out + size_t(begin) * 2 * kFc1 + kFc1;   // h part of the new state rows
      hl.fc1_stride = 2 * kFc1;
      const int o2[3] = {0, 48, 144}, o3[3] = {0, 1, 5};
      for (int k = 0; k < 3; ++k) {
        hl.head[k] = ldp->lm->hw[k];
        hl.head[k].b2 = ldp->d_b2eff + o2[k], hl.head[k].w2q = ldp->d_zero;
        hl.head[k].b3 = ldp->d_b3eff + o3[k], hl.head[k].w3q = ldp->d_zero;
      }
      hl.q = 0.f;
      hl.prob = d_out, hl.flags = c.flags, hl.t1 = h->t1, hl.t2 = h->t2;
      hl.n_ctus = n, hl.ctu_begin = int(begin), hl.ctus_per_frame = ctus_per_frame, hl.chunks_per_frame = chunks_per_frame;
      CUDA_TRY(launch_heads(hl, stream));
      ++h->launches;
      continue;
    }
    if (h->fc1_path >= 2) {  // FC1 + FC2 + FC3 in one tcgen05 kernel; a1 never leaves the SM
      FusedParams fp;
      const float q = scaled_qp(h->mode, qp);
      for (int i = 0; i < 336; ++i) fp.b2eff[i] = std::fmaf(q, m->h_w2q[i], m->h_b2[i]);
      for (int i = 0; i < 21; ++i) fp.b3eff[i] = std::fmaf(q, m->h_w3q[i], m->h_b3[i]);
      fp.unscale1 = std::ldexp(1.0f, -(m->feat_exp + m->w_exp));
      fp.a1_scale = std::ldexp(1.0f, m->a1_exp);
      fp.unscale2 = std::ldexp(1.0f, -(m->a1_exp + m->w2_exp));
      fp.t1 = h->t1, fp.t2 = h->t2;
      fp.b1 = m->b1, fp.w3 = m->w3_packed;
      fp.prob = fc1_out ? nullptr : d_out;
      fp.fc1_out = fc1_out ? fc1_dst : nullptr;
      fp.flags = gated ? c.flags : nullptr;
      fp.n_ctus = n, fp.ctu_begin = int(begin), fp.ctus_per_frame = ctus_per_frame, fp.chunks_per_frame = chunks_per_frame;
      StageTimer t(c, stream, ETHCNN_STAGE_FC1);
      CUDA_TRY(launch_fc_fused(c.feat_hi, c.feat_lo, m->fused, fp, h->fc1_path == 3 ? 2 : 1, c.sm_count, stream));
      ++h->launches;
      continue;
    }
    {
      StageTimer t(c, stream, ETHCNN_STAGE_FC1);
      if (h->fc1_path == 1) {
        CUDA_TRY(launch_fc1_tc(c.feat_hi, c.feat_lo, m->tc, m->b1, std::ldexp(1.0f, -(m->feat_exp + m->w_exp)), fc1_dst, n,
                               c.sm_count, stream));
      } else {
        CUDA_TRY(launch_fc1_simt(c.feat_hi, c.feat_lo, std::ldexp(1.0f, -m->feat_exp), m->w1, m->b1, fc1_dst, n, stream));
      }
      ++h->launches;
    }
    if (fc1_out) continue;
    HeadsLaunch hl{};
    hl.fc1 = c.fc1;
    hl.fc1_stride = kFc1;
    for (int k = 0; k < 3; ++k) hl.head[k] = m->hw[k];
    hl.q = scaled_qp(h->mode, qp);
    hl.prob = d_out;
    hl.flags = gated ? c.flags : nullptr;
    hl.t1 = h->t1, hl.t2 = h->t2;
    hl.n_ctus = n, hl.ctu_begin = int(begin), hl.ctus_per_frame = ctus_per_frame, hl.chunks_per_frame = chunks_per_frame;
    {
      StageTimer t(c, stream, ETHCNN_STAGE_HEADS);
      CUDA_TRY(launch_heads(hl, stream));
      ++h->launches;
    }
  }
  if (d_map) {   // gates + HM's threshold rule in one pass: float rows (in place or exported) and the packed 2-bit map
    StageTimer t(c, stream, ETHCNN_STAGE_GATE);
    CUDA_TRY(launch_gate_map(d_out, final_out, gated ? c.flags : nullptr, h->t2, total, ctus_per_frame, chunks_per_frame, d_map, h->thr6,
                             stream));
    ++h->launches;
  } else if (staged) {
    StageTimer t(c, stream, ETHCNN_STAGE_GATE);
    CUDA_TRY(launch_gate_export(d_out, final_out, gated ? c.flags : nullptr, h->t2, total, ctus_per_frame, chunks_per_frame, stream));
    ++h->launches;
  } else if (gated) {
    StageTimer t(c, stream, ETHCNN_STAGE_GATE);
    CUDA_TRY(launch_gate(d_out, c.flags, h->t2, total, ctus_per_frame, chunks_per_frame, stream));
    ++h->launches;
  }
  if (!c.ev_last) CUDA_TRY(cudaEventCreateWithFlags(&c.ev_last, cudaEventDisableTiming));
  CUDA_TRY(cudaEventRecord(c.ev_last, stream));
  return ETHCNN_OK;
}

bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

void parallel_copy_frames(uint8_t* dst, const uint8_t* src, size_t frame_bytes, size_t src_stride, int n_frames) {
  const size_t total = frame_bytes * size_t(n_frames);
  int nt = int(std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), 8));
  if (total < (size_t(4) << 20)) nt = 1;
  auto work = [&](int t) {
    // split by bytes so a single huge frame is still shared between threads
    const size_t lo = total * t / nt, hi = total * (t + 1) / nt;
    size_t pos = lo;
    while (pos < hi) {
      const size_t f = pos / frame_bytes, o = pos - f * frame_bytes;
      const size_t len = std::min(frame_bytes - o, hi - pos);
      memcpy(dst + pos, src + f * src_stride + o, len);
      pos += len;
    }
  };
  if (nt == 1) {
    work(0);
    return;
  }
  std::vector<std::thread> th;
  for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
}

int ensure_staging(DeviceCtx& c, size_t slab_bytes, size_t slab_prob_floats, bool need_stage, bool need_hprob, size_t map_words = 0,
                   bool need_hmap = false) {
  if (map_words > c.slab_map_words) {
    for (int i = 0; i < DeviceCtx::kSlabs; ++i) {
      cudaFree(c.d_map[i]);
      c.d_map[i] = nullptr;
      CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c.d_map[i]), map_words * 8));
    }
    c.slab_map_words = map_words;
  }
  if (need_hmap && map_words > c.hmap_words) {
    for (int i = 0; i < DeviceCtx::kSlabs; ++i) {
      cudaFreeHost(c.h_map[i]);
      c.h_map[i] = nullptr;
      CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&c.h_map[i]), map_words * 8));
    }
    c.hmap_words = map_words;
  }
  if (slab_bytes > c.slab_bytes) {
    for (int i = 0; i < DeviceCtx::kSlabs; ++i) {
      cudaFree(c.d_slab[i]);
      c.d_slab[i] = nullptr;
      CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c.d_slab[i]), slab_bytes));
    }
    c.slab_bytes = slab_bytes;
  }
  if (slab_prob_floats > c.slab_prob_floats) {
    for (int i = 0; i < DeviceCtx::kSlabs; ++i) {
      cudaFree(c.d_prob[i]);
      c.d_prob[i] = nullptr;
      CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c.d_prob[i]), slab_prob_floats * 4));
    }
    c.slab_prob_floats = slab_prob_floats;
  }
  if (need_stage && slab_bytes > c.stage_bytes) {
    for (int i = 0; i < DeviceCtx::kSlabs; ++i) {
      cudaFreeHost(c.h_stage[i]);
      c.h_stage[i] = nullptr;
      CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&c.h_stage[i]), slab_bytes));
    }
    c.stage_bytes = slab_bytes;
  }
  if (need_hprob && slab_prob_floats > c.hprob_floats) {
    for (int i = 0; i < DeviceCtx::kSlabs; ++i) {
      cudaFreeHost(c.h_prob[i]);
      c.h_prob[i] = nullptr;
      CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&c.h_prob[i]), slab_prob_floats * 4));
    }
    c.hprob_floats = slab_prob_floats;
  }
  for (int i = 0; i < DeviceCtx::kSlabs; ++i) {
    if (!c.ev_h2d[i]) CUDA_TRY(cudaEventCreateWithFlags(&c.ev_h2d[i], cudaEventDisableTiming));
    if (!c.ev_comp[i]) CUDA_TRY(cudaEventCreateWithFlags(&c.ev_comp[i], cudaEventDisableTiming));
    if (!c.ev_d2h[i]) CUDA_TRY(cudaEventCreateWithFlags(&c.ev_d2h[i], cudaEventDisableTiming));
  }
  return ETHCNN_OK;
}

// Host-input pipeline on ONE device for frames [0, n_frames) at y / out (already offset by the caller).
// per_ctu_out = 21 (probabilities) or 448 (FC1 tap).
int run_host_pipeline(ethcnn_handle* h, DeviceCtx& c, const uint8_t* y, int width, int height, size_t frame_stride,
                      int n_frames, int qp, float* out, bool fc1_tap, unsigned long long* map_out = nullptr) {
  if (n_frames <= 0) return ETHCNN_OK;
  CUDA_TRY(cudaSetDevice(c.device));
  const int per_ctu = fc1_tap ? kFc1 : kProbs;
  const size_t luma_bytes = size_t(width) * height;
  const size_t pitch = (size_t(width) + 15) / 16 * 16;
  const size_t dev_frame = pitch * height;  // multiple of 16
  const int ctus_per_frame = ((width + kCtu - 1) / kCtu) * ((height + kCtu - 1) / kCtu);
  // slabs of ~16 MB: small enough that the first H2D and the last kernels/D2H (the un-overlapped ends of the
  // pipeline) are short, large enough to amortise the ~10 launches per slab
  const size_t target = size_t(16) << 20;
  int slab_frames = int(std::max<size_t>(1, std::min<size_t>(size_t(n_frames), target / dev_frame)));
  if (n_frames > 1 && slab_frames > (n_frames + 2) / 3) slab_frames = (n_frames + 2) / 3;
  const bool src_pinned = is_pinned(y);
  const bool dst_pinned = is_pinned(out);
  trace("host pipeline: start");
  const bool map_pinned = map_out ? is_pinned(map_out) : true;
  int rc = ensure_staging(c, dev_frame * slab_frames, size_t(slab_frames) * ctus_per_frame * per_ctu, !src_pinned, !dst_pinned,
                          map_out ? size_t(slab_frames) * ctus_per_frame : 0, !map_pinned);
  if (rc) return rc;
  trace("host pipeline: staging buffers ready");

  struct Pending {
    int slab = -1, f0 = 0, nf = 0;
  } pend[DeviceCtx::kSlabs];
  auto drain = [&](int b) -> int {  // finish the D2H of buffer b and hand the rows to the caller
    if (pend[b].slab < 0) return ETHCNN_OK;
    CUDA_TRY(cudaEventSynchronize(c.ev_d2h[b]));
    if (!dst_pinned)
      memcpy(out + size_t(pend[b].f0) * ctus_per_frame * per_ctu, c.h_prob[b], size_t(pend[b].nf) * ctus_per_frame * per_ctu * 4);
    if (map_out && !map_pinned)
      memcpy(map_out + size_t(pend[b].f0) * ctus_per_frame, c.h_map[b], size_t(pend[b].nf) * ctus_per_frame * 8);
    pend[b].slab = -1;
    return ETHCNN_OK;
  };

  int slab_idx = 0;
  for (int f0 = 0; f0 < n_frames; f0 += slab_frames, ++slab_idx) {
    const int nf = std::min(slab_frames, n_frames - f0);
    const int b = slab_idx % DeviceCtx::kSlabs;
    if ((rc = drain(b))) return rc;  // also guarantees buffer b (device slab, staging, prob) is free
    const uint8_t* src = y + size_t(f0) * frame_stride;
    size_t src_stride = frame_stride;
    if (!src_pinned) {
      parallel_copy_frames(c.h_stage[b], src, luma_bytes, frame_stride, nf);
      src = c.h_stage[b];
      src_stride = luma_bytes;
    }
    if (pitch == size_t(width) && src_stride == luma_bytes) {   // luma-only clip: one contiguous copy
      CUDA_TRY(cudaMemcpyAsync(c.d_slab[b], src, luma_bytes * nf, cudaMemcpyHostToDevice, c.s_h2d));
    } else if (pitch == size_t(width)) {
      CUDA_TRY(cudaMemcpy2DAsync(c.d_slab[b], dev_frame, src, src_stride, luma_bytes, nf, cudaMemcpyHostToDevice, c.s_h2d));
    } else {
      for (int f = 0; f < nf; ++f)
        CUDA_TRY(cudaMemcpy2DAsync(c.d_slab[b] + size_t(f) * dev_frame, pitch, src + size_t(f) * src_stride, width, width, height,
                                   cudaMemcpyHostToDevice, c.s_h2d));
    }
    CUDA_TRY(cudaEventRecord(c.ev_h2d[b], c.s_h2d));
    CUDA_TRY(cudaStreamWaitEvent(c.s_compute, c.ev_h2d[b], 0));
    rc = run_device(h, c, c.d_slab[b], width, height, pitch, dev_frame, nf, qp, fc1_tap ? nullptr : c.d_prob[b],
                    fc1_tap ? c.d_prob[b] : nullptr, c.s_compute, nullptr, map_out ? c.d_map[b] : nullptr);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(c.ev_comp[b], c.s_compute));
    CUDA_TRY(cudaStreamWaitEvent(c.s_d2h, c.ev_comp[b], 0));
    const size_t nfl = size_t(nf) * ctus_per_frame * per_ctu;
    float* dst = dst_pinned ? out + size_t(f0) * ctus_per_frame * per_ctu : c.h_prob[b];
    CUDA_TRY(cudaMemcpyAsync(dst, c.d_prob[b], nfl * 4, cudaMemcpyDeviceToHost, c.s_d2h));
    if (map_out) {
      unsigned long long* mdst = map_pinned ? map_out + size_t(f0) * ctus_per_frame : c.h_map[b];
      CUDA_TRY(cudaMemcpyAsync(mdst, c.d_map[b], size_t(nf) * ctus_per_frame * 8, cudaMemcpyDeviceToHost, c.s_d2h));
    }
    CUDA_TRY(cudaEventRecord(c.ev_d2h[b], c.s_d2h));
    pend[b].slab = slab_idx, pend[b].f0 = f0, pend[b].nf = nf;
  }
  for (int b = 0; b < DeviceCtx::kSlabs; ++b)
    if ((rc = drain(b))) return rc;
  trace("host pipeline: all slabs done");
  return ETHCNN_OK;
}

// Contiguous frame ranges per device: the first (n % G) devices take one extra frame.
void frame_range(int n_frames, int n_dev, int r, int* f0, int* nf) {
  const int base = n_frames / n_dev, rem = n_frames % n_dev;
  *f0 = r * base + std::min(r, rem);
  *nf = base + (r < rem ? 1 : 0);
}

int run_host_all_devices(ethcnn_handle* h, const uint8_t* y, int width, int height, size_t frame_stride, int n_frames, int qp,
                         float* out, bool fc1_tap, unsigned long long* map_out = nullptr) {
  const int per_ctu = fc1_tap ? kFc1 : kProbs;
  const size_t ctus_per_frame = size_t((width + kCtu - 1) / kCtu) * ((height + kCtu - 1) / kCtu);
  const int nd = int(h->devs.size());
  if (nd == 1 || n_frames < 2)
    return run_host_pipeline(h, *h->devs[0], y, width, height, frame_stride, n_frames, qp, out, fc1_tap, map_out);
  std::vector<int> rcs(nd, 0);
  std::vector<std::string> errs(nd);
  std::vector<std::thread> th;
  for (int r = 0; r < nd; ++r) {
    th.emplace_back([&, r]() {
      int f0, nf;
      frame_range(n_frames, nd, r, &f0, &nf);
      rcs[r] = run_host_pipeline(h, *h->devs[r], y + size_t(f0) * frame_stride, width, height, frame_stride, nf, qp,
                                 out + size_t(f0) * ctus_per_frame * per_ctu, fc1_tap, map_out ? map_out + size_t(f0) * ctus_per_frame : nullptr);
      if (rcs[r]) errs[r] = g_last_error;
    });
  }
  for (auto& t : th) t.join();
  for (int r = 0; r < nd; ++r)
    if (rcs[r]) return fail(rcs[r], "device " + std::to_string(h->devs[r]->device) + ": " + errs[r]);
  return ETHCNN_OK;
}

// LSTM checkpoint selection (resi_to_cu_depth_LDP.py:169-177).
std::string lstm_prefix(int qp) {
  if (qp < 25) return "model_LDP_200000_qp22.dat";
  if (qp < 30) return "model_LDP_200000_qp27.dat";
  if (qp < 35) return "model_LDP_200000_qp32.dat";
  return "model_LDP_200000_qp37.dat";
}

int get_lstm_model(ethcnn_handle* h, DeviceCtx& c, int qp, LstmDeviceModel** out) {
  const std::string prefix = lstm_prefix(qp);
  auto it = c.lstm_models.find(prefix);
  if (it != c.lstm_models.end()) {
    *out = &it->second;
    return ETHCNN_OK;
  }
  std::map<std::string, BundleTensor> t;
  std::string err;
  const std::string path = h->model_dir + "/" + prefix;
  if (!read_tf_bundle(path, &t, &err)) return fail(err.find("cannot open") != std::string::npos ? ETHCNN_E_IO : ETHCNN_E_FORMAT, err);
  static const char* hn[3] = {"64", "32", "16"};
  const int n[3] = {64, 128, 256}, n2[3] = {48, 96, 192}, n3[3] = {1, 4, 16};
  LstmDeviceModel m;
  std::vector<float> blob;
  size_t off_k[3], off_b[3], off_w2[3], off_w3[3];
  for (int k = 0; k < 3; ++k) {
    const std::string pre = std::string("RNN") + hn[k] + "/";
    auto need = [&](const std::string& name, std::vector<int64_t> shape) -> const BundleTensor* {
      auto f = t.find(pre + name);
      if (f == t.end() || f->second.shape != shape) {
        err = path + ": tensor " + pre + name + " missing or of unexpected shape";
        return nullptr;
      }
      return &f->second;
    };
    const BundleTensor* kr = need("multi_rnn_cell/cell_0/lstm_cell/kernel", {2 * n[k], 4 * n[k]});
    const BundleTensor* bs = need("multi_rnn_cell/cell_0/lstm_cell/bias", {4 * n[k]});
    const BundleTensor* w2 = need("fc2/full_connect_w", {n[k] + 5, n2[k]});
    const BundleTensor* b2 = need("fc2/full_connect_b", {n2[k]});
    const BundleTensor* w3 = need("fc3/full_connect_w", {n2[k] + 5, n3[k]});
    const BundleTensor* b3 = need("fc3/full_connect_b", {n3[k]});
    if (!kr || !bs || !w2 || !b2 || !w3 || !b3) return fail(ETHCNN_E_FORMAT, err);
    auto push = [&](const float* p, size_t cnt) {
      while (blob.size() % 4) blob.push_back(0.f);
      const size_t o = blob.size();
      blob.insert(blob.end(), p, p + cnt);
      return o;
    };
    off_k[k] = push(kr->data.data(), kr->data.size());
    off_b[k] = push(bs->data.data(), bs->data.size());
    off_w2[k] = push(w2->data.data(), size_t(n[k]) * n2[k]);     // rows of h; the 5 extra-feature rows stay on the host
    off_w3[k] = push(w3->data.data(), size_t(n2[k]) * n3[k]);
    m.h_w2e.insert(m.h_w2e.end(), w2->data.begin() + size_t(n[k]) * n2[k], w2->data.end());
    m.h_b2.insert(m.h_b2.end(), b2->data.begin(), b2->data.end());
    m.h_w3e.insert(m.h_w3e.end(), w3->data.begin() + size_t(n2[k]) * n3[k], w3->data.end());
    m.h_b3.insert(m.h_b3.end(), b3->data.begin(), b3->data.end());
  }
  int rc = upload(&m.blob, blob.data(), blob.size() * 4);
  if (rc) return rc;
  for (int k = 0; k < 3; ++k) {
    m.kernel[k] = m.blob + off_k[k], m.bias[k] = m.blob + off_b[k];
    m.hw[k].w2 = m.blob + off_w2[k], m.hw[k].w3 = m.blob + off_w3[k];
  }
  CUDA_TRY(cudaDeviceSynchronize());   // as in get_model: the upload must have landed before non-blocking streams use it
  auto ins = c.lstm_models.emplace(prefix, std::move(m));
  *out = &ins.first->second;
  return ETHCNN_OK;
}

// One frame of the deployed LDP predictor: residual ETH-CNN + one LSTM step + heads + gates
// (resi_to_cu_depth_LDP.py:114-129, net_CNN_LSTM_one_step.py:266-323).  Host pointers; state_in may be NULL (zeros).
int run_ldp_step(ethcnn_handle* h, const uint8_t* y, int width, int height, int qp, int i_frame, const float* state_in,
                 float* state_out, float* prob) {
  if (h->mode != ETHCNN_MODE_LDP) return fail(ETHCNN_E_ARG, "the LSTM step requires ETHCNN_MODE_LDP");
  if (!h->have_thr) return fail(ETHCNN_E_IO, "Thr_info.txt was not found: the LSTM step needs its gate thresholds");
  if (width <= 0 || height <= 0) return fail(ETHCNN_E_ARG, "bad frame geometry");
  DeviceCtx& c = *h->devs[0];
  CUDA_TRY(cudaSetDevice(c.device));
  LstmDeviceModel* lm = nullptr;
  int rc = get_lstm_model(h, c, qp, &lm);
  if (rc) return rc;
  const size_t rows = size_t((width + kCtu - 1) / kCtu) * ((height + kCtu - 1) / kCtu);
  const size_t pitch = (size_t(width) + 15) / 16 * 16, dev_frame = pitch * height;
  if ((rc = ensure_staging(c, dev_frame, rows * kProbs, false, false))) return rc;
  if (rows > c.lstm_rows) {
    cudaFree(c.d_state_in), cudaFree(c.d_state_out), cudaFree(c.d_z);
    c.d_state_in = c.d_state_out = c.d_z = nullptr;
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c.d_state_in), rows * 2 * kFc1 * 4));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c.d_state_out), rows * 2 * kFc1 * 4));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c.d_z), std::min(rows, h->chunk_ctus) * 4 * kFc1 * 4));
    c.lstm_rows = rows;
  }
  constexpr int kBeff = 336 + 24 + 336;
  if (!c.d_beff) {
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&c.d_beff), kBeff * 4));
    CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&c.h_beff), kBeff * 4));
  }
  // the five extra inputs of FC2 / FC3 are identical for every CTU of the frame: fold their rows into the biases
  const float efs[5] = {scaled_qp(ETHCNN_MODE_LDP, qp), (i_frame % 4 + 4) % 4 == 0 ? 1.f : 0.f, (i_frame % 4 + 4) % 4 == 1 ? 1.f : 0.f,
                        (i_frame % 4 + 4) % 4 == 2 ? 1.f : 0.f, (i_frame % 4 + 4) % 4 == 3 ? 1.f : 0.f};
  const int n2[3] = {48, 96, 192}, n3[3] = {1, 4, 16};
  memset(c.h_beff, 0, kBeff * 4);
  {
    size_t o2 = 0, o3 = 0, e2 = 0, e3 = 0;
    for (int k = 0; k < 3; ++k) {
      for (int j = 0; j < n2[k]; ++j) {
        float v = lm->h_b2[o2 + j];
        for (int e = 0; e < 5; ++e) v = std::fmaf(efs[e], lm->h_w2e[e2 + size_t(e) * n2[k] + j], v);
        c.h_beff[o2 + j] = v;
      }
      for (int j = 0; j < n3[k]; ++j) {
        float v = lm->h_b3[o3 + j];
        for (int e = 0; e < 5; ++e) v = std::fmaf(efs[e], lm->h_w3e[e3 + size_t(e) * n3[k] + j], v);
        c.h_beff[336 + o3 + j] = v;
      }
      o2 += n2[k], o3 += n3[k], e2 += 5 * size_t(n2[k]), e3 += 5 * size_t(n3[k]);
    }
  }
  cudaStream_t s = c.s_compute;
  CUDA_TRY(cudaMemcpyAsync(c.d_beff, c.h_beff, kBeff * 4, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpy2DAsync(c.d_slab[0], pitch, y, width, width, height, cudaMemcpyHostToDevice, s));
  if (state_in) {
    CUDA_TRY(cudaMemcpyAsync(c.d_state_in, state_in, rows * 2 * kFc1 * 4, cudaMemcpyHostToDevice, s));
  } else {
    CUDA_TRY(cudaMemsetAsync(c.d_state_in, 0, rows * 2 * kFc1 * 4, s));
  }
  LdpStep ls{lm, c.d_state_in, c.d_state_out, c.d_z, c.d_beff, c.d_beff + 336, c.d_beff + 360};
  if ((rc = run_device(h, c, c.d_slab[0], width, height, pitch, dev_frame, 1, qp, c.d_prob[0], nullptr, s, &ls))) return rc;
  CUDA_TRY(cudaMemcpyAsync(prob, c.d_prob[0], rows * kProbs * 4, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(state_out, c.d_state_out, rows * 2 * kFc1 * 4, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return ETHCNN_OK;
}

bool file_exists(const std::string& p) {
  struct stat st;
  return stat(p.c_str(), &st) == 0;
}

bool read_exact(const std::string& path, void* dst, size_t bytes) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  const size_t got = fread(dst, 1, bytes, f);
  fclose(f);
  return got == bytes;
}

bool write_all(const std::string& path, const void* src, size_t bytes) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  const size_t put = bytes ? fwrite(src, 1, bytes, f) : 0;
  return (fclose(f) == 0) && put == bytes;
}

int open_device(ethcnn_handle* h, int device) {
  int count = 0;
  trace("open_device: start");
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(ETHCNN_E_CUDA, "no CUDA device available (this library has no CPU fallback)");
  }
  if (device < 0 || device >= count) return fail(ETHCNN_E_ARG, "CUDA device " + std::to_string(device) + " does not exist");
  trace("open_device: driver initialised (cudaGetDeviceCount)");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  trace("open_device: device selected");
  if (prop.major != 10) return fail(ETHCNN_E_CUDA, std::string("device is ") + prop.name + " (sm_" + std::to_string(prop.major) +
                                                        std::to_string(prop.minor) + "); this build targets sm_100a only");
  std::unique_ptr<DeviceCtx> c(new DeviceCtx());
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  CUDA_TRY(cudaStreamCreateWithFlags(&c->s_compute, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
  CUDA_TRY(conv_features_configure());
  CUDA_TRY(conv_tc_configure());
  CUDA_TRY(fc1_tc_configure());
  CUDA_TRY(heads_configure());
  CUDA_TRY(fc_fused_configure());
  trace("open_device: context + streams + kernel attributes");
  h->devs.push_back(std::move(c));
  return ETHCNN_OK;
}

void close_device(DeviceCtx& c) {
  cudaSetDevice(c.device);
  cudaDeviceSynchronize();
  for (auto& kv : c.models) free_model(kv.second);
  for (auto& kv : c.lstm_models) cudaFree(kv.second.blob);
  cudaFree(c.d_state_in), cudaFree(c.d_state_out), cudaFree(c.d_z), cudaFree(c.d_beff), cudaFreeHost(c.h_beff);
  cudaFree(c.feat_hi), cudaFree(c.feat_lo), cudaFree(c.fc1), cudaFree(c.flags), cudaFree(c.gprob);
  for (void* p : c.peer_opened) cudaIpcCloseMemHandle(p);
  for (void* p : c.peer_owned) cudaFree(p);
  for (int i = 0; i < DeviceCtx::kSlabs; ++i) {
    cudaFree(c.d_slab[i]), cudaFree(c.d_prob[i]), cudaFreeHost(c.h_stage[i]), cudaFreeHost(c.h_prob[i]);
    cudaFree(c.d_map[i]), cudaFreeHost(c.h_map[i]);
    if (c.ev_h2d[i]) cudaEventDestroy(c.ev_h2d[i]);
    if (c.ev_comp[i]) cudaEventDestroy(c.ev_comp[i]);
    if (c.ev_d2h[i]) cudaEventDestroy(c.ev_d2h[i]);
  }
  for (auto& e : c.prof) cudaEventDestroy(e.a), cudaEventDestroy(e.b);
  for (auto& e : c.ev_pool) cudaEventDestroy(e);
  if (c.ev_last) cudaEventDestroy(c.ev_last);
  if (c.s_compute) cudaStreamDestroy(c.s_compute);
  if (c.s_h2d) cudaStreamDestroy(c.s_h2d);
  if (c.s_d2h) cudaStreamDestroy(c.s_d2h);
}

int create_common(const char* model_dir, const char* thr_path, int mode, const std::vector<int>& devices, ethcnn_handle** out) {
  if (!out) return fail(ETHCNN_E_ARG, "out is NULL");
  *out = nullptr;
  if (mode != ETHCNN_MODE_AI && mode != ETHCNN_MODE_LDP) return fail(ETHCNN_E_ARG, "unknown mode");
  std::unique_ptr<ethcnn_handle> h(new ethcnn_handle());
  h->mode = mode;
  h->model_dir = (model_dir && *model_dir) ? model_dir : ".";
  if (const char* e = getenv("ETHCNN_CHUNK_CTUS")) {
    const long v = atol(e);
    if (v >= 1 && v <= (1 << 22)) h->chunk_ctus = size_t(v);
  }
  if (const char* e = getenv("ETHCNN_CONV")) h->conv_path = (strcmp(e, "tc") == 0) ? 1 : 0;
  if (const char* e = getenv("ETHCNN_FC1")) h->fc1_path = (strcmp(e, "simt") == 0) ? 0 : (strcmp(e, "tc") == 0 ? 1 : (strcmp(e, "pair") == 0 ? 3 : 2));
  {
    const std::string tp = thr_path ? std::string(thr_path) : h->model_dir + "/Thr_info.txt";
    float tok6[6];
    bool have6 = false;
    int rc = read_thresholds(tp, &h->t1, &h->t2, tok6, &have6);
    h->have_thr = (rc == ETHCNN_OK);
    if (rc == ETHCNN_OK && have6) {
      // AI file order: up, down per depth (TEncCu.cpp:250); LDP file order: down, up (TEncGOP.cpp(LDP):1449)
      for (int d = 0; d < 3; ++d) {
        h->thr6[2 * d] = tok6[2 * d + (mode == ETHCNN_MODE_LDP ? 1 : 0)];
        h->thr6[2 * d + 1] = tok6[2 * d + (mode == ETHCNN_MODE_LDP ? 0 : 1)];
      }
      h->have_thr6 = true;
    }
    // the AI script cannot run without Thr_info.txt (net_CNN.py:47); the LDP CNN-only entry points can
    if (rc && (mode == ETHCNN_MODE_AI || thr_path)) return rc;
  }
  for (int d : devices) {
    int rc = open_device(h.get(), d);
    if (rc) {
      for (auto& c : h->devs) close_device(*c);
      return rc;
    }
  }
  *out = h.release();
  return ETHCNN_OK;
}

}  // namespace
}  // namespace ethcnn

// =================================================================================================
extern "C" {

int ethcnn_abi_version(void) { return ETHCNN_ABI_VERSION; }

const char* ethcnn_last_error(void) { return g_last_error.c_str(); }

int ethcnn_create(const char* model_dir, const char* thr_path, int mode, int n_gpus, ethcnn_handle** out) {
  if (n_gpus < 1) return fail(ETHCNN_E_ARG, "n_gpus must be >= 1");
  std::vector<int> devs;
  for (int i = 0; i < n_gpus; ++i) devs.push_back(i);
  return create_common(model_dir, thr_path, mode, devs, out);
}

int ethcnn_create_on_device(const char* model_dir, const char* thr_path, int mode, int device, ethcnn_handle** out) {
  return create_common(model_dir, thr_path, mode, std::vector<int>{device}, out);
}

void ethcnn_destroy(ethcnn_handle* h) {
  if (!h) return;
  trace("destroy: start");
  for (auto& c : h->devs) close_device(*c);
  delete h;
  trace("destroy: done");
}

int ethcnn_predict_luma_device(ethcnn_handle* h, const uint8_t* d_y, int width, int height, size_t pitch, size_t frame_stride,
                               int n_frames, int qp, float* d_out, void* stream) {
  if (!h || !d_y || !d_out) return fail(ETHCNN_E_ARG, "NULL argument");
  std::lock_guard<std::mutex> lock(h->mu);
  return run_device(h, *h->devs[0], d_y, width, height, pitch, frame_stride, n_frames, qp, d_out, nullptr,
                    static_cast<cudaStream_t>(stream));
}

int ethcnn_predict_luma(ethcnn_handle* h, const uint8_t* y, int width, int height, size_t frame_stride, int n_frames, int qp,
                        float* out) {
  if (!h || (!y && n_frames > 0) || (!out && n_frames > 0)) return fail(ETHCNN_E_ARG, "NULL argument");
  if (width <= 0 || height <= 0 || n_frames < 0 || frame_stride < size_t(width) * height) return fail(ETHCNN_E_ARG, "bad frame geometry");
  std::lock_guard<std::mutex> lock(h->mu);
  return run_host_all_devices(h, y, width, height, frame_stride, n_frames, qp, out, false);
}

int ethcnn_predict_luma_device_map(ethcnn_handle* h, const uint8_t* d_y, int width, int height, size_t pitch, size_t frame_stride,
                                   int n_frames, int qp, float* d_out, uint64_t* d_map, void* stream) {
  if (!h || !d_y || !d_out || !d_map) return fail(ETHCNN_E_ARG, "NULL argument");
  std::lock_guard<std::mutex> lock(h->mu);
  return run_device(h, *h->devs[0], d_y, width, height, pitch, frame_stride, n_frames, qp, d_out, nullptr,
                    static_cast<cudaStream_t>(stream), nullptr, reinterpret_cast<unsigned long long*>(d_map));
}

int ethcnn_predict_luma_map(ethcnn_handle* h, const uint8_t* y, int width, int height, size_t frame_stride, int n_frames, int qp,
                            float* out, uint64_t* map) {
  if (!h || ((!y || !out || !map) && n_frames > 0)) return fail(ETHCNN_E_ARG, "NULL argument");
  if (width <= 0 || height <= 0 || n_frames < 0 || frame_stride < size_t(width) * height) return fail(ETHCNN_E_ARG, "bad frame geometry");
  std::lock_guard<std::mutex> lock(h->mu);
  if (!h->have_thr6)
    return fail(ETHCNN_E_FORMAT, "the decision map needs the six thresholds of Thr_info.txt (or ethcnn_set_decision_thresholds)");
  return run_host_all_devices(h, y, width, height, frame_stride, n_frames, qp, out, false, reinterpret_cast<unsigned long long*>(map));
}

int ethcnn_set_decision_thresholds(ethcnn_handle* h, const float thr6[6]) {
  if (!h || !thr6) return fail(ETHCNN_E_ARG, "NULL argument");
  std::lock_guard<std::mutex> lock(h->mu);
  for (int i = 0; i < 6; ++i) h->thr6[i] = thr6[i];
  h->have_thr6 = true;
  return ETHCNN_OK;
}

int ethcnn_get_decision_thresholds(ethcnn_handle* h, float thr6[6]) {
  if (!h || !thr6) return fail(ETHCNN_E_ARG, "NULL argument");
  std::lock_guard<std::mutex> lock(h->mu);
  if (!h->have_thr6) return fail(ETHCNN_E_FORMAT, "Thr_info.txt did not carry six thresholds");
  for (int i = 0; i < 6; ++i) thr6[i] = h->thr6[i];
  return ETHCNN_OK;
}

int ethcnn_export_fc1(ethcnn_handle* h, const uint8_t* y, int width, int height, size_t frame_stride, int n_frames, float* out) {
  if (!h || (!y && n_frames > 0) || (!out && n_frames > 0)) return fail(ETHCNN_E_ARG, "NULL argument");
  if (h->mode != ETHCNN_MODE_LDP) return fail(ETHCNN_E_ARG, "ethcnn_export_fc1 requires ETHCNN_MODE_LDP");
  if (width <= 0 || height <= 0 || n_frames < 0 || frame_stride < size_t(width) * height) return fail(ETHCNN_E_ARG, "bad frame geometry");
  std::lock_guard<std::mutex> lock(h->mu);
  return run_host_all_devices(h, y, width, height, frame_stride, n_frames, 0, out, true);
}

int ethcnn_predict_yuv_file(ethcnn_handle* h, const char* yuv_path, int width, int height, int qp, const char* out_path) {
  if (!h || !yuv_path || !out_path) return fail(ETHCNN_E_ARG, "NULL argument");
  if (width <= 0 || height <= 0) return fail(ETHCNN_E_ARG, "bad frame geometry");
  const int fd = open(yuv_path, O_RDONLY);
  if (fd < 0) return fail(ETHCNN_E_IO, std::string("cannot open ") + yuv_path);
  struct stat st;
  if (fstat(fd, &st) != 0) {
    close(fd);
    return fail(ETHCNN_E_IO, std::string("cannot stat ") + yuv_path);
  }
  const size_t file_bytes = size_t(st.st_size);
  const size_t frame_bytes = size_t(width) * height * 3 / 2;
  if (frame_bytes == 0 || file_bytes % frame_bytes != 0) {  // video_to_cu_depth.py:137
    close(fd);
    return fail(ETHCNN_E_ARG, "file size is not a whole number of " + std::to_string(width) + "x" + std::to_string(height) +
                                  " 4:2:0 frames (file_bytes % frame_bytes != 0)");
  }
  const size_t n_frames = file_bytes / frame_bytes;
  if (n_frames > size_t(0x7fffffff)) {
    close(fd);
    return fail(ETHCNN_E_ARG, "too many frames");
  }
  const uint8_t* base = nullptr;
  if (file_bytes) {
    void* mp = mmap(nullptr, file_bytes, PROT_READ, MAP_PRIVATE, fd, 0);
    if (mp == MAP_FAILED) {
      close(fd);
      return fail(ETHCNN_E_IO, std::string("cannot map ") + yuv_path);
    }
    madvise(mp, file_bytes, MADV_SEQUENTIAL);
    base = static_cast<const uint8_t*>(mp);
  }
  trace("predict_yuv_file: file mapped");
  const size_t ctus = size_t((width + kCtu - 1) / kCtu) * ((height + kCtu - 1) / kCtu);
  std::vector<float> prob(n_frames * ctus * kProbs);
  int rc;
  {
    std::lock_guard<std::mutex> lock(h->mu);
    rc = run_host_all_devices(h, base, width, height, frame_bytes, int(n_frames), qp, prob.data(), false);
  }
  if (base) munmap(const_cast<uint8_t*>(base), file_bytes);
  close(fd);
  if (rc) return rc;
  trace("predict_yuv_file: probabilities on the host");
  // single write at the end (video_to_cu_depth.py:114-116), via a temporary so failures leave no partial file
  const std::string tmp = std::string(out_path) + ".tmp." + std::to_string(getpid());
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return fail(ETHCNN_E_IO, "cannot create " + tmp);
  const size_t wrote = prob.empty() ? 0 : fwrite(prob.data(), sizeof(float), prob.size(), f);
  const bool ok = (wrote == prob.size()) && (fclose(f) == 0);
  if (!ok) {
    remove(tmp.c_str());
    return fail(ETHCNN_E_IO, "short write on " + tmp);
  }
  if (rename(tmp.c_str(), out_path) != 0) {
    remove(tmp.c_str());
    return fail(ETHCNN_E_IO, std::string("cannot rename onto ") + out_path);
  }
  trace("predict_yuv_file: cu_depth.dat written");
  return ETHCNN_OK;
}

int ethcnn_reload_thresholds(ethcnn_handle* h, const char* thr_path) {
  if (!h) return fail(ETHCNN_E_ARG, "NULL handle");
  std::lock_guard<std::mutex> lock(h->mu);
  const std::string tp = (thr_path && *thr_path) ? std::string(thr_path) : h->model_dir + "/Thr_info.txt";
  float t1, t2, tok6[6];
  bool have6 = false;
  int rc = read_thresholds(tp, &t1, &t2, tok6, &have6);
  if (rc) return rc;
  h->t1 = t1, h->t2 = t2, h->have_thr = true;
  h->have_thr6 = have6;
  for (int d = 0; d < 3 && have6; ++d) {
    h->thr6[2 * d] = tok6[2 * d + (h->mode == ETHCNN_MODE_LDP ? 1 : 0)];
    h->thr6[2 * d + 1] = tok6[2 * d + (h->mode == ETHCNN_MODE_LDP ? 0 : 1)];
  }
  return ETHCNN_OK;
}

int ethcnn_predict_yuv_file_from(ethcnn_handle* h, const char* client_dir, const char* yuv_path, int width, int height, int qp,
                                 const char* out_path) {
  if (!h || !client_dir || !yuv_path || !out_path) return fail(ETHCNN_E_ARG, "NULL argument");
  // What the reference script does on EVERY invocation, from the encoder's cwd: read Thr_info.txt (net_CNN.py:47) and restore
  // the checkpoint of the QP range (video_to_cu_depth.py:126-133).  A resident handle must not answer with stale copies.
  const std::string dir = *client_dir ? client_dir : ".";
  int rc = ethcnn_reload_thresholds(h, (dir + "/Thr_info.txt").c_str());
  if (rc) return rc;
  const std::string prefix = model_prefix(h->mode, qp);
  std::string theirs, ours;
  if (!slurp(dir + "/" + prefix + ".index", &theirs)) return fail(ETHCNN_E_IO, "cannot open " + dir + "/" + prefix + ".index");
  {
    std::lock_guard<std::mutex> lock(h->mu);
    bool stale = false;
    for (auto& dc : h->devs) {
      auto it = dc->models.find(prefix);
      if (it != dc->models.end() && it->second.index_bytes != theirs) stale = true;
    }
    if (stale || !slurp(h->model_dir + "/" + prefix + ".index", &ours) || ours != theirs) {
      // the client's weights are not the ones this handle holds (or would load)
      char a[4096], b[4096];
      const bool same_dir = realpath(dir.c_str(), a) && realpath(h->model_dir.c_str(), b) && strcmp(a, b) == 0;
      if (!same_dir)
        return fail(ETHCNN_E_FORMAT, "checkpoint " + prefix + " in " + dir + " differs from the one of the resident handle (" + h->model_dir +
                                         "): run the predictor in-process or start the server in that directory");
      for (auto& dc : h->devs) {   // same directory, file replaced since it was loaded: drop the resident copy, it is re-read below
        auto it = dc->models.find(prefix);
        if (it != dc->models.end()) {
          cudaSetDevice(dc->device);
          cudaDeviceSynchronize();
          free_model(it->second);
          dc->models.erase(it);
        }
      }
    }
  }
  return ethcnn_predict_yuv_file(h, yuv_path, width, height, qp, out_path);
}

int ethcnn_ldp_step(ethcnn_handle* h, const uint8_t* y, int width, int height, int qp, int i_frame, const float* state_in,
                    float* state_out, float* prob) {
  if (!h || !y || !state_out || !prob) return fail(ETHCNN_E_ARG, "NULL argument");
  std::lock_guard<std::mutex> lock(h->mu);
  return run_ldp_step(h, y, width, height, qp, i_frame, state_in, state_out, prob);
}

int ethcnn_ldp_serve(ethcnn_handle* h, const char* dir, int max_frames, int idle_timeout_ms) {
  if (!h) return fail(ETHCNN_E_ARG, "NULL handle");
  // Fail before HM starts waiting on pred_end.sig (its wait is unbounded, TEncGOP.cpp(LDP):1487): the reference daemon dies at
  // import when Thr_info.txt is missing (net_CNN_LSTM_one_step.py:67-68) and at restore when the CNN checkpoint is.
  if (h->mode != ETHCNN_MODE_LDP) return fail(ETHCNN_E_ARG, "the LDP daemon requires ETHCNN_MODE_LDP");
  if (!h->have_thr) return fail(ETHCNN_E_IO, "Thr_info.txt was not found or is malformed: the LDP daemon needs its gate thresholds");
  {
    std::lock_guard<std::mutex> lock(h->mu);
    DeviceCtx& c0 = *h->devs[0];
    if (cudaSetDevice(c0.device) != cudaSuccess) return fail(ETHCNN_E_CUDA, "cannot select the device");
    DeviceModel* m0 = nullptr;
    int rc0 = get_model(h, c0, 32, &m0);   // one CNN checkpoint serves every QP (resi_to_cu_depth_LDP.py:158-159)
    if (rc0) return rc0;
  }
  const std::string d = (dir && *dir) ? std::string(dir) + "/" : std::string();
  const std::string start_file = d + "pred_start.sig", end_file = d + "pred_end.sig", command_file = d + "command.dat";
  const std::string yuv_file = d + "resi.yuv", state_file = d + "state.dat", save_file = d + "cu_depth.dat";
  int served = 0;
  auto last = std::chrono::steady_clock::now();
  std::vector<uint8_t> luma;
  std::vector<float> state_in, state_out, prob;
  while (max_frames <= 0 || served < max_frames) {
    if (!file_exists(start_file)) {
      if (idle_timeout_ms > 0 &&
          std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - last).count() > idle_timeout_ms)
        break;
      std::this_thread::sleep_for(std::chrono::microseconds(200));
      continue;
    }
    // command.dat: "<i_frame> <width> <height> <qp> [end]" (resi_to_cu_depth_LDP.py:54-70); anything else = not ready yet
    int i_frame = -1, w = -1, hgt = -1, qp = -1;
    {
      FILE* f = fopen(command_file.c_str(), "r");
      char tail[16] = {0};
      if (f) {
        char line[256] = {0};
        if (fgets(line, sizeof(line), f) && sscanf(line, "%d %d %d %d %15s", &i_frame, &w, &hgt, &qp, tail) == 5 &&
            strcmp(tail, "[end]") == 0) {
        } else {
          i_frame = -1;
        }
        fclose(f);
      }
    }
    if (i_frame < 0) {
      std::this_thread::sleep_for(std::chrono::microseconds(200));
      continue;
    }
    remove(start_file.c_str());
    if (w <= 0 || hgt <= 0) return fail(ETHCNN_E_ARG, "command.dat carries a bad frame size");
    const size_t rows = size_t((w + kCtu - 1) / kCtu) * ((hgt + kCtu - 1) / kCtu);
    luma.resize(size_t(w) * hgt);
    if (!read_exact(yuv_file, luma.data(), luma.size())) return fail(ETHCNN_E_IO, "cannot read " + yuv_file);
    state_in.assign(rows * 2 * kFc1, 0.f);
    state_out.resize(rows * 2 * kFc1);
    prob.resize(rows * kProbs);
    if (i_frame > 1 && !read_exact(state_file, state_in.data(), state_in.size() * 4))   // :103-112: zeros when i_frame <= 1
      return fail(ETHCNN_E_IO, "cannot read " + state_file);
    int rc = ethcnn_ldp_step(h, luma.data(), w, hgt, qp, i_frame, state_in.data(), state_out.data(), prob.data());
    if (rc) return rc;
    // same order as save_cu_depth_and_state (:131-144): state, probabilities, then the end signal
    if (!write_all(state_file, state_out.data(), state_out.size() * 4) || !write_all(save_file, prob.data(), prob.size() * 4) ||
        !write_all(end_file, nullptr, 0))
      return fail(ETHCNN_E_IO, "cannot write the result files");
    ++served;
    last = std::chrono::steady_clock::now();
  }
  return served;
}

int ethcnn_decisions(ethcnn_handle* h, const float* prob, size_t n_ctus, const float thr6[6], uint8_t* decision) {
  if (!h || !thr6 || ((!prob || !decision) && n_ctus)) return fail(ETHCNN_E_ARG, "NULL argument");
  if (n_ctus == 0) return ETHCNN_OK;
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceCtx& c = *h->devs[0];
  CUDA_TRY(cudaSetDevice(c.device));
  const size_t n = n_ctus * kProbs;
  float *d_p = nullptr, *d_t = nullptr;
  unsigned char* d_d = nullptr;
  CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&d_p), n * 4));
  CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&d_t), 6 * 4));
  CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&d_d), n));
  cudaError_t e = cudaMemcpyAsync(d_p, prob, n * 4, cudaMemcpyHostToDevice, c.s_compute);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_t, thr6, 24, cudaMemcpyHostToDevice, c.s_compute);
  if (e == cudaSuccess) e = launch_decisions(d_p, d_d, (long long)n, d_t, c.s_compute);
  ++h->launches;
  if (e == cudaSuccess) e = cudaMemcpyAsync(decision, d_d, n, cudaMemcpyDeviceToHost, c.s_compute);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c.s_compute);
  cudaFree(d_p), cudaFree(d_t), cudaFree(d_d);
  if (e != cudaSuccess) return fail(ETHCNN_E_CUDA, std::string("decisions: ") + cudaGetErrorString(e));
  return ETHCNN_OK;
}

int ethcnn_query(ethcnn_handle* h, int what, int64_t* value) {
  if (!h || !value) return fail(ETHCNN_E_ARG, "NULL argument");
  switch (what) {
    case ETHCNN_Q_KERNEL_LAUNCHES: *value = h->launches.load(); break;
    case ETHCNN_Q_N_DEVICES: *value = int64_t(h->devs.size()); break;
    case ETHCNN_Q_FC1_PATH: *value = h->fc1_path; break;
    case ETHCNN_Q_TMA_LOADER_USED: *value = h->devs[0]->last_used_tma; break;
    case ETHCNN_Q_SM_COUNT: *value = h->devs[0]->sm_count; break;
    case ETHCNN_Q_CONV_PATH: *value = h->conv_path; break;
    default: return fail(ETHCNN_E_ARG, "unknown query");
  }
  return ETHCNN_OK;
}

int ethcnn_set_option(ethcnn_handle* h, int option, int64_t value) {
  if (!h) return fail(ETHCNN_E_ARG, "NULL handle");
  std::lock_guard<std::mutex> lock(h->mu);
  switch (option) {
    case ETHCNN_OPT_FC1_PATH:
      if (value < 0 || value > 3) return fail(ETHCNN_E_ARG, "FC path must be 0, 1, 2 or 3");
      h->fc1_path = int(value);
      break;
    case ETHCNN_OPT_CHUNK_CTUS:
      if (value < 1 || value > (1 << 22)) return fail(ETHCNN_E_ARG, "chunk size out of range");
      h->chunk_ctus = size_t(value);
      break;
    case ETHCNN_OPT_CONV_PATH:
      if (value < 0 || value > 1) return fail(ETHCNN_E_ARG, "conv path must be 0 or 1");
      h->conv_path = int(value);
      break;
    case ETHCNN_OPT_STAGED_OUTPUT:
      h->staged_output = value != 0;
      break;
    default: return fail(ETHCNN_E_ARG, "unknown option");
  }
  return ETHCNN_OK;
}

// ---- gather buffers in peer memory (one process per GPU; the rows travel over NVLink as the kernels store them)
int ethcnn_peer_buffer_create(ethcnn_handle* h, size_t bytes, void** d_ptr, uint8_t handle_out[ETHCNN_IPC_HANDLE_BYTES]) {
  if (!h || !d_ptr || !handle_out || bytes == 0) return fail(ETHCNN_E_ARG, "NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == ETHCNN_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceCtx& c = *h->devs[0];
  CUDA_TRY(cudaSetDevice(c.device));
  void* p = nullptr;
  CUDA_TRY(cudaMalloc(&p, bytes));   // its own allocation: an IPC handle always names a whole cudaMalloc block
  cudaIpcMemHandle_t mh;
  cudaError_t e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&mh, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(ETHCNN_E_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
  }
  memcpy(handle_out, &mh, sizeof(mh));
  c.peer_owned.push_back(p);
  *d_ptr = p;
  return ETHCNN_OK;
}

int ethcnn_peer_buffer_open(ethcnn_handle* h, const uint8_t handle[ETHCNN_IPC_HANDLE_BYTES], void** d_ptr) {
  if (!h || !d_ptr || !handle) return fail(ETHCNN_E_ARG, "NULL argument");
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceCtx& c = *h->devs[0];
  CUDA_TRY(cudaSetDevice(c.device));
  cudaIpcMemHandle_t mh;
  memcpy(&mh, handle, sizeof(mh));
  void* p = nullptr;
  // maps the exporter's allocation into this process and enables peer access between the two devices
  cudaError_t e = cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(ETHCNN_E_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
  }
  c.peer_opened.push_back(p);
  *d_ptr = p;
  return ETHCNN_OK;
}

int ethcnn_peer_buffer_release(ethcnn_handle* h, void* d_ptr) {
  if (!h || !d_ptr) return fail(ETHCNN_E_ARG, "NULL argument");
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceCtx& c = *h->devs[0];
  CUDA_TRY(cudaSetDevice(c.device));
  CUDA_TRY(cudaDeviceSynchronize());
  for (size_t i = 0; i < c.peer_opened.size(); ++i)
    if (c.peer_opened[i] == d_ptr) {
      c.peer_opened.erase(c.peer_opened.begin() + i);
      CUDA_TRY(cudaIpcCloseMemHandle(d_ptr));
      return ETHCNN_OK;
    }
  for (size_t i = 0; i < c.peer_owned.size(); ++i)
    if (c.peer_owned[i] == d_ptr) {
      c.peer_owned.erase(c.peer_owned.begin() + i);
      CUDA_TRY(cudaFree(d_ptr));
      return ETHCNN_OK;
    }
  return fail(ETHCNN_E_ARG, "not a peer buffer of this handle");
}

int ethcnn_profile_enable(ethcnn_handle* h, int on) {
  if (!h) return fail(ETHCNN_E_ARG, "NULL handle");
  std::lock_guard<std::mutex> lock(h->mu);
  for (auto& c : h->devs) {
    c->profiling = on != 0;
    if (on && c->ev_pool.size() < 64) {  // pre-create the timing events outside any timed region
      CUDA_TRY(cudaSetDevice(c->device));
      while (c->ev_pool.size() < 64) {
        cudaEvent_t e = nullptr;
        CUDA_TRY(cudaEventCreate(&e));
        c->ev_pool.push_back(e);
      }
    }
  }
  return ETHCNN_OK;
}

int ethcnn_profile_read(ethcnn_handle* h, int stage, double* ms_total, int64_t* launches, int reset) {
  if (!h || stage < 0 || stage >= ETHCNN_N_STAGES) return fail(ETHCNN_E_ARG, "bad stage");
  std::lock_guard<std::mutex> lock(h->mu);
  double ms = 0;
  int64_t n = 0;
  for (auto& cp : h->devs) {
    DeviceCtx& c = *cp;
    CUDA_TRY(cudaSetDevice(c.device));
    for (auto& e : c.prof) {  // fold finished event pairs into the per-stage totals
      CUDA_TRY(cudaEventSynchronize(e.b));
      float t = 0.f;
      CUDA_TRY(cudaEventElapsedTime(&t, e.a, e.b));
      c.prof_ms[e.stage] += t;
      c.prof_launches[e.stage] += 1;
      c.ev_pool.push_back(e.a), c.ev_pool.push_back(e.b);
    }
    c.prof.clear();
    ms += c.prof_ms[stage];
    n += c.prof_launches[stage];
    if (reset) c.prof_ms[stage] = 0, c.prof_launches[stage] = 0;
  }
  if (ms_total) *ms_total = ms;
  if (launches) *launches = n;
  return ETHCNN_OK;
}

int ethcnn_debug_pack_model(const char* ckpt_prefix, float input_bound, float* conv, float* w1, float* b1, uint16_t* w1_hi,
                            uint16_t* w1_lo, int exps[16], float* feat_bound) {
  if (!ckpt_prefix) return fail(ETHCNN_E_ARG, "NULL argument");
  std::map<std::string, BundleTensor> tensors;
  std::string err;
  if (!read_tf_bundle(ckpt_prefix, &tensors, &err))
    return fail(err.find("cannot open") != std::string::npos ? ETHCNN_E_IO : ETHCNN_E_FORMAT, err);
  PackedModel pm;
  if (!pack_model(tensors, input_bound, &pm, &err)) return fail(ETHCNN_E_FORMAT, err);
  if (conv) memcpy(conv, pm.conv.data(), pm.conv.size() * 4);
  if (w1) memcpy(w1, pm.w1.data(), pm.w1.size() * 4);
  if (b1) memcpy(b1, pm.b1.data(), pm.b1.size() * 4);
  if (w1_hi) memcpy(w1_hi, pm.w1_hi.data(), pm.w1_hi.size() * 2);
  if (w1_lo) memcpy(w1_lo, pm.w1_lo.data(), pm.w1_lo.size() * 2);
  if (exps) {
    exps[0] = pm.feat_exp, exps[1] = pm.w_exp, exps[2] = pm.a1_exp, exps[3] = pm.w2_exp;
    for (int br = 0; br < 3; ++br)
      for (int k = 0; k < 4; ++k) exps[4 + 4 * br + k] = pm.conv_exp[br][k];
  }
  if (feat_bound) *feat_bound = pm.feat_bound;
  return ETHCNN_OK;
}

int ethcnn_debug_pack_conv_tc(const char* ckpt_prefix, float input_bound, uint8_t* blob) {
  if (!ckpt_prefix || !blob) return fail(ETHCNN_E_ARG, "NULL argument");
  std::map<std::string, BundleTensor> tensors;
  std::string err;
  if (!read_tf_bundle(ckpt_prefix, &tensors, &err))
    return fail(err.find("cannot open") != std::string::npos ? ETHCNN_E_IO : ETHCNN_E_FORMAT, err);
  PackedModel pm;
  if (!pack_model(tensors, input_bound, &pm, &err)) return fail(ETHCNN_E_FORMAT, err);
  memcpy(blob, pm.conv_tc.data(), pm.conv_tc.size());
  return ETHCNN_OK;
}

int ethcnn_debug_read_thresholds(const char* thr_path, float thr[2]) {
  if (!thr_path || !thr) return fail(ETHCNN_E_ARG, "NULL argument");
  return read_thresholds(thr_path, &thr[0], &thr[1]);
}

uint16_t ethcnn_debug_f32_to_f16(float v) { return f32_to_f16_bits(v); }

int ethcnn_debug_read_scratch(ethcnn_handle* h, int what, size_t n_ctus, float* out) {
  if (!h || !out) return fail(ETHCNN_E_ARG, "NULL argument");
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceCtx& c = *h->devs[0];
  if (!c.feat_hi || n_ctus > c.chunk_ctus) return fail(ETHCNN_E_ARG, "no scratch of that size");
  CUDA_TRY(cudaSetDevice(c.device));
  CUDA_TRY(cudaDeviceSynchronize());
  if (what == 1) {
    CUDA_TRY(cudaMemcpy(out, c.fc1, n_ctus * kFc1 * 4, cudaMemcpyDeviceToHost));
    return ETHCNN_OK;
  }
  if (what != 0) return fail(ETHCNN_E_ARG, "unknown scratch id");
  std::vector<uint16_t> hi(n_ctus * kFeat), lo(n_ctus * kFeat);
  CUDA_TRY(cudaMemcpy(hi.data(), c.feat_hi, hi.size() * 2, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(lo.data(), c.feat_lo, lo.size() * 2, cudaMemcpyDeviceToHost));
  const float inv = std::ldexp(1.0f, -c.last_feat_exp);
  for (size_t i = 0; i < hi.size(); ++i) out[i] = (f16_bits_to_f32(hi[i]) + f16_bits_to_f32(lo[i])) * inv;
  return ETHCNN_OK;
}

void* ethcnn_alloc_pinned(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
    fail(ETHCNN_E_NOMEM, "cudaMallocHost failed");
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

void ethcnn_free_pinned(void* p) {
  if (p) cudaFreeHost(p);
}

}  // extern "C"
