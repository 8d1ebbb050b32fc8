/*
 * ethcnn.h -- C ABI of libethcnn_b200.so, the B200-native ETH-CNN CU-partition predictor.
 *
 * The reference (tianyili2017/HEVC-Complexity-Reduction) has no C API for this path: the patched HM
 * encoder shells out to a Python script and reads a file back.  The entry points below are what a
 * binding for that path would call; each cites the reference interface it replaces (paths relative
 * to the reference root).  All functions are blocking unless stated, return 0 on success and a
 * negative ETHCNN_E_* code on failure; ethcnn_last_error() returns the message of the last failure
 * on the calling thread.  No CPU fallback exists: without a CUDA device every compute entry point
 * fails with ETHCNN_E_CUDA.
 */
#ifndef ETHCNN_H_
#define ETHCNN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ETHCNN_ABI_VERSION 4 /* 2: staged output option + peer gather buffers; 3: resident server; 4: decision map */

/* Which network/deployment is evaluated. */
#define ETHCNN_MODE_AI  0 /* HM-16.5_Test_AI/bin/net_CNN.py:103-195 (x/255, qp/51, batch-level gates)      */
#define ETHCNN_MODE_LDP 1 /* ETH-CNN_Training_LDP/net_CTU64.py:100-179 ((x-128)/255*10, qp/51*0.18, no gates) */

#define ETHCNN_OK            0
#define ETHCNN_E_ARG        -1 /* bad argument (python: AssertionError at video_to_cu_depth.py:120,137) */
#define ETHCNN_E_IO         -2 /* file missing / unreadable / short write                              */
#define ETHCNN_E_FORMAT     -3 /* checkpoint or Thr_info.txt malformed (python: saver.restore raises)   */
#define ETHCNN_E_CUDA       -4 /* CUDA runtime/driver failure or no device                             */
#define ETHCNN_E_NOMEM      -5

#define ETHCNN_PROBS_PER_CTU 21  /* video_to_cu_depth.py:63: y64(1) | y32(4) | y16(16) */
#define ETHCNN_FC1_WIDTH     448 /* HM-16.5_Test_LDP/bin/net_CNN_LSTM_one_step.py:187-199 */

typedef struct ethcnn_handle ethcnn_handle;

/*
 * Replaces the module-level set-up of video_to_cu_depth.py:14-29 and net_CNN.py:38-47: opens the
 * device(s), reads Thr_info.txt (tokens [1] and [3], split on single spaces) and remembers where the
 * TF-Saver-V2 checkpoints live.  Checkpoints are parsed, packed and uploaded on first use of their
 * QP range (video_to_cu_depth.py:126-133) and stay resident.
 *   model_dir  directory holding model_2000000_qp{20~25,25~30,30~35,35~40}.dat.{index,data-00000-of-00001}
 *              (AI) or model_LDP_2000000_qp22~37.dat.* (LDP); the reference uses the cwd, i.e. ".".
 *   thr_path   path of Thr_info.txt; NULL = <model_dir>/Thr_info.txt.  Ignored for ETHCNN_MODE_LDP.
 *   n_gpus     >= 1: use CUDA devices 0..n_gpus-1, frames are sharded across them in contiguous
 *              ranges and gathered into the caller's buffer (single process, HM forks one child).
 */
int ethcnn_create(const char* model_dir, const char* thr_path, int mode, int n_gpus, ethcnn_handle** out);

/* As ethcnn_create with n_gpus = 1 but on an explicit CUDA device: the one-process-per-GPU form used
 * under torchrun / torch.distributed (rank r opens device LOCAL_RANK). */
int ethcnn_create_on_device(const char* model_dir, const char* thr_path, int mode, int device, ethcnn_handle** out);

void ethcnn_destroy(ethcnn_handle* h);

/*
 * Replaces the whole script run `python video_to_cu_depth.py <yuv> <W> <H> <QP>`
 * (HM-16.5_Test_AI/source/App/TAppEncoder/TAppEncCfg.cpp:2317-2321; video_to_cu_depth.py:120-145):
 * reads every frame of the 8-bit 4:2:0 planar file (luma only), and writes out_path ("cu_depth.dat"
 * in the reference) as n_frames * ceil(W/64) * ceil(H/64) * 21 little-endian float32, frame-major, CTU
 * raster order, no header (consumer: TLibEncoder/TEncCu.cpp:257-258,434-447).  Fails with ETHCNN_E_ARG
 * when file_bytes % (W*H*3/2) != 0 (video_to_cu_depth.py:137).  The output is written to a temporary
 * file and renamed, so a failure never leaves a truncated cu_depth.dat.
 */
int ethcnn_predict_yuv_file(ethcnn_handle* h, const char* yuv_path, int width, int height, int qp,
                            const char* out_path);

/*
 * Replaces get_prob() (video_to_cu_depth.py:75-118) for luma already in HOST memory.
 *   y            first luma sample of frame 0; rows are `width` bytes, contiguous
 *   frame_stride bytes between the first luma samples of consecutive frames (W*H*3/2 for a YUV 4:2:0
 *                buffer, W*H for luma-only)
 *   out          n_frames * nCTU * 21 floats (host)
 * Pinned (page-locked) buffers are copied directly; pageable ones are staged through the library's
 * own pinned ring.  Host<->device copies are pipelined against the kernels.
 */
int ethcnn_predict_luma(ethcnn_handle* h, const uint8_t* y, int width, int height, size_t frame_stride,
                        int n_frames, int qp, float* out);

/*
 * Same computation with DEVICE pointers on the handle's (first) device, enqueued on `stream`
 * (a cudaStream_t passed as void*; NULL = the legacy default stream) and NOT synchronised: the
 * kernel-only path used for roofline measurement and by callers that keep frames resident.
 *   pitch        bytes between rows (>= width).  pitch % 16 == 0, frame_stride % 16 == 0 and a 16-byte
 *                aligned d_y select the TMA tile loader; anything else falls back to a plain-load
 *                variant of the same kernel (still on the GPU).
 *   d_out        n_frames * nCTU * 21 floats (device)
 */
int ethcnn_predict_luma_device(ethcnn_handle* h, const uint8_t* d_y, int width, int height, size_t pitch,
                               size_t frame_stride, int n_frames, int qp, float* d_out, void* stream);

/*
 * LDP deployment tap: the 448-vector [fc1_64 | fc1_32 | fc1_16] that resi_cnn() hands to the LSTM
 * (HM-16.5_Test_LDP/bin/net_CNN_LSTM_one_step.py:151-199).  Host pointers; out has
 * n_frames * nCTU * 448 floats.  Requires ETHCNN_MODE_LDP.
 */
int ethcnn_export_fc1(ethcnn_handle* h, const uint8_t* y, int width, int height, size_t frame_stride,
                      int n_frames, float* out);

/*
 * The deployed inter-mode (LDP) predictor for ONE frame: residual ETH-CNN -> one step of the three ETH-LSTM cells
 * -> FC2 / FC3 with the five extra features [qp/51*0.18, one_hot(i_frame % 4)] -> gates.  Replaces
 * predict_cu_depth() of HM-16.5_Test_LDP/bin/resi_to_cu_depth_LDP.py:114-129 (graph: net_CNN_LSTM_one_step.py:266-323).
 *   y          luma of the residue frame (resi.yuv: resi + 128 clipped), width * height bytes, host
 *   state_in   nCTU * 2 * 448 floats ([c | h] per CTU, heads 64|128|256 side by side) or NULL for zeros
 *   state_out  same shape; prob nCTU * 21 floats.  LSTM checkpoints model_LDP_200000_qp{22,27,32,37}.dat are chosen
 *              by QP range (resi_to_cu_depth_LDP.py:169-177) from model_dir; Thr_info.txt supplies the gate thresholds.
 * Requires ETHCNN_MODE_LDP.
 */
int ethcnn_ldp_step(ethcnn_handle* h, const uint8_t* y, int width, int height, int qp, int i_frame,
                    const float* state_in, float* state_out, float* prob);

/*
 * The file-signal daemon of the LDP encoder (README.md:64-84; HM side TEncGOP.cpp(LDP):1471-1505; Python side
 * resi_to_cu_depth_LDP.py:146-187): polls <dir>/pred_start.sig, reads command.dat ("<frame> <W> <H> <QP> [end]"),
 * resi.yuv and (for frame > 1) state.dat, runs ethcnn_ldp_step, writes state.dat, cu_depth.dat, then pred_end.sig.
 * Weights stay on the device between frames.  Returns the number of frames served (>= 0) once max_frames (> 0) have
 * been served or nothing arrived for idle_timeout_ms (> 0); with both <= 0 it serves forever like the reference.
 */
int ethcnn_ldp_serve(ethcnn_handle* h, const char* dir, int max_frames, int idle_timeout_ms);

/*
 * Resident server for the All-Intra drop-in.  HM starts the predictor afresh for every encode (TAppEncCfg.cpp:2317-2321) and a
 * fresh process pays 2 - 4 s of CUDA initialisation before ~0.1 s of work; the reference solves the same start-up problem of
 * TensorFlow with a resident daemon for its inter-mode path (README.md:64-84).  ethcnn_serve keeps the handle (context, packed
 * weights, scratch) alive and answers requests on a Unix-domain stream socket until max_requests (> 0) PREDICT requests have
 * been served, nothing arrived for idle_timeout_ms (> 0) or a client asked it to quit; returns the number of requests served
 * (>= 0) or a negative code.  Every request re-reads the CLIENT directory's Thr_info.txt and checks its checkpoint against the
 * resident one (ethcnn_predict_yuv_file_from below); relative paths are resolved against the client's working directory.  The
 * socket is created with mode 0600 and only peers with the server's uid (or root) are served; a second server on a socket that a
 * live server answers refuses to start (ETHCNN_E_IO).  Protocol: csrc/serve.cpp.
 * ethcnn_request is the client (no handle, no CUDA): returns the server's code for ethcnn_predict_yuv_file, or ETHCNN_E_IO when
 * nobody listens on socket_path (the CLI and the Python shim then work in-process); ethcnn_request_error() has the message.
 */
int ethcnn_serve(ethcnn_handle* h, const char* socket_path, int max_requests, int idle_timeout_ms);

/*
 * What a RESIDENT handle must redo per request to behave like a fresh run of the reference script from `client_dir` (the
 * encoder's cwd): re-read <client_dir>/Thr_info.txt (net_CNN.py:47 reads it at every invocation, and HM reads the same file
 * itself, TEncCu.cpp:250) and make sure the checkpoint of the QP range in client_dir (video_to_cu_depth.py:126-133) is the one
 * the handle holds: the .index files (which carry the crc32c of every tensor) must be identical, otherwise the call fails with
 * ETHCNN_E_FORMAT -- or, when client_dir IS the handle's model directory and the file was replaced, the resident copy is dropped
 * and re-read.  Then ethcnn_predict_yuv_file.  ethcnn_serve answers every request through this.
 * ethcnn_reload_thresholds re-reads a Thr_info.txt (NULL = <model_dir>/Thr_info.txt) into the handle.
 */
int ethcnn_predict_yuv_file_from(ethcnn_handle* h, const char* client_dir, const char* yuv_path, int width, int height, int qp,
                                 const char* out_path);
int ethcnn_reload_thresholds(ethcnn_handle* h, const char* thr_path);
int ethcnn_request(const char* socket_path, const char* yuv_path, int width, int height, int qp, const char* out_path);
int ethcnn_request_quit(const char* socket_path);
const char* ethcnn_request_error(void);

/*
 * Decision map: HM's threshold rule (TLibEncoder/TEncCu.cpp:434-462) applied ON THE DEVICE by the
 * gate kernel, next to the float rows.  One 64-bit word per CTU: bits [2k, 2k+1] hold the decision for entry k of the CTU's
 * cu_depth.dat row (k = 0: the 64x64 CU; 1 + x/32 + 2 (y/32): 32x32; 5 + x/16 + 4 (y/16): 16x16 -- the indices HM computes at
 * TEncCu.cpp:434-447): 2 = split only (p > up), 0 = no split (p <= down), 1 = check both; bits 42..63 are zero.  8 bytes per
 * CTU instead of 84.  The thresholds are the six numbers of Thr_info.txt read at create (AI file order up,down per depth,
 * TEncCu.cpp:250; LDP file order down,up, TEncGOP.cpp(LDP):1449) unless replaced with ethcnn_set_decision_thresholds
 * (thr6 = up0, down0, up1, down1, up2, down2).  The float rows are produced exactly as by the calls without _map.
 *   ethcnn_predict_luma_map          host pointers (out: n*21 floats, map: n words)
 *   ethcnn_predict_luma_device_map   device pointers, enqueued on `stream`, not synchronised
 */
int ethcnn_predict_luma_map(ethcnn_handle* h, const uint8_t* y, int width, int height, size_t frame_stride, int n_frames, int qp,
                            float* out, uint64_t* map);
int ethcnn_predict_luma_device_map(ethcnn_handle* h, const uint8_t* d_y, int width, int height, size_t pitch, size_t frame_stride,
                                   int n_frames, int qp, float* d_out, uint64_t* d_map, void* stream);
int ethcnn_set_decision_thresholds(ethcnn_handle* h, const float thr6[6]);
int ethcnn_get_decision_thresholds(ethcnn_handle* h, float thr6[6]);

/* HM's use of a probability (TLibEncoder/TEncCu.cpp:448-462): 2 = split only (p > up), 0 = no split
 * (p <= down), 1 = check both.  thr6 = the six numbers of Thr_info.txt (up,down per depth,
 * TEncCu.cpp:250).  Runs on the device; prob/decision are HOST arrays of n_ctus*21 entries. */
int ethcnn_decisions(ethcnn_handle* h, const float* prob, size_t n_ctus, const float thr6[6], uint8_t* decision);

/* Introspection (for bench.py and tests). */
#define ETHCNN_Q_KERNEL_LAUNCHES   1 /* kernels launched by this handle since creation            */
#define ETHCNN_Q_N_DEVICES         2
#define ETHCNN_Q_FC1_PATH          3 /* dense path, see ETHCNN_OPT_FC1_PATH                           */
#define ETHCNN_Q_TMA_LOADER_USED   4 /* 1 if the last device call used the TMA tile loader         */
#define ETHCNN_Q_SM_COUNT          5
#define ETHCNN_Q_CONV_PATH         6 /* conv stage, see ETHCNN_OPT_CONV_PATH                        */
int ethcnn_query(ethcnn_handle* h, int what, int64_t* value);

/* Per-kernel device timing: when enabled, every launch is bracketed by CUDA events on the launching
 * stream; ethcnn_profile_read() synchronises and returns, for stage s (ETHCNN_STAGE_*), the summed
 * milliseconds and launch count since the last reset. */
#define ETHCNN_STAGE_CONV   0 /* luma tiles -> 2688 features (conv stack)   */
#define ETHCNN_STAGE_FC1    1 /* 2688 -> 448 dense contraction              */
#define ETHCNN_STAGE_HEADS  2 /* FC2 + FC3 + sigmoid                        */
#define ETHCNN_STAGE_GATE   3 /* per-sub-batch gates                        */
#define ETHCNN_N_STAGES     4
int ethcnn_profile_enable(ethcnn_handle* h, int on);
int ethcnn_profile_read(ethcnn_handle* h, int stage, double* ms_total, int64_t* launches, int reset);

/* Tuning knobs (before the first predict call): see ETHCNN_OPT_*. */
#define ETHCNN_OPT_FC1_PATH    1 /* 0 = SIMT fp32 FC1 + heads kernel, 1 = tcgen05 FC1 + heads kernel,
                                    2 = fused tcgen05 FC1+FC2+FC3, one CTA per 128 CTUs,
                                    3 = the fused kernel on CTA pairs (tcgen05 cta_group::2, 256 CTUs per pair) */
#define ETHCNN_OPT_CHUNK_CTUS  2 /* CTUs per feature-buffer chunk                    */
#define ETHCNN_OPT_STAGED_OUTPUT 3 /* ethcnn_predict_luma_device only. 1 = the dense kernel writes its raw probabilities to a
                                    library-owned local buffer and the gate kernel copies the finished rows to d_out with
                                    coalesced stores: set it when d_out is PEER memory (ethcnn_peer_buffer_open), so that the
                                    rows cross NVLink as full 128-byte transactions.  0 (default) = rows are written in place */
#define ETHCNN_OPT_CONV_PATH   4 /* 0 = conv stage on mma.sync (register-chained fragments), 1 = conv stage on tcgen05 with the
                                    activations chained through tensor memory (needs the TMA tile loader; falls back to 0) */
int ethcnn_set_option(ethcnn_handle* h, int option, int64_t value);

/*
 * Multi-GPU gather without a collective (one process per GPU).  The reference serialises all
 * probabilities in one process (video_to_cu_depth.py:114-116); when frames are sharded over ranks the per-rank rows have
 * to reach rank 0.  Instead of a gather AFTER the kernels, rank 0 exports one buffer for the whole sequence, every other
 * rank maps it (CUDA IPC; peer access over NVLink / NVSwitch is enabled by the mapping) and passes
 * `peer + first_row_of_rank * 21` as d_out of ethcnn_predict_luma_device with ETHCNN_OPT_STAGED_OUTPUT = 1: the gate
 * kernel's stores ARE the transfer.  The rows are complete on rank 0 once every rank has synchronised its stream and
 * the ranks have met at a barrier.  The 64-byte handle travels through whatever the ranks share (torch.distributed
 * broadcast in sharding.py).
 *   ethcnn_peer_buffer_create   cudaMalloc on the handle's device + export; the buffer is zero-filled
 *   ethcnn_peer_buffer_open     map another process's buffer; fails with ETHCNN_E_CUDA when the devices cannot reach each
 *                               other (the caller then falls back to an NCCL gather)
 *   ethcnn_peer_buffer_release  unmap (opened) or free (created); ethcnn_destroy releases what is left
 */
#define ETHCNN_IPC_HANDLE_BYTES 64
int ethcnn_peer_buffer_create(ethcnn_handle* h, size_t bytes, void** d_ptr, uint8_t handle_out[ETHCNN_IPC_HANDLE_BYTES]);
int ethcnn_peer_buffer_open(ethcnn_handle* h, const uint8_t handle[ETHCNN_IPC_HANDLE_BYTES], void** d_ptr);
int ethcnn_peer_buffer_release(ethcnn_handle* h, void* d_ptr);

/* Pinned host memory helpers (so foreign-language callers can hand over page-locked buffers). */
void* ethcnn_alloc_pinned(size_t bytes);
void ethcnn_free_pinned(void* p);

/* Host-side testing hooks (no GPU needed): the checkpoint reader + weight packer and the Thr_info.txt
 * parser, exposed so the CPU test-suite can check them against the oracle.
 *   conv     3 * 4976 32-bit words (branch S, M, L blocks in the shared-memory layout of csrc/kernels.h:
 *            header, biases, filters as mma.sync B fragments in fp16 hi/lo)
 *   w1       2688 * 448 floats (heads 64 | 32 | 16 side by side), b1 448 floats
 *   w1_hi/lo 448 * 2688 fp16 bit patterns of w1 * 2^exps[1] (K-major)
 *   exps     16 ints: feat_exp, w_exp, a1_exp, w2_exp, then per branch (e1w, e_c1, e2w, e3w)
 * Any output pointer may be NULL. */
int ethcnn_debug_pack_model(const char* ckpt_prefix, float input_bound, float* conv, float* w1, float* b1,
                            uint16_t* w1_hi, uint16_t* w1_lo, int exps[16], float* feat_bound);
/* The conv filters as the tcgen05 conv stage consumes them (csrc/conv_tc.h): 3 branches (S, M, L) x 29696 bytes of
 * 128-byte-swizzled K-major fp16 hi / lo tiles + the pre-scaled bias tables. */
#define ETHCNN_CONV_TC_BLOB_BYTES (3 * 29696)
int ethcnn_debug_pack_conv_tc(const char* ckpt_prefix, float input_bound, uint8_t* blob);
int ethcnn_debug_read_thresholds(const char* thr_path, float thr[2]);
uint16_t ethcnn_debug_f32_to_f16(float v);
/* Device-side testing hook: copy back intermediates of the LAST chunk processed on device 0.
 * what = 0: the 2688 conv features per CTU (hi + lo halves recombined and unscaled), out[n_ctus*2688];
 * what = 1: the FC1 activations, out[n_ctus*448].  n_ctus must not exceed the chunk size. */
int ethcnn_debug_read_scratch(ethcnn_handle* h, int what, size_t n_ctus, float* out);

const char* ethcnn_last_error(void);
int ethcnn_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ETHCNN_H_ */
