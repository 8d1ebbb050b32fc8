"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the reference's ETH-CNN CU-partition path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module; the product (hevc-complexity-reduction_b200/) never does and has no CPU
fallback.

Parity pin: the reference's arithmetic lives in TensorFlow 1.x (un-vendored third party,
README.md:42; checkpoints written by TF 1.4.1) which is not installable here, and the reference
ships no golden vectors for this path.  The restatement is pinned instead by
  (1) executing the reference's OWN UNMODIFIED net_CNN.py / video_to_cu_depth.py /
      ETH-CNN_Training_LDP/net_CTU64.py graph code from /root/reference on top of a numpy
      stand-in for the handful of TF ops they call (oracle/tf_shim), and committing the
      resulting vectors as tests/golden/*.npz (generator: oracle/make_golden.py);
  (2) the reference's own accuracy logs on its 15 000 labelled CTUs
      (ETH-CNN_Training_AI/Models/loss_accuracy_list_*.dat:1002), a statistical pin;
  (3) the unmodified prebuilt HM encoder consuming the emitted cu_depth.dat (tests/test_hm_e2e.py).
Because (1) still relies on restated TF op semantics, the header says it plainly:
PARITY IS PINNED TO THE REFERENCE'S GRAPH CODE, NOT TO TENSORFLOW'S OWN KERNELS.

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from . import tf_bundle

IMAGE_SIZE = 64           # HM-16.5_Test_AI/bin/net_CNN.py:8
SUB_BATCH = 1024          # HM-16.5_Test_AI/bin/video_to_cu_depth.py:64
N_FEATURES = 2688         # net_CNN.py:27
N_OUT = 21                # video_to_cu_depth.py:63 (1 + 4 + 16)
LEAKY_ALPHA = 0.2         # tf.nn.leaky_relu default; `.meta` Const LeakyRelu/alpha = 0x3e4ccccd

MODE_AI = 0
MODE_LDP = 1

# Variable creation order in net_CNN.py:126-141 (L first, then M, then S), unnamed tf.Variables.
BRANCH_VARS = {"L": 0, "M": 6, "S": 12}
HEADS = (("64", 64, 48, 1), ("32", 128, 96, 4), ("16", 256, 192, 16))  # net_CNN.py:29-36,156-185


def _vname(i: int) -> str:
    return "Variable" if i == 0 else "Variable_%d" % i


# ----------------------------------------------------------------------------- model files
def ai_model_prefix(qp: int) -> str:
    """Model selection by QP range, HM-16.5_Test_AI/bin/video_to_cu_depth.py:126-133."""
    if qp < 25:
        return "model_2000000_qp20~25.dat"
    elif qp < 30:
        return "model_2000000_qp25~30.dat"
    elif qp < 35:
        return "model_2000000_qp30~35.dat"
    return "model_2000000_qp35~40.dat"


LDP_MODEL_PREFIX = "model_LDP_2000000_qp22~37.dat"  # HM-16.5_Test_LDP/bin/resi_to_cu_depth_LDP.py:158-159


def get_thresholds(thr_file: str) -> Tuple[float, float]:
    """net_CNN.py:38-45 -- tokens [1] and [3] of the first line split on single spaces."""
    with open(thr_file, "r") as f:
        line = f.readline()
    str_arr = line.split(" ")
    return float(str_arr[1]), float(str_arr[3])


def load_weights(prefix: str) -> Dict[str, np.ndarray]:
    return tf_bundle.read_bundle(prefix)


def random_weights(seed: int, scale: float = 1.0) -> Dict[str, np.ndarray]:
    """Synthetic checkpoint with the reference's 36-tensor layout.
    Fan-in scaled normal weights so activations stay O(1) and probabilities spread over (0,1)."""
    rng = np.random.default_rng(seed)
    w: Dict[str, np.ndarray] = {}
    for base in BRANCH_VARS.values():
        for li, shape in enumerate(((4, 4, 1, 16), (2, 2, 16, 24), (2, 2, 24, 32))):
            fan_in = shape[0] * shape[1] * shape[2]
            w[_vname(base + 2 * li)] = (rng.standard_normal(shape) * scale * (1.6 / math.sqrt(fan_in))).astype(np.float32)
            w[_vname(base + 2 * li + 1)] = (rng.standard_normal(shape[3]) * 0.05).astype(np.float32)
    for h, n1, n2, n3 in HEADS:
        for nm, (ni, no) in (("h_fc1__%s__" % h, (N_FEATURES, n1)), ("h_fc2__%s__" % h, (n1 + 1, n2)),
                             ("y_conv_flat__%s__" % h, (n2 + 1, n3))):
            w[nm + "w"] = (rng.standard_normal((ni, no)) * scale * (1.3 / math.sqrt(ni))).astype(np.float32)
            w[nm + "b"] = (rng.standard_normal(no) * 0.05).astype(np.float32)
    return w


# ----------------------------------------------------------------------------- network pieces
def _leaky(x):
    # tf.nn.leaky_relu == Maximum(alpha * x, x)  (graph in model_2000000_qp30~35.dat.meta)
    return np.maximum(x.dtype.type(LEAKY_ALPHA) * x, x)


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def aver_pool(x: np.ndarray, k: int) -> np.ndarray:
    """net_CNN.py:62-63: avg_pool ksize=stride=k, SAME (64 divides evenly, so no edge effects).
    x: [B, H, W]."""
    b, h, w = x.shape
    return x.reshape(b, h // k, k, w // k, k).mean(axis=(2, 4), dtype=x.dtype)


def zero_mean_norm_local(x: np.ndarray, kernel_width: int = 16) -> np.ndarray:
    """net_CNN.py:78-84: VALID conv with a constant 1/k^2 kernel, stride k; nearest-neighbour
    expansion (align_corners=false => output pixel i takes source i // k); subtract."""
    b, h, w = x.shape
    k = kernel_width
    wn = x.dtype.type(1.0 / (k * k))
    m = (x.reshape(b, h // k, k, w // k, k) * wn).sum(axis=(2, 4), dtype=x.dtype)
    return x - np.repeat(np.repeat(m, k, axis=1), k, axis=2)


def non_overlap_conv(x: np.ndarray, w: np.ndarray, bias: np.ndarray) -> np.ndarray:
    """net_CNN.py:86-92: VALID conv, kernel = stride = k, + bias, leaky_relu.
    x: [B, H, W, Cin]; w: [k, k, Cin, Cout] (TF HWIO)."""
    k = w.shape[0]
    b, h, ww, cin = x.shape
    patches = x.reshape(b, h // k, k, ww // k, k, cin).transpose(0, 1, 3, 2, 4, 5).reshape(b, h // k, ww // k, k * k * cin)
    out = patches @ w.reshape(k * k * cin, -1).astype(x.dtype) + bias.astype(x.dtype)
    return _leaky(out)


def input_scaling(ctus: np.ndarray, qp: float, mode: int, dtype) -> Tuple[np.ndarray, np.ndarray]:
    """AI: net_CNN.py:105-106 (x * 1/255, qp * 1/51).
    LDP: ETH-CNN_Training_LDP/net_CTU64.py:102-103 ((x - 128) / 255 * 10, qp / 51 * 0.18)."""
    x = ctus.astype(dtype)
    q = np.full((x.shape[0], 1), qp, dtype=dtype)
    if mode == MODE_AI:
        x = x * dtype(1.0 / 255.0)
        q = q * dtype(1 / 51.0)
    else:
        x = (x - dtype(128)) / dtype(255.0) * dtype(10)
        q = q / dtype(51.0) * dtype(0.18)
    return x, q


def conv_features(x: np.ndarray, weights: Dict[str, np.ndarray]) -> np.ndarray:
    """net_CNN.py:124-150: three branches and the 2688-wide concat
    [c3_S 512 | c3_M 128 | c3_L 32 | c2_S 1536 | c2_M 384 | c2_L 96], each NHWC-flattened."""
    b = x.shape[0]
    feats = {}
    for br, img in (("L", zero_mean_norm_local(aver_pool(x, 4))),
                    ("M", zero_mean_norm_local(aver_pool(x, 2))),
                    ("S", zero_mean_norm_local(x))):
        v = BRANCH_VARS[br]
        c1 = non_overlap_conv(img[..., None], weights[_vname(v)], weights[_vname(v + 1)])
        c2 = non_overlap_conv(c1, weights[_vname(v + 2)], weights[_vname(v + 3)])
        c3 = non_overlap_conv(c2, weights[_vname(v + 4)], weights[_vname(v + 5)])
        feats[br] = (c2.reshape(b, -1), c3.reshape(b, -1))
    return np.concatenate([feats["S"][1], feats["M"][1], feats["L"][1],
                           feats["S"][0], feats["M"][0], feats["L"][0]], axis=1)


def fc_heads(f: np.ndarray, q: np.ndarray, weights: Dict[str, np.ndarray], return_fc1: bool = False):
    """net_CNN.py:156-185: per head  leaky(f W1 + b1) -> concat qp -> leaky(. W2 + b2) -> concat qp
    -> sigmoid(. W3 + b3)."""
    dt = f.dtype
    outs = []
    fc1s = []
    for h, _n1, _n2, _n3 in HEADS:
        a1 = _leaky(f @ weights["h_fc1__%s__w" % h].astype(dt) + weights["h_fc1__%s__b" % h].astype(dt))
        fc1s.append(a1)
        a1q = np.concatenate([a1, q], axis=1)
        a2 = _leaky(a1q @ weights["h_fc2__%s__w" % h].astype(dt) + weights["h_fc2__%s__b" % h].astype(dt))
        a2q = np.concatenate([a2, q], axis=1)
        y = _sigmoid(a2q @ weights["y_conv_flat__%s__w" % h].astype(dt) + weights["y_conv_flat__%s__b" % h].astype(dt))
        outs.append(y.astype(dt))
    if return_fc1:
        return outs, np.concatenate(fc1s, axis=1)
    return outs


def net_forward(ctus: np.ndarray, qp: float, weights: Dict[str, np.ndarray], mode: int = MODE_AI,
                thresholds: Optional[Tuple[float, float]] = None, dtype=np.float32) -> np.ndarray:
    """One `sess.run([y64, y32, y16])` over ONE sub-batch (net_CNN.py:103-195).
    ctus: [B, 64, 64] (uint8 or float). Returns [B, 21] = [y64 | y32 | y16].
    thresholds = (THR_L1_LOWER, THR_L2_LOWER) enables the batch-level gates of net_CNN.py:175,187
    (AI deployment); None = ungated (training nets / LDP net_CTU64.py)."""
    x, q = input_scaling(np.asarray(ctus), qp, mode, dtype)
    f = conv_features(x, weights)
    y64, y32, y16 = fc_heads(f, q, weights)
    if thresholds is not None:
        t1, t2 = np.float32(thresholds[0]), np.float32(thresholds[1])
        # net_CNN.py:175  tf.cond(count_nonzero(y64 > THR_L1_LOWER) > 0, y32, zeros)
        if np.count_nonzero(y64.astype(np.float32) > t1) == 0:
            y32 = np.zeros_like(y32)
        # net_CNN.py:187  the y16 gate looks at the ALREADY GATED y32
        if np.count_nonzero(y32.astype(np.float32) > t2) == 0:
            y16 = np.zeros_like(y16)
    return np.concatenate([y64, y32, y16], axis=1)


def fc1_export(ctus: np.ndarray, qp: float, weights: Dict[str, np.ndarray], dtype=np.float32) -> np.ndarray:
    """LDP deployment tap: the 448-vector [fc1_64 | fc1_32 | fc1_16] handed to the LSTM
    (HM-16.5_Test_LDP/bin/net_CNN_LSTM_one_step.py:187-199)."""
    x, q = input_scaling(np.asarray(ctus), qp, MODE_LDP, dtype)
    _, fc1 = fc_heads(conv_features(x, weights), q, weights, return_fc1=True)
    return fc1


# ----------------------------------------------------------------------------- LDP: one-step ETH-LSTM heads
LSTM_HEADS = (("64", 64, 48, 1), ("32", 128, 96, 4), ("16", 256, 192, 16))
LDP_LSTM_MODELS = {22: "model_LDP_200000_qp22.dat", 27: "model_LDP_200000_qp27.dat",
                   32: "model_LDP_200000_qp32.dat", 37: "model_LDP_200000_qp37.dat"}


def ldp_lstm_model_prefix(qp: int) -> str:
    """LSTM checkpoint selection, HM-16.5_Test_LDP/bin/resi_to_cu_depth_LDP.py:169-177."""
    if qp < 25:
        return LDP_LSTM_MODELS[22]
    elif qp < 30:
        return LDP_LSTM_MODELS[27]
    elif qp < 35:
        return LDP_LSTM_MODELS[32]
    return LDP_LSTM_MODELS[37]


def random_lstm_weights(seed: int) -> Dict[str, np.ndarray]:
    """Synthetic LSTM checkpoint in the reference's 18-tensor layout (names as written by TF for
    net_CNN_LSTM_one_step.py:201-264)."""
    rng = np.random.default_rng(seed)
    w: Dict[str, np.ndarray] = {}
    for h, n, n2, n3 in LSTM_HEADS:
        pre = "RNN%s/" % h
        w[pre + "multi_rnn_cell/cell_0/lstm_cell/kernel"] = (rng.standard_normal((2 * n, 4 * n)) * (1.0 / math.sqrt(2 * n))).astype(np.float32)
        w[pre + "multi_rnn_cell/cell_0/lstm_cell/bias"] = (rng.standard_normal(4 * n) * 0.1).astype(np.float32)
        w[pre + "fc2/full_connect_w"] = (rng.standard_normal((n + 5, n2)) * (1.3 / math.sqrt(n))).astype(np.float32)
        w[pre + "fc2/full_connect_b"] = (rng.standard_normal(n2) * 0.05).astype(np.float32)
        w[pre + "fc3/full_connect_w"] = (rng.standard_normal((n2 + 5, n3)) * (1.3 / math.sqrt(n2))).astype(np.float32)
        w[pre + "fc3/full_connect_b"] = (rng.standard_normal(n3) * 0.05).astype(np.float32)
    return w


def lstm_cell_step(x, c_prev, h_prev, kernel, bias, forget_bias=1.0, cell_clip=5.0):
    """tf.contrib.rnn.LSTMCell (TensorFlow 1.x rnn_cell_impl.LSTMCell.call, no peepholes, no projection), as
    instantiated at net_CNN_LSTM_one_step.py:205-206: gate columns in the order (i, j, f, o);
    c = sigmoid(f + forget_bias) * c_prev + sigmoid(i) * tanh(j), clipped to +-cell_clip; h = sigmoid(o) * tanh(c)."""
    dt = x.dtype
    z = np.concatenate([x, h_prev], axis=1) @ kernel.astype(dt) + bias.astype(dt)
    i, j, f, o = np.split(z, 4, axis=1)
    c = _sigmoid(f + dt.type(forget_bias)) * c_prev + _sigmoid(i) * np.tanh(j)
    c = np.clip(c, dt.type(-cell_clip), dt.type(cell_clip))
    h = _sigmoid(o) * np.tanh(c)
    return c.astype(dt), h.astype(dt)


def ldp_lstm_forward(ctus: np.ndarray, qp: int, i_frame: int, state_in: np.ndarray, cnn_w, lstm_w,
                     thresholds: Tuple[float, float], dtype=np.float32):
    """One `sess.run` of the deployed LDP predictor on ONE sub-batch (net_CNN_LSTM_one_step.py:266-323):
    residual ETH-CNN up to FC1 (:151-199) -> three one-step LSTMs (:201-264) -> FC2 / FC3 with the five extra
    features [qp/51*0.18, one_hot(i_frame % 4)] -> gates on THR_L1_LOWER / THR_L2_LOWER (:304,310).
    state_in: [B, 1, 2, 448] (c then h, heads 64|128|256 side by side).  Returns ([B,21], state_out)."""
    ctus = np.asarray(ctus)
    n = ctus.shape[0]
    x, _ = input_scaling(ctus, qp, MODE_LDP, dtype)
    _, vec = fc_heads(conv_features(x, cnn_w), np.zeros((n, 1), dtype), cnn_w, return_fc1=True)   # resi_cnn(): the 448-vector
    q = dtype(qp) / dtype(51.0) * dtype(0.18)                                                    # :281
    efs = np.zeros((n, 5), dtype)
    efs[:, 0] = q
    efs[:, 1 + int(i_frame) % 4] = 1                                                             # tf.one_hot(i_frame_in_GOP, depth=4)
    state_in = np.asarray(state_in, dtype).reshape(n, 2, 448)
    state_out = np.zeros((n, 2, 448), dtype)
    ys = []
    off = 0
    for h, nh, _n2, _n3 in LSTM_HEADS:
        pre = "RNN%s/" % h
        c, hh = lstm_cell_step(vec[:, off:off + nh].astype(dtype), state_in[:, 0, off:off + nh], state_in[:, 1, off:off + nh],
                               lstm_w[pre + "multi_rnn_cell/cell_0/lstm_cell/kernel"], lstm_w[pre + "multi_rnn_cell/cell_0/lstm_cell/bias"])
        state_out[:, 0, off:off + nh], state_out[:, 1, off:off + nh] = c, hh
        a2 = _leaky(np.concatenate([hh, efs], 1) @ lstm_w[pre + "fc2/full_connect_w"].astype(dtype) + lstm_w[pre + "fc2/full_connect_b"].astype(dtype))
        y = _sigmoid(np.concatenate([a2, efs], 1) @ lstm_w[pre + "fc3/full_connect_w"].astype(dtype) + lstm_w[pre + "fc3/full_connect_b"].astype(dtype))
        ys.append(y.astype(dtype))
        off += nh
    y64, y32, y16 = ys
    t1, t2 = np.float32(thresholds[0]), np.float32(thresholds[1])
    if np.count_nonzero(y64.astype(np.float32) > t1) == 0:      # :304
        y32 = np.zeros_like(y32)
    if np.count_nonzero(y32.astype(np.float32) > t2) == 0:      # :310 (sees the gated y32)
        y16 = np.zeros_like(y16)
    return np.concatenate([y64, y32, y16], axis=1), state_out.reshape(n, 1, 2, 448)


def ldp_predict_frame(luma: np.ndarray, qp: int, i_frame: int, state_in: Optional[np.ndarray], cnn_w, lstm_w,
                      thresholds: Tuple[float, float], dtype=np.float32):
    """resi_to_cu_depth_LDP.py:72-129 for one residue frame: zero-pad, slice CTUs in raster order, mini-batches of
    1024 through the net; state_in None = zeros (the script uses zeros when i_frame <= 1, :103-112)."""
    h, w = luma.shape
    vh, vw = math.ceil(h / 64) * 64, math.ceil(w / 64) * 64
    pad = np.zeros((vh, vw), np.uint8)
    pad[:h, :w] = luma
    ctus = frame_to_ctus(pad)
    n = ctus.shape[0]
    if state_in is None:
        state_in = np.zeros((n, 1, 2, 448), dtype)
    prob = np.zeros((n, N_OUT), dtype)
    state_out = np.zeros((n, 1, 2, 448), dtype)
    for s in range(0, n, SUB_BATCH):
        e = min(s + SUB_BATCH, n)
        prob[s:e], state_out[s:e] = ldp_lstm_forward(ctus[s:e], qp, i_frame, state_in[s:e], cnn_w, lstm_w, thresholds, dtype)
    return prob.astype(np.float32), state_out.astype(np.float32)


# ----------------------------------------------------------------------------- driver restatement
def get_Y_for_one_frame(buf: memoryview, frame_index: int, frame_width: int, frame_height: int,
                        image_size: int = IMAGE_SIZE) -> np.ndarray:
    """video_to_cu_depth.py:46-59: take W*H luma bytes of the frame (the W*H/2 chroma bytes are read
    and dropped), zero-pad bottom then right to multiples of 64."""
    frame_bytes = frame_width * frame_height * 3 // 2
    off = frame_index * frame_bytes
    data = np.frombuffer(buf, dtype=np.uint8, count=frame_width * frame_height, offset=off)
    data = data.reshape(frame_height, frame_width)
    valid_height = math.ceil(frame_height / image_size) * image_size
    valid_width = math.ceil(frame_width / image_size) * image_size
    if valid_height > frame_height or valid_width > frame_width:
        out = np.zeros((valid_height, valid_width), dtype=np.uint8)
        out[:frame_height, :frame_width] = data
        return out
    return data


def frame_to_ctus(valid_luma: np.ndarray, image_size: int = IMAGE_SIZE) -> np.ndarray:
    """video_to_cu_depth.py:94-104: raster-order (row-major) 64x64 tiles -> [nCTU, 64, 64]."""
    vh, vw = valid_luma.shape
    r, c = vh // image_size, vw // image_size
    return valid_luma.reshape(r, image_size, c, image_size).transpose(0, 2, 1, 3).reshape(r * c, image_size, image_size)


def predict_frame(valid_luma: np.ndarray, qp: float, weights, mode: int, thresholds, dtype=np.float32) -> np.ndarray:
    """video_to_cu_depth.py:61-73: sub-batches of <= 1024 CTUs of one frame, each through net_forward
    (so the gates act per sub-batch)."""
    ctus = frame_to_ctus(valid_luma)
    n = ctus.shape[0]
    out = np.zeros((n, N_OUT), dtype=dtype)
    for i in range(math.ceil(n / SUB_BATCH)):
        s, e = i * SUB_BATCH, min((i + 1) * SUB_BATCH, n)
        out[s:e] = net_forward(ctus[s:e], qp, weights, mode, thresholds, dtype)
    return out


def get_prob(yuv_bytes, frame_width: int, frame_height: int, qp: int, weights, mode: int = MODE_AI,
             thresholds: Optional[Tuple[float, float]] = (0.5, 0.5), dtype=np.float32,
             n_frames: Optional[int] = None) -> np.ndarray:
    """video_to_cu_depth.py:75-118 + :135-140: all frames of an 8-bit 4:2:0 buffer ->
    float32 [n_frames * nCTU, 21], frame-major, CTU raster order."""
    buf = memoryview(yuv_bytes)
    frame_bytes = frame_width * frame_height * 3 // 2
    if len(buf) % frame_bytes != 0:
        raise AssertionError("file_bytes % frame_bytes != 0")  # video_to_cu_depth.py:137
    total = len(buf) // frame_bytes
    if n_frames is None:
        n_frames = total
    rows = []
    for k in range(n_frames):
        luma = get_Y_for_one_frame(buf, k, frame_width, frame_height)
        rows.append(predict_frame(luma, qp, weights, mode, thresholds, dtype))
    if not rows:
        return np.zeros((0, N_OUT), dtype=np.float32)
    return np.concatenate(rows, axis=0).astype(np.float32)


def video_to_cu_depth(yuv_path: str, width: int, height: int, qp: int, model_dir: str = ".",
                      thr_path: Optional[str] = None, out_path: str = "cu_depth.dat",
                      mode: int = MODE_AI) -> np.ndarray:
    """The whole script (video_to_cu_depth.py:120-145) as a function: same inputs from the same
    places, writes the same float32 blob."""
    thr = get_thresholds(thr_path or os.path.join(model_dir, "Thr_info.txt")) if mode == MODE_AI else None
    prefix = ai_model_prefix(qp) if mode == MODE_AI else LDP_MODEL_PREFIX
    weights = load_weights(os.path.join(model_dir, prefix))
    with open(yuv_path, "rb") as f:
        data = f.read()
    prob = get_prob(data, width, height, qp, weights, mode, thr)
    with open(out_path, "wb") as f:
        f.write(prob.astype("<f4").tobytes())
    return prob


# ----------------------------------------------------------------------------- consumer-side quantiser
def decisions(prob: np.ndarray, thr6: Sequence[float] = (0.5,) * 6) -> np.ndarray:
    """HM's use of each probability, HM-16.5_Test_AI/source/Lib/TLibEncoder/TEncCu.cpp:448-462:
    p > up -> 2 (split only), p <= down -> 0 (no split), else 1 (check both).  Thr_info.txt order
    is up,down per depth (TEncCu.cpp:250).  prob: [..., 21]; returns uint8 of the same shape."""
    p = np.asarray(prob, dtype=np.float32)
    up = np.empty(21, dtype=np.float32)
    down = np.empty(21, dtype=np.float32)
    for lvl, sl in enumerate((slice(0, 1), slice(1, 5), slice(5, 21))):
        up[sl] = np.float32(thr6[2 * lvl])
        down[sl] = np.float32(thr6[2 * lvl + 1])
    d = np.ones(p.shape, dtype=np.uint8)
    d[p > up] = 2
    d[(p <= down) & ~(p > up)] = 0
    return d


def hm_cu_index(depth: int, x: int, y: int) -> int:
    """Index into the 21-vector for a CU at pixel offset (x, y) inside its CTU, TEncCu.cpp:434-447."""
    if depth == 0:
        return 0
    if depth == 1:
        return 1 + x // 32 + 2 * (y // 32)
    return 5 + x // 16 + 4 * (y // 16)


# ----------------------------------------------------------------------------- synthetic content
def synth_frame(width: int, height: int, seed: int) -> np.ndarray:
    """Procedural multi-scale luma frame: 128 + six random sinusoids + per-32x32-cell Gaussian noise
    + random rectangles, clipped to uint8. Deterministic in (width, height, seed)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    img = np.full((height, width), 128.0, dtype=np.float32)
    for _ in range(6):
        freq = rng.uniform(0.002, 0.08)
        ang = rng.uniform(0, 2 * np.pi)
        amp = rng.uniform(5, 40)
        ph = rng.uniform(0, 2 * np.pi)
        img += (amp * np.sin(2 * np.pi * freq * (xx * np.cos(ang) + yy * np.sin(ang)) + ph)).astype(np.float32)
    ch, cw = (height + 31) // 32, (width + 31) // 32
    sig = rng.choice(np.array([0, 0, 0, 1, 2, 4, 8, 16, 25], dtype=np.float32), size=(ch, cw))
    sig_full = np.repeat(np.repeat(sig, 32, axis=0), 32, axis=1)[:height, :width]
    img += rng.standard_normal((height, width), dtype=np.float32) * sig_full
    for _ in range(max(1, width * height // 20000)):
        rw, rh = int(rng.integers(4, 201)), int(rng.integers(4, 201))
        x0, y0 = int(rng.integers(0, width)), int(rng.integers(0, height))
        img[y0:y0 + rh, x0:x0 + rw] += rng.uniform(-60, 60)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def synth_residue_frame(width: int, height: int, seed: int) -> np.ndarray:
    """Residue-like luma (config 5): clip(128 + r), r ~ Laplace(0, b) with b varying per 32x32 cell and
    flat (all-128) regions -- the on-disk form of resi.yuv (HM-16.5_Test_LDP TEncSearch.cpp:4548-4557)."""
    rng = np.random.default_rng(seed)
    ch, cw = (height + 31) // 32, (width + 31) // 32
    b = rng.choice(np.array([0, 0, 0.5, 1, 2, 3, 4, 6, 10], dtype=np.float32), size=(ch, cw))
    b_full = np.repeat(np.repeat(b, 32, axis=0), 32, axis=1)[:height, :width]
    r = rng.laplace(0.0, 1.0, size=(height, width)).astype(np.float32) * b_full
    return np.clip(np.rint(128.0 + r), 0, 255).astype(np.uint8)


def synth_yuv(width: int, height: int, n_frames: int, seed0: int = 0, residue: bool = False) -> bytes:
    """8-bit 4:2:0 planar stream: Y from synth_frame(seed0 + k), U/V constant 128."""
    uv = bytes([128]) * (width * height // 2)
    gen = synth_residue_frame if residue else synth_frame
    return b"".join(gen(width, height, seed0 + k).tobytes() + uv for k in range(n_frames))


def known_answer_ctus() -> np.ndarray:
    """The two data-free CTUs of SURVEY.md section 4 item 5."""
    y, x = np.mgrid[0:64, 0:64]
    c0 = (7 * x + 13 * y + (x * y) // 8) % 256
    c1 = 128 + 20 * ((x // 16 + y // 16) % 2) + 6 * ((x % 8) < 4)
    return np.stack([c0, c1]).astype(np.uint8)
