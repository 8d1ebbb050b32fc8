#!/usr/bin/env python
"""TEST INFRASTRUCTURE (oracle) -- generate tests/golden/*.npz by executing the reference's OWN,
UNMODIFIED inference code from /root/reference on top of oracle/tf_shim (numpy stand-in for the TF
ops; TensorFlow itself is not installable here).

    python oracle/make_golden.py            # needs /root/reference; rewrites tests/golden/

What runs unmodified:
  * HM-16.5_Test_AI/bin/video_to_cu_depth.py (whole script, via runpy, argv = <yuv> W H QP, cwd holding
    links to the deployed checkpoints + Thr_info.txt) -> the cu_depth.dat it writes is the golden
    output.  Covers frame reading, zero padding, CTU slicing, <=1024 sub-batching, gates, writer.
  * ETH-CNN_Training_LDP/net_CTU64.py net() with HM-16.5_Test_LDP/bin/model_LDP_2000000_qp22~37.dat
    -> golden 21-vectors for residue CTUs (BASELINE config 5).
  * HM-16.5_Test_LDP/bin/net_CNN_LSTM_one_step.py resi_cnn() -> golden 448-vectors (the FC1 tap).

Each .npz stores the exact input bytes (or, for the one large case, the recipe + sha256) and the
reference output, so the GPU box (which has no /root/reference) can check against them.
"""
from __future__ import annotations

import hashlib
import importlib
import os
import runpy
import shutil
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import ethcnn_oracle as eo  # noqa: E402

REF = os.environ.get("ETHCNN_REFERENCE", "/root/reference")
AI_BIN = os.path.join(REF, "HM-16.5_Test_AI", "bin")
LDP_BIN = os.path.join(REF, "HM-16.5_Test_LDP", "bin")
LDP_TRAIN = os.path.join(REF, "ETH-CNN_Training_LDP")
SHIM = os.path.join(REPO, "oracle", "tf_shim")
OUT = os.path.join(REPO, "tests", "golden")


def _fresh_tf():
    for m in [k for k in sys.modules if k == "tensorflow" or k.startswith("tensorflow.")]:
        del sys.modules[m]
    if SHIM not in sys.path:
        sys.path.insert(0, SHIM)
    tf = importlib.import_module("tensorflow")
    tf.reset_default_graph()
    return tf


def run_reference_script(yuv_bytes: bytes, width: int, height: int, qp: int, thr_line: str = None) -> np.ndarray:
    """Execute the unmodified video_to_cu_depth.py the way HM does (TAppEncCfg.cpp:2319) and return
    the float32 contents of the cu_depth.dat it wrote."""
    work = tempfile.mkdtemp(prefix="ethcnn_golden_")
    old_cwd, old_argv, old_path = os.getcwd(), sys.argv, list(sys.path)
    try:
        for fn in os.listdir(AI_BIN):
            if fn.startswith("model_") and not fn.endswith(".meta"):
                os.symlink(os.path.join(AI_BIN, fn), os.path.join(work, fn))
        if thr_line is None:
            shutil.copy(os.path.join(AI_BIN, "Thr_info.txt"), os.path.join(work, "Thr_info.txt"))
        else:
            with open(os.path.join(work, "Thr_info.txt"), "w") as f:
                f.write(thr_line)
        with open(os.path.join(work, "in.yuv"), "wb") as f:
            f.write(yuv_bytes)
        os.chdir(work)
        _fresh_tf()
        for m in ("net_CNN",):
            sys.modules.pop(m, None)
        sys.path.insert(0, AI_BIN)
        sys.argv = ["video_to_cu_depth.py", "in.yuv", str(width), str(height), str(qp)]
        runpy.run_path(os.path.join(AI_BIN, "video_to_cu_depth.py"), run_name="__main__")
        return np.fromfile(os.path.join(work, "cu_depth.dat"), dtype="<f4").reshape(-1, 21)
    finally:
        os.chdir(old_cwd)
        sys.argv = old_argv
        sys.path[:] = old_path
        shutil.rmtree(work, ignore_errors=True)


def run_reference_ldp_net(ctus: np.ndarray, qp: int) -> np.ndarray:
    """ETH-CNN_Training_LDP/net_CTU64.py net() (unmodified) with the deployed LDP CNN checkpoint."""
    old_path = list(sys.path)
    try:
        tf = _fresh_tf()
        sys.modules.pop("net_CTU64", None)
        sys.path.insert(0, LDP_TRAIN)
        nt = importlib.import_module("net_CTU64")
        x = tf.placeholder("float", [None, 64, 64, 1])
        y_ = tf.placeholder("float", [None, 16])
        qp_ph = tf.placeholder("float", [None, 1])
        outs = nt.net(x, y_, qp_ph, 0, 0, 0.01, 0.9, 1000, 0.3, 0)
        y64, y32, y16, opt_vars_all = outs[3], outs[4], outs[5], outs[-1]
        sess = tf.Session()
        tf.train.Saver(opt_vars_all).restore(sess, os.path.join(LDP_BIN, "model_LDP_2000000_qp22~37.dat"))
        n = ctus.shape[0]
        r = sess.run([y64, y32, y16], feed_dict={x: ctus.reshape(n, 64, 64, 1).astype(np.float32),
                                                 y_: np.zeros((n, 16)), qp_ph: np.full((n, 1), qp)})
        return np.concatenate(r, axis=1).astype(np.float32)
    finally:
        sys.path[:] = old_path
        sys.modules.pop("net_CTU64", None)


def run_reference_resi_cnn(ctus: np.ndarray) -> np.ndarray:
    """HM-16.5_Test_LDP/bin/net_CNN_LSTM_one_step.py resi_cnn() (unmodified): the 448-vector."""
    old_path, old_cwd = list(sys.path), os.getcwd()
    work = tempfile.mkdtemp(prefix="ethcnn_golden_")
    try:
        # the module reads 'Thr_info.txt' from cwd at import time (net_CNN_LSTM_one_step.py:72, mode 'r+')
        shutil.copy(os.path.join(LDP_BIN, "Thr_info.txt"), os.path.join(work, "Thr_info.txt"))
        os.chdir(work)
        tf = _fresh_tf()
        for m in ("net_CNN_LSTM_one_step", "config"):
            sys.modules.pop(m, None)
        sys.path.insert(0, LDP_BIN)
        nt = importlib.import_module("net_CNN_LSTM_one_step")
        x = tf.placeholder("float", [None, 64, 64, 1])
        vector, opt_vars = nt.resi_cnn(x)
        sess = tf.Session()
        tf.train.Saver(opt_vars).restore(sess, os.path.join(LDP_BIN, "model_LDP_2000000_qp22~37.dat"))
        n = ctus.shape[0]
        v = sess.run(vector, feed_dict={x: ctus.reshape(n, 64, 64, 1).astype(np.float32)})
        return np.asarray(v, dtype=np.float32).reshape(n, 448)
    finally:
        os.chdir(old_cwd)
        shutil.rmtree(work, ignore_errors=True)
        sys.path[:] = old_path
        for m in ("net_CNN_LSTM_one_step", "config"):
            sys.modules.pop(m, None)


def run_reference_ldp_daemon_functions(frames, width, height, qps):
    """HM-16.5_Test_LDP/bin/resi_to_cu_depth_LDP.py (unmodified): import the module (builds the CNN + one-step LSTM
    graph through net_CNN_LSTM_one_step.net), restore the CNN and the per-QP LSTM checkpoints the way its main loop
    does (:158-178), then call its own get_images_from_one_file / get_state_in_from_one_file / predict_cu_depth /
    save_cu_depth_and_state for a sequence of residue frames.  Returns per frame (cu_depth [n,21], state [n,1,2,448])."""
    old_path, old_cwd = list(sys.path), os.getcwd()
    work = tempfile.mkdtemp(prefix="ethcnn_golden_")
    try:
        shutil.copy(os.path.join(LDP_BIN, "Thr_info.txt"), os.path.join(work, "Thr_info.txt"))
        for fn in os.listdir(LDP_BIN):
            if fn.startswith("model_") and not fn.endswith(".meta"):
                os.symlink(os.path.join(LDP_BIN, fn), os.path.join(work, fn))
        os.chdir(work)
        tf = _fresh_tf()
        for m in ("resi_to_cu_depth_LDP", "net_CNN_LSTM_one_step", "config"):
            sys.modules.pop(m, None)
        sys.path.insert(0, LDP_BIN)
        mod = importlib.import_module("resi_to_cu_depth_LDP")
        mod.state_file = "state.dat"   # a global its __main__ block defines (:150) and get_state_in_from_one_file reads
        mod.saver_CNN.restore(mod.sess, "model_LDP_2000000_qp22~37.dat")
        out = []
        qp_last = None
        for i_frame, (luma, qp) in enumerate(zip(frames, qps), start=1):
            if qp != qp_last:   # :165-178
                name = ("model_LDP_200000_qp22.dat" if qp < 25 else "model_LDP_200000_qp27.dat" if qp < 30 else
                        "model_LDP_200000_qp32.dat" if qp < 35 else "model_LDP_200000_qp37.dat")
                mod.saver_LSTM.restore(mod.sess, name)
                qp_last = qp
            with open("resi.yuv", "wb") as f:
                f.write(luma.tobytes() + bytes([128]) * (width * height // 2))
            images, num_vectors = mod.get_images_from_one_file("resi.yuv", width, height, 64)
            state_in = mod.get_state_in_from_one_file("state.dat", num_vectors, i_frame)
            depth_out, state_out = mod.predict_cu_depth(images, state_in, qp, i_frame)
            mod.save_cu_depth_and_state(depth_out, state_out, "cu_depth.dat", "state.dat", "pred_end.sig", num_vectors)
            out.append((np.fromfile("cu_depth.dat", "<f4").reshape(-1, 21), np.fromfile("state.dat", "<f4").reshape(-1, 1, 2, 448)))
        return out
    finally:
        os.chdir(old_cwd)
        shutil.rmtree(work, ignore_errors=True)
        sys.path[:] = old_path
        for m in ("resi_to_cu_depth_LDP", "net_CNN_LSTM_one_step", "config"):
            sys.modules.pop(m, None)


def demo_ctus() -> np.ndarray:
    d = np.fromfile(os.path.join(REF, "ETH-CNN_Training_AI", "Data", "AI_Test_5000.dat_shuffled"), dtype=np.uint8)
    return d.reshape(5000, 4992)[:, :4096].reshape(5000, 64, 64)


def mosaic(ctus: np.ndarray, rows: int, cols: int) -> np.ndarray:
    return ctus[:rows * cols].reshape(rows, cols, 64, 64).transpose(0, 2, 1, 3).reshape(rows * 64, cols * 64)


def yuv_from_luma(frames) -> bytes:
    out = []
    for fr in frames:
        h, w = fr.shape
        out.append(np.ascontiguousarray(fr, dtype=np.uint8).tobytes())
        out.append(bytes([128]) * (w * h // 2))
    return b"".join(out)


QPS = (22, 27, 32, 37)


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="", help="comma-separated fixture stems to (re)generate; default all")
    only = set(x for x in ap.parse_args().only.split(",") if x)

    def want(stem):
        return not only or stem in only
    os.makedirs(OUT, exist_ok=True)
    real = demo_ctus()

    # A. padding in both directions, two frames, procedural content
    w, h = 200, 136
    yuv = eo.synth_yuv(w, h, 2, seed0=100)
    if want("ai_pad_200x136_f2"):
      np.savez_compressed(os.path.join(OUT, "ai_pad_200x136_f2.npz"), yuv=np.frombuffer(yuv, np.uint8), width=w, height=h,
                        qps=np.array(QPS), **{"prob_qp%d" % q: run_reference_script(yuv, w, h, q) for q in QPS})

    # B. mosaic of real CTUs (realistic spread of probabilities)
    luma = mosaic(real[:40], 5, 8)
    yuv = yuv_from_luma([luma])
    if want("ai_mosaic_512x320"):
      np.savez_compressed(os.path.join(OUT, "ai_mosaic_512x320.npz"), yuv=np.frombuffer(yuv, np.uint8), width=512, height=320,
                        qps=np.array(QPS), **{"prob_qp%d" % q: run_reference_script(yuv, 512, 320, q) for q in QPS})

    # C. gates: 64x64 "video" -- every frame is its own sub-batch of one CTU.
    wq = eo.load_weights(os.path.join(AI_BIN, eo.ai_model_prefix(32)))
    p = eo.net_forward(real[:2000], 32, wq)
    g1_off = np.where(p[:, 0] <= 0.45)[0][:3]                                   # y64 small -> y32, y16 zeroed
    g2_off = np.where((p[:, 0] > 0.55) & (p[:, 1:5].max(1) <= 0.45))[0][:3]     # y64 on, y32 all small -> y16 zeroed
    both_on = np.where((p[:, 0] > 0.55) & (p[:, 1:5].max(1) > 0.55))[0][:3]
    assert len(g1_off) == 3 and len(g2_off) == 3 and len(both_on) == 3
    sel = np.concatenate([g1_off, g2_off, both_on])
    flat = np.full((1, 64, 64), 128, np.uint8)
    frames = list(real[sel]) + [flat[0]]
    yuv = yuv_from_luma(frames)
    if want("ai_gates_64x64_f10"):
      np.savez_compressed(os.path.join(OUT, "ai_gates_64x64_f10.npz"), yuv=np.frombuffer(yuv, np.uint8), width=64, height=64,
                        qps=np.array([32]), prob_qp32=run_reference_script(yuv, 64, 64, 32))

    # C2. a frame with more than 1024 CTUs: 33 x 32 = 1056 -> sub-batches 1024 + 32; textured band at the
    # top, flat elsewhere, so the second sub-batch is entirely gated while the first is not.
    w, h = 2112, 2048
    luma = np.full((h, w), 128, np.uint8)
    luma[:128, :1024] = mosaic(real[100:132], 2, 16)
    yuv = yuv_from_luma([luma])
    if want("ai_subbatch_2112x2048"):
      np.savez_compressed(os.path.join(OUT, "ai_subbatch_2112x2048.npz"), yuv=np.frombuffer(yuv, np.uint8), width=w, height=h,
                        qps=np.array([32]), prob_qp32=run_reference_script(yuv, w, h, 32))

    # C3. non-default thresholds (tokens [1] and [3] are the ones the script reads)
    luma = mosaic(real[200:212], 3, 4)
    yuv = yuv_from_luma([luma, np.full_like(luma, 77)])
    thr_line = "0.9 0.95 0.8 0.7 0.6 0.4"
    if want("ai_thr_256x192_f2"):
      np.savez_compressed(os.path.join(OUT, "ai_thr_256x192_f2.npz"), yuv=np.frombuffer(yuv, np.uint8), width=256, height=192,
                        qps=np.array([27]), thr_line=np.array(thr_line), prob_qp27=run_reference_script(yuv, 256, 192, 27, thr_line))

    # D. BASELINE config 1: 768x512, 1 frame, QP 32 -- input regenerated from the recipe at test time
    w, h = 768, 512
    yuv = eo.synth_yuv(w, h, 1, seed0=1)
    if want("ai_cfg1_768x512"):
      np.savez_compressed(os.path.join(OUT, "ai_cfg1_768x512.npz"), width=w, height=h, seed0=1, n_frames=1,
                        yuv_sha256=np.array(hashlib.sha256(yuv).hexdigest()), qps=np.array([32]),
                        prob_qp32=run_reference_script(yuv, w, h, 32))

    # E. LDP residual CNN (config 5 arithmetic): residue-like CTUs + a few real ones
    resi = eo.frame_to_ctus(eo.synth_residue_frame(512, 256, 7))
    ctus = np.concatenate([resi, eo.known_answer_ctus(), real[:6]])
    if want("ldp_ctus"):
      np.savez_compressed(os.path.join(OUT, "ldp_ctus.npz"), ctus=ctus, qps=np.array([22, 37]),
                        prob_qp22=run_reference_ldp_net(ctus, 22), prob_qp37=run_reference_ldp_net(ctus, 37),
                        fc1_vector=run_reference_resi_cnn(ctus))

    # E2. the deployed LDP predictor (CNN + one-step ETH-LSTM) over a 5-frame residue sequence with a QP switch,
    # state carried through state.dat exactly as the daemon does
    w, h = 200, 136
    frames = [eo.synth_residue_frame(w, h, 20 + k) for k in range(5)]
    frames[3] = np.full((h, w), 128, np.uint8)          # an all-flat residue frame (gates close)
    qps = [32, 32, 32, 32, 22]
    if want("ldp_lstm_200x136_f5"):
        res = run_reference_ldp_daemon_functions(frames, w, h, qps)
        np.savez_compressed(os.path.join(OUT, "ldp_lstm_200x136_f5.npz"), frames=np.stack(frames), width=w, height=h, qps=np.array(qps),
                            cu_depth=np.stack([r[0] for r in res]), state=np.stack([r[1] for r in res]))
    # E3. a first frame that is completely flat (zero LSTM state): y64 stays under THR_L1_LOWER, both gates close;
    # and a 2112x2048 residue frame (1056 CTUs: mini-batches of 1024 + 32, the second one flat)
    if want("ldp_lstm_gates"):
        flat = [np.full((64, 128), 128, np.uint8)]
        res_flat = run_reference_ldp_daemon_functions(flat, 128, 64, [37])
        big = np.full((2048, 2112), 128, np.uint8)
        big[:128, :1024] = eo.synth_residue_frame(1024, 128, 31)
        res_big = run_reference_ldp_daemon_functions([big], 2112, 2048, [27])
        np.savez_compressed(os.path.join(OUT, "ldp_lstm_gates.npz"), flat=flat[0], flat_cu_depth=res_flat[0][0], flat_state=res_flat[0][1],
                            big=big, big_cu_depth=res_big[0][0], big_state=res_big[0][1])

    # F. raw-CTU AI vectors (ungated single sub-batch semantics are covered by C; these pin the net itself)
    ctus = np.concatenate([eo.known_answer_ctus(), real[300:330]])
    if want("ai_ctus_ungated"):
        out = {}
        for q in QPS:
            yuv = yuv_from_luma(list(ctus))  # 64x64 frames
            out["prob_qp%d" % q] = run_reference_script(yuv, 64, 64, q, "0.5 -1 0.5 -1 0.5 -1")  # gates always open
        np.savez_compressed(os.path.join(OUT, "ai_ctus_ungated.npz"), ctus=ctus, qps=np.array(QPS), **out)
    # G. the two checkpoints the GPU box needs for real-weight parity, as {tensor name: float32 array}
    # (re-serialised into TF bundles at test time by oracle/assets.materialize)
    from oracle import assets, tf_bundle
    for name, npz in assets.NPZ.items():
        if not want(npz[:-4]):
            continue
        src = os.path.join(AI_BIN if name.startswith("model_2000000") else LDP_BIN, name)
        np.savez_compressed(os.path.join(OUT, npz), **tf_bundle.read_bundle(src, verify_crc=True))
    print("golden fixtures written to", OUT)
    for fn in sorted(os.listdir(OUT)):
        print("  %-32s %8d B" % (fn, os.path.getsize(os.path.join(OUT, fn))))


if __name__ == "__main__":
    main()
