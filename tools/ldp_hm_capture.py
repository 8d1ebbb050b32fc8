#!/usr/bin/env python
"""Phase 1 of the LDP end-to-end check (runs where /root/reference exists): encode a short synthetic clip with the
UNMODIFIED prebuilt LDP encoder (HM-16.5_Test_LDP/bin/TAppEncoderStatic) while an ORACLE stand-in daemon answers its
file-signal handshake, and record the residue frames HM produced, the oracle's answers and the bitstream digest into
tests/golden/ldp_hm_capture.npz.  Phase 2 (tools/make_cuda_fixture.py, GPU box) runs the CUDA predictor over the
recorded residue frames; phase 3 (tests/test_hm_e2e.py) replays the encoder with the CUDA-made answers."""
import hashlib
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import assets  # noqa: E402
from oracle import ethcnn_oracle as eo  # noqa: E402

LDP_BIN, _WHY = assets.hm_dir("LDP")   # None unless running the prebuilt encoder was opted into (oracle/assets.py)
W, H, NF, QP = 416, 240, 5, 32


def clip():
    """Temporally coherent luma: one procedural frame panned by a few pixels per frame plus mild noise."""
    base = eo.synth_frame(W + 64, H + 64, 900)
    rng = np.random.default_rng(901)
    frames = []
    for k in range(NF):
        f = base[8 + 3 * k: 8 + 3 * k + H, 5 + 4 * k: 5 + 4 * k + W].astype(np.int16) + rng.integers(-2, 3, (H, W))
        frames.append(np.clip(f, 0, 255).astype(np.uint8))
    uv = bytes([128]) * (W * H // 2)
    return b"".join(f.tobytes() + uv for f in frames)


def run_hm_ldp(work, answer):
    """Run the encoder in `work`; `answer(i_frame, w, h, qp, resi_luma) -> (prob, state)` is called per P frame."""
    if LDP_BIN is None:
        raise RuntimeError(_WHY)
    os.makedirs(work, exist_ok=True)
    hm = os.path.join(work, "TAppEncoderStatic")
    shutil.copyfile(os.path.join(LDP_BIN, "TAppEncoderStatic"), hm)
    os.chmod(hm, 0o755)
    for fn in ("encoder_lowdelay_P_main.cfg", "Thr_info.txt"):
        shutil.copyfile(os.path.join(LDP_BIN, fn), os.path.join(work, fn))
    with open(os.path.join(work, "in.yuv"), "wb") as f:
        f.write(clip())
    stop = threading.Event()
    seen = []

    def daemon():
        while not stop.is_set():
            if os.path.exists(os.path.join(work, "pred_start.sig")):
                t = open(os.path.join(work, "command.dat")).readline().split(" ")
                if len(t) == 5 and t[4] == "[end]":
                    i_frame, fw, fh, q = map(int, t[:4])
                    os.remove(os.path.join(work, "pred_start.sig"))
                    luma = np.frombuffer(open(os.path.join(work, "resi.yuv"), "rb").read(fw * fh), np.uint8).reshape(fh, fw).copy()
                    prob, state = answer(i_frame, fw, fh, q, luma)
                    state.astype("<f4").tofile(os.path.join(work, "state.dat"))
                    prob.astype("<f4").tofile(os.path.join(work, "cu_depth.dat"))
                    open(os.path.join(work, "pred_end.sig"), "wb").close()
                    seen.append((i_frame, luma, prob))
            time.sleep(0.0005)
    th = threading.Thread(target=daemon)
    th.start()
    try:
        r = subprocess.run([hm, "-c", "encoder_lowdelay_P_main.cfg", "-i", "in.yuv", "-wdt", str(W), "-hgt", str(H), "-fr", "30",
                            "-f", str(NF), "-q", str(QP), "-b", "str.bin", "-o", ""], cwd=work, capture_output=True, timeout=900)
    finally:
        stop.set()
        th.join()
    assert r.returncode == 0, r.stdout.decode()[-1500:] + r.stderr.decode()[-1500:]
    data = open(os.path.join(work, "str.bin"), "rb").read()
    return hashlib.md5(data).hexdigest(), len(data), seen


def oracle_answerer():
    cnn = assets.load_weights(assets.LDP_MODEL)
    lw = assets.load_weights(eo.ldp_lstm_model_prefix(QP))
    st = {"state": None}

    def answer(i_frame, fw, fh, q, luma):
        prob, st["state"] = eo.ldp_predict_frame(luma, q, i_frame, st["state"] if i_frame > 1 else None, cnn, lw, (0.6, 0.7))
        return prob, st["state"]
    return answer


def main():
    work = tempfile.mkdtemp(prefix="ldp_hm_capture_")
    try:
        md5, size, seen = run_hm_ldp(work, oracle_answerer())
    finally:
        shutil.rmtree(work, ignore_errors=True)
    out = os.path.join(ROOT, "tests", "golden", "ldp_hm_capture.npz")
    np.savez_compressed(out, width=W, height=H, qp=QP, n_frames=NF, i_frames=np.array([s[0] for s in seen]),
                        resi=np.stack([s[1] for s in seen]), oracle_prob=np.stack([s[2] for s in seen]),
                        str_md5=np.array(md5), str_size=size)
    print("captured %d P frames, str.bin %d B md5 %s -> %s (%d B)" % (len(seen), size, md5, out, os.path.getsize(out)))


if __name__ == "__main__":
    main()
