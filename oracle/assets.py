"""TEST INFRASTRUCTURE (oracle) -- where the checkers find the reference's trained checkpoints.

/root/reference does not exist on the GPU box, so the deployed checkpoints travel in two ways:
  * oracle/_ref/checkpoints/{AI,LDP}/  -- byte copies staged by __graft_entry__.build() when
    /root/reference is present (git-ignored, NOT gpurun-ignored, like the built .so files);
  * tests/golden/weights_*.npz         -- committed fixtures (tensor name -> float32 array) written by
    oracle/make_golden.py for the QP-32 AI model and the LDP CNN, re-serialised as TF bundles on demand
    with oracle/tf_bundle.write_bundle.
"""
from __future__ import annotations

import os
import shutil
from typing import Dict, List

import numpy as np

from . import tf_bundle

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ETHCNN_REFERENCE", "/root/reference")
STAGED = os.path.join(REPO, "oracle", "_ref", "checkpoints")
GOLDEN = os.path.join(REPO, "tests", "golden")

AI_MODELS = {22: "model_2000000_qp20~25.dat", 27: "model_2000000_qp25~30.dat",
             32: "model_2000000_qp30~35.dat", 37: "model_2000000_qp35~40.dat"}
LDP_MODEL = "model_LDP_2000000_qp22~37.dat"
LDP_LSTM_MODELS = {22: "model_LDP_200000_qp22.dat", 27: "model_LDP_200000_qp27.dat",
                   32: "model_LDP_200000_qp32.dat", 37: "model_LDP_200000_qp37.dat"}
NPZ = {"model_2000000_qp30~35.dat": "weights_ai_qp30_35.npz", LDP_MODEL: "weights_ldp_cnn.npz",
       LDP_LSTM_MODELS[32]: "weights_ldp_lstm_qp32.npz"}
SUFFIXES = (".index", ".data-00000-of-00001")


def _source_dirs(kind: str) -> List[str]:
    ref_bin = os.path.join(REF, "HM-16.5_Test_AI" if kind == "AI" else "HM-16.5_Test_LDP", "bin")
    return [ref_bin, os.path.join(STAGED, kind)]


def stage_from_reference() -> int:
    """Copy the 5 CNN checkpoints + Thr_info.txt from /root/reference into oracle/_ref/checkpoints.
    Returns the number of files copied (0 when the reference is absent)."""
    n = 0
    for kind, names in (("AI", list(AI_MODELS.values())), ("LDP", [LDP_MODEL] + list(LDP_LSTM_MODELS.values()))):
        src_dir = _source_dirs(kind)[0]
        if not os.path.isdir(src_dir):
            continue
        dst_dir = os.path.join(STAGED, kind)
        os.makedirs(dst_dir, exist_ok=True)
        for name in names:
            for suf in SUFFIXES:
                s, d = os.path.join(src_dir, name + suf), os.path.join(dst_dir, name + suf)
                if os.path.exists(s) and not (os.path.exists(d) and os.path.getsize(d) == os.path.getsize(s)):
                    shutil.copyfile(s, d)
                    n += 1
        thr = os.path.join(src_dir, "Thr_info.txt")
        if os.path.exists(thr):
            shutil.copyfile(thr, os.path.join(dst_dir, "Thr_info.txt"))
    return n


# ----------------------------------------------------------------------------- labelled demo CTUs
DEMO_SETS = ("AI_Train_5000.dat_shuffled", "AI_Valid_5000.dat_shuffled", "AI_Test_5000.dat_shuffled")
DEMO_STAGED = os.path.join(REPO, "oracle", "_ref", "data")
SAMPLE_BYTES = 4992   # 4096 luma + 64 info + 52 x 16 labels (ETH-CNN_Training_AI/input_data.py:16)


def stage_demo_data() -> int:
    """Copy the reference's 15 000 labelled CTU samples (ETH-CNN_Training_AI/Data, 3 x 24.96 MB) into oracle/_ref/data
    so that the real-content parity / accuracy tests can run on the GPU box.  Returns the number of files copied."""
    src_dir = os.path.join(REF, "ETH-CNN_Training_AI", "Data")
    n = 0
    if not os.path.isdir(src_dir):
        return 0
    os.makedirs(DEMO_STAGED, exist_ok=True)
    for name in DEMO_SETS:
        s, d = os.path.join(src_dir, name), os.path.join(DEMO_STAGED, name)
        if os.path.exists(s) and not (os.path.exists(d) and os.path.getsize(d) == os.path.getsize(s)):
            shutil.copyfile(s, d)
            n += 1
    return n


def demo_set_path(name: str):
    for d in (os.path.join(REF, "ETH-CNN_Training_AI", "Data"), DEMO_STAGED):
        p = os.path.join(d, name)
        if os.path.exists(p):
            return p
    return None


def load_demo_set(name: str):
    """(luma [n,64,64] uint8, labels {qp: [n,16] uint8}) of one demo file; label row of QP q sits at byte 4160 + 16 q
    (input_data.py:101-109, writer Extract_Data/extract_data_AI.py:103-110)."""
    p = demo_set_path(name)
    if p is None:
        raise FileNotFoundError("demo set %s is not available on this box" % name)
    raw = np.fromfile(p, dtype=np.uint8).reshape(-1, SAMPLE_BYTES)
    luma = raw[:, :4096].reshape(-1, 64, 64).copy()
    labels = {qp: raw[:, 4160 + 16 * qp: 4160 + 16 * qp + 16].copy() for qp in (22, 27, 32, 37)}
    return luma, labels


# ----------------------------------------------------------------------------- the reference's PREBUILT encoders
# tests/test_hm_e2e.py and tests/test_gpu_hm_live.py execute the reference's prebuilt x86-64 HM binaries (an opaque
# third-party ELF).  That never happens implicitly: the binaries are only staged, and only run, after an explicit
# opt-in -- `python -m oracle.assets --stage-hm --allow-execute` (which records their sha256 in
# oracle/_ref/hm/EXECUTION_ALLOWED), or ETHCNN_RUN_REFERENCE_HM=1 in the environment of the test run.
HM_STAGED = os.path.join(REPO, "oracle", "_ref", "hm")
HM_FILES = {"AI": ("TAppEncoderStatic", "encoder_intra_main.cfg", "Thr_info.txt"),
            "LDP": ("TAppEncoderStatic", "encoder_lowdelay_P_main.cfg", "Thr_info.txt")}


def _sha256(path: str) -> str:
    import hashlib
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def stage_hm(allow_execute: bool) -> int:
    """Copy the prebuilt encoders + their cfg files into oracle/_ref/hm/{AI,LDP}; with allow_execute, record the opt-in."""
    n = 0
    lines = []
    for kind, files in HM_FILES.items():
        src_dir = _source_dirs(kind)[0]
        if not os.path.isdir(src_dir):
            continue
        dst_dir = os.path.join(HM_STAGED, kind)
        os.makedirs(dst_dir, exist_ok=True)
        for fn in files:
            s, d = os.path.join(src_dir, fn), os.path.join(dst_dir, fn)
            if os.path.exists(s):
                shutil.copyfile(s, d)
                n += 1
        lines.append("%s %s" % (kind, _sha256(os.path.join(dst_dir, "TAppEncoderStatic"))))
    if allow_execute and lines:
        with open(os.path.join(HM_STAGED, "EXECUTION_ALLOWED"), "w") as f:
            f.write("\n".join(lines) + "\n")
    return n


def hm_dir(kind: str):
    """Directory holding the prebuilt encoder of `kind` and its cfg files, or None.  None as well when running the binary
    has not been opted into (see above); the returned string `reason` says which."""
    allowed = {}
    marker = os.path.join(HM_STAGED, "EXECUTION_ALLOWED")
    if os.path.exists(marker):
        for line in open(marker):
            t = line.split()
            if len(t) == 2:
                allowed[t[0]] = t[1]
    env_ok = os.environ.get("ETHCNN_RUN_REFERENCE_HM", "") == "1"
    for d in (os.path.join(HM_STAGED, kind), _source_dirs(kind)[0]):
        exe = os.path.join(d, "TAppEncoderStatic")
        if not os.path.exists(exe):
            continue
        if env_ok or allowed.get(kind) == _sha256(exe):
            return d, "ok"
        return None, ("the prebuilt HM encoder is present but running it needs an explicit opt-in: "
                      "`python -m oracle.assets --stage-hm --allow-execute` or ETHCNN_RUN_REFERENCE_HM=1")
    return None, "prebuilt HM encoder not on this box"


LDP_THR_LINE = "0.4 0.6 0.3 0.7 0.2 0.8"   # HM-16.5_Test_LDP/bin/Thr_info.txt (down, up per depth)


def materialize(dst_dir: str, kind: str = "AI", thr_line: str = None) -> List[str]:
    """Populate dst_dir with every available checkpoint of `kind` (+ Thr_info.txt) the way the encoder's
    working directory holds them.  Returns the model prefixes now present."""
    os.makedirs(dst_dir, exist_ok=True)
    names = list(AI_MODELS.values()) if kind == "AI" else [LDP_MODEL] + list(LDP_LSTM_MODELS.values())
    present = []
    for name in names:
        done = False
        for src_dir in _source_dirs(kind):
            if all(os.path.exists(os.path.join(src_dir, name + suf)) for suf in SUFFIXES):
                for suf in SUFFIXES:
                    d = os.path.join(dst_dir, name + suf)
                    if not os.path.exists(d):
                        os.symlink(os.path.join(src_dir, name + suf), d)
                done = True
                break
        if not done and name in NPZ and os.path.exists(os.path.join(GOLDEN, NPZ[name])):
            z = np.load(os.path.join(GOLDEN, NPZ[name]))
            tf_bundle.write_bundle(os.path.join(dst_dir, name), {k: z[k] for k in z.files})
            done = True
        if done:
            present.append(name)
    with open(os.path.join(dst_dir, "Thr_info.txt"), "w") as f:
        f.write(thr_line if thr_line is not None else ("0.5 0.5 0.5 0.5 0.5 0.5" if kind == "AI" else LDP_THR_LINE))
    return present


def load_weights(name: str) -> Dict[str, np.ndarray]:
    """Checkpoint `name` (a model prefix) as {tensor: float32 array} from whichever source exists."""
    kind = "LDP" if name.startswith("model_LDP") else "AI"
    for src_dir in _source_dirs(kind):
        if all(os.path.exists(os.path.join(src_dir, name + suf)) for suf in SUFFIXES):
            return tf_bundle.read_bundle(os.path.join(src_dir, name))
    if name in NPZ and os.path.exists(os.path.join(GOLDEN, NPZ[name])):
        z = np.load(os.path.join(GOLDEN, NPZ[name]))
        return {k: z[k] for k in z.files}
    raise FileNotFoundError("checkpoint %s is not available on this box" % name)


def available_ai_qps() -> List[int]:
    out = []
    for qp, name in AI_MODELS.items():
        try:
            load_weights(name)
            out.append(qp)
        except FileNotFoundError:
            pass
    return out


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser(description="stage reference-derived checker assets under oracle/_ref (git-ignored)")
    ap.add_argument("--stage-hm", action="store_true", help="copy the reference's prebuilt HM encoders + cfg files")
    ap.add_argument("--allow-execute", action="store_true", help="opt in to RUNNING those prebuilt binaries in the tests")
    ap.add_argument("--stage-data", action="store_true", help="copy the 15 000 labelled demo CTUs")
    a = ap.parse_args()
    print("checkpoints:", stage_from_reference())
    if a.stage_data:
        print("demo data files:", stage_demo_data())
    if a.stage_hm:
        print("HM files:", stage_hm(a.allow_execute))
