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
