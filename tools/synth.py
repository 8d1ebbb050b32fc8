"""Workload generators for the product arm of bench.py and the profiling tools (no dependency on oracle/).

  * synthetic luma / residue frames with the statistics SURVEY.md section 8(d) asks for (procedural multi-scale
    content: sinusoids + per-cell noise + rectangles; residue: clip(128 + Laplace) with flat regions);
  * an encoder-like working directory: the deployed checkpoints (data files staged by __graft_entry__.build() under
    oracle/_ref/checkpoints, or $ETHCNN_MODEL_DIR) + Thr_info.txt; QP ranges whose checkpoint is missing get a synthetic
    checkpoint in the same 36-tensor TF-Saver-V2 layout, written by the small bundle writer below.
"""
from __future__ import annotations

import math
import os
import struct
from typing import Dict, List

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AI_MODELS = {22: "model_2000000_qp20~25.dat", 27: "model_2000000_qp25~30.dat",
             32: "model_2000000_qp30~35.dat", 37: "model_2000000_qp35~40.dat"}
LDP_MODEL = "model_LDP_2000000_qp22~37.dat"
SUFFIXES = (".index", ".data-00000-of-00001")


# ------------------------------------------------------------------------------------------- frames
def synth_frame(width: int, height: int, seed: int) -> np.ndarray:
    """uint8 [height, width]: mid grey + six oriented sinusoids + Gaussian noise whose sigma is drawn per 32x32 cell
    + about one rectangle per 20 000 pixels.  Deterministic in its arguments."""
    rng = np.random.default_rng(seed)
    y = np.arange(height, dtype=np.float32)[:, None]
    x = np.arange(width, dtype=np.float32)[None, :]
    img = np.full((height, width), 128.0, np.float32)
    for _ in range(6):
        f, a = rng.uniform(0.002, 0.08), rng.uniform(0.0, 2.0 * math.pi)
        amp, ph = rng.uniform(5.0, 40.0), rng.uniform(0.0, 2.0 * math.pi)
        img += np.float32(amp) * np.sin(np.float32(2.0 * math.pi * f) * (x * np.float32(math.cos(a)) + y * np.float32(math.sin(a)))
                                        + np.float32(ph))
    cells = rng.choice(np.array([0, 0, 0, 1, 2, 4, 8, 16, 25], np.float32), size=(-(-height // 32), -(-width // 32)))
    sigma = np.kron(cells, np.ones((32, 32), np.float32))[:height, :width]
    img += rng.standard_normal((height, width), dtype=np.float32) * sigma
    for _ in range(max(1, width * height // 20000)):
        w, h = int(rng.integers(4, 201)), int(rng.integers(4, 201))
        x0, y0 = int(rng.integers(0, width)), int(rng.integers(0, height))
        img[y0:y0 + h, x0:x0 + w] += np.float32(rng.uniform(-60.0, 60.0))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def synth_residue_frame(width: int, height: int, seed: int) -> np.ndarray:
    """Residue-like luma of BASELINE config 5: clip(128 + r), r Laplace with a scale drawn per 32x32 cell, a third of
    the cells flat -- what HM's pre-encode writes into resi.yuv (HM-16.5_Test_LDP TEncSearch.cpp:4548-4557)."""
    rng = np.random.default_rng(seed)
    cells = rng.choice(np.array([0, 0, 0.5, 1, 2, 3, 4, 6, 10], np.float32), size=(-(-height // 32), -(-width // 32)))
    scale = np.kron(cells, np.ones((32, 32), np.float32))[:height, :width]
    r = rng.laplace(0.0, 1.0, size=(height, width)).astype(np.float32) * scale
    return np.clip(np.rint(128.0 + r), 0, 255).astype(np.uint8)


def make_clip(width: int, height: int, frames: int, seed0: int, n_base: int = 5, residue: bool = False) -> np.ndarray:
    """[frames, height, width] uint8: n_base generated frames and shifted copies of them (generation is the slow part)."""
    gen = synth_residue_frame if residue else synth_frame
    nb = max(1, min(n_base, frames))
    base = [gen(width, height, seed0 + k) for k in range(nb)]
    out = np.empty((frames, height, width), np.uint8)
    for k in range(frames):
        out[k] = np.roll(base[k % nb], shift=(8 * (k // nb), 16 * (k // nb)), axis=(0, 1))
    return out


# ------------------------------------------------------------------------------------------- checkpoints
def _staged_dirs(kind: str) -> List[str]:
    dirs = []
    if os.environ.get("ETHCNN_MODEL_DIR"):
        dirs.append(os.environ["ETHCNN_MODEL_DIR"])
    dirs.append(os.path.join(REPO, "oracle", "_ref", "checkpoints", kind))   # data files only (staged by build())
    return dirs


def prepare_models(dst_dir: str, kind: str = "AI", thr_line: str = None) -> List[int]:
    """Make dst_dir look like the encoder's bin/ directory.  Returns the QPs (AI) whose checkpoint had to be synthesised."""
    os.makedirs(dst_dir, exist_ok=True)
    names = dict(AI_MODELS) if kind == "AI" else {37: LDP_MODEL}
    synthetic = []
    for qp, name in names.items():
        found = False
        for src in _staged_dirs(kind):
            if all(os.path.exists(os.path.join(src, name + s)) for s in SUFFIXES):
                for s in SUFFIXES:
                    dst = os.path.join(dst_dir, name + s)
                    if not os.path.lexists(dst):
                        os.symlink(os.path.join(src, name + s), dst)
                found = True
                break
        if not found:
            write_bundle(os.path.join(dst_dir, name), random_cnn_weights(100 + qp))
            synthetic.append(qp)
    with open(os.path.join(dst_dir, "Thr_info.txt"), "w") as f:
        f.write(thr_line if thr_line is not None else ("0.5 0.5 0.5 0.5 0.5 0.5" if kind == "AI" else "0.4 0.6 0.3 0.7 0.2 0.8"))
    return synthetic


def random_cnn_weights(seed: int) -> Dict[str, np.ndarray]:
    """A checkpoint with the deployed 36-tensor layout: unnamed conv variables Variable[_k]
    in creation order L, M, S (net_CNN.py:126-141) and the named FC tensors; fan-in scaled so activations stay O(1)."""
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = {}
    k = 0
    for _branch in range(3):
        for shape in ((4, 4, 1, 16), (2, 2, 16, 24), (2, 2, 24, 32)):
            fan = shape[0] * shape[1] * shape[2]
            for arr in ((rng.standard_normal(shape) * (1.6 / math.sqrt(fan))), rng.standard_normal(shape[3]) * 0.05):
                out["Variable" if k == 0 else "Variable_%d" % k] = arr.astype(np.float32)
                k += 1
    for tag, n1, n2, n3 in (("64", 64, 48, 1), ("32", 128, 96, 4), ("16", 256, 192, 16)):
        for stem, ni, no in (("h_fc1__%s__" % tag, 2688, n1), ("h_fc2__%s__" % tag, n1 + 1, n2), ("y_conv_flat__%s__" % tag, n2 + 1, n3)):
            out[stem + "w"] = (rng.standard_normal((ni, no)) * (1.3 / math.sqrt(ni))).astype(np.float32)
            out[stem + "b"] = (rng.standard_normal(no) * 0.05).astype(np.float32)
    return out


# --- minimal TF Saver-V2 bundle writer (one shard, float32 tensors, uncompressed LevelDB-style table)
def _crc32c_tables():
    t = np.zeros((8, 256), np.uint32)
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        t[0, i] = c
    for k in range(1, 8):
        t[k] = t[0][t[k - 1] & 0xFF] ^ (t[k - 1] >> 8)
    return [row.tolist() for row in t]


_T = None


def crc32c(data: bytes) -> int:
    global _T
    if _T is None:
        _T = _crc32c_tables()
    t0, t1, t2, t3, t4, t5, t6, t7 = _T
    crc, n8 = 0xFFFFFFFF, len(data) // 8 * 8
    words = np.frombuffer(data[:n8], "<u4").reshape(-1, 2)
    for lo, hi in zip(words[:, 0].tolist(), words[:, 1].tolist()):
        lo ^= crc
        crc = (t7[lo & 255] ^ t6[(lo >> 8) & 255] ^ t5[(lo >> 16) & 255] ^ t4[lo >> 24]
               ^ t3[hi & 255] ^ t2[(hi >> 8) & 255] ^ t1[(hi >> 16) & 255] ^ t0[hi >> 24])
    for b in data[n8:]:
        crc = t0[(crc ^ b) & 255] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def _masked(crc: int) -> int:
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def _varint(v: int) -> bytes:
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _block(entries) -> bytes:
    """One table block without key sharing (every entry is a restart point), + type byte 0 + masked crc."""
    body, restarts = bytearray(), []
    for key, val in entries:
        restarts.append(len(body))
        body += _varint(0) + _varint(len(key)) + _varint(len(val)) + key + val
    for r in restarts or [0]:
        body += struct.pack("<I", r)
    body += struct.pack("<I", max(1, len(restarts)))
    return bytes(body) + b"\x00" + struct.pack("<I", _masked(crc32c(bytes(body) + b"\x00")))


def write_bundle(prefix: str, tensors: Dict[str, np.ndarray]) -> None:
    names = sorted(tensors, key=lambda s: s.encode())
    blob = bytearray()
    entries = [(b"", b"\x08\x01\x1a\x02\x08\x01")]            # BundleHeaderProto: num_shards 1, version.producer 1
    for name in names:
        raw = np.ascontiguousarray(tensors[name], "<f4").tobytes()
        dims = b"".join(b"\x12" + _varint(len(d)) + d for d in (b"\x08" + _varint(int(n)) for n in tensors[name].shape))
        e = b"\x08\x01" + b"\x12" + _varint(len(dims)) + dims
        if blob:
            e += b"\x20" + _varint(len(blob))
        e += b"\x28" + _varint(len(raw)) + b"\x35" + struct.pack("<I", _masked(crc32c(raw)))
        entries.append((name.encode(), e))
        blob += raw
    data_blk, meta_blk = _block(entries), _block([])
    index_blk = _block([((names[-1].encode() if names else b"") + b"\x00", _varint(0) + _varint(len(data_blk) - 5))])
    footer = (_varint(len(data_blk)) + _varint(len(meta_blk) - 5) + _varint(len(data_blk) + len(meta_blk)) + _varint(len(index_blk) - 5))
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57)
    with open(prefix + ".index", "wb") as f:
        f.write(data_blk + meta_blk + index_blk + footer)
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(blob))
