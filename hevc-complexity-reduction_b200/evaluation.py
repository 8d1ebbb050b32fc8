"""Evaluation harness of the reference for scoring predicted split probabilities against the
ground-truth CU depths of labelled CTUs: confusion matrices per level, accuracy and "tendency".

Mirrors ETH-CNN_Training_AI/train_CNN_CTU64.py:103-147 (get_class_matrices, get_tendency_2x2) and :159-214
(get_accuracy_on_large_data) -- same names, same argument meaning -- vectorised with numpy instead of per-sample Python
loops.  It consumes [n, 21] probability rows in the cu_depth.dat layout (what the CUDA path emits) or the reference's three
arrays.  Sample layout of the labelled sets: ETH-CNN_Training_AI/input_data.py:16,92-116 (4992 bytes: 4096 luma, 64 info,
52 label rows of 16 depths; the row of QP q starts at byte 4160 + 16 q).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np

DEFAULT_THR_LIST = (0.5, 1.5, 2.5)          # input_data.py:18: label depth thresholds per level
INDEX_32_LIST = ((0, 1, 4, 5), (2, 3, 6, 7), (8, 9, 12, 13), (10, 11, 14, 15))   # train_CNN_CTU64.py:121
NUM_SAMPLE_LENGTH = 4992
LABEL_OFFSET, NUM_LABEL_BYTES = 4160, 16


def read_samples(path: str, qp: int) -> Tuple[np.ndarray, np.ndarray]:
    """(luma uint8 [n, 64, 64], labels uint8 [n, 16]) of a *.dat_shuffled file for one QP (input_data.py:92-109)."""
    raw = np.fromfile(path, dtype=np.uint8)
    assert raw.size % NUM_SAMPLE_LENGTH == 0
    raw = raw.reshape(-1, NUM_SAMPLE_LENGTH)
    lo = LABEL_OFFSET + qp * NUM_LABEL_BYTES
    return raw[:, :4096].reshape(-1, 64, 64).copy(), raw[:, lo:lo + NUM_LABEL_BYTES].copy()


def split_rows(prob_rows: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """[n, 21] cu_depth rows -> (y64 [n, 1], y32 [n, 4], y16 [n, 16]) as the reference's sess.run returns them."""
    p = np.asarray(prob_rows).reshape(-1, 21)
    return p[:, 0:1], p[:, 1:5], p[:, 5:21]


def get_class_matrices(y_truth, y_predict_64, y_predict_32, y_predict_16, thr_list: Sequence[float]):
    """train_CNN_CTU64.py:103-137.  y_truth [n, 16] depths 0..3; returns three 2x2 matrices [[n00, n01], [n10, n11]]
    (n_xy: ground truth x, prediction y).  Level 32 is only scored inside CTUs whose truth is split at level 64, level 16 only
    inside 32x32 CUs whose truth is split -- the prediction of the coarser level plays no role."""
    y_truth = np.asarray(y_truth, dtype=np.float64).reshape(-1, 16)
    p64 = np.asarray(y_predict_64, dtype=np.float64).reshape(len(y_truth), -1).mean(axis=1)
    p32 = np.asarray(y_predict_32, dtype=np.float64).reshape(-1, 4)
    p16 = np.asarray(y_predict_16, dtype=np.float64).reshape(-1, 16)
    idx = np.array(INDEX_32_LIST)                                        # [4, 4]
    t64 = y_truth.mean(axis=1) > DEFAULT_THR_LIST[0]
    c64 = p64 > thr_list[0]
    t32 = y_truth[:, idx].mean(axis=2) > DEFAULT_THR_LIST[1]             # [n, 4]
    c32 = p32 > thr_list[1]
    t16 = y_truth[:, idx] > DEFAULT_THR_LIST[2]                          # [n, 4, 4], element (j, k) = CU index_32_list[j][k]
    c16 = p16[:, idx] > thr_list[2]
    m32 = np.broadcast_to(t64[:, None], t32.shape)                       # scored where the 64x64 truth is split
    m16 = np.broadcast_to((t64[:, None] & t32)[:, :, None], t16.shape)   # ... and the 32x32 truth is split

    def matrix(truth, pred, mask):
        return [[int(np.sum(mask & ~truth & ~pred)), int(np.sum(mask & ~truth & pred))],
                [int(np.sum(mask & truth & ~pred)), int(np.sum(mask & truth & pred))]]
    return matrix(t64, c64, np.ones_like(t64)), matrix(t32, c32, m32), matrix(t16, c16, m16)


def get_tendency_2x2(m) -> float:
    """train_CNN_CTU64.py:139-147."""
    if m[0][1] == 0 and m[1][0] == 0:
        return 0
    if m[0][1] == 0 or m[1][1] == 0:
        return -100
    if m[1][0] == 0 or m[0][0] == 0:
        return 100
    return -math.log10((m[0][0] / m[0][1]) / (m[1][1] / m[1][0]))


def get_accuracy_on_large_data(prob_rows: np.ndarray, labels: np.ndarray, thr_list: Sequence[float] = (0.5, 0.5, 0.5)) -> Dict[str, object]:
    """train_CNN_CTU64.py:159-214 (the part that does not need the network): matrices summed over the set, per-level
    accuracy (n00 + n11) / sum and tendency."""
    y64, y32, y16 = split_rows(prob_rows)
    ms = get_class_matrices(labels, y64, y32, y16, thr_list)
    acc = [(m[0][0] + m[1][1]) / max(1, m[0][0] + m[0][1] + m[1][0] + m[1][1]) for m in ms]
    return {"matrices": ms, "accuracy": acc, "tendency": [get_tendency_2x2(m) for m in ms]}


def evaluate(predict_ctus, sets: Dict[str, str], qps: Sequence[int] = (22, 27, 32, 37), thr_list: Sequence[float] = (0.5, 0.5, 0.5)) -> List[dict]:
    """train_CNN_CTU64.py:275-281 for a predictor `predict_ctus(luma [n,64,64] uint8, qp) -> [n, 21]` (e.g. EthCnn.predict_ctus on
    a handle whose Thr_info.txt keeps the batch gates open): one record per (set, QP)."""
    out = []
    for name, path in sets.items():
        for qp in qps:
            luma, labels = read_samples(path, qp)
            r = get_accuracy_on_large_data(predict_ctus(luma, qp), labels, thr_list)
            r.update({"set": name, "qp": qp, "n": int(len(luma))})
            out.append(r)
    return out
