"""TEST INFRASTRUCTURE (oracle) -- literal, loop-by-loop restatement of the reference's evaluation functions
(ETH-CNN_Training_AI/train_CNN_CTU64.py:64-147, input_data.py:18), kept deliberately slow and un-vectorised: it is the
checker for the product's vectorised hevc-complexity-reduction_b200/evaluation.py."""
import math

import numpy as np

DEFAULT_THR_LIST = [0.5, 1.5, 2.5]   # input_data.py:18


def is_sep_64(y_one_sample, thr):    # train_CNN_CTU64.py:64-69
    return 1 if np.mean(y_one_sample) > thr else 0


def is_sep_32(y_one_sample, thr):    # :77-82
    return 1 if np.mean(y_one_sample) > thr else 0


def is_sep_16(y_one_sample, thr):    # :90-94
    return 1 if y_one_sample > thr else 0


def get_class_matrices(y_truth, y_predict_64, y_predict_32, y_predict_16, thr_list):   # :103-137
    matrix_64 = [[0, 0], [0, 0]]
    matrix_32 = [[0, 0], [0, 0]]
    matrix_16 = [[0, 0], [0, 0]]
    assert y_truth.shape[0] == y_predict_16.shape[0]
    index_32_list = [[0, 1, 4, 5], [2, 3, 6, 7], [8, 9, 12, 13], [10, 11, 14, 15]]
    for i in range(y_truth.shape[0]):
        class_64_truth = is_sep_64(y_truth[i], DEFAULT_THR_LIST[0])
        matrix_64[class_64_truth][is_sep_64(y_predict_64[i], thr_list[0])] += 1
        if class_64_truth == 1:
            for j in range(4):
                class_32_truth = is_sep_32(y_truth[i][index_32_list[j]], DEFAULT_THR_LIST[1])
                matrix_32[class_32_truth][is_sep_32(y_predict_32[i][j], thr_list[1])] += 1
                if class_32_truth == 1:
                    for k in range(4):
                        class_16_truth = is_sep_16(y_truth[i][index_32_list[j][k]], DEFAULT_THR_LIST[2])
                        matrix_16[class_16_truth][is_sep_16(y_predict_16[i][index_32_list[j][k]], thr_list[2])] += 1
    return matrix_64, matrix_32, matrix_16


def get_tendency_2x2(m):             # :139-147
    if m[0][1] == 0 and m[1][0] == 0:
        return 0
    elif m[0][1] == 0 or m[1][1] == 0:
        return -100
    elif m[1][0] == 0 or m[0][0] == 0:
        return 100
    return -math.log10((m[0][0] / m[0][1]) / (m[1][1] / m[1][0]))
