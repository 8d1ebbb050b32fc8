"""Mirror of the constants and helpers of HM-16.5_Test_AI/bin/net_CNN.py that callers of the path
touch.  The graph itself (net_CNN.py:103-195) lives in csrc/*.cu."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .binding import EthCnnError, load_library

IMAGE_SIZE = 64                       # net_CNN.py:8
NUM_CHANNELS = 1                      # net_CNN.py:10
NUM_EXT_FEATURES = 1                  # net_CNN.py:12
NUM_LABEL_BYTES = 16                  # net_CNN.py:13
NUM_CONVLAYER_FLAT_FILTERS = 2688     # net_CNN.py:27


def get_thresholds(thr_file: str):
    """net_CNN.py:38-45 -- (THR_L1_LOWER, THR_L2_LOWER) = tokens [1] and [3] of the first line, parsed by
    the library's own reader (so tests exercise the product code)."""
    lib = load_library()
    thr = np.zeros(2, dtype=np.float32)
    rc = lib.ethcnn_debug_read_thresholds(thr_file.encode(), C.c_void_p(thr.ctypes.data))
    if rc != 0:
        raise EthCnnError(rc, lib.ethcnn_last_error().decode("utf-8", "replace"))
    return float(thr[0]), float(thr[1])
