#!/usr/bin/env python
"""Drop-in for HM-16.5_Test_AI/bin/video_to_cu_depth.py.

The patched HM encoder hard-codes `python video_to_cu_depth.py <yuv> <W> <H> <QP>`
(TAppEncCfg.cpp:2319), so a file of this name has to exist in the encoder's working directory; copy or
symlink this one there.  It contains no arithmetic: it calls libethcnn_b200.so, which reads
Thr_info.txt and the model_2000000_qp*.dat.* checkpoints from the cwd and writes cu_depth.dat, exactly
like the reference script.  Exit status 0 on success, 1 on any failure (HM asserts on it,
TAppEncCfg.cpp:2321).  There is no CPU fallback.
"""
from __future__ import annotations

import os
import sys
import time

SAVE_FILE = "cu_depth.dat"   # video_to_cu_depth.py:20
IMAGE_SIZE = 64


def _binding():
    try:
        from . import binding  # imported as part of the package
        return binding
    except ImportError:
        # executed as a plain script from the encoder's cwd: locate the package next to the real file
        here = os.path.dirname(os.path.realpath(__file__))
        sys.path.insert(0, os.path.dirname(here))
        sys.path.insert(0, here)
        import binding  # type: ignore
        return binding


def get_file_size(path):
    """video_to_cu_depth.py:39-44"""
    return os.path.getsize(path)


def get_prob(yuv_name, image_size, save_file, qp_seq, n_frames_start, n_frames_end, frame_width, frame_height,
             model_dir=".", n_gpus=None):
    """video_to_cu_depth.py:75-118.  The reference always calls it with n_frames_start = 0 and
    n_frames_end = all frames of the file (:139-143); only that use is supported."""
    b = _binding()
    if image_size != IMAGE_SIZE:
        raise ValueError("image_size must be 64")
    frame_bytes = frame_width * frame_height * 3 // 2
    if n_frames_start != 0 or n_frames_end != get_file_size(yuv_name) // frame_bytes:
        raise ValueError("only whole-file prediction is supported (as the reference script does)")
    if n_gpus is None:
        n_gpus = int(os.environ.get("ETHCNN_GPUS", "1") or "1")
    server = os.environ.get("ETHCNN_SERVER", "")
    if server and b.request(server, yuv_name, frame_width, frame_height, qp_seq, save_file):
        return   # a resident server (video_to_cu_depth --serve) did it: no CUDA start-up in this process
    with b.EthCnn(model_dir, None, b.MODE_AI, n_gpus=n_gpus) as net:
        net.predict_yuv_file(yuv_name, frame_width, frame_height, qp_seq, save_file)


def main(argv=None):
    argv = sys.argv if argv is None else argv
    assert len(argv) == 5                      # video_to_cu_depth.py:120
    yuv_file = argv[1]
    width = int(argv[2])
    height = int(argv[3])
    qp_seq = int(argv[4])
    file_bytes = get_file_size(yuv_file)
    frame_bytes = width * height * 3 // 2
    assert file_bytes % frame_bytes == 0       # video_to_cu_depth.py:137
    t1 = time.time()
    get_prob(yuv_file, IMAGE_SIZE, SAVE_FILE, qp_seq, 0, file_bytes // frame_bytes, width, height,
             model_dir=os.environ.get("ETHCNN_MODEL_DIR", "."))
    t2 = time.time()
    print('--------\n\nPredicting Time: %.3f sec.\n\n--------' % float(t2 - t1))
    return 0


if __name__ == "__main__":
    try:
        sys.exit(main())
    except AssertionError:
        raise
    except Exception as e:  # any failure -> non-zero exit status, message on stderr
        sys.stderr.write("video_to_cu_depth: %s\n" % (e,))
        sys.exit(1)
