"""Diagnostic for the in-library multi-GPU path (run on a box with >= 2 GPUs): where do the rows of an n_gpus=2 handle
differ from the single-GPU rows, is a handle on device 1 alone identical to device 0, are repeated calls stable?"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ethcnn_b200 as eb  # noqa: E402
from oracle import assets  # noqa: E402
from oracle import ethcnn_oracle as eo  # noqa: E402

d = tempfile.mkdtemp(prefix="diag_")
assets.materialize(d, "AI")
W, H, nf, qp = 1920, 1080, 7, 32
yuv = np.frombuffer(eo.synth_yuv(W, H, nf, seed0=2), np.uint8)


def report(tag, a, b):
    diff = np.abs(a.astype(np.float64) - b.astype(np.float64))
    rows = np.nonzero((a != b).any(axis=1))[0]
    print("%-28s equal=%s differing rows=%d max|d|=%.3g frames=%s first rows=%s" % (
        tag, np.array_equal(a, b), len(rows), diff.max(), sorted(set((rows // 510).tolist())), rows[:8].tolist()))
    if len(rows):
        r = rows[0]
        print("   row", r, "slots differing", np.nonzero(a[r] != b[r])[0].tolist(), a[r][:6], b[r][:6])


with eb.EthCnn(d, None, eb.MODE_AI, n_gpus=1) as n1:
    a = n1.predict_yuv_buffer(yuv, W, H, qp)
    a2 = n1.predict_yuv_buffer(yuv, W, H, qp)
report("dev0 vs dev0 again", a, a2)
with eb.EthCnn(d, None, eb.MODE_AI, device=1) as nd1:
    e = nd1.predict_yuv_buffer(yuv, W, H, qp)
    e2 = nd1.predict_yuv_buffer(yuv, W, H, qp)
report("dev0 vs dev1 alone", a, e)
report("dev1 vs dev1 again", e, e2)
with eb.EthCnn(d, None, eb.MODE_AI, n_gpus=2) as n2:
    b = n2.predict_yuv_buffer(yuv, W, H, qp)
    c = n2.predict_yuv_buffer(yuv, W, H, qp)
report("dev0 vs n_gpus=2", a, b)
report("n_gpus=2 vs again", b, c)
for path in (2, 0):
    with eb.EthCnn(d, None, eb.MODE_AI, device=1) as nd1:
        nd1.set_option(eb.OPT_FC1_PATH, path)
        f = nd1.predict_yuv_buffer(yuv, W, H, qp)
    with eb.EthCnn(d, None, eb.MODE_AI, device=0) as nd0:
        nd0.set_option(eb.OPT_FC1_PATH, path)
        g = nd0.predict_yuv_buffer(yuv, W, H, qp)
    report("fc path %d: dev0 vs dev1" % path, g, f)
