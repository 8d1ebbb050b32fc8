"""Diagnostic for the tcgen05 conv stage (ETHCNN_OPT_CONV_PATH = 1): features and probabilities against the mma.sync conv
stage on the same handle, segment by segment, then a timing comparison on a 1080p clip."""
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
SEG = (("c3_S", 0, 512), ("c3_M", 512, 640), ("c3_L", 640, 672), ("c2_S", 672, 2208), ("c2_M", 2208, 2592), ("c2_L", 2592, 2688))


def run(net, path, luma, W, H, nf, qp):
    net.set_option(eb.OPT_CONV_PATH, path)
    prob = net.predict_luma(luma, W, H, nf, qp)
    r, c = eb.ctu_grid(W, H)
    n = min(nf * r * c, 37888)
    return prob, net.debug_read_scratch(0, n)


for (W, H, nf) in ((64, 64, 1), (768, 512, 2), (1920, 1080, 3)):
    luma = np.stack([eo.synth_frame(W, H, 40 + k) for k in range(nf)])
    with eb.EthCnn(d, None, eb.MODE_AI, device=0) as net:      # a handle per path: fresh (zeroed) feature buffers
        p0, f0 = run(net, 0, luma, W, H, nf, 32)
    with eb.EthCnn(d, None, eb.MODE_AI, device=0) as net:
        p1, f1 = run(net, 1, luma, W, H, nf, 32)
    print("%dx%d x%d: max|dprob| = %.3g   max|dfeat| = %.3g (max|feat| = %.3g)" % (W, H, nf, np.abs(p0 - p1).max(), np.abs(f0 - f1).max(), np.abs(f0).max()))
    for name, a, b in SEG:
        df = np.abs(f0[:, a:b] - f1[:, a:b])
        bad = np.argwhere(df > 1e-3 * max(1e-6, np.abs(f0[:, a:b]).max()))
        print("   %-5s max|d| = %.3g  rows bad = %d of %d   first bad (row, col) = %s  ref %s got %s" % (
            name, df.max(), len(set(bad[:, 0].tolist())), f0.shape[0], bad[:3].tolist(),
            [float(f0[i, a + j]) for i, j in bad[:3]], [float(f1[i, a + j]) for i, j in bad[:3]]))

W, H, nf = 1920, 1080, 50
luma = np.stack([eo.synth_frame(W, H, 70 + (k % 5)) for k in range(nf)])
import torch  # noqa: E402

dev = torch.device("cuda", 0)
dl = torch.from_numpy(luma).to(dev)
out = torch.empty((nf * 510, 21), dtype=torch.float32, device=dev)
res = {}
for path, mask in ((0, 7), (1, 7), (1, 1), (1, 2), (1, 4)):
    os.environ["ETHCNN_TC_PHASES"] = str(mask)
    with eb.EthCnn(d, None, eb.MODE_AI, device=0) as net:
        net.set_option(eb.OPT_CONV_PATH, path)
        s = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            net.predict_luma_device(dl.data_ptr(), W, H, W, W * H, nf, 32, out.data_ptr(), s)
        torch.cuda.synchronize()
        net.profile_enable(True)
        for st in range(4):
            net.profile_read(st, reset=True)
        for _ in range(20):
            net.predict_luma_device(dl.data_ptr(), W, H, W, W * H, nf, 32, out.data_ptr(), s)
        torch.cuda.synchronize()
        ms = [net.profile_read(st, reset=True) for st in range(4)]
        net.profile_enable(False)
        if mask == 7:
            res[path] = out.cpu().numpy().copy()
        print("conv path %d (branch mask %d): conv %.4f ms  fc %.4f ms  gate %.4f ms per 25 500 CTUs" % (path, mask, ms[0][0] / 20, ms[1][0] / 20, ms[3][0] / 20))
print("1080p x50: max|dprob| between conv paths = %.3g" % np.abs(res[0] - res[1]).max())
