"""A few device-resident passes over a 1080p x 50 clip with the tcgen05 conv stage (for ncu)."""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ethcnn_b200 as eb  # noqa: E402
from oracle import assets  # noqa: E402
from oracle import ethcnn_oracle as eo  # noqa: E402

d = tempfile.mkdtemp(prefix="prof_")
assets.materialize(d, "AI")
W, H, nf = 1920, 1080, 50
luma = np.stack([eo.synth_frame(W, H, 70 + (k % 5)) for k in range(nf)])
dev = torch.device("cuda", 0)
dl = torch.from_numpy(luma).to(dev)
out = torch.empty((nf * 510, 21), dtype=torch.float32, device=dev)
with eb.EthCnn(d, None, eb.MODE_AI, device=0) as net:
    net.set_option(eb.OPT_CONV_PATH, int(os.environ.get("CONV_PATH", "1")))
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        net.predict_luma_device(dl.data_ptr(), W, H, W, W * H, nf, 32, out.data_ptr(), s)
    torch.cuda.synchronize()
