"""How much of the conv stage's exact hi / lo operand splitting does the 1e-4 contract really need?  Numpy emulation on the
reference's labelled CTUs (fp64 everywhere except where stated): round the conv1 activations (the A operand of conv2), the
conv2 features as conv3's A operand, or the features handed to FC1 to a single fp16 value and look at the probability error
and at HM decision flips.  Background for DESIGN.md section 3.1b: the epilogues' split (about 2.5 of 6 instructions per
activation) is what keeps the conv stage SIMT-bound.  Needs /root/reference; CPU only."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import assets  # noqa: E402
from oracle import ethcnn_oracle as eo  # noqa: E402


def f16(v):
    return v.astype(np.float16).astype(np.float64)


def features(x, w, round_c1=False, round_c2_for_c3=False):
    b = x.shape[0]
    feats = {}
    for br, img in (("L", eo.zero_mean_norm_local(eo.aver_pool(x, 4))), ("M", eo.zero_mean_norm_local(eo.aver_pool(x, 2))),
                    ("S", eo.zero_mean_norm_local(x))):
        v = eo.BRANCH_VARS[br]
        c1 = eo.non_overlap_conv(img[..., None], w[eo._vname(v)], w[eo._vname(v + 1)])
        if round_c1:
            c1 = f16(c1 * 64.0) / 64.0
        c2 = eo.non_overlap_conv(c1, w[eo._vname(v + 2)], w[eo._vname(v + 3)])
        c2in = f16(c2 * 64.0) / 64.0 if round_c2_for_c3 else c2
        c3 = eo.non_overlap_conv(c2in, w[eo._vname(v + 4)], w[eo._vname(v + 5)])
        feats[br] = (c2.reshape(b, -1), c3.reshape(b, -1))
    return np.concatenate([feats["S"][1], feats["M"][1], feats["L"][1], feats["S"][0], feats["M"][0], feats["L"][0]], axis=1)


def main():
    d = "/root/reference/ETH-CNN_Training_AI/Data/AI_Test_5000.dat_shuffled"
    ctus = np.fromfile(d, np.uint8).reshape(-1, 4992)[:3000, :4096].reshape(-1, 64, 64)
    for qp in (22, 32, 37):
        w = {k: v.astype(np.float64) for k, v in assets.load_weights(assets.AI_MODELS[qp]).items()}
        x, q = eo.input_scaling(ctus, qp, eo.MODE_AI, np.float64)
        ref = np.concatenate(eo.fc_heads(features(x, w), q, w), axis=1)
        out = []
        for name, kw, round_feat in (("conv1 activations as one fp16", dict(round_c1=True), False),
                                     ("conv3's input as one fp16", dict(round_c2_for_c3=True), False),
                                     ("features as one fp16 (FC1's A operand)", {}, True)):
            f = features(x, w, **kw)
            if round_feat:
                f = f16(f * 64.0) / 64.0
            p = np.concatenate(eo.fc_heads(f, q, w), axis=1)
            out.append("%s: max|dp| %.2g, flips %d" % (name, np.abs(p - ref).max(), int(((p > 0.5) != (ref > 0.5)).sum())))
        print("qp %d, %d probabilities | %s" % (qp, ref.size, " | ".join(out)))


if __name__ == "__main__":
    main()
