#!/usr/bin/env python
"""Stage-by-stage diagnosis of the CUDA path against the oracle (run on the GPU box):
conv features, FC1 activations (tcgen05 and SIMT), final probabilities.  Writes gpurun_out/diag.json."""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ethcnn_b200 as eb  # noqa: E402
from oracle import assets, tf_bundle  # noqa: E402
from oracle import ethcnn_oracle as eo  # noqa: E402


def main():
    out = {}
    work = tempfile.mkdtemp()
    present = assets.materialize(work, "AI")
    real = assets.AI_MODELS[32] in present
    w = assets.load_weights(assets.AI_MODELS[32]) if real else eo.random_weights(1)
    if not real:
        tf_bundle.write_bundle(os.path.join(work, assets.AI_MODELS[32]), w)
    out["weights"] = "deployed" if real else "synthetic"
    W, H, nf, qp = 1920, 1080, 1, 32   # one frame = one slab = one chunk, so the scratch holds all of it
    yuv = eo.synth_yuv(W, H, nf, seed0=11)
    n = nf * 510
    ctus = np.concatenate([eo.frame_to_ctus(eo.get_Y_for_one_frame(memoryview(yuv), k, W, H)) for k in range(nf)])
    x, q = eo.input_scaling(ctus, qp, eo.MODE_AI, np.float64)
    f64 = eo.conv_features(x, w)
    outs, a1_64 = eo.fc_heads(f64, q, w, return_fc1=True)
    p64 = eo.get_prob(yuv, W, H, qp, w, eo.MODE_AI, (0.5, 0.5), dtype=np.float64)
    p32 = eo.get_prob(yuv, W, H, qp, w, eo.MODE_AI, (0.5, 0.5))
    out["oracle_f32_vs_f64_max_dp"] = float(np.abs(p32 - p64).max())
    for path, name in ((0, "simt"), (1, "tcgen05"), (2, "fused")):
        try:
            with eb.EthCnn(work, None, eb.MODE_AI, device=0) as net:
                net.set_option(1, path)
                t = time.time()
                got = net.predict_yuv_buffer(np.frombuffer(yuv, np.uint8), W, H, qp)
                dt = time.time() - t
                feat = net.debug_read_scratch(0, n)
                fc1 = net.debug_read_scratch(1, n) if path != 2 else a1_64.astype(np.float32)  # fused: a1 stays on chip
            r = {
                "first_call_s": dt,
                "feat_max_abs_err": float(np.abs(feat - f64).max()),
                "feat_max_abs": float(np.abs(f64).max()),
                "fc1_max_abs_err": float(np.abs(fc1 - a1_64).max()),
                "fc1_max_abs": float(np.abs(a1_64).max()),
                "prob_max_abs_err_vs_f64": float(np.abs(got - p64).max()),
                "prob_max_abs_err_vs_f32": float(np.abs(got - p32).max()),
                "decision_flips_vs_f32": int((eo.decisions(got) != eo.decisions(p32)).sum()),
                "gate_zero_mismatch": int(((got == 0) != (p32 == 0)).sum()),
                "nan": int(np.isnan(got).sum()),
            }
            if r["feat_max_abs_err"] > 1e-3:
                bad = np.argwhere(np.abs(feat - f64) > 1e-3)
                r["feat_bad_count"] = int(len(bad))
                r["feat_bad_first"] = [[int(a), int(b), float(feat[a, b]), float(f64[a, b])] for a, b in bad[:12]]
                r["feat_bad_cols_hist"] = np.histogram(bad[:, 1], bins=[0, 512, 640, 672, 2208, 2592, 2688])[0].tolist()
            if r["fc1_max_abs_err"] > 1e-3:
                bad = np.argwhere(np.abs(fc1 - a1_64) > 1e-3)
                r["fc1_bad_count"] = int(len(bad))
                r["fc1_bad_first"] = [[int(a), int(b), float(fc1[a, b]), float(a1_64[a, b])] for a, b in bad[:12]]
        except Exception as e:  # keep going: the other path may still tell us something
            r = {"error": repr(e)}
        out[name] = r
        print(name, json.dumps(r)[:1500], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "diag.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
