#!/usr/bin/env python
"""Run on the GPU box: produce cu_depth.dat files with the CUDA path (through the video_to_cu_depth CLI,
the way HM launches it) for the inputs the HM end-to-end test uses, into gpurun_out/.
tests/test_hm_e2e.py (CPU box, where the prebuilt HM binary lives) compares the bitstreams HM produces
from these against the ones it produces from the oracle's cu_depth.dat."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import assets  # noqa: E402
from oracle import ethcnn_oracle as eo  # noqa: E402

CASES = (("cfg1_768x512_f1_qp32", 768, 512, 1, 1, 32), ("cfg2crop_1920x1080_f2_qp32", 1920, 1080, 2, 300, 32),
         ("pad_200x136_f2_qp32", 200, 136, 2, 100, 32))


def main():
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    work = tempfile.mkdtemp()
    assets.materialize(work, "AI")
    cli = os.path.join(ROOT, "hevc-complexity-reduction_b200", "bin", "video_to_cu_depth")
    for name, w, h, nf, seed, qp in CASES:
        yuv = os.path.join(work, name + ".yuv")
        with open(yuv, "wb") as f:
            f.write(eo.synth_yuv(w, h, nf, seed0=seed))
        subprocess.run([cli, yuv, str(w), str(h), str(qp)], cwd=work, check=True)
        os.replace(os.path.join(work, "cu_depth.dat"), os.path.join(out_dir, "cuda_%s.cu_depth.dat" % name))
        print("wrote", name)
    # LDP: the CUDA predictor over the residue frames the prebuilt LDP encoder produced (tools/ldp_hm_capture.py)
    import numpy as np

    import ethcnn_b200 as eb
    cap = np.load(os.path.join(ROOT, "tests", "golden", "ldp_hm_capture.npz"))
    lwork = tempfile.mkdtemp()
    present = assets.materialize(lwork, "LDP")
    if eo.ldp_lstm_model_prefix(int(cap["qp"])) in present:
        probs, state = [], None
        with eb.EthCnn(lwork, None, eb.MODE_LDP, device=0) as net:
            for i_frame, luma in zip(cap["i_frames"], cap["resi"]):
                prob, state = net.ldp_step(luma, int(cap["qp"]), int(i_frame), state if i_frame > 1 else None)
                probs.append(prob)
        np.save(os.path.join(out_dir, "cuda_ldp_hm_prob.npy"), np.stack(probs))
        print("wrote cuda_ldp_hm_prob.npy")


if __name__ == "__main__":
    main()
