"""End to end through the UNMODIFIED prebuilt HM encoder of the reference
(/root/reference/HM-16.5_Test_AI/bin/TAppEncoderStatic): HM forks `python video_to_cu_depth.py ...`
(TAppEncCfg.cpp:2319), reads cu_depth.dat and Thr_info.txt (TEncCu.cpp:237-261) and its RDO decisions --
hence the bitstream -- depend on the probabilities.  The bitstream produced from the CUDA-made
cu_depth.dat (tests/golden/cuda_*.cu_depth.dat, generated on the B200 box by tools/make_cuda_fixture.py
through the product's CLI) must be identical to the one produced from the oracle-made file.
Only runs where the prebuilt HM binary exists AND executing it has been opted into explicitly
(`python -m oracle.assets --stage-hm --allow-execute`, or ETHCNN_RUN_REFERENCE_HM=1): it is an opaque third-party ELF
from the reference tree, so a plain `pytest` never runs it by itself (oracle/assets.py:hm_dir)."""
import hashlib
import os
import shutil
import stat
import subprocess

import numpy as np
import pytest

from oracle import assets
from oracle import ethcnn_oracle as eo

REF_BIN, _WHY = assets.hm_dir("AI")
HM = os.path.join(REF_BIN, "TAppEncoderStatic") if REF_BIN else None
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

pytestmark = pytest.mark.skipif(HM is None, reason=_WHY)

CASES = {"cfg1_768x512_f1_qp32": (768, 512, 1, 1, 32), "pad_200x136_f2_qp32": (200, 136, 2, 100, 32),
         "cfg2crop_1920x1080_f2_qp32": (1920, 1080, 2, 300, 32)}

STAND_IN = """import shutil, sys
assert len(sys.argv) == 5
shutil.copyfile({src!r}, 'cu_depth.dat')
"""


def run_hm(work, yuv_path, w, h, nf, qp, cu_depth_src):
    """Encode with the prebuilt HM; its system('python video_to_cu_depth.py ...') call finds a stand-in script
    that drops the given cu_depth.dat into the cwd (the GPU is not available on this box)."""
    os.makedirs(work, exist_ok=True)
    hm = os.path.join(work, "TAppEncoderStatic")
    shutil.copyfile(HM, hm)
    os.chmod(hm, os.stat(hm).st_mode | stat.S_IXUSR)
    shutil.copyfile(os.path.join(REF_BIN, "encoder_intra_main.cfg"), os.path.join(work, "encoder_intra_main.cfg"))
    shutil.copyfile(os.path.join(REF_BIN, "Thr_info.txt"), os.path.join(work, "Thr_info.txt"))
    with open(os.path.join(work, "video_to_cu_depth.py"), "w") as f:
        f.write(STAND_IN.format(src=cu_depth_src))
    cmd = [hm, "-c", "encoder_intra_main.cfg", "-i", yuv_path, "-wdt", str(w), "-hgt", str(h), "-fr", "30", "-f", str(nf),
           "-q", str(qp), "-b", "str.bin", "-o", ""]
    r = subprocess.run(cmd, cwd=work, capture_output=True, timeout=900)
    assert r.returncode == 0, r.stdout.decode()[-2000:] + r.stderr.decode()[-2000:]
    data = open(os.path.join(work, "str.bin"), "rb").read()
    return hashlib.md5(data).hexdigest(), len(data)


@pytest.mark.parametrize("name", sorted(CASES))
def test_hm_bitstream_identical_for_cuda_and_oracle_probabilities(tmp_path, name):
    fixture = os.path.join(GOLDEN, "cuda_%s.cu_depth.dat" % name)
    if not os.path.exists(fixture):
        pytest.skip("no CUDA-made fixture committed yet (tools/make_cuda_fixture.py on the GPU box)")
    w, h, nf, seed, qp = CASES[name]
    yuv = eo.synth_yuv(w, h, nf, seed0=seed)
    yuv_path = str(tmp_path / "in.yuv")
    open(yuv_path, "wb").write(yuv)
    weights = assets.load_weights(assets.AI_MODELS[qp])
    oracle_dat = str(tmp_path / "oracle.dat")
    p_oracle = eo.get_prob(yuv, w, h, qp, weights, eo.MODE_AI, (0.5, 0.5))
    p_oracle.astype("<f4").tofile(oracle_dat)
    p_cuda = np.fromfile(fixture, dtype="<f4").reshape(-1, 21)
    assert p_cuda.shape == p_oracle.shape
    assert np.abs(p_cuda - p_oracle).max() <= 1e-4
    assert np.array_equal(eo.decisions(p_cuda), eo.decisions(p_oracle))
    md5_o, size_o = run_hm(str(tmp_path / "o"), yuv_path, w, h, nf, qp, oracle_dat)
    md5_c, size_c = run_hm(str(tmp_path / "c"), yuv_path, w, h, nf, qp, fixture)
    assert (md5_c, size_c) == (md5_o, size_o)
    # and HM really is sensitive to the file: constant probabilities give a different bitstream
    const_dat = str(tmp_path / "const.dat")
    np.ones_like(p_oracle).astype("<f4").tofile(const_dat)
    md5_k, _ = run_hm(str(tmp_path / "k"), yuv_path, w, h, nf, qp, const_dat)
    assert md5_k != md5_o


LDP_DIR, _LDP_WHY = assets.hm_dir("LDP")


@pytest.mark.skipif(LDP_DIR is None, reason=_LDP_WHY)
def test_hm_ldp_bitstream_identical_with_cuda_made_answers(tmp_path):
    """Inter mode: the unmodified prebuilt LDP encoder (two passes per frame, file-signal handshake, README.md:64-84)
    answered (a) by the oracle and (b) by the probabilities the CUDA predictor computed on the B200 for the very residue
    frames HM produced (tests/golden/cuda_ldp_hm_prob.npy via tools/make_cuda_fixture.py).  The residue frames of run
    (b) must equal the recorded ones frame by frame and the bitstream must be identical."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import ldp_hm_capture as cap

    g = np.load(os.path.join(GOLDEN, "ldp_hm_capture.npz"))
    fixture = os.path.join(GOLDEN, "cuda_ldp_hm_prob.npy")
    md5_o, size_o, seen_o = cap.run_hm_ldp(str(tmp_path / "o"), cap.oracle_answerer())
    os.makedirs(tmp_path / "o", exist_ok=True)
    assert (md5_o, size_o) == (str(g["str_md5"]), int(g["str_size"]))            # the capture is reproducible
    assert all(np.array_equal(s[1], r) for s, r in zip(seen_o, g["resi"]))
    if not os.path.exists(fixture):
        pytest.skip("no CUDA-made LDP fixture committed yet")
    cuda_prob = np.load(fixture)
    assert np.abs(cuda_prob - g["oracle_prob"]).max() <= 1e-4
    thr6 = (0.6, 0.4, 0.7, 0.3, 0.8, 0.2)
    assert np.array_equal(eo.decisions(cuda_prob, thr6), eo.decisions(g["oracle_prob"], thr6))
    served = {"k": 0}

    def replay(i_frame, fw, fh, q, luma):
        k = served["k"]
        if not (int(g["i_frames"][k]) == i_frame and np.array_equal(luma, g["resi"][k])):
            served.setdefault("diverged_at", i_frame)      # checked after the run (an assert here would hang HM)
        served["k"] += 1
        return cuda_prob[k], np.zeros((cuda_prob[k].shape[0], 1, 2, 448), np.float32)   # HM never reads state.dat
    md5_c, size_c, _ = cap.run_hm_ldp(str(tmp_path / "c"), replay)
    assert "diverged_at" not in served, "HM's residue differs from the recorded run at frame %d" % served["diverged_at"]
    assert (md5_c, size_c) == (md5_o, size_o)
    # sensitivity: all-ones probabilities change the bitstream
    md5_k, _, _ = cap.run_hm_ldp(str(tmp_path / "k"), lambda i, fw, fh, q, luma: (np.ones((cuda_prob[0].shape[0], 21), np.float32),
                                                                                   np.zeros((cuda_prob[0].shape[0], 1, 2, 448), np.float32)))
    assert md5_k != md5_o
