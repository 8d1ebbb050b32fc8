"""End to end through the UNMODIFIED prebuilt HM encoder of the reference
(/root/reference/HM-16.5_Test_AI/bin/TAppEncoderStatic): HM forks `python video_to_cu_depth.py ...`
(TAppEncCfg.cpp:2319), reads cu_depth.dat and Thr_info.txt (TEncCu.cpp:237-261) and its RDO decisions --
hence the bitstream -- depend on the probabilities.  The bitstream produced from the CUDA-made
cu_depth.dat (tests/golden/cuda_*.cu_depth.dat, generated on the B200 box by tools/make_cuda_fixture.py
through the product's CLI) must be identical to the one produced from the oracle-made file.
Only runs where the reference tree (and hence the HM binary) exists."""
import hashlib
import os
import shutil
import stat
import subprocess

import numpy as np
import pytest

from oracle import assets
from oracle import ethcnn_oracle as eo

REF_BIN = "/root/reference/HM-16.5_Test_AI/bin"
HM = os.path.join(REF_BIN, "TAppEncoderStatic")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

pytestmark = pytest.mark.skipif(not os.path.exists(HM), reason="prebuilt HM encoder not on this box")

CASES = {"cfg1_768x512_f1_qp32": (768, 512, 1, 1, 32), "pad_200x136_f2_qp32": (200, 136, 2, 100, 32)}

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
