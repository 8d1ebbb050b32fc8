"""The LIVE chain on a GPU box: the reference's unmodified prebuilt HM encoder forks `python video_to_cu_depth.py <yuv> <W> <H>
<QP>` (TAppEncCfg.cpp:2317-2321), finds OUR drop-in of that name in its working directory, which calls libethcnn_b200.so on the
GPU (in-process, or through the resident server when ETHCNN_SERVER names its socket), and HM then reads the cu_depth.dat it wrote
(TEncCu.cpp:237-261).  The bitstream must be identical to the one HM produces when fed the ORACLE's cu_depth.dat.
Needs the staged HM binary and the explicit opt-in to run it (oracle/assets.py:hm_dir)."""
import hashlib
import os
import shutil
import stat
import subprocess
import sys
import time

import numpy as np
import pytest

from oracle import assets
from oracle import ethcnn_oracle as eo

HM_DIR, _WHY = assets.hm_dir("AI")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(HM_DIR is None, reason=_WHY)]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "hevc-complexity-reduction_b200", "video_to_cu_depth.py")
CLI = os.path.join(ROOT, "hevc-complexity-reduction_b200", "bin", "video_to_cu_depth")

CASES = {"cfg1_768x512_f1_qp32": (768, 512, 1, 1, 32), "cfg2crop_1920x1080_f2_qp32": (1920, 1080, 2, 300, 32)}
STAND_IN = "import shutil, sys\nassert len(sys.argv) == 5\nshutil.copyfile(%r, 'cu_depth.dat')\n"


def encoder_dir(path, script_text=None):
    """A directory laid out like HM-16.5_Test_AI/bin: encoder, cfg, Thr_info.txt, checkpoints and a video_to_cu_depth.py."""
    os.makedirs(path, exist_ok=True)
    assets.materialize(path, "AI")
    hm = os.path.join(path, "TAppEncoderStatic")
    shutil.copyfile(os.path.join(HM_DIR, "TAppEncoderStatic"), hm)
    os.chmod(hm, os.stat(hm).st_mode | stat.S_IXUSR)
    shutil.copyfile(os.path.join(HM_DIR, "encoder_intra_main.cfg"), os.path.join(path, "encoder_intra_main.cfg"))
    shutil.copyfile(os.path.join(HM_DIR, "Thr_info.txt"), os.path.join(path, "Thr_info.txt"))
    script = os.path.join(path, "video_to_cu_depth.py")
    if script_text is None:
        os.symlink(SHIM, script)          # the product's drop-in (it locates the package next to the real file)
    else:
        open(script, "w").write(script_text)
    return hm


def encode(work, yuv_path, w, h, nf, qp, env=None):
    cmd = [os.path.join(work, "TAppEncoderStatic"), "-c", "encoder_intra_main.cfg", "-i", yuv_path, "-wdt", str(w), "-hgt", str(h),
           "-fr", "30", "-f", str(nf), "-q", str(qp), "-b", "str.bin", "-o", ""]
    e = dict(os.environ)
    e["PATH"] = os.path.dirname(sys.executable) + os.pathsep + e.get("PATH", "")    # HM runs `python ...` through system()
    e.pop("ETHCNN_SERVER", None)
    e.update(env or {})
    r = subprocess.run(cmd, cwd=work, capture_output=True, timeout=1200, env=e)
    assert r.returncode == 0, r.stdout.decode()[-2000:] + r.stderr.decode()[-2000:]
    data = open(os.path.join(work, "str.bin"), "rb").read()
    return hashlib.md5(data).hexdigest(), len(data), r.stdout.decode()


@pytest.mark.parametrize("name", sorted(CASES))
def test_hm_runs_our_drop_in_on_the_gpu_and_the_bitstream_matches_the_oracle_fed_run(tmp_path, name):
    w, h, nf, seed, qp = CASES[name]
    yuv = eo.synth_yuv(w, h, nf, seed0=seed)
    yuv_path = str(tmp_path / "in.yuv")
    open(yuv_path, "wb").write(yuv)
    p_oracle = eo.get_prob(yuv, w, h, qp, assets.load_weights(assets.AI_MODELS[qp]), eo.MODE_AI, (0.5, 0.5))
    oracle_dat = str(tmp_path / "oracle.dat")
    p_oracle.astype("<f4").tofile(oracle_dat)
    # (a) HM fed by the oracle through a stand-in script
    wo = str(tmp_path / "oracle_fed")
    encoder_dir(wo, STAND_IN % oracle_dat)
    md5_o, size_o, _ = encode(wo, yuv_path, w, h, nf, qp)
    # (b) HM -> our video_to_cu_depth.py -> libethcnn_b200.so on the GPU, in-process
    wl = str(tmp_path / "live")
    encoder_dir(wl)
    md5_l, size_l, log = encode(wl, yuv_path, w, h, nf, qp)
    assert "Predicting Time" in log                                  # the drop-in ran inside HM's system() call
    p_live = np.fromfile(os.path.join(wl, "cu_depth.dat"), "<f4").reshape(-1, 21)
    assert p_live.shape == p_oracle.shape and np.abs(p_live - p_oracle).max() <= 2e-5
    assert np.array_equal(eo.decisions(p_live), eo.decisions(p_oracle))
    assert (md5_l, size_l) == (md5_o, size_o)
    # (c) the same through the resident server (no CUDA start-up inside HM's child process)
    ws = str(tmp_path / "served")
    encoder_dir(ws)
    sock = str(tmp_path / "s.sock")
    srv = subprocess.Popen([CLI, "--serve", sock], cwd=ws, stderr=subprocess.PIPE)
    try:
        for _ in range(1200):
            if os.path.exists(sock) or srv.poll() is not None:
                break
            time.sleep(0.05)
        assert os.path.exists(sock), "server did not come up"
        md5_s, size_s, log_s = encode(ws, yuv_path, w, h, nf, qp, env={"ETHCNN_SERVER": sock})
        assert (md5_s, size_s) == (md5_o, size_o)
        assert open(os.path.join(ws, "cu_depth.dat"), "rb").read() == open(os.path.join(wl, "cu_depth.dat"), "rb").read()
        assert subprocess.run([CLI, "--quit", sock]).returncode == 0
        assert srv.wait(timeout=30) == 0
        assert b"served 1 requests" in srv.stderr.read()
    finally:
        if srv.poll() is None:
            srv.kill()
