"""Wall clock of the drop-in as the HM encoder uses it (TAppEncCfg.cpp:2319: `python video_to_cu_depth.py <yuv> <W> <H> <QP>`
in the encoder's working directory) and of the C++ CLI, on synthetic BASELINE-sized files, next to the reference-style CPU
port (oracle, one process per core) on the same box.  Everything is inside the timed region: process start, CUDA context,
checkpoint parsing / packing / upload, file read, H2D, kernels, D2H, the cu_depth.dat write."""
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import synth  # noqa: E402

PKG = os.path.join(ROOT, "hevc-complexity-reduction_b200")


def make_file(path, W, H, nf):
    base = [synth.synth_frame(W, H, 500 + k) for k in range(5)]
    uv = bytes([128]) * (W * H // 2)
    with open(path, "wb") as f:
        for k in range(nf):
            f.write(base[k % 5].tobytes())
            f.write(uv)


def main():
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="config1,config2,config3")
    ap.add_argument("--skip-inprocess", action="store_true", help="only the resident-server clients (ETHCNN_GPUS applies to the server)")
    args = ap.parse_args()
    work = tempfile.mkdtemp(prefix="cli_wall_")
    synth.prepare_models(work)
    os.symlink(os.path.join(PKG, "video_to_cu_depth.py"), os.path.join(work, "video_to_cu_depth.py"))   # as INTEGRATION.md says
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + PKG)
    cases = [c for c in (("config1", 768, 512, 1, 32), ("config2", 1920, 1080, 50, 32), ("config3", 4928, 3264, 50, 32),
                         ("config4", 2880, 1920, 425, 27)) if c[0] in args.cases.split(",")]
    print("ETHCNN_GPUS = %s" % os.environ.get("ETHCNN_GPUS", "1"))
    # a resident server (video_to_cu_depth --serve): the clients below hand their request to it
    sock = os.path.join(work, "ethcnn.sock")
    srv = subprocess.Popen([os.path.join(PKG, "bin", "video_to_cu_depth"), "--serve", sock], cwd=work, stderr=subprocess.DEVNULL)
    for _ in range(1200):
        if os.path.exists(sock):
            break
        time.sleep(0.05)
    env_srv = dict(env, ETHCNN_SERVER=sock)
    for name, W, H, nf, qp in cases:
        yuv = os.path.join(work, name + ".yuv")
        make_file(yuv, W, H, nf)
        r, c = -(-H // 64), -(-W // 64)
        n = nf * r * c
        py = [sys.executable, "video_to_cu_depth.py", yuv, str(W), str(H), str(qp)]
        cc = [os.path.join(PKG, "bin", "video_to_cu_depth"), yuv, str(W), str(H), str(qp)]
        runs = (("python shim", py, env), ("C++ CLI", cc, env), ("shim->server", py, env_srv), ("CLI->server", cc, env_srv))
        for label, cmd, run_env in (runs[2:] if args.skip_inprocess else runs):
            times = []
            for _ in range(3):
                out = os.path.join(work, "cu_depth.dat")
                if os.path.exists(out):
                    os.remove(out)
                t = time.time()
                rc = subprocess.run(cmd, cwd=work, env=run_env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
                times.append(time.time() - t)
                assert rc.returncode == 0, rc.stderr.decode()[-400:]
                assert os.path.getsize(out) == n * 84
            print("%-8s %4dx%-4d x%-3d %7d CTUs  %-12s wall %.3f s (best of 3: %s)  %.3g CTU/s" % (
                name, W, H, nf, n, label, min(times), " ".join("%.3f" % t for t in times), n / min(times)))
        os.remove(yuv)
    subprocess.run([os.path.join(PKG, "bin", "video_to_cu_depth"), "--quit", sock])
    srv.wait(timeout=30)
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
