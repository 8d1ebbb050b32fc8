"""The inter-mode (LDP) predictor on the GPU: residual ETH-CNN + one-step ETH-LSTM (ethcnn_ldp_step) and the
file-signal daemon (ethcnn_ldp_serve / bin/resi_to_cu_depth_LDP) against the oracle and the vectors produced by the
reference's own unmodified functions."""
import os
import subprocess
import time

import numpy as np
import pytest

from oracle import assets, tf_bundle
from oracle import ethcnn_oracle as eo

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
THR = (0.6, 0.7)
THR6 = (0.6, 0.4, 0.7, 0.3, 0.8, 0.2)   # LDP order in the file is down,up; decisions() wants up,down


def close(got, want, tol, what):
    err = np.abs(got.astype(np.float64) - want.astype(np.float64)).max()
    assert err <= tol, "%s: %g" % (what, err)


def test_ldp_step_matches_reference_sequence(eb, ldp_model_dir, golden_dir):
    d, present = ldp_model_dir
    g = np.load(os.path.join(golden_dir, "ldp_lstm_200x136_f5.npz"))
    with eb.EthCnn(d, None, eb.MODE_LDP, device=0) as net:
        state = None
        done = 0
        for k, (luma, qp) in enumerate(zip(g["frames"], g["qps"]), start=1):
            if eo.ldp_lstm_model_prefix(int(qp)) not in present:
                break
            prob, state_out = net.ldp_step(luma, int(qp), k, state)
            close(prob, g["cu_depth"][k - 1], 2e-5, "frame %d prob" % k)
            close(state_out, g["state"][k - 1], 1e-4, "frame %d state" % k)
            assert np.array_equal(eo.decisions(prob, THR6), eo.decisions(g["cu_depth"][k - 1], THR6))
            assert np.array_equal(prob == 0, g["cu_depth"][k - 1] == 0)
            state = g["state"][k - 1]
            done += 1
        assert done >= 4


def test_ldp_step_gates_and_mini_batches(eb, ldp_model_dir, golden_dir):
    d, present = ldp_model_dir
    g = np.load(os.path.join(golden_dir, "ldp_lstm_gates.npz"))
    with eb.EthCnn(d, None, eb.MODE_LDP, device=0) as net:
        for key, qp in (("flat", 37), ("big", 27)):
            if eo.ldp_lstm_model_prefix(qp) not in present:
                continue
            prob, state = net.ldp_step(g[key], qp, 1, None)
            close(prob, g[key + "_cu_depth"], 2e-5, key)
            assert np.array_equal(prob == 0, g[key + "_cu_depth"] == 0), key
            close(state, g[key + "_state"], 1e-4, key + " state")


def test_ldp_step_synthetic_weights_vs_oracle(eb, tmp_path):
    d = str(tmp_path)
    cnn, lstm = eo.random_weights(5), eo.random_lstm_weights(6)
    tf_bundle.write_bundle(os.path.join(d, assets.LDP_MODEL), cnn)
    for name in assets.LDP_LSTM_MODELS.values():
        tf_bundle.write_bundle(os.path.join(d, name), lstm)
    open(os.path.join(d, "Thr_info.txt"), "w").write(assets.LDP_THR_LINE)
    w, h = 712, 328
    with eb.EthCnn(d, None, eb.MODE_LDP, device=0) as net:
        state_o = state_c = None
        for k in range(1, 4):
            luma = eo.synth_residue_frame(w, h, 40 + k)
            want_p, state_o = eo.ldp_predict_frame(luma, 27, k, state_o, cnn, lstm, THR)
            got_p, state_c = net.ldp_step(luma, 27, k, state_c)
            close(got_p, want_p, 3e-5, "synthetic frame %d" % k)
            close(state_c, state_o, 2e-4, "synthetic state %d" % k)
        with pytest.raises(eb.EthCnnError):
            net.ldp_step(luma, 27, 4, np.zeros(10, np.float32))


def test_ldp_daemon_file_protocol(eb, ldp_model_dir, tmp_path):
    """A stand-in for the HM side (TEncGOP.cpp(LDP):1471-1505): write resi.yuv + command.dat, raise pred_start.sig,
    wait for pred_end.sig, read cu_depth.dat -- against the CLI daemon started in the encoder's directory."""
    d, present = ldp_model_dir
    if eo.ldp_lstm_model_prefix(32) not in present:
        pytest.skip("LSTM checkpoint for QP 32 not on this box")
    work = str(tmp_path)
    for fn in os.listdir(d):
        os.symlink(os.path.join(d, fn), os.path.join(work, fn))
    cli = os.path.join(ROOT, "hevc-complexity-reduction_b200", "bin", "resi_to_cu_depth_LDP")
    env = dict(os.environ, ETHCNN_DAEMON_MAX_FRAMES="3", ETHCNN_DAEMON_IDLE_MS="60000")
    proc = subprocess.Popen([cli], cwd=work, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    try:
        cnn = assets.load_weights(assets.LDP_MODEL)
        lw = assets.load_weights(eo.ldp_lstm_model_prefix(32))
        w, h = 416, 240
        state = None
        for i_frame in range(1, 4):
            luma = eo.synth_residue_frame(w, h, 60 + i_frame)
            with open(os.path.join(work, "resi.yuv"), "wb") as f:
                f.write(luma.tobytes() + bytes([128]) * (w * h // 2))
            with open(os.path.join(work, "command.dat"), "w") as f:
                f.write("%d %d %d %d [end]" % (i_frame, w, h, 32))
            open(os.path.join(work, "pred_start.sig"), "w").close()
            t0 = time.time()
            while not os.path.exists(os.path.join(work, "pred_end.sig")):
                assert proc.poll() is None, proc.stderr.read().decode()
                assert time.time() - t0 < 120, "daemon did not answer"
                time.sleep(0.001)
            os.remove(os.path.join(work, "pred_end.sig"))          # HM removes it (TEncGOP.cpp(LDP):1497)
            assert not os.path.exists(os.path.join(work, "pred_start.sig"))
            got = np.fromfile(os.path.join(work, "cu_depth.dat"), "<f4").reshape(-1, 21)
            want, state = eo.ldp_predict_frame(luma, 32, i_frame, state, cnn, lw, THR)
            close(got, want, 3e-5, "daemon frame %d" % i_frame)
            got_state = np.fromfile(os.path.join(work, "state.dat"), "<f4").reshape(-1, 1, 2, 448)
            close(got_state, state, 2e-4, "daemon state %d" % i_frame)
            state = got_state                                        # the daemon reads its own state.dat back
        assert proc.wait(timeout=60) == 0
        assert b"3 frames predicted" in proc.stdout.read()
    finally:
        if proc.poll() is None:
            proc.kill()


def test_ldp_daemon_fails_fast_without_thresholds_or_checkpoint(eb, ldp_model_dir, tmp_path):
    """The HM side waits for pred_end.sig without a time-out (TEncGOP.cpp(LDP):1487): a daemon that cannot work must fail at
    start-up -- like the reference, which dies at import without Thr_info.txt (net_CNN_LSTM_one_step.py:67-68) and at restore
    without the CNN checkpoint -- not after it has consumed pred_start.sig."""
    import os
    import shutil

    d, present = ldp_model_dir
    # (a) checkpoints present, Thr_info.txt missing
    a = tmp_path / "no_thr"
    a.mkdir()
    for fn in os.listdir(d):
        if fn != "Thr_info.txt":
            os.symlink(os.path.realpath(os.path.join(d, fn)), a / fn)
    open(a / "pred_start.sig", "wb").close()
    with eb.EthCnn(str(a), None, eb.MODE_LDP, device=0) as net:
        with pytest.raises(eb.EthCnnError):
            net.ldp_serve(str(a), max_frames=1, idle_timeout_ms=300)
    assert (a / "pred_start.sig").exists()          # untouched: nothing was "accepted"
    # (b) Thr_info.txt present, CNN checkpoint missing
    b = tmp_path / "no_ckpt"
    b.mkdir()
    shutil.copyfile(os.path.join(d, "Thr_info.txt"), b / "Thr_info.txt")
    open(b / "pred_start.sig", "wb").close()
    with eb.EthCnn(str(b), None, eb.MODE_LDP, device=0) as net:
        with pytest.raises(eb.EthCnnError) as e:
            net.ldp_serve(str(b), max_frames=1, idle_timeout_ms=300)
        assert e.value.code == -2
    assert (b / "pred_start.sig").exists()
