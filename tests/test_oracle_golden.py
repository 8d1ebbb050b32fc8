"""The oracle against (a) the vectors produced by the reference's own unmodified scripts on the numpy TF
shim (tests/golden/*.npz, generator oracle/make_golden.py), (b) the known answers of SURVEY.md section 4,
(c) the reference's accuracy logs when its demo data is present."""
import glob
import hashlib
import os

import numpy as np
import pytest

from oracle import assets
from oracle import ethcnn_oracle as eo

REF = "/root/reference"


def _thr(g):
    if "thr_line" in g.files:
        t = str(g["thr_line"]).split(" ")
        return float(t[1]), float(t[3])
    return 0.5, 0.5


def _golden_cases(golden_dir):
    for fn in sorted(glob.glob(os.path.join(golden_dir, "ai_*.npz"))):
        g = np.load(fn)
        for q in g["qps"]:
            yield os.path.basename(fn), g, int(q)


def golden_input(g):
    """(yuv bytes, W, H, thresholds) of an AI golden file."""
    if "yuv" in g.files:
        return g["yuv"].tobytes(), int(g["width"]), int(g["height"]), _thr(g)
    if "ctus" in g.files:  # each CTU is a 64x64 frame; gates forced open by thresholds of -1
        yuv = b"".join(c.tobytes() + bytes([128]) * 2048 for c in g["ctus"])
        return yuv, 64, 64, (-1.0, -1.0)
    yuv = eo.synth_yuv(int(g["width"]), int(g["height"]), int(g["n_frames"]), int(g["seed0"]))
    if hashlib.sha256(yuv).hexdigest() != str(g["yuv_sha256"]):
        pytest.skip("numpy RNG stream changed: recipe no longer reproduces the golden input")
    return yuv, int(g["width"]), int(g["height"]), (0.5, 0.5)


def test_oracle_matches_reference_vectors(golden_dir):
    n = 0
    for name, g, qp in _golden_cases(golden_dir):
        try:
            w = assets.load_weights(assets.AI_MODELS[qp])
        except FileNotFoundError:
            continue
        yuv, W, H, thr = golden_input(g)
        got = eo.get_prob(yuv, W, H, qp, w, eo.MODE_AI, thr)
        ref = g["prob_qp%d" % qp]
        assert got.shape == ref.shape, name
        assert np.abs(got - ref).max() <= 2e-5, (name, qp)            # fp32 evaluation-order noise only
        assert np.array_equal(eo.decisions(got), eo.decisions(ref)), (name, qp)
        assert np.array_equal(got == 0, ref == 0), (name, qp)          # gated zeros identical
        n += 1
    assert n >= 6


def test_gate_fixture_really_exercises_gates(golden_dir):
    g = np.load(os.path.join(golden_dir, "ai_gates_64x64_f10.npz"))
    p = g["prob_qp32"]
    closed1 = (p[:, 1:] == 0).all(1)
    closed2 = (p[:, 5:] == 0).all(1) & ~closed1
    assert closed1.sum() >= 3 and closed2.sum() >= 3 and (~closed1 & ~closed2).sum() >= 3
    g = np.load(os.path.join(golden_dir, "ai_subbatch_2112x2048.npz"))
    p = g["prob_qp32"]
    assert p.shape[0] == 1056
    assert (p[1024:, 1:] == 0).all() and (p[:1024, 1:5] != 0).any()   # second sub-batch gated, first not


def test_ldp_oracle_matches_reference_vectors(golden_dir):
    g = np.load(os.path.join(golden_dir, "ldp_ctus.npz"))
    w = assets.load_weights(assets.LDP_MODEL)
    for qp in g["qps"]:
        got = eo.net_forward(g["ctus"], int(qp), w, eo.MODE_LDP)
        assert np.abs(got - g["prob_qp%d" % qp]).max() <= 2e-5
    fc1 = eo.fc1_export(g["ctus"], 22, w)
    assert np.abs(fc1 - g["fc1_vector"]).max() <= 1e-4 * max(1.0, np.abs(g["fc1_vector"]).max())


def test_survey_known_answers():
    """SURVEY.md section 4 item 5 (data-free known answers, ungated)."""
    c = eo.known_answer_ctus()
    w = assets.load_weights(assets.AI_MODELS[32])
    p = eo.net_forward(c, 32, w)
    assert np.allclose(p[1, :6], [0.434744, 0.035414, 0.027328, 0.035052, 0.036708, 0.011265], atol=2e-5)
    assert abs(p[0].sum() - 20.984108) < 2e-4 and abs(p[1].sum() - 0.693440) < 2e-4
    wl = assets.load_weights(assets.LDP_MODEL)
    pl = eo.net_forward(c, 37, wl, eo.MODE_LDP)
    assert np.allclose(pl[1, :6], [0.439766, 0.060379, 0.067282, 0.060659, 0.069460, 0.023417], atol=2e-5)
    assert abs(pl[0].sum() - 20.632744) < 2e-4 and abs(pl[1].sum() - 1.165507) < 2e-4


def test_fp32_vs_fp64_oracle():
    w = eo.random_weights(7)
    ctus = eo.frame_to_ctus(eo.synth_frame(512, 256, 3))
    p32 = eo.net_forward(ctus, 27, w)
    p64 = eo.net_forward(ctus, 27, w, dtype=np.float64)
    assert np.abs(p32 - p64).max() < 5e-6


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "ETH-CNN_Training_AI/Data/AI_Test_5000.dat_shuffled")),
                    reason="reference demo data not on this box")
def test_accuracy_rows_of_the_reference_logs():
    """Statistical pin: accuracy on the reference's 5000 labelled test CTUs at QP 32 (SURVEY.md section 4
    item 4; the reference's own log row ends at 0.8382 / 0.8211 / 0.7787 on its validation split)."""
    d = np.fromfile(os.path.join(REF, "ETH-CNN_Training_AI/Data/AI_Test_5000.dat_shuffled"), dtype=np.uint8).reshape(5000, 4992)
    ctus = d[:, :4096].reshape(5000, 64, 64)
    label = d[:, 4160 + 16 * 32: 4160 + 16 * 32 + 16].astype(np.float32).reshape(5000, 4, 4)  # input_data.py:101-109
    w = assets.load_weights(assets.AI_MODELS[32])
    p = np.concatenate([eo.net_forward(ctus[i:i + 1024], 32, w) for i in range(0, 5000, 1024)])
    assert abs(p[:, 0].mean() - 0.547301) < 1e-4 and abs(p[:, 1:5].mean() - 0.336638) < 1e-4
    # level-64 accuracy: label depth > 0.5 anywhere <=> split (train_CNN_CTU64.py:103-137)
    y64 = np.minimum(label.reshape(5000, 16).mean(1), 1.0)          # net_CNN.py:110 relu(avg)-relu(avg-1)
    acc64 = ((p[:, 0] > 0.5) == (y64 > 0.5)).mean()
    assert abs(acc64 - 0.8276) < 2e-3


def test_decision_quantiser_edges():
    p = np.array([[0.5, 0.50000006, 0.49999997] + [0.0] * 18], dtype=np.float32)
    d = eo.decisions(p)
    assert d[0, 0] == 0 and d[0, 1] == 2 and d[0, 2] == 0            # "<= down" wins at exactly 0.5 (TEncCu.cpp:453)
    d = eo.decisions(p, (0.7, 0.3, 0.7, 0.3, 0.7, 0.3))
    assert d[0, 0] == 1
    assert [eo.hm_cu_index(0, 0, 0), eo.hm_cu_index(1, 32, 32), eo.hm_cu_index(2, 48, 16)] == [0, 4, 12]


def test_driver_restatement_edge_cases():
    w = eo.random_weights(3)
    with pytest.raises(AssertionError):
        eo.get_prob(bytes(100), 64, 64, 32, w)                       # not a whole number of frames
    assert eo.get_prob(b"", 64, 64, 32, w).shape == (0, 21)          # empty file: zero frames
    luma = eo.get_Y_for_one_frame(memoryview(eo.synth_yuv(72, 40, 1)), 0, 72, 40)
    assert luma.shape == (64, 128) and (luma[40:] == 0).all() and (luma[:, 72:] == 0).all()


def test_ldp_lstm_oracle_matches_reference_daemon_functions(golden_dir):
    """The one-step ETH-LSTM restatement against the vectors produced by the reference's unmodified
    resi_to_cu_depth_LDP.py functions (5-frame residue sequence, state carried through state.dat, QP switch)."""
    g = np.load(os.path.join(golden_dir, "ldp_lstm_200x136_f5.npz"))
    cnn = assets.load_weights(assets.LDP_MODEL)
    thr = (0.6, 0.7)   # tokens [1], [3] of HM-16.5_Test_LDP/bin/Thr_info.txt
    state = None
    checked = 0
    for k, (luma, qp) in enumerate(zip(g["frames"], g["qps"]), start=1):
        try:
            lw = assets.load_weights(eo.ldp_lstm_model_prefix(int(qp)))
        except FileNotFoundError:
            break
        prob, state = eo.ldp_predict_frame(luma, int(qp), k, state if k > 1 else None, cnn, lw, thr)
        assert np.abs(prob - g["cu_depth"][k - 1]).max() <= 2e-5
        assert np.abs(state - g["state"][k - 1]).max() <= 5e-5
        assert np.array_equal(eo.decisions(prob, (0.6, 0.4, 0.7, 0.3, 0.8, 0.2)), eo.decisions(g["cu_depth"][k - 1], (0.6, 0.4, 0.7, 0.3, 0.8, 0.2)))
        state = g["state"][k - 1]   # continue from the reference's own state so errors do not compound
        checked += 1
    assert checked >= 4


def test_ldp_lstm_gates(golden_dir):
    g = np.load(os.path.join(golden_dir, "ldp_lstm_gates.npz"))
    assert (g["flat_cu_depth"][:, 1:] == 0).all() and (g["flat_cu_depth"][:, 0] < 0.6).all()
    cnn = assets.load_weights(assets.LDP_MODEL)
    for key, qp in (("flat", 37), ("big", 27)):
        try:
            lw = assets.load_weights(eo.ldp_lstm_model_prefix(qp))
        except FileNotFoundError:
            continue
        prob, state = eo.ldp_predict_frame(g[key], qp, 1, None, cnn, lw, (0.6, 0.7))
        assert np.abs(prob - g[key + "_cu_depth"]).max() <= 2e-5 and np.array_equal(prob == 0, g[key + "_cu_depth"] == 0)
        assert np.abs(state - g[key + "_state"]).max() <= 5e-5
