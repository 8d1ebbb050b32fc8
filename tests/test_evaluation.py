"""The evaluation harness: the product's vectorised get_class_matrices / accuracy / tendency
(hevc-complexity-reduction_b200/evaluation.py) against the literal restatement of the reference's loops
(oracle/evaluation_ref.py, train_CNN_CTU64.py:103-147), and the accuracy rows of SURVEY.md section 4 item 4 reproduced
from the oracle's probabilities on the reference's labelled demo CTUs.  Also the product-side workload tools."""
import os

import numpy as np
import pytest

from oracle import assets, evaluation_ref as er, tf_bundle
from oracle import ethcnn_oracle as eo

# accuracy 64 / 32 / 16 on AI_Test_5000.dat_shuffled, ungated, deployed checkpoints
SURVEY_ROWS = {22: (0.7212, 0.7776, 0.7739), 27: (0.7942, 0.8287, 0.7758), 32: (0.8276, 0.8150, 0.7750), 37: (0.8642, 0.8137, 0.7683)}


def test_vectorised_matrices_equal_the_reference_loops_on_random_data(eb):
    ev = eb.evaluation
    rng = np.random.default_rng(5)
    for trial in range(4):
        n = 300
        labels = rng.integers(0, 4, size=(n, 16)).astype(np.float64)
        if trial == 1:
            labels[:] = 0            # nothing split: levels 32 / 16 never scored
        if trial == 2:
            labels[:] = 3            # everything split
        rows = rng.random((n, 21)).astype(np.float32)
        rows[rng.random((n, 21)) < 0.05] = 0.5       # values sitting exactly on the prediction threshold
        thr = (0.5, 0.5, 0.5) if trial != 3 else (0.3, 0.6, 0.45)
        y64, y32, y16 = ev.split_rows(rows)
        want = er.get_class_matrices(labels, y64, y32, y16, list(thr))
        got = ev.get_class_matrices(labels, y64, y32, y16, thr)
        assert [list(map(list, m)) for m in got] == [list(map(list, m)) for m in want]
        for g, w in zip(got, want):
            assert ev.get_tendency_2x2(g) == er.get_tendency_2x2(w)


def test_tendency_corner_cases(eb):
    ev = eb.evaluation
    for m in ([[5, 0], [0, 7]], [[5, 0], [3, 7]], [[5, 2], [3, 0]], [[5, 2], [0, 7]], [[0, 2], [3, 7]], [[5, 2], [3, 7]]):
        assert ev.get_tendency_2x2(m) == er.get_tendency_2x2(m)


@pytest.mark.skipif(assets.demo_set_path("AI_Test_5000.dat_shuffled") is None, reason="reference demo data not on this box")
@pytest.mark.parametrize("qp", [22, 27, 32, 37])
def test_accuracy_rows_of_the_survey_from_the_oracle(eb, qp):
    """Pins the harness AND the oracle: the four rows of SURVEY.md section 4 item 4 to the 4th digit."""
    ev = eb.evaluation
    try:
        w = assets.load_weights(assets.AI_MODELS[qp])
    except FileNotFoundError:
        pytest.skip("checkpoint for QP %d not on this box" % qp)
    luma, labels = ev.read_samples(assets.demo_set_path("AI_Test_5000.dat_shuffled"), qp)
    luma2, labels2 = assets.load_demo_set("AI_Test_5000.dat_shuffled")
    assert np.array_equal(luma, luma2) and np.array_equal(labels, labels2[qp])
    p = np.concatenate([eo.net_forward(luma[i:i + 1024], qp, w) for i in range(0, len(luma), 1024)])
    r = ev.get_accuracy_on_large_data(p, labels)
    assert [round(a, 4) for a in r["accuracy"]] == list(SURVEY_ROWS[qp]), r["accuracy"]
    y64, y32, y16 = ev.split_rows(p)
    want = er.get_class_matrices(labels.astype(np.float64), y64, y32, y16, [0.5, 0.5, 0.5])
    assert [list(map(list, m)) for m in r["matrices"]] == [list(map(list, m)) for m in want]


def test_unpack_decisions_matches_the_oracle_quantiser(eb):
    rng = np.random.default_rng(3)
    p = rng.random((500, 21)).astype(np.float32)
    thr6 = (0.7, 0.3, 0.6, 0.4, 0.55, 0.45)
    d = eo.decisions(p, thr6)
    words = np.zeros(500, np.uint64)
    for k in range(21):
        words |= d[:, k].astype(np.uint64) << np.uint64(2 * k)
    assert np.array_equal(eb.unpack_decisions(words), d)


def test_product_side_synth_tools(tmp_path):
    """tools/synth.py (what bench.py's GPU arm uses instead of oracle/): its bundle writer is readable by the oracle's
    independent reader (crc-checked) and by the C++ reader; frames are deterministic and have a usable spread."""
    import ctypes as C
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from tools import synth
    import ethcnn_b200 as eb

    w = synth.random_cnn_weights(9)
    assert len(w) == 36 and sum(v.size for v in w.values()) == 1288210
    prefix = str(tmp_path / "m.dat")
    synth.write_bundle(prefix, w)
    back = tf_bundle.read_bundle(prefix, verify_crc=True)
    assert set(back) == set(w) and all(np.array_equal(back[k], w[k]) for k in w)
    rc = eb.load_library().ethcnn_debug_pack_model(prefix.encode(), C.c_float(1.0), None, None, None, None, None, None, None)
    assert rc == 0
    a, b = synth.synth_frame(320, 200, 4), synth.synth_frame(320, 200, 4)
    assert np.array_equal(a, b) and a.dtype == np.uint8 and 20 < a.std() < 90
    r = synth.synth_residue_frame(320, 200, 4)
    assert abs(float(r.mean()) - 128.0) < 2.0 and (r == 128).mean() > 0.2
    d = str(tmp_path / "bin")
    synth.prepare_models(d)
    assert open(os.path.join(d, "Thr_info.txt")).read() == "0.5 0.5 0.5 0.5 0.5 0.5"
    assert all(os.path.exists(os.path.join(d, n + ".index")) for n in synth.AI_MODELS.values())
