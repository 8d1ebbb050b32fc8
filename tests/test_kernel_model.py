"""The lane-level model of the conv kernel and the three-pass FC1 arithmetic against the oracle (CPU)."""
import ctypes as C
import os

import numpy as np
import pytest

import kernel_model as km
from oracle import assets, tf_bundle
from oracle import ethcnn_oracle as eo


def _packed(eb, prefix, bound):
    lib = eb.load_library()
    conv = np.zeros(3 * 4976, np.uint32)
    b1 = np.zeros(448, np.float32)
    hi = np.zeros((448, 2688), np.uint16)
    lo = np.zeros((448, 2688), np.uint16)
    exps = np.zeros(16, np.int32)
    rc = lib.ethcnn_debug_pack_model(prefix.encode(), C.c_float(bound), C.c_void_p(conv.ctypes.data), None,
                                     C.c_void_p(b1.ctypes.data), C.c_void_p(hi.ctypes.data), C.c_void_p(lo.ctypes.data),
                                     C.c_void_p(exps.ctypes.data), None)
    assert rc == 0
    return conv.reshape(3, 4976), b1, hi, lo, exps


@pytest.mark.parametrize("mode", [eo.MODE_AI, eo.MODE_LDP])
def test_conv_fragment_chaining_matches_oracle(eb, tmp_path, mode):
    """The kernel's dataflow (lane -> pixels -> conv1 A fragments -> C fragments re-used as conv2 / conv3 A
    fragments -> feature offsets) on the conv blocks produced by the C++ packer, against the oracle."""
    w = eo.random_weights(21)
    prefix = str(tmp_path / "m.dat")
    tf_bundle.write_bundle(prefix, w)
    conv, _, _, _, exps = _packed(eb, prefix, 10.0 if mode == eo.MODE_LDP else 1.0)
    conv_exps = [tuple(int(v) for v in exps[4 + 4 * br: 8 + 4 * br]) for br in range(3)]
    frame = eo.synth_residue_frame(1024, 64, 4) if mode == eo.MODE_LDP else eo.synth_frame(1024, 64, 4)
    tiles = eo.frame_to_ctus(frame)
    tiles[7] = eo.known_answer_ctus()[0]
    scale = 10.0 / 255.0 if mode == eo.MODE_LDP else 1.0 / 255.0
    cst3 = [scale / 256, scale / 1024, scale / 4096]
    feat = km.conv_features_group(tiles, conv, conv_exps, cst3)
    assert not np.isnan(feat).any()                        # every one of the 2688 slots was written by some lane
    x, _ = eo.input_scaling(tiles, 32, mode, np.float64)
    want = eo.conv_features(x, w)
    assert np.abs(feat - want).max() <= 2e-6 * max(1.0, np.abs(want).max())


def test_three_pass_fc1_meets_fp32_class_accuracy(eb, tmp_path):
    d = str(tmp_path)
    assets.materialize(d, "AI")
    name = assets.AI_MODELS[32]
    w = assets.load_weights(name)
    conv, b1, hi, lo, exps = _packed(eb, os.path.join(d, name), 1.0)
    fe, we = int(exps[0]), int(exps[1])
    ctus = np.concatenate([eo.frame_to_ctus(eo.synth_frame(1024, 512, 9)), eo.known_answer_ctus()])
    x, q = eo.input_scaling(ctus, 32, eo.MODE_AI, np.float32)
    f = eo.conv_features(x, w)
    a1 = km.fc1_three_pass(f, hi, lo, b1, fe, we)
    outs64, a1_ref = eo.fc_heads(f.astype(np.float64), q.astype(np.float64), w, return_fc1=True)
    assert np.abs(a1 - a1_ref).max() <= 3e-6 * max(1.0, np.abs(a1_ref).max())
    # and through the heads: probabilities within a few 1e-7 of the fp64 oracle
    p = []
    col = 0
    for h, n1, n2, n3 in eo.HEADS:
        a = a1[:, col:col + n1].astype(np.float64)
        col += n1
        a2 = km.leaky(np.concatenate([a, q], 1) @ w["h_fc2__%s__w" % h] + w["h_fc2__%s__b" % h])
        p.append(1 / (1 + np.exp(-(np.concatenate([a2, q], 1) @ w["y_conv_flat__%s__w" % h] + w["y_conv_flat__%s__b" % h]))))
    p = np.concatenate(p, 1)
    assert np.abs(p - np.concatenate(outs64, 1)).max() <= 1e-6
