"""the packed 2-bit decision map produced on the device by the gate kernel next to the float rows
(HM's rule, TLibEncoder/TEncCu.cpp:434-462)."""
import os

import numpy as np
import pytest

from oracle import assets
from oracle import ethcnn_oracle as eo

pytestmark = pytest.mark.gpu


def test_map_equals_hm_rule_on_the_emitted_rows_and_rows_are_unchanged(eb, ai_model_dir, tmp_path):
    d, present = ai_model_dir
    W, H, nf, qp = 456, 264, 3, 32
    yuv = np.frombuffer(eo.synth_yuv(W, H, nf, seed0=40), np.uint8)
    luma = np.stack([yuv[k * W * H * 3 // 2: k * W * H * 3 // 2 + W * H] for k in range(nf)])
    with eb.EthCnn(d, None, eb.MODE_AI, device=0) as net:
        plain = net.predict_luma(luma, W, H, nf, qp)
        rows, words = net.predict_luma_map(luma, W, H, nf, qp)
        assert np.array_equal(rows, plain)                                       # the float file is what it always was
        assert np.array_equal(net.get_decision_thresholds(), np.full(6, 0.5, np.float32))
        assert np.array_equal(eb.unpack_decisions(words), eo.decisions(rows))
        assert (words >> np.uint64(42)).max() == 0
        thr6 = (0.8, 0.2, 0.7, 0.35, 0.6, 0.4)
        net.set_decision_thresholds(thr6)
        rows2, words2 = net.predict_luma_map(luma, W, H, nf, qp)
        assert np.array_equal(rows2, plain)
        assert np.array_equal(eb.unpack_decisions(words2), eo.decisions(rows2, thr6))
        assert np.array_equal(net.decisions(rows2, thr6), eo.decisions(rows2, thr6))
    # closed gates: the zeros the gate writes are what the map is computed from
    td = str(tmp_path / "gated")
    assets.materialize(td, "AI", thr_line="0.9 2.0 0.9 2.0 0.9 2.0")             # lower thresholds no probability exceeds
    with eb.EthCnn(td, None, eb.MODE_AI, device=0) as net:
        rows, words = net.predict_luma_map(luma, W, H, nf, qp)
        assert (rows[:, 1:] == 0).all() and (rows[:, 0] > 0).all()
        want = eo.decisions(rows, (0.9, 2.0, 0.9, 2.0, 0.9, 2.0))
        assert np.array_equal(eb.unpack_decisions(words), want) and (want[:, 1:] == 0).all()


def test_map_on_the_device_path_and_on_threshold_edges(eb, ai_model_dir):
    import torch

    d, _ = ai_model_dir
    W, H, nf, qp = 1920, 1080, 2, 27
    luma = np.stack([eo.synth_frame(W, H, 60 + k) for k in range(nf)])
    dev = torch.device("cuda", 0)
    dl = torch.from_numpy(luma).to(dev)
    n = nf * 510
    out = torch.zeros((n, 21), dtype=torch.float32, device=dev)
    out2 = torch.zeros((n, 21), dtype=torch.float32, device=dev)
    dmap = torch.zeros((n,), dtype=torch.int64, device=dev)
    with eb.EthCnn(d, None, eb.MODE_AI, device=0) as net:
        s = torch.cuda.current_stream().cuda_stream
        net.predict_luma_device(dl.data_ptr(), W, H, W, W * H, nf, qp, out.data_ptr(), s)
        net.predict_luma_device_map(dl.data_ptr(), W, H, W, W * H, nf, qp, out2.data_ptr(), dmap.data_ptr(), s)
        torch.cuda.synchronize()
        rows = out2.cpu().numpy()
        assert np.array_equal(rows, out.cpu().numpy())
        assert np.array_equal(eb.unpack_decisions(dmap.cpu().numpy().view(np.uint64)), eo.decisions(rows))
        # thresholds placed exactly ON emitted probabilities and one ulp to either side: "> up" and "<= down" are strict / inclusive
        p = float(rows[7, 0])
        for up, down in ((p, p), (np.nextafter(np.float32(p), np.float32(0)), np.nextafter(np.float32(p), np.float32(0))),
                         (np.nextafter(np.float32(p), np.float32(1)), np.nextafter(np.float32(p), np.float32(1)))):
            thr6 = (float(up), float(down), 0.5, 0.5, 0.5, 0.5)
            net.set_decision_thresholds(thr6)
            net.predict_luma_device_map(dl.data_ptr(), W, H, W, W * H, nf, qp, out2.data_ptr(), dmap.data_ptr(), s)
            torch.cuda.synchronize()
            assert np.array_equal(eb.unpack_decisions(dmap.cpu().numpy().view(np.uint64)), eo.decisions(rows, thr6))


def test_map_needs_six_thresholds(eb, tmp_path):
    td = str(tmp_path)
    assets.materialize(td, "AI", thr_line="0.5 0.5 0.5 0.5")      # enough for the gates (tokens [1], [3]), not for HM's rule
    luma = np.zeros((1, 64, 64), np.uint8)
    with eb.EthCnn(td, None, eb.MODE_AI, device=0) as net:
        net.predict_luma(luma, 64, 64, 1, 32)
        with pytest.raises(eb.EthCnnError):
            net.predict_luma_map(luma, 64, 64, 1, 32)
        net.set_decision_thresholds((0.5,) * 6)
        net.predict_luma_map(luma, 64, 64, 1, 32)
