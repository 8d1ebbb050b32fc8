"""CPU-side checks of the product's host code through the C ABI: the library loads and exports every
symbol include/ethcnn.h declares, the C++ TF-bundle reader + weight packer agree with the oracle's
independent reader, Thr_info.txt parsing follows net_CNN.py:38-45, and error behaviour (no compute
without a GPU; non-zero exit codes)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import assets, tf_bundle
from oracle import ethcnn_oracle as eo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol(eb):
    lib = eb.load_library()
    hdr = open(os.path.join(ROOT, "include", "ethcnn.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(ethcnn_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), "library lacks %s declared in include/ethcnn.h" % n
    assert lib.ethcnn_abi_version() == 4


def test_missing_library_fails_loudly(eb, monkeypatch, tmp_path):
    from hevc_complexity_reduction_b200 import binding
    monkeypatch.setattr(binding, "_LIB", None)
    monkeypatch.setenv("ETHCNN_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(eb.EthCnnError):
        binding.load_library()


def _pack(eb, prefix, input_bound=1.0):
    lib = eb.load_library()
    conv = np.zeros(3 * 4976, np.uint32)
    w1 = np.zeros((2688, 448), np.float32)
    b1 = np.zeros(448, np.float32)
    hi = np.zeros((448, 2688), np.uint16)
    lo = np.zeros((448, 2688), np.uint16)
    exps = np.zeros(16, np.int32)
    fb = np.zeros(1, np.float32)
    rc = lib.ethcnn_debug_pack_model(prefix.encode(), C.c_float(input_bound), *[C.c_void_p(a.ctypes.data) for a in
                                                                                   (conv, w1, b1, hi, lo, exps, fb)])
    return rc, conv.reshape(3, 4976), w1, b1, hi, lo, exps, float(fb[0])


@pytest.mark.parametrize("which", ["real", "synthetic"])
def test_cpp_reader_and_packer_match_oracle(eb, tmp_path, which):
    if which == "real":
        d = str(tmp_path)
        assets.materialize(d, "AI")
        prefix = os.path.join(d, assets.AI_MODELS[32])
        w = assets.load_weights(assets.AI_MODELS[32])
    else:
        w = eo.random_weights(11)
        prefix = str(tmp_path / "m.dat")
        tf_bundle.write_bundle(prefix, w)
    rc, conv, w1, b1, hi, lo, exps, fbound = _pack(eb, prefix)
    assert rc == 0, eb.load_library().ethcnn_last_error()
    # conv blocks: branch order S, M, L <- Variable_12.., Variable_6.., Variable..; the layout (biases + filters
    # as mma.sync B fragments in fp16 hi/lo) is stated independently in tests/kernel_model.py
    import kernel_model as km
    for br, base in enumerate((12, 6, 0)):
        want, ex = km.pack_conv_block_reference(w, base, 1.0)
        assert tuple(exps[4 + 4 * br: 8 + 4 * br]) == ex
        assert np.array_equal(conv[br][16:], want[16:])                      # everything but the header
        hdr = conv[br][:4].view(np.float32)
        assert hdr[0] == np.float32(32.0 * 2.0 ** -ex[0]) and hdr[1] == np.float32(2.0 ** -(ex[1] + ex[2]))
        assert hdr[2] == np.float32(2.0 ** -(int(exps[0]) + ex[3])) and hdr[3] == np.float32(2.0 ** ex[1])
    ref_w1 = np.concatenate([w["h_fc1__%s__w" % h] for h in ("64", "32", "16")], axis=1)
    assert np.array_equal(w1, ref_w1)
    assert np.array_equal(b1, np.concatenate([w["h_fc1__%s__b" % h] for h in ("64", "32", "16")]))
    # hi/lo split reproduces the scaled weights to ~2^-22 relative
    ws = np.float32(2.0 ** exps[1])
    rec = (hi.view(np.float16).astype(np.float64) + lo.view(np.float16).astype(np.float64)).T / ws
    assert np.abs(rec - ref_w1).max() <= np.abs(ref_w1).max() * 2.0 ** -21
    assert np.abs(ref_w1).max() * ws < 65504 / 1.9
    # feature bound is a true bound: oracle features on extreme content stay below it
    ctus = np.concatenate([eo.known_answer_ctus(), (np.indices((64, 64)).sum(0) % 2 * 255).astype(np.uint8)[None]])
    x, _ = eo.input_scaling(ctus, 32, eo.MODE_AI, np.float32)
    f = eo.conv_features(x, w)
    assert np.abs(f).max() <= fbound
    assert fbound * 2.0 ** exps[0] <= 32768.0


def test_conv_tc_weight_image_matches_the_fragment_packing(eb, tmp_path):
    """The tcgen05 conv stage reads the SAME hi / lo filter values as the mma.sync stage, re-arranged as 128-byte-swizzled
    K-major UMMA tiles (csrc/conv_tc.h): element (n, k) of a [..][64] tile sits at
    (n / 8) * 1024 + (n % 8) * 128 + ((k / 8) ^ (n % 8)) * 16 + (k % 8) * 2.  Decode every tile and compare with the scaled
    weights' fp16 split; the tables hold the biases pre-scaled by the same powers of two."""
    w = eo.random_weights(21)
    prefix = str(tmp_path / "m.dat")
    tf_bundle.write_bundle(prefix, w)
    rc, conv, _w1, _b1, _hi, _lo, exps, _fb = _pack(eb, prefix)
    assert rc == 0
    lib = eb.load_library()
    blob = np.zeros(3 * 29696, np.uint8)
    assert lib.ethcnn_debug_pack_conv_tc(prefix.encode(), np.float32(1.0), blob.ctypes.data) == 0

    def tile(raw, rows):   # -> [rows][64] uint16
        out = np.zeros((rows, 64), np.uint16)
        u16 = raw.view(np.uint16)
        for n in range(rows):
            for k in range(64):
                out[n, k] = u16[((n >> 3) * 1024 + (n & 7) * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2) // 2]
        return out

    def split(v):
        h = v.astype(np.float32).astype(np.float16)
        l = (v.astype(np.float32) - h.astype(np.float32)).astype(np.float16)
        return h.view(np.uint16), l.view(np.uint16)

    names = lambda i: "Variable" if i == 0 else "Variable_%d" % i
    for br, base in enumerate((12, 6, 0)):
        b = blob[br * 29696:(br + 1) * 29696]
        e1w, ec1, e2w, e3w = (int(x) for x in exps[4 + 4 * br: 8 + 4 * br])
        w1 = w[names(base)].reshape(16, 16) * np.float32(2.0 ** e1w)          # [tap][co]
        w2 = w[names(base + 2)].reshape(64, 24) * np.float32(2.0 ** e2w)      # [(ky, kx, ci)][co]
        w3 = w[names(base + 4)].reshape(96, 32) * np.float32(2.0 ** e3w)
        h, l = split(w1.T)
        assert np.array_equal(tile(b[0:2048], 16)[:, :16], h) and np.array_equal(tile(b[2048:4096], 16)[:, :16], l)
        assert not tile(b[0:2048], 16)[:, 16:].any()
        h, l = split(w2.T)
        t_hi, t_lo = tile(b[4096:8192], 32), tile(b[8192:12288], 32)
        assert np.array_equal(t_hi[:24], h) and np.array_equal(t_lo[:24], l) and not t_hi[24:].any()
        h, l = split(w3.T)
        for t in range(2):
            kk = slice(64 * t, min(96, 64 * t + 64))
            n_k = kk.stop - kk.start
            assert np.array_equal(tile(b[12288 + 4096 * t:12288 + 4096 * (t + 1)], 32)[:, :n_k], h[:, kk])
            assert np.array_equal(tile(b[20480 + 4096 * t:20480 + 4096 * (t + 1)], 32)[:, :n_k], l[:, kk])
        tab = b[28672:28672 + 88 * 4].view(np.float32)
        fs = np.float32(2.0 ** int(exps[0]))
        assert np.array_equal(tab[0:16], w[names(base + 1)] * np.float32(2.0 ** ec1))
        assert np.array_equal(tab[16:32], conv[br][4960:4976].view(np.float32))          # the tap sums of the fragment block
        assert np.array_equal(tab[32:56], w[names(base + 3)] * fs) and np.array_equal(tab[56:88], w[names(base + 5)] * fs)


def test_cpp_reader_rejects_corruption(eb, tmp_path):
    w = eo.random_weights(5)
    prefix = str(tmp_path / "m.dat")
    tf_bundle.write_bundle(prefix, w)
    lib = eb.load_library()
    assert _pack(eb, prefix)[0] == 0
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[1000] ^= 0x40
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    assert _pack(eb, prefix)[0] == -3 and b"crc" in lib.ethcnn_last_error()
    assert _pack(eb, str(tmp_path / "absent.dat"))[0] == -2
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[-1] ^= 0xFF
    open(prefix + ".index", "wb").write(bytes(idx))
    assert _pack(eb, prefix)[0] == -3 and b"magic" in lib.ethcnn_last_error()
    # a checkpoint lacking a tensor
    w2 = dict(w)
    del w2["h_fc2__32__w"]
    tf_bundle.write_bundle(prefix, w2)
    assert _pack(eb, prefix)[0] == -3 and b"h_fc2__32__w" in lib.ethcnn_last_error()


def test_f16_conversion_matches_numpy(eb):
    lib = eb.load_library()
    rng = np.random.default_rng(0)
    vals = np.concatenate([
        rng.standard_normal(2000).astype(np.float32) * np.float32(10.0) ** rng.integers(-9, 5, 2000).astype(np.float32),
        np.array([0.0, -0.0, 65504.0, 65519.9, 65520.0, 1e6, -1e6, 6.1035156e-05, 6.0975552e-05, 5.9604645e-08,
                  2.9802322e-08, 2.98023259e-08, 1.0, 1.00048828125, 1.0009765625, 0.33325195], dtype=np.float32)])
    for v in vals:
        want = np.float32(v).astype(np.float16).view(np.uint16)
        got = lib.ethcnn_debug_f32_to_f16(C.c_float(float(v)))
        assert int(got) == int(want), (float(v), hex(got), hex(int(want)))


def test_threshold_parser_follows_python_split(eb, tmp_path):
    p = tmp_path / "Thr_info.txt"
    p.write_text("0.9 0.95 0.8 0.7 0.6 0.4")
    assert eb.net_CNN.get_thresholds(str(p)) == pytest.approx((0.95, 0.7))
    assert eo.get_thresholds(str(p)) == pytest.approx((0.95, 0.7))
    p.write_text("0.5 0.25 0.5 0.125\n")                     # trailing newline on the 4th token is fine for float()
    assert eb.net_CNN.get_thresholds(str(p)) == (0.25, 0.125)
    p.write_text("0.5  0.25 0.5 0.125")                      # double space -> token[1] == '' -> float('') raises
    with pytest.raises(eb.EthCnnError):
        eb.net_CNN.get_thresholds(str(p))
    with pytest.raises(ValueError):
        eo.get_thresholds(str(p))
    p.write_text("0.5 0.5")
    with pytest.raises(eb.EthCnnError):
        eb.net_CNN.get_thresholds(str(p))
    with pytest.raises(eb.EthCnnError):
        eb.net_CNN.get_thresholds(str(tmp_path / "missing.txt"))


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback(eb, ai_model_dir):
    d, _ = ai_model_dir
    with pytest.raises(eb.EthCnnError) as e:
        eb.EthCnn(d)
    assert e.value.code == -4 and "no CPU fallback" in str(e.value)


def test_cli_usage_and_failure_exit_codes(tmp_path):
    cli = os.path.join(ROOT, "hevc-complexity-reduction_b200", "bin", "video_to_cu_depth")
    assert os.path.exists(cli), "CLI not built"
    r = subprocess.run([cli, "a.yuv", "64"], capture_output=True, cwd=tmp_path)
    assert r.returncode == 1 and b"usage" in r.stderr
    r = subprocess.run([cli, "a.yuv", "sixty", "64", "32"], capture_output=True, cwd=tmp_path)
    assert r.returncode == 1
    # no Thr_info.txt in the cwd -> failure, and no cu_depth.dat left behind
    r = subprocess.run([cli, "a.yuv", "64", "64", "32"], capture_output=True, cwd=tmp_path)
    assert r.returncode == 1 and not os.path.exists(tmp_path / "cu_depth.dat")


def test_python_drop_in_script_has_no_arithmetic():
    src = open(os.path.join(ROOT, "hevc-complexity-reduction_b200", "video_to_cu_depth.py")).read()
    assert "numpy" not in src and "oracle" not in src
    for f in ("binding.py", "__init__.py", "net_CNN.py", "sharding.py"):
        assert "oracle" not in open(os.path.join(ROOT, "hevc-complexity-reduction_b200", f)).read().replace("oracle/", "")


def test_server_client_protocol_against_a_fake_server(eb, tmp_path):
    """ethcnn_request (the C client of the resident server, csrc/serve.cpp) against a Python stand-in that speaks the
    protocol: the request line carries the client's cwd and the arguments, a "0" reply is success, an error reply
    surfaces code and message, and a missing server is reported as such (the drop-in then works in-process)."""
    import socket
    import threading

    sock_path = str(tmp_path / "srv.sock")
    seen = []

    def fake_server(replies):
        srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        srv.bind(sock_path)
        srv.listen(4)
        for reply in replies:
            c, _ = srv.accept()
            buf = b""
            while not buf.endswith(b"\n"):
                buf += c.recv(4096)
            seen.append(buf.decode())
            c.sendall(reply)
            c.close()
        srv.close()

    assert eb.request(sock_path, "clip.yuv", 64, 64, 32) is False          # nobody listens
    t = threading.Thread(target=fake_server, args=([b"0\n", b"-1\tfile size is not a whole number of frames\n"],))
    t.start()
    import time
    for _ in range(200):
        if os.path.exists(sock_path):
            break
        time.sleep(0.01)
    assert eb.request(sock_path, "clip.yuv", 1920, 1080, 27, "cu_depth.dat") is True
    with pytest.raises(eb.EthCnnError) as ei:
        eb.request(sock_path, "/abs/bad.yuv", 10, 10, 32, "out.dat")
    t.join()
    assert ei.value.code == -1 and "whole number" in str(ei.value)
    f = seen[0].rstrip("\n").split("\t")
    assert f[0] == "PREDICT" and f[1] == os.getcwd() and f[2:] == ["clip.yuv", "1920", "1080", "27", "cu_depth.dat"]
    assert seen[1].split("\t")[2] == "/abs/bad.yuv"


def test_corrupt_checkpoint_index_is_rejected_not_read_out_of_bounds(eb, tmp_path):
    """A crafted .index (huge varints in block handles / entries, negative dims, extents past the data file) must come back as
    ETHCNN_E_FORMAT from the C++ reader: its bounds checks compare without adding attacker-controlled 64-bit quantities."""
    import struct

    w = eo.random_weights(5)
    good = str(tmp_path / "good.dat")
    tf_bundle.write_bundle(good, w)
    idx = bytearray(open(good + ".index", "rb").read())
    data = open(good + ".data-00000-of-00001", "rb").read()
    lib = eb.load_library()

    def try_index(blob, tag):
        pre = str(tmp_path / ("bad_%s.dat" % tag))
        open(pre + ".index", "wb").write(bytes(blob))
        open(pre + ".data-00000-of-00001", "wb").write(data)
        rc = lib.ethcnn_debug_pack_model(pre.encode(), C.c_float(1.0), None, None, None, None, None, None, None)
        assert rc in (-2, -3), (tag, rc)

    huge = b"\xff\xff\xff\xff\xff\xff\xff\xff\xff\x01"                      # varint 2^64 - 1
    footer = len(idx) - 48
    try_index(idx[:footer] + huge + huge + huge + huge + bytes(40 - 40) + idx[footer + 40:], "footer_handles")   # offsets wrap
    f2 = bytearray(idx)
    f2[footer:footer + 40] = (huge + b"\x04" + b"\x00" + huge)[:40].ljust(40, b"\x00")
    try_index(f2, "index_handle")
    # flip bytes inside the data block: crc mismatch or corrupt entries, never a crash
    rng = np.random.default_rng(1)
    for trial in range(24):
        b = bytearray(idx)
        for pos in rng.integers(0, footer, size=3):
            b[int(pos)] ^= int(rng.integers(1, 256))
        try_index(b, "flip%d" % trial)
    try_index(idx[:20], "truncated")
