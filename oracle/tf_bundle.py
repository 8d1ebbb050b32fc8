"""TEST INFRASTRUCTURE (oracle) -- reader/writer for TensorFlow "Saver V2" tensor bundles.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  The product reads checkpoints with its own C++ reader
(hevc-complexity-reduction_b200/csrc/tf_bundle.cpp); this file is the independent checker.

What it restates: the on-disk format the reference's `saver.restore(sess, 'model_...dat')`
consumes (HM-16.5_Test_AI/bin/video_to_cu_depth.py:126-133, written by
ETH-CNN_Training_AI/train_CNN_CTU64.py:331-332 via tf.train.Saver V2).  TensorFlow itself is a
third-party dependency absent from /root/reference (README.md:42 "TensorFlow >= 1.8.0"; the
deployed checkpoints were written by TF 1.4.1), so the format is restated from its published
layout (tensorflow/core/util/tensor_bundle + tensorflow/core/lib/io/table, a LevelDB-style
table):

  <prefix>.index                     table of key -> serialized proto
      key ""      -> BundleHeaderProto {1: num_shards, 2: endianness, 3: version}
      key <name>  -> BundleEntryProto  {1: dtype, 2: TensorShapeProto{2: dim{1: size}},
                                        3: shard_id, 4: offset, 5: size, 6: fixed32 crc32c}
      blocks: entries (varint shared, varint non_shared, varint value_len, key delta, value),
              then uint32 restart offsets, uint32 num_restarts; block trailer = 1 byte
              compression type (0) + 4 byte masked crc32c.
      footer (48 B): metaindex handle, index handle (varint offset,size each), zero padding,
              magic 0xdb4775248b80fb57 little-endian.
  <prefix>.data-00000-of-00001       raw little-endian tensors at [offset, offset+size).
"""
from __future__ import annotations

import struct
from typing import Dict, List, Tuple

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
DT_FLOAT = 1


# ----------------------------------------------------------------------------- crc32c
def _make_crc_table() -> List[int]:
    poly = 0x82F63B78
    tbl = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ poly if c & 1 else c >> 1
        tbl.append(c)
    return tbl


_CRC_TABLE = np.array(_make_crc_table(), dtype=np.uint32)


def crc32c(data: bytes) -> int:
    """CRC-32C (Castagnoli), as used by the bundle's per-tensor and per-block checksums."""
    crc = 0xFFFFFFFF
    tbl = _CRC_TABLE
    for b in data:
        crc = int(tbl[(crc ^ b) & 0xFF]) ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def crc32c_fast(data: bytes) -> int:
    """Same value as crc32c(); slicing-by-8 in numpy so 5 MB tensors check in well under a second."""
    n = len(data)
    if n < 4096:
        return crc32c(data)
    # Build 8 tables once.
    global _CRC_T8
    try:
        t8 = _CRC_T8
    except NameError:
        t = np.zeros((8, 256), dtype=np.uint32)
        t[0] = _CRC_TABLE
        for k in range(1, 8):
            t[k] = t[0][t[k - 1] & 0xFF] ^ (t[k - 1] >> 8)
        _CRC_T8 = t8 = t
    # Process sequentially in 8-byte words (python loop over words would be slow); instead use the
    # linearity of CRC: split into chunks, crc each with table lookups vectorised across chunks.
    # Simpler and fast enough: process in a python loop over 64 KiB blocks using the byte table via
    # a cumulative trick is not possible, so fall back to a tight loop on memoryview in blocks.
    crc = 0xFFFFFFFF
    mv = memoryview(data)
    full = n // 8 * 8
    words = np.frombuffer(mv[:full], dtype="<u4").reshape(-1, 2)
    t0, t1, t2, t3, t4, t5, t6, t7 = (t8[k].tolist() for k in range(8))
    lo_list = words[:, 0].tolist()
    hi_list = words[:, 1].tolist()
    for lo, hi in zip(lo_list, hi_list):
        lo ^= crc
        crc = (t7[lo & 0xFF] ^ t6[(lo >> 8) & 0xFF] ^ t5[(lo >> 16) & 0xFF] ^ t4[lo >> 24]
               ^ t3[hi & 0xFF] ^ t2[(hi >> 8) & 0xFF] ^ t1[(hi >> 16) & 0xFF] ^ t0[hi >> 24])
    tb = _CRC_TABLE.tolist()
    for b in mv[full:]:
        crc = tb[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def _mask_crc(crc: int) -> int:
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def _unmask_crc(m: int) -> int:
    rot = (m - 0xA282EAD8) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ----------------------------------------------------------------------------- varint / proto
def _get_varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_proto(buf: bytes) -> Dict[int, list]:
    """Minimal protobuf wire parser: field number -> list of raw values (int or bytes)."""
    out: Dict[int, list] = {}
    pos = 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        out.setdefault(field, []).append(v)
    return out


class BundleEntry:
    __slots__ = ("name", "dtype", "shape", "shard", "offset", "size", "crc")

    def __init__(self, name, dtype, shape, shard, offset, size, crc):
        self.name, self.dtype, self.shape = name, dtype, tuple(shape)
        self.shard, self.offset, self.size, self.crc = shard, offset, size, crc

    def __repr__(self):
        return "BundleEntry(%r, shape=%r, offset=%d, size=%d)" % (self.name, self.shape, self.offset, self.size)


def _parse_entry(name: str, raw: bytes) -> BundleEntry:
    p = _parse_proto(raw)
    dtype = p.get(1, [0])[0]
    shape: List[int] = []
    if 2 in p:
        sp = _parse_proto(p[2][0])
        for dim_raw in sp.get(2, []):
            dp = _parse_proto(dim_raw)
            shape.append(dp.get(1, [0])[0])
    return BundleEntry(name, dtype, shape, p.get(3, [0])[0], p.get(4, [0])[0], p.get(5, [0])[0],
                       p.get(6, [0])[0])


# ----------------------------------------------------------------------------- table reader
def _read_block(buf: bytes, offset: int, size: int, verify: bool) -> bytes:
    block = buf[offset:offset + size]
    trailer = buf[offset + size:offset + size + 5]
    if len(block) != size or len(trailer) != 5:
        raise ValueError("truncated table block")
    if trailer[0] != 0:
        raise ValueError("compressed table blocks are not supported (type %d)" % trailer[0])
    if verify:
        want = _unmask_crc(struct.unpack("<I", trailer[1:])[0])
        got = crc32c(block + trailer[:1])
        if want != got:
            raise ValueError("table block crc mismatch")
    return block


def _iter_block(block: bytes):
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * num_restarts
    pos = 0
    key = b""
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_index(index_path: str, verify_crc: bool = True) -> Dict[str, BundleEntry]:
    """Parse `<prefix>.index` -> {tensor name: BundleEntry}."""
    with open(index_path, "rb") as f:
        buf = f.read()
    if len(buf) < 48:
        raise ValueError("index file too small")
    footer = buf[-48:]
    if struct.unpack("<Q", footer[40:])[0] != TABLE_MAGIC:
        raise ValueError("bad table magic in %s" % index_path)
    pos = 0
    _mi_off, pos = _get_varint(footer, pos)
    _mi_size, pos = _get_varint(footer, pos)
    idx_off, pos = _get_varint(footer, pos)
    idx_size, pos = _get_varint(footer, pos)
    entries: Dict[str, BundleEntry] = {}
    header_seen = False
    for _sep_key, handle in _iter_block(_read_block(buf, idx_off, idx_size, verify_crc)):
        b_off, p = _get_varint(handle, 0)
        b_size, p = _get_varint(handle, p)
        for key, value in _iter_block(_read_block(buf, b_off, b_size, verify_crc)):
            if key == b"":
                hdr = _parse_proto(value)
                if hdr.get(1, [1])[0] != 1:
                    raise ValueError("multi-shard bundles are not supported")
                if hdr.get(2, [0])[0] != 0:
                    raise ValueError("big-endian bundles are not supported")
                header_seen = True
            else:
                name = key.decode("utf-8")
                entries[name] = _parse_entry(name, value)
    if not header_seen:
        raise ValueError("bundle header missing")
    return entries


def read_bundle(prefix: str, verify_crc: bool = False) -> Dict[str, np.ndarray]:
    """Load every float tensor of `<prefix>.index` / `<prefix>.data-00000-of-00001`."""
    entries = read_index(prefix + ".index")
    with open(prefix + ".data-00000-of-00001", "rb") as f:
        data = f.read()
    out: Dict[str, np.ndarray] = {}
    for name, e in entries.items():
        if e.dtype != DT_FLOAT:
            raise ValueError("tensor %s: dtype %d is not DT_FLOAT" % (name, e.dtype))
        n = int(np.prod(e.shape)) if e.shape else 1
        if e.size != 4 * n or e.offset + e.size > len(data):
            raise ValueError("tensor %s: bad extent" % name)
        raw = data[e.offset:e.offset + e.size]
        if verify_crc and _unmask_crc(e.crc) != crc32c_fast(raw):
            raise ValueError("tensor %s: crc32c mismatch" % name)
        out[name] = np.frombuffer(raw, dtype="<f4").reshape(e.shape).copy()
    return out


# ----------------------------------------------------------------------------- writer (synthetic checkpoints)
def _block_bytes(items: List[Tuple[bytes, bytes]], restart_interval: int = 16) -> bytes:
    out = bytearray()
    restarts = []
    last = b""
    for i, (k, v) in enumerate(items):
        if i % restart_interval == 0:
            restarts.append(len(out))
            shared = 0
        else:
            shared = 0
            m = min(len(k), len(last))
            while shared < m and k[shared] == last[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v))
        out += k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _with_trailer(block: bytes) -> bytes:
    return block + b"\x00" + struct.pack("<I", _mask_crc(crc32c(block + b"\x00")))


def _field_varint(field: int, v: int) -> bytes:
    return _put_varint(field << 3) + _put_varint(v)


def _field_bytes(field: int, b: bytes) -> bytes:
    return _put_varint((field << 3) | 2) + _put_varint(len(b)) + b


def write_bundle(prefix: str, tensors: Dict[str, np.ndarray]) -> None:
    """Write a single-shard V2 bundle holding float32 `tensors` (used to make synthetic
    checkpoints with the reference's 36-tensor layout for tests and for bench.py)."""
    names = sorted(tensors.keys(), key=lambda s: s.encode("utf-8"))
    data = bytearray()
    items: List[Tuple[bytes, bytes]] = []
    header = _field_varint(1, 1) + _field_bytes(3, _field_varint(1, 1))  # num_shards=1, version{producer=1}
    items.append((b"", header))
    for name in names:
        arr = np.ascontiguousarray(tensors[name], dtype="<f4")
        raw = arr.tobytes()
        shape = b"".join(_field_bytes(2, _field_varint(1, int(d))) for d in arr.shape)
        entry = _field_varint(1, DT_FLOAT) + _field_bytes(2, shape)
        if len(data):
            entry += _field_varint(4, len(data))
        entry += _field_varint(5, len(raw))
        entry += _put_varint((6 << 3) | 5) + struct.pack("<I", _mask_crc(crc32c_fast(raw)))
        items.append((name.encode("utf-8"), entry))
        data += raw
    data_block = _with_trailer(_block_bytes(items))
    meta_block = _with_trailer(_block_bytes([]))
    out = bytearray(data_block)
    meta_off = len(out)
    out += meta_block
    idx_off = len(out)
    # index block: one entry, separator key >= last key, value = handle of the data block
    handle = _put_varint(0) + _put_varint(len(data_block) - 5)
    sep = names[-1].encode("utf-8") + b"\x00" if names else b"\x00"
    idx_block = _with_trailer(_block_bytes([(sep, handle)], restart_interval=1))
    out += idx_block
    footer = (_put_varint(meta_off) + _put_varint(len(meta_block) - 5)
              + _put_varint(idx_off) + _put_varint(len(idx_block) - 5))
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    out += footer
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))
