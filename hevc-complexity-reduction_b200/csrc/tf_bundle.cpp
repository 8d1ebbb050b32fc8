#include "tf_bundle.h"

#include <cstdio>
#include <cstring>

namespace ethcnn {
namespace {

const uint64_t kTableMagic = 0xdb4775248b80fb57ull;

struct Crc32cTables {
  uint32_t t[8][256];
  Crc32cTables() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82f63b78u : c >> 1;
      t[0][i] = c;
    }
    for (int k = 1; k < 8; ++k)
      for (uint32_t i = 0; i < 256; ++i) t[k][i] = t[0][t[k - 1][i] & 0xff] ^ (t[k - 1][i] >> 8);
  }
};

uint32_t unmask_crc(uint32_t m) {
  uint32_t rot = m - 0xa282ead8u;
  return (rot >> 17) | (rot << 15);
}

bool get_varint(const uint8_t* buf, size_t len, size_t* pos, uint64_t* v) {
  uint64_t r = 0;
  for (int shift = 0; shift < 64; shift += 7) {
    if (*pos >= len) return false;
    uint8_t b = buf[(*pos)++];
    r |= uint64_t(b & 0x7f) << shift;
    if (!(b & 0x80)) {
      *v = r;
      return true;
    }
  }
  return false;
}

bool read_file(const std::string& path, std::vector<uint8_t>* out, std::string* err) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) {
    *err = "cannot open " + path;
    return false;
  }
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  out->resize(n > 0 ? size_t(n) : 0);
  size_t got = out->empty() ? 0 : fread(out->data(), 1, out->size(), f);
  fclose(f);
  if (got != out->size()) {
    *err = "short read on " + path;
    return false;
  }
  return true;
}

// One table block: verifies the trailer, returns [begin, limit) of the entry area.
bool open_block(const std::vector<uint8_t>& buf, uint64_t off, uint64_t size, const uint8_t** begin, size_t* limit,
                std::string* err) {
  // overflow-safe: the handle comes from the file, off + size + 5 may wrap
  if (size < 4 || buf.size() < 5 || size > buf.size() - 5 || off > buf.size() - 5 - size) {
    *err = "table block outside file";
    return false;
  }
  const uint8_t* b = buf.data() + off;
  if (b[size] != 0) {
    *err = "compressed table block (type " + std::to_string(int(b[size])) + ") not supported";
    return false;
  }
  uint32_t stored;
  memcpy(&stored, b + size + 1, 4);
  if (unmask_crc(stored) != crc32c(b, size + 1)) {
    *err = "table block crc mismatch";
    return false;
  }
  uint32_t num_restarts;
  memcpy(&num_restarts, b + size - 4, 4);
  if (uint64_t(num_restarts) * 4 + 4 > size) {
    *err = "bad restart array";
    return false;
  }
  *begin = b;
  *limit = size_t(size - 4 - 4ull * num_restarts);
  return true;
}

template <class Fn>
bool for_each_entry(const uint8_t* b, size_t limit, std::string* err, Fn fn) {
  size_t pos = 0;
  std::string key;
  while (pos < limit) {
    uint64_t shared, non_shared, vlen;
    if (!get_varint(b, limit, &pos, &shared) || !get_varint(b, limit, &pos, &non_shared) ||
        !get_varint(b, limit, &pos, &vlen) || shared > key.size() || non_shared > limit - pos || vlen > limit - pos - non_shared) {
      *err = "corrupt table entry";
      return false;
    }
    key.resize(shared);
    key.append(reinterpret_cast<const char*>(b + pos), non_shared);
    pos += non_shared;
    if (!fn(key, b + pos, size_t(vlen))) return false;
    pos += vlen;
  }
  return true;
}

struct Entry {
  int dtype = 0;
  std::vector<int64_t> shape;
  uint64_t shard = 0, offset = 0, size = 0;
  uint32_t crc = 0;
};

// Minimal protobuf walk; `on_field(field, wire_type, varint_value, bytes_ptr, bytes_len)`.
template <class Fn>
bool walk_proto(const uint8_t* p, size_t n, Fn on_field) {
  size_t pos = 0;
  while (pos < n) {
    uint64_t tag;
    if (!get_varint(p, n, &pos, &tag)) return false;
    int field = int(tag >> 3), wt = int(tag & 7);
    uint64_t v = 0;
    const uint8_t* bp = nullptr;
    size_t bl = 0;
    if (wt == 0) {
      if (!get_varint(p, n, &pos, &v)) return false;
    } else if (wt == 1) {
      if (pos + 8 > n) return false;
      memcpy(&v, p + pos, 8);
      pos += 8;
    } else if (wt == 2) {
      uint64_t l;
      if (!get_varint(p, n, &pos, &l) || l > n - pos) return false;
      bp = p + pos;
      bl = size_t(l);
      pos += l;
    } else if (wt == 5) {
      if (pos + 4 > n) return false;
      uint32_t v32;
      memcpy(&v32, p + pos, 4);
      v = v32;
      pos += 4;
    } else {
      return false;
    }
    if (!on_field(field, wt, v, bp, bl)) return false;
  }
  return true;
}

bool parse_entry(const uint8_t* p, size_t n, Entry* e) {
  return walk_proto(p, n, [&](int field, int, uint64_t v, const uint8_t* bp, size_t bl) {
    switch (field) {
      case 1: e->dtype = int(v); break;
      case 2:  // TensorShapeProto { repeated Dim dim = 2 { int64 size = 1 } }
        return walk_proto(bp, bl, [&](int f2, int, uint64_t, const uint8_t* dp, size_t dl) {
          if (f2 != 2) return true;
          int64_t sz = 0;
          bool ok = walk_proto(dp, dl, [&](int f3, int, uint64_t v3, const uint8_t*, size_t) {
            if (f3 == 1) sz = int64_t(v3);
            return true;
          });
          e->shape.push_back(sz);
          return ok;
        });
      case 3: e->shard = v; break;
      case 4: e->offset = v; break;
      case 5: e->size = v; break;
      case 6: e->crc = uint32_t(v); break;
      default: break;
    }
    return true;
  });
}

}  // namespace

uint32_t crc32c(const uint8_t* p, size_t n) {
  static const Crc32cTables T;
  uint32_t crc = 0xffffffffu;
  while (n >= 8) {
    uint32_t lo, hi;
    memcpy(&lo, p, 4);
    memcpy(&hi, p + 4, 4);
    lo ^= crc;
    crc = T.t[7][lo & 0xff] ^ T.t[6][(lo >> 8) & 0xff] ^ T.t[5][(lo >> 16) & 0xff] ^ T.t[4][lo >> 24] ^
          T.t[3][hi & 0xff] ^ T.t[2][(hi >> 8) & 0xff] ^ T.t[1][(hi >> 16) & 0xff] ^ T.t[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) crc = T.t[0][(crc ^ *p++) & 0xff] ^ (crc >> 8);
  return crc ^ 0xffffffffu;
}

bool read_tf_bundle(const std::string& prefix, std::map<std::string, BundleTensor>* out, std::string* err) {
  std::vector<uint8_t> idx, data;
  if (!read_file(prefix + ".index", &idx, err)) return false;
  if (!read_file(prefix + ".data-00000-of-00001", &data, err)) return false;
  if (idx.size() < 48) {
    *err = prefix + ".index: too small";
    return false;
  }
  const uint8_t* footer = idx.data() + idx.size() - 48;
  uint64_t magic;
  memcpy(&magic, footer + 40, 8);
  if (magic != kTableMagic) {
    *err = prefix + ".index: bad table magic";
    return false;
  }
  size_t pos = 0;
  uint64_t mi_off, mi_size, ix_off, ix_size;
  if (!get_varint(footer, 40, &pos, &mi_off) || !get_varint(footer, 40, &pos, &mi_size) ||
      !get_varint(footer, 40, &pos, &ix_off) || !get_varint(footer, 40, &pos, &ix_size)) {
    *err = prefix + ".index: bad footer";
    return false;
  }
  const uint8_t* ib;
  size_t ilimit;
  if (!open_block(idx, ix_off, ix_size, &ib, &ilimit, err)) return false;
  bool header_seen = false;
  std::map<std::string, Entry> entries;
  bool ok = for_each_entry(ib, ilimit, err, [&](const std::string&, const uint8_t* hv, size_t hl) {
    size_t hp = 0;
    uint64_t boff, bsize;
    if (!get_varint(hv, hl, &hp, &boff) || !get_varint(hv, hl, &hp, &bsize)) {
      *err = "bad block handle";
      return false;
    }
    const uint8_t* bb;
    size_t blimit;
    if (!open_block(idx, boff, bsize, &bb, &blimit, err)) return false;
    return for_each_entry(bb, blimit, err, [&](const std::string& key, const uint8_t* v, size_t vl) {
      if (key.empty()) {
        uint64_t shards = 1, endian = 0;
        walk_proto(v, vl, [&](int f, int, uint64_t val, const uint8_t*, size_t) {
          if (f == 1) shards = val;
          if (f == 2) endian = val;
          return true;
        });
        if (shards != 1 || endian != 0) {
          *err = "only single-shard little-endian bundles are supported";
          return false;
        }
        header_seen = true;
        return true;
      }
      Entry e;
      if (!parse_entry(v, vl, &e)) {
        *err = "bad BundleEntryProto for " + key;
        return false;
      }
      entries[key] = e;
      return true;
    });
  });
  if (!ok) {
    *err = prefix + ".index: " + *err;
    return false;
  }
  if (!header_seen) {
    *err = prefix + ".index: bundle header missing";
    return false;
  }
  out->clear();
  for (auto& kv : entries) {
    const Entry& e = kv.second;
    if (e.dtype != 1) {
      *err = "tensor " + kv.first + ": dtype is not DT_FLOAT";
      return false;
    }
    uint64_t n = 1;
    bool dims_ok = true;
    for (int64_t d : e.shape) {   // no negative dims, no element count beyond what the data file could hold
      if (d < 0 || (d > 0 && n > data.size() / uint64_t(d))) {
        dims_ok = false;
        break;
      }
      n *= uint64_t(d);
    }
    if (!dims_ok || e.shard != 0 || n > data.size() / 4 || e.size != 4 * n || e.size > data.size() || e.offset > data.size() - e.size) {
      *err = "tensor " + kv.first + ": bad extent";
      return false;
    }
    if (unmask_crc(e.crc) != crc32c(data.data() + e.offset, size_t(e.size))) {
      *err = "tensor " + kv.first + ": crc32c mismatch";
      return false;
    }
    BundleTensor& t = (*out)[kv.first];
    t.shape = e.shape;
    t.data.resize(size_t(n));
    memcpy(t.data.data(), data.data() + e.offset, size_t(e.size));
  }
  return true;
}

}  // namespace ethcnn
