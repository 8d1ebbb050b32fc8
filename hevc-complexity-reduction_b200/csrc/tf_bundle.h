// Reader for TensorFlow "Saver V2" tensor bundles (<prefix>.index + <prefix>.data-00000-of-00001).
// Replaces `saver.restore(sess, 'model_2000000_qp..~...dat')` of the reference
// (HM-16.5_Test_AI/bin/video_to_cu_depth.py:126-133).  Format notes: SURVEY.md section 8(c).
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace ethcnn {

struct BundleTensor {
  std::vector<int64_t> shape;
  std::vector<float> data;  // DT_FLOAT only
  size_t count() const { return data.size(); }
};

// Loads every tensor of the bundle.  Returns false and fills *err on any malformation
// (bad magic, compressed block, crc mismatch, non-float dtype, extent outside the data file).
bool read_tf_bundle(const std::string& prefix, std::map<std::string, BundleTensor>* out, std::string* err);

uint32_t crc32c(const uint8_t* p, size_t n);

}  // namespace ethcnn
