// CPU twin of the device-side clip_stream decoder: the SAME header the kernel compiles (csrc/zstd_lit.cuh),
// built for the host by tests/test_zstd_lit.py so that the decoder logic is checked against libzstd without a GPU.
// Test infrastructure only — nothing in the product links this.
#include "../searchable-generative-image-compression_b200/csrc/zstd_lit.cuh"

extern "C" int zl_decode(const uint8_t* s, uint32_t n, uint8_t* dst, uint32_t cap, uint32_t* out_size) {
  static thread_local uint16_t tab[1u << sgic::zl::kHufMaxLog];
  static thread_local uint8_t lit[4096];
  static thread_local sgic::zl::SeqTables seq;
  return sgic::zl::decode_frame_serial(s, n, dst, cap, tab, out_size, lit, sizeof(lit), &seq);
}
extern "C" int zl_classify(const uint8_t* s, uint32_t n) {
  sgic::zl::FrameInfo f;
  return sgic::zl::parse_frame(s, n, f);
}
