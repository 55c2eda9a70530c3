// Memory-safety check of the batched .c2df walker (csrc/c2df_walk.cpp) on hostile input: every golden fuzz vector and
// random mutations of it (bytes replaced, inserted, files cut) are walked from an exact-size heap copy under
// AddressSanitizer + UBSan.  Built and run by tests/test_cabi.py (CPU tier).  usage: walk_sanitize blob.bin offsets.bin
#include <cstdio>
#include <string>
#include <vector>
#include <fstream>
#include <random>
namespace sgic { void set_error(const std::string&) {} }
#include "../../searchable-generative-image-compression_b200/csrc/c2df_walk.cpp"
int main(int argc, char** argv) {
  std::ifstream f(argv[1], std::ios::binary);
  std::vector<uint8_t> blob((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  std::ifstream g(argv[2], std::ios::binary);
  std::vector<int64_t> offs; int64_t v;
  while (g.read(reinterpret_cast<char*>(&v), 8)) offs.push_back(v);
  const size_t n = offs.size() - 1;
  std::mt19937_64 rng(1);
  size_t kept = 0, total = 0;
  for (int round = 0; round < 12; ++round)
    for (size_t i = 0; i < n; ++i) {
      // exact-size heap copy so that ASan sees any read past the end of the file
      const size_t len = offs[i + 1] - offs[i];
      std::vector<uint8_t> file(blob.begin() + offs[i], blob.begin() + offs[i + 1]);
      if (round > 0 && len > 0) {
        const int k = 1 + rng() % 4;
        for (int j = 0; j < k; ++j) {
          const int op = rng() % 3;
          if (op == 0) file[rng() % file.size()] = static_cast<uint8_t>(rng());
          else if (op == 1) file.resize(rng() % (file.size() + 1));
          else file.insert(file.begin() + rng() % (file.size() + 1), static_cast<uint8_t>(rng()));
          if (file.empty()) break;
        }
      }
      std::vector<uint8_t> exact(file);  exact.shrink_to_fit();
      sgic::Found fd;
      sgic::WalkMemo memo;
      int st = sgic::walk_one(exact.data(), exact.size(), &fd, (i & 1) ? &memo : nullptr);
      if (st == 0 && fd.has_stream && fd.stream_is_bytes && fd.stream_len <= sgic::kZlMaxFrameBytes) {
        // the frame classifier of the device-decode route reads the stream on the host as well; an exact-size copy again
        std::vector<uint8_t> frame(fd.stream, fd.stream + fd.stream_len);
        frame.shrink_to_fit();
        sgic::zl::FrameInfo fi;
        (void)sgic::zl::parse_frame(frame.data(), static_cast<uint32_t>(frame.size()), fi);
      }
      long long dim = 0;
      if (st == 0 && fd.has_meta && fd.meta_type == 4 && sgic::clip_meta_dim(fd.meta, fd.meta_dim, &dim)) ++kept;
      ++total;
    }
  std::printf("%zu files walked, %zu with a readable dim, no sanitizer report\n", total, kept);
}
