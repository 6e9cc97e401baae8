#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include <zlib.h>
#include "fast_inflate.h"
// Mutation fuzz of fast_inflate_gzip under ASan/UBSan: whatever the input, no out-of-bounds access; when it accepts,
// zlib must accept too and agree byte for byte.
static std::vector<unsigned char> gz(const std::vector<unsigned char>& data, int level, int strategy) {
  z_stream zs{}; deflateInit2(&zs, level, Z_DEFLATED, 31, 8, strategy);
  std::vector<unsigned char> out(deflateBound(&zs, data.size()) + 64);
  zs.next_in = const_cast<unsigned char*>(data.data()); zs.avail_in = data.size(); zs.next_out = out.data(); zs.avail_out = out.size();
  deflate(&zs, Z_FINISH); out.resize(zs.total_out); deflateEnd(&zs); return out;
}
static bool zl(const std::vector<unsigned char>& in, std::vector<unsigned char>& out) {
  z_stream zs{}; inflateInit2(&zs, 47); out.assign(1 << 22, 0);
  zs.next_in = const_cast<unsigned char*>(in.data()); zs.avail_in = in.size(); zs.next_out = out.data(); zs.avail_out = out.size();
  int rc = inflate(&zs, Z_FINISH); out.resize(zs.total_out); inflateEnd(&zs); return rc == Z_STREAM_END;
}
int main(int argc, char** argv) {
  std::mt19937 rng(12345);
  long accepted = 0, declined = 0, iters = 0;
  for (int base = 0; base < (argc > 1 ? atoi(argv[1]) : 60); base++) {
    std::vector<unsigned char> data(rng() % 20000 + 1);
    int kind = base % 4;
    for (auto& c : data) c = kind == 0 ? "0123456789.,-\n"[rng() % 14] : kind == 1 ? (unsigned char)(rng() % 4 + 'a') : kind == 2 ? (unsigned char)rng() : (unsigned char)('a' + (rng() % 100 < 97 ? 0 : rng() % 26));
    auto good = gz(data, (int)(rng() % 10), (int)(rng() % 5));
    for (int m = 0; m < 1500; m++) {
      auto b = good;
      int nmut = 1 + rng() % 3;
      for (int k = 0; k < nmut; k++) {
        switch (rng() % 4) {
          case 0: b[rng() % b.size()] ^= (unsigned char)(1u << (rng() % 8)); break;
          case 1: b[rng() % b.size()] = (unsigned char)rng(); break;
          case 2: b.resize(rng() % b.size() + 1); break;
          default: { size_t i = rng() % b.size(); b.insert(b.begin() + i, (unsigned char)rng()); }
        }
      }
      // exact-size heap copy so that ASan sees any read past the end of the input
      unsigned char* heap = (unsigned char*)malloc(b.size()); memcpy(heap, b.data(), b.size());
      std::vector<char> out; size_t got = 0;
      bool ok = fmio::fast_inflate_gzip(heap, b.size(), out, &got);
      free(heap);
      iters++;
      if (ok) {
        accepted++;
        std::vector<unsigned char> ref;
        if (!zl(b, ref) || ref.size() != got || memcmp(ref.data(), out.data(), got) != 0) { printf("MISMATCH base %d mut %d\n", base, m); return 1; }
      } else declined++;
    }
  }
  printf("fuzz ok: %ld inputs, %ld accepted (all equal to zlib), %ld declined\n", iters, accepted, declined);
}
