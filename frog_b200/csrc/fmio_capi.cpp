// fmio_capi.cpp -- C wrappers over the host-side readers / pruning / pairs.bin writer so the
// CPU test-suite (and Python callers) can exercise exactly the code bin/match runs.
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <zlib.h>

#include <string>
#include <vector>

#include "fast_inflate.h"
#include "keypoint_io.h"

extern "C" {

// Test hook for fast_inflate.cpp: 1 and *produced bytes in out (cap permitting) when the decoder accepted the
// member, 0 when it declined (the readers then use zlib), -2 when cap is too small.
int fmio_fast_inflate(const unsigned char* src, size_t n, char* out, size_t cap, size_t* produced) {
  std::vector<char> buf;
  size_t got = 0;
  if (!fmio::fast_inflate_gzip(src, n, buf, &got)) return 0;
  *produced = got;
  if (got > cap) return -2;
  memcpy(out, buf.data(), got);
  return 1;
}

// Reads a keypoint file by extension (match.cpp:514,528-536).  Returns the number of records,
// -1 on failure (message in err), -2 if the caller's buffers are too small.
int64_t fmio_read(const char* path, float* head, float* desc, int64_t cap_rows, uint32_t d_cap, uint32_t* d_out,
                  char* err, size_t errlen) {
  fmio::KeypointSet k;
  std::string e;
  if (!fmio::read_keypoints(path, k, e)) {
    if (err && errlen) { strncpy(err, e.c_str(), errlen - 1); err[errlen - 1] = 0; }
    return -1;
  }
  if (d_out) *d_out = k.d;
  if ((int64_t)k.n > cap_rows || k.d > d_cap) return -2;
  if (k.n) {
    memcpy(head, k.head.data(), (size_t)k.n * 6 * sizeof(float));
    for (uint32_t i = 0; i < k.n; i++)
      memcpy(desc + (size_t)i * d_cap, k.desc.data() + (size_t)i * k.d, k.d * sizeof(float));
  }
  return k.n;
}

// z-window filter then response pruning, in place (match.cpp:538-546, 585-595).  Returns new n.
int64_t fmio_filter_prune(float* head, float* desc, int64_t n, uint32_t d, float zT, float zmin, float zmax,
                          float sp, int np) {
  fmio::KeypointSet k;
  k.n = (uint32_t)n;
  k.d = d;
  k.head.assign(head, head + n * 6);
  k.desc.assign(desc, desc + n * d);
  fmio::filter_z(k, zT, zmin, zmax);
  fmio::prune(k, sp, np);
  if (k.n) {
    memcpy(head, k.head.data(), (size_t)k.n * 6 * sizeof(float));
    memcpy(desc, k.desc.data(), (size_t)k.n * d * sizeof(float));
  }
  return k.n;
}

// pairs.bin writer (match.cpp:675-742).  head: per-image [n][6] floats concatenated by offsets.
int fmio_write_pairs_bin(const char* path, int n_images, const char* const* filenames, const double* rigids,
                         const int64_t* offsets, const float* head, int64_t n_blocks, const int* first,
                         const int* second, const int64_t* counts, const int64_t* pair_offsets,
                         const uint32_t* pairs) {
  std::vector<std::string> names(filenames, filenames + n_images);
  std::vector<std::array<double, 3>> rg;
  if (rigids)
    for (int i = 0; i < n_images; i++) rg.push_back({rigids[3 * i], rigids[3 * i + 1], rigids[3 * i + 2]});
  std::vector<fmio::KeypointSet> images(n_images);
  for (int i = 0; i < n_images; i++) {
    images[i].n = (uint32_t)(offsets[i + 1] - offsets[i]);
    images[i].head.assign(head + offsets[i] * 6, head + offsets[i + 1] * 6);
  }
  std::vector<fmio::PairBlock> blocks(n_blocks);
  for (int64_t b = 0; b < n_blocks; b++) {
    blocks[b].first = first[b];
    blocks[b].second = second[b];
    blocks[b].count = (uint32_t)counts[b];
    blocks[b].pairs = pairs + 2 * pair_offsets[b];
  }
  return fmio::write_pairs_bin(path, names, rg, images, blocks) ? 0 : 1;
}

// Synthetic-data writer for the tests and the bench (not used by bin/match): the .csv / .csv.gz text surf3d writes
// (vtkOpenSURF3D/vtk3DSURF.cxx:451-484): "%f" per value, the laplacian sign as "%d", one keypoint per line.
// head: n x 6 floats, desc: n x d floats.  gz_level < 0: plain text; else gzip at that level.  Returns 0 on success.
// printf("%f", (double)v) without printf.  A float is k * 2^e with k < 2^24 and 10^6 = 2^6 * 5^6 (5^6 < 2^14), so
// v * 1e6 is exact in double; rounding that to an integer with ties-to-even is what glibc's correctly rounded "%f"
// prints.  Values the shortcut does not cover (>= 1e12, inf, nan) go through snprintf.
static inline char* fmt_f6(char* o, float v) {
  const double x = (double)v * 1e6;
  const double ax = x < 0 ? -x : x;
  if (!(ax < 1e18)) return o + snprintf(o, 64, "%f", (double)v);
  if (std::signbit(v)) *o++ = '-';
  unsigned long long q = (unsigned long long)std::nearbyint(ax);
  unsigned long long ip = q / 1000000ull;
  unsigned fp = (unsigned)(q % 1000000ull);
  char tmp[24];
  int k = 0;
  do { tmp[k++] = (char)('0' + ip % 10); ip /= 10; } while (ip);
  while (k) *o++ = tmp[--k];
  *o++ = '.';
  for (int i = 5; i >= 0; i--) { o[i] = (char)('0' + fp % 10); fp /= 10; }
  return o + 6;
}

// Test hook: fmt_f6 against snprintf on n random floats of the magnitudes keypoint files hold; returns mismatches.
int64_t fmio_fuzz_fmt(uint64_t seed, int64_t n) {
  uint64_t s = seed * 0x9E3779B97F4A7C15ull + 7;
  int64_t bad = 0;
  char a[80], b[80];
  for (int64_t i = 0; i < n; i++) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    uint32_t bits = (uint32_t)(s >> 32);
    float v;
    if (i & 1) {  // any finite bit pattern
      memcpy(&v, &bits, 4);
      if (!std::isfinite(v)) v = 0.5f;
    } else {  // typical magnitudes, many of them 6-decimal ties
      const double scale[4] = {1.0, 10.0, 2000.0, 1e4};
      v = (float)(((double)(bits >> 8) / 16777216.0 - 0.5) * 2 * scale[(s >> 20) & 3]);
      if ((s >> 24) & 1) v = (float)(std::nearbyint((double)v * 2e6) / 2e6);
    }
    *fmt_f6(a, v) = 0;
    snprintf(b, sizeof b, "%f", (double)v);
    if (strcmp(a, b) != 0) bad++;
  }
  return bad;
}

int fmio_write_csv(const char* path, const float* head, const float* desc, int64_t n, uint32_t d, int gz_level) {
  std::vector<char> text((size_t)n * ((size_t)d + 6) * 11 + 4096);  // typical cell: "-0.123456," ; grown when a row may not fit
  const size_t row_max = ((size_t)d + 6) * 48 + 16;
  char* o = text.data();
  for (int64_t r = 0; r < n; r++) {
    if ((size_t)(text.data() + text.size() - o) < row_max) {
      const size_t used = (size_t)(o - text.data());
      text.resize(text.size() + text.size() / 2 + row_max);
      o = text.data() + used;
    }
    const float* h = head + r * 6;
    for (int k = 0; k < 4; k++) { o = fmt_f6(o, h[k]); *o++ = ','; }
    o += snprintf(o, 16, "%d", (int)h[4]);
    *o++ = ',';
    o = fmt_f6(o, h[5]);
    const float* dd = desc + r * (int64_t)d;
    for (uint32_t k = 0; k < d; k++) { *o++ = ','; o = fmt_f6(o, dd[k]); }
    *o++ = '\n';
  }
  text.resize((size_t)(o - text.data()));
  FILE* f = fopen(path, "wb");
  if (!f) return 1;
  bool ok = true;
  if (gz_level < 0) {
    ok = fwrite(text.data(), 1, text.size(), f) == text.size();
  } else {
    z_stream zs{};
    if (deflateInit2(&zs, gz_level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) { fclose(f); return 2; }
    std::vector<unsigned char> out(deflateBound(&zs, (uLong)text.size()) + 64);
    zs.next_in = reinterpret_cast<Bytef*>(const_cast<char*>(text.data()));
    zs.avail_in = (uInt)text.size();
    zs.next_out = out.data();
    zs.avail_out = (uInt)out.size();
    ok = deflate(&zs, Z_FINISH) == Z_STREAM_END;
    const size_t got = out.size() - zs.avail_out;
    deflateEnd(&zs);
    ok = ok && fwrite(out.data(), 1, got, f) == got;
  }
  ok = (fclose(f) == 0) && ok;
  return ok ? 0 : 3;
}

// Fuzz the libc-free decimal parser against strtof: `n` random cells in the formats surf3d and
// hand-edited files produce.  Returns the number of cells whose value bits or consumed length
// differ from strtof's (must be 0); *n_fast = cells the fast path handled itself.
int64_t fmio_fuzz_floats(uint64_t seed, int64_t n, int64_t* n_fast, char* first_bad, size_t badlen) {
  uint64_t s = seed * 0x9E3779B97F4A7C15ull + 1;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
  int64_t bad = 0, fast = 0;
  char buf[96];
  for (int64_t i = 0; i < n; i++) {
    const int kind = (int)(rnd() % 8);
    int len = 0;
    if (rnd() % 3 == 0) buf[len++] = '-';
    if (kind <= 3) {  // "%f": integer part of 0..6 digits, 6 decimals
      const int id = (int)(rnd() % 7);
      if (id == 0) buf[len++] = '0';
      for (int k = 0; k < id; k++) buf[len++] = (char)('0' + (k == 0 ? 1 + rnd() % 9 : rnd() % 10));
      buf[len++] = '.';
      for (int k = 0; k < 6; k++) buf[len++] = (char)('0' + rnd() % 10);
    } else if (kind == 4) {  // many digits
      const int nd = 1 + (int)(rnd() % 24), dot = (int)(rnd() % (nd + 1));
      for (int k = 0; k < nd; k++) { if (k == dot) buf[len++] = '.'; buf[len++] = (char)('0' + rnd() % 10); }
    } else if (kind == 5) {  // exponent forms, some broken ("1e", "1e+")
      const int nd = 1 + (int)(rnd() % 9);
      for (int k = 0; k < nd; k++) { if (k == 1) buf[len++] = '.'; buf[len++] = (char)('0' + rnd() % 10); }
      buf[len++] = (rnd() & 1) ? 'e' : 'E';
      const int r = (int)(rnd() % 4);
      if (r == 0) buf[len++] = '-'; else if (r == 1) buf[len++] = '+';
      const int ed = (int)(rnd() % 3);
      for (int k = 0; k < ed; k++) buf[len++] = (char)('0' + rnd() % 10);
    } else if (kind == 6) {  // floats near midpoints: a float's midpoint with its successor, printed exactly-ish
      uint32_t fb = (uint32_t)(rnd() % 0x7F000000u);
      float f0, f1;
      memcpy(&f0, &fb, 4);
      fb++;
      memcpy(&f1, &fb, 4);
      const double mid = ((double)f0 + (double)f1) * 0.5;
      len += snprintf(buf + len, 60, (rnd() & 1) ? "%.17g" : "%.25g", mid);
    } else {  // trailing junk / odd shapes
      static const char* odd[] = {"1.5abc", ".5", "5.", "0x1p3", "inf", "nan", "-.25e1x", "00012.500", "+7", "1e400", "1e-50", "4e-39", "3.5e38", "."};
      const char* o = odd[rnd() % (sizeof odd / sizeof odd[0])];
      len = (int)strlen(o);
      memcpy(buf, o, (size_t)len);
    }
    buf[len] = 0;
    float got = 0, want = 0;
    int used = -1;
    const int how = fmio::debug_cell_to_float(buf, (size_t)len, &got, &used);
    char* endp = nullptr;
    errno = 0;
    want = strtof(buf, &endp);
    const bool want_ok = endp != buf && errno != ERANGE;
    bool ok;
    if (how < 0) ok = !want_ok;
    else {
      uint32_t a, b;
      memcpy(&a, &got, 4);
      memcpy(&b, &want, 4);
      ok = want_ok && a == b && (how == 0 || used == (int)(endp - buf));
    }
    if (how == 1) fast++;
    if (!ok) {
      if (bad == 0 && first_bad && badlen) { strncpy(first_bad, buf, badlen - 1); first_bad[badlen - 1] = 0; }
      bad++;
    }
  }
  if (n_fast) *n_fast = fast;
  return bad;
}

}  // extern "C"
