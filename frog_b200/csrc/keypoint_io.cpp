// keypoint_io.cpp -- see keypoint_io.h.
#include "keypoint_io.h"

#include <zlib.h>

#include "fast_inflate.h"

#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace fmio {

namespace {

bool slurp(const std::string& path, std::vector<char>& buf, std::string& err) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) { err = "cannot open " + path; return false; }
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  buf.resize((size_t)std::max(sz, 0L) + 1);
  size_t got = sz > 0 ? fread(buf.data(), 1, (size_t)sz, f) : 0;
  fclose(f);
  buf.resize(got + 1);
  buf[got] = 0;
  return true;
}

// Decimal text -> float without libc, bit-identical to strtof on everything it accepts and
// "don't know" (false) on the rest.  surf3d writes "%f" cells (vtk3DSURF.cxx:451-478): a sign,
// digits, a point, six digits -- at most 19 significant digits m and a decimal exponent e10 with
// |e10| <= 22.  Then m and 10^|e10| are exact doubles, so one IEEE multiply/divide gives the
// correctly rounded DOUBLE d of the exact value x (Clinger's fast path).  Rounding d once more to
// float equals rounding x directly unless d sits exactly on the midpoint M of two adjacent floats
// (every such M is itself a double, so no M can lie strictly between x and its nearest double):
// those cases, and anything outside the normal float range, are left to strtof.
// *endp = one past the longest prefix strtof would consume.
inline bool fast_strtof(const char* c, const char* ce, float& v, const char*& endp) {
  const char* p = c;
  bool neg = false;
  if (p < ce && (*p == '-' || *p == '+')) { neg = *p == '-'; p++; }
  // At most 19 digits in all (leading zeros included): m cannot overflow 64 bits, so the digit loops carry no
  // bookkeeping; longer cells are left to strtof.
  uint64_t m = 0;
  const char* const d0 = p;
  while (p < ce && (unsigned)(*p - '0') <= 9u) { m = m * 10 + (uint64_t)(*p - '0'); p++; }
  int digits = (int)(p - d0), frac = 0;
  if (p < ce && *p == '.') {
    p++;
    const char* const f0 = p;
    while (p < ce && (unsigned)(*p - '0') <= 9u) { m = m * 10 + (uint64_t)(*p - '0'); p++; }
    frac = (int)(p - f0);
    digits += frac;
  }
  if (digits == 0) return false;  // "inf", "nan", ".", junk: strtof decides
  if (digits > 19) return false;
  int e10 = -frac;
  if (p < ce && (*p == 'e' || *p == 'E')) {
    const char* q = p + 1;
    bool eneg = false;
    if (q < ce && (*q == '-' || *q == '+')) { eneg = *q == '-'; q++; }
    if (q < ce && *q >= '0' && *q <= '9') {
      int ex = 0;
      while (q < ce && *q >= '0' && *q <= '9') { if (ex < 10000) ex = ex * 10 + (*q - '0'); q++; }
      e10 += eneg ? -ex : ex;
      p = q;
    }  // else: "1e" / "1e+" -- the exponent marker is not part of the number
  } else if (p < ce && (*p == 'x' || *p == 'X' || *p == 'p' || *p == 'P')) {
    return false;  // hex float: strtof decides
  }
  endp = p;
  if (m == 0) { v = neg ? -0.0f : 0.0f; return true; }
  if (m >= (1ull << 53) || e10 < -22 || e10 > 22) return false;
  static const double kPow10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11,
                                    1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
  const double d = e10 < 0 ? (double)m / kPow10[-e10] : (double)m * kPow10[e10];
  if (!(d >= 1.1754943508222875e-38 && d <= 3.4028234663852886e38)) return false;  // subnormal / overflow: ERANGE rules
  uint64_t bits;
  memcpy(&bits, &d, 8);
  if ((bits & 0x1FFFFFFFull) == 0x10000000ull) return false;  // exactly on a float midpoint
  v = neg ? -(float)d : (float)d;
  return true;
}

// The cell shape that makes up 48 of the 54 cells of a surf3d row: an L2-normalised descriptor component written with
// "%f" -- optional '-', then "0." and six digits, then ',' or the end of the line.  Eight bytes are loaded at once,
// the '.' is overwritten by '0' ("00dddddd"), all eight bytes are checked to be digits and converted with three
// multiplications (SWAR); the value is m / 10^6 rounded to float (see below).  `q` points past the sign; at least 8
// readable bytes follow it.
inline bool descriptor_cell(const char* q, bool neg, float& v) {
  uint64_t w;
  memcpy(&w, q, 8);
  if ((w & 0xFFFFu) != 0x2E30u) return false;            // "0."
  w = (w & ~0xFF00ull) | 0x3000ull;                       // "00dddddd"
  if ((((w & 0xF0F0F0F0F0F0F0F0ull) | (((w + 0x0606060606060606ull) & 0xF0F0F0F0F0F0F0F0ull) >> 4)) != 0x3333333333333333ull))
    return false;                                          // a byte that is not '0'..'9'
  w = (w & 0x0F0F0F0F0F0F0F0Full) * 2561 >> 8;            // pairs of digits
  w = (w & 0x00FF00FF00FF00FFull) * 6553601 >> 16;        // groups of four
  const uint64_t m = (w & 0x0000FFFF0000FFFFull) * 42949672960001ull >> 32;
  // m / 10^6 rounded to float.  No division: m * fl(1e-6) is within 2^-52 (relative) of the exact quotient, while the
  // quotient of an integer below 10^6 by 10^6 is either a float itself (m a multiple of 5^6) or at least 2^-24 / 10^6
  // = 6e-14 (relative) away from every midpoint of two floats (m 2^t - k 10^6 is a non-zero integer), so both round
  // to the same float.  tests/test_hostio.py checks all 10^6 values against strtof.  (m == 0 gives +0.0, signed below
  // like strtof's "-0.000000".)
  const float f = (float)((double)m * 1e-6);
  uint32_t fb;
  memcpy(&fb, &f, 4);
  fb |= (uint32_t)neg << 31;  // the sign is as random as the data: no branch on it
  memcpy(&v, &fb, 4);
  return true;
}

// std::stof semantics (match.cpp:69,152) on the cell [c, ce): strtof after leading blanks, longest
// valid prefix, trailing junk ignored; no conversion or ERANGE make std::stof throw, which
// terminates the reference -- reported here as an error.
inline bool cell_to_float(const char* c, const char* ce, float& v) {
  while (c < ce && (*c == ' ' || (*c >= '\t' && *c <= '\r'))) c++;
  if (c == ce) return false;
  const char* fe = nullptr;
  if (fast_strtof(c, ce, v, fe)) return true;
  char* endp = nullptr;
  errno = 0;
  v = strtof(c, &endp);  // stops at ',' '\n' or NUL at the latest: none can continue a number
  if (endp == c || errno == ERANGE) return false;
  return true;
}

}  // namespace

// Test hook: 1 = the libc-free path handled the cell, 0 = it deferred to strtof, -1 = not a number.
// *endp_off = characters consumed (fast path only).
int debug_cell_to_float(const char* c, size_t len, float* v, int* endp_off) {
  const char* fe = c;
  float x = 0;
  if (fast_strtof(c, c + len, x, fe)) { *v = x; *endp_off = (int)(fe - c); return 1; }
  *endp_off = -1;
  return cell_to_float(c, c + len, *v) ? 0 : -1;
}

bool parse_csv_text(const char* text, size_t len, KeypointSet& out, std::string& err) {
  out = KeypointSet{};
  const char* p = text;
  const char* const end_all = text + len;
  std::vector<float> row;
  row.reserve(64);
  out.desc.reserve(len / 9 + 64);  // a "%f" descriptor cell is 9-10 bytes: no regrowth copies for a surf3d file
  out.head.reserve(len / 72 + 64);
  size_t line_no = 0;
  while (p < end_all) {
    const char* eol = static_cast<const char*>(memchr(p, '\n', (size_t)(end_all - p)));
    const char* end = eol ? eol : end_all;
    line_no++;
    row.clear();
    const char* c = p;
    // std::getline(lineStream, cell, ',') loop of match.cpp:150: an exhausted stream yields no
    // further (empty) cell, a cell starting with CR ends the row.
    while (c < end) {
      if (*c == 13) break;
      float v;
      {
        // descriptor cells, eight bytes at a time (the text is NUL-terminated, so c[0] is always readable)
        const bool neg = *c == '-';
        const char* q = c + neg;
        if (end - q >= 8 && (q + 8 == end || q[8] == ',') && descriptor_cell(q, neg, v)) {
          row.push_back(v);
          if (q + 8 == end) break;
          c = q + 9;
          continue;
        }
      }
      {
        // Common case, no search for the cell's end: the number runs right up to the next ',' (or the end of the
        // line).  Then the cell-limited conversion below would see exactly the same characters.
        const char* q = c;
        while (q < end && (*q == ' ' || (*q >= '\t' && *q <= '\r'))) q++;
        const char* fe = nullptr;
        if (q < end && fast_strtof(q, end, v, fe) && (fe == end || *fe == ',')) {
          row.push_back(v);
          if (fe == end) break;
          c = fe + 1;
          continue;
        }
      }
      const char* comma = static_cast<const char*>(memchr(c, ',', (size_t)(end - c)));
      const char* ce = comma ? comma : end;
      if (!cell_to_float(c, ce, v)) {
        err = "stof failed at line " + std::to_string(line_no) + " cell " + std::to_string(row.size());
        return false;
      }
      row.push_back(v);
      if (!comma) break;
      c = comma + 1;
    }
    if (row.size() > 6) {  // match.cpp:170
      const uint32_t d = (uint32_t)row.size() - 6;
      if (out.n == 0) out.d = d;
      if (d != out.d) {
        err = "descriptor length changes inside the file (line " + std::to_string(line_no) + ")";
        return false;
      }
      out.head.insert(out.head.end(), row.begin(), row.begin() + 6);
      out.desc.insert(out.desc.end(), row.begin() + 6, row.end());
      out.n++;
    }
    if (!eol) break;
    p = eol + 1;
  }
  return true;
}

bool read_csv(const std::string& path, KeypointSet& out, std::string& err) {
  static thread_local std::vector<char> buf;  // reused from file to file, like read_csv_gz's
  if (!slurp(path, buf, err)) return false;
  return parse_csv_text(buf.data(), buf.size() - 1, out, err);
}

// Inflate every gzip member of `raw` (n bytes) into `text`.  boost::iostreams::gzip_decompressor -- what the reference
// reads .csv.gz through, match.cpp:55-58 -- continues into the members that follow the first one, so concatenated
// gzip files load as the concatenation of their texts.  *clean = false when the stream ended in an error (truncated
// file, corrupt data, trailing bytes that are not a gzip member).
static void inflate_members(const char* raw, size_t n, std::vector<char>& text, size_t& produced, bool& clean) {
  produced = 0;
  clean = true;
  size_t pos = 0;
  {
    // one well-formed member spanning the file (what surf3d writes): the table-driven decoder; anything else: zlib
    size_t got = 0;
    if (n > 0 && fast_inflate_gzip(reinterpret_cast<const uint8_t*>(raw), n, text, &got)) { produced = got; return; }
    text.resize(std::max<size_t>(n * 4, 1 << 16));
  }
  while (pos < n) {
    z_stream zs{};
    if (inflateInit2(&zs, 15 + 16) != Z_OK) { clean = false; return; }
    zs.next_in = reinterpret_cast<Bytef*>(const_cast<char*>(raw + pos));
    zs.avail_in = (uInt)std::min<size_t>(n - pos, 0xFFFFFFF0u);
    const size_t in0 = zs.avail_in;
    int rc = Z_OK;
    while (rc != Z_STREAM_END) {
      if (produced == text.size()) text.resize(text.size() * 2);
      const size_t chunk = std::min<size_t>(text.size() - produced, 1u << 30);
      zs.next_out = reinterpret_cast<Bytef*>(text.data() + produced);
      zs.avail_out = (uInt)chunk;
      rc = inflate(&zs, Z_NO_FLUSH);
      produced += chunk - zs.avail_out;
      if (rc != Z_OK && rc != Z_STREAM_END) break;                      // corrupt: keep what was inflated
      if (rc == Z_OK && zs.avail_in == 0 && zs.avail_out != 0) break;   // truncated: keep what was inflated
    }
    pos += in0 - zs.avail_in;
    inflateEnd(&zs);
    if (rc != Z_STREAM_END) { clean = false; return; }
  }
}

bool read_csv_gz(const std::string& path, KeypointSet& out, std::string& err) {
  // file image and inflated text live in per-thread buffers that are reused from file to file: a 10 MB vector costs
  // its zero-fill and ~2500 page faults every time it is created
  static thread_local std::vector<char> raw, text;
  if (!slurp(path, raw, err)) return false;
  size_t produced = 0;
  bool clean = true;
  inflate_members(raw.data(), raw.size() - 1, text, produced, clean);
  if (!clean) {
    // The reference's getline loop (match.cpp:61) ends when the decompressor throws inside the stream: the line being
    // read at that moment -- the unterminated tail of what could be inflated -- is never processed.
    while (produced > 0 && text[produced - 1] != '\n') produced--;
  }
  if (text.size() < produced + 1) text.resize(produced + 1);
  text[produced] = 0;
  return parse_csv_text(text.data(), produced, out, err);
}

bool read_bin(const std::string& path, KeypointSet& out, std::string& err) {
  std::vector<char> buf;
  if (!slurp(path, buf, err)) return false;
  const size_t len = buf.size() - 1;
  out = KeypointSet{};
  out.d = 48;  // match.cpp:201
  // Replay of `while(!feof(file))` (match.cpp:184-205) over an in-memory image of the file:
  // fread(&valF,4,1) returns 0 at EOF and leaves valF as it was (a short tail is copied
  // partially, as glibc does), so after the last full record one more iteration runs whose six
  // header fields all equal the last response and whose descriptor is 48 zeros.
  size_t pos = 0;
  bool eof = false;
  float valF = 0.f;
  auto read_float = [&]() {
    size_t avail = len - pos;
    if (avail >= 4) { memcpy(&valF, buf.data() + pos, 4); pos += 4; }
    else { if (avail) memcpy(&valF, buf.data() + pos, avail); pos = len; eof = true; }
  };
  while (!eof) {
    float h[6];
    for (int k = 0; k < 6; k++) { read_float(); h[k] = valF; }
    float desc[48] = {0};
    size_t avail = len - pos;
    if (avail >= sizeof desc) { memcpy(desc, buf.data() + pos, sizeof desc); pos += sizeof desc; }
    else { if (avail) memcpy(desc, buf.data() + pos, avail); pos = len; eof = true; }
    out.head.insert(out.head.end(), h, h + 6);
    out.desc.insert(out.desc.end(), desc, desc + 48);
    out.n++;
  }
  return true;
}

bool read_keypoints(const std::string& path, KeypointSet& out, std::string& err) {
  const std::string ext = path.substr(path.find_last_of('.') + 1);  // match.cpp:514
  if (ext == "csv") return read_csv(path, out, err);
  if (ext == "bin") return read_bin(path, out, err);
  if (ext == "gz") return read_csv_gz(path, out, err);
  err = "Bad file format : " + ext;  // match.cpp:535 (the reference then dereferences garbage)
  return false;
}

static void keep_rows(KeypointSet& k, const std::vector<uint32_t>& rows) {
  KeypointSet o;
  o.d = k.d;
  o.n = (uint32_t)rows.size();
  o.head.resize((size_t)o.n * 6);
  o.desc.resize((size_t)o.n * k.d);
  for (uint32_t i = 0; i < o.n; i++) {
    memcpy(o.head.data() + (size_t)i * 6, k.head.data() + (size_t)rows[i] * 6, 6 * sizeof(float));
    if (k.d) memcpy(o.desc.data() + (size_t)i * k.d, k.desc.data() + (size_t)rows[i] * k.d, k.d * sizeof(float));
  }
  k = std::move(o);
}

void filter_z(KeypointSet& k, float zT, float zmin, float zmax) {
  std::vector<uint32_t> rows;
  rows.reserve(k.n);
  for (uint32_t i = 0; i < k.n; i++) {
    float z = k.row_head(i)[2] + zT;  // match.cpp:542
    if (!(z < zmin || z > zmax)) rows.push_back(i);
  }
  if (rows.size() != k.n) keep_rows(k, rows);
}

namespace {
struct RespIdx {
  float response;
  uint32_t idx;
};
// match.cpp:338 compareCSVrow: by-value compare on response, descending
bool by_response_desc(RespIdx i, RespIdx j) { return i.response > j.response; }
}  // namespace

void prune(KeypointSet& k, float sp, int np) {
  std::vector<RespIdx> v;
  v.reserve(k.n);
  for (uint32_t i = 0; i < k.n; i++) {
    float resp = k.row_head(i)[5];
    if (!(resp < sp)) v.push_back(RespIdx{resp, i});  // remove_if(response < sp), match.cpp:585-589
  }
  bool changed = v.size() != k.n;
  if (np >= 0 && v.size() > (size_t)np) {  // match.cpp:592-594
    std::partial_sort(v.begin(), v.begin() + np, v.end(), by_response_desc);
    v.resize((size_t)np);
    changed = true;
  }
  if (changed) {
    std::vector<uint32_t> rows(v.size());
    for (size_t i = 0; i < v.size(); i++) rows[i] = v[i].idx;
    keep_rows(k, rows);
  }
}

bool write_pairs_bin(const std::string& path, const std::vector<std::string>& filenames,
                     const std::vector<std::array<double, 3>>& rigids, const std::vector<KeypointSet>& images,
                     const std::vector<PairBlock>& blocks) {
  FILE* file = fopen(path.c_str(), "wb");
  if (!file) return false;
  static char iobuf[1 << 22];
  setvbuf(file, iobuf, _IOFBF, sizeof iobuf);
  unsigned short nbAcq = (unsigned short)filenames.size();  // match.cpp:684
  fwrite(&nbAcq, sizeof nbAcq, 1, file);
  for (size_t it = 0; it < images.size(); it++) {
    size_t found = filenames[it].find_last_of("/\\");  // match.cpp:690-694
    std::string cur = filenames[it].substr(found + 1);
    unsigned short len = (unsigned short)cur.size();
    fwrite(&len, sizeof len, 1, file);
    fwrite(cur.c_str(), 1, cur.size(), file);
    std::array<double, 3> tmp = {0.0, 0.0, 0.0};  // match.cpp:697-708
    if (!rigids.empty()) tmp = rigids[it];
    fwrite(tmp.data(), sizeof(double), 3, file);
    uint32_t nbPoints = images[it].n;  // pointIdType = unsigned int (INT_PTIDS), match.cpp:711-713
    fwrite(&nbPoints, sizeof nbPoints, 1, file);
    if (nbPoints) fwrite(images[it].head.data(), sizeof(float), (size_t)nbPoints * 6, file);  // :715-723
  }
  for (const PairBlock& b : blocks) {  // match.cpp:727-742
    fwrite(&b.first, sizeof(unsigned short), 1, file);
    fwrite(&b.second, sizeof(unsigned short), 1, file);
    unsigned int size = b.count;
    fwrite(&size, sizeof size, 1, file);
    if (size) fwrite(b.pairs, 8, size, file);
  }
  fclose(file);
  return true;
}

}  // namespace fmio
