// fast_inflate.cpp -- a table-driven DEFLATE / gzip decoder for the `.csv.gz` keypoint files
// (match.cpp:51-92 reads them through boost::iostreams::gzip_decompressor).
//
// After the libc-free float parser, inflating the text was the larger half of loading a keypoint
// file (zlib: ~0.2 GB/s of output on this text).  This decoder trades generality for speed: the
// whole member is in memory, the output size is known from the gzip trailer, so it runs one tight
// loop with a 64-bit bit buffer refilled by unaligned 8-byte loads, two-level Huffman tables
// (10-bit root for literals/lengths, 8-bit for distances) and word-wise match copies.  Keypoint text
// is match-heavy (99.8 % of a surf3d file's bytes come out of 3-5 byte matches: ",0." / ",-0." and digit groups that
// occurred in the last 32 KB), so the decode rate is set by the serial chain lookup -> shift -> lookup of each
// length / distance pair, not by table size: 1.3x zlib with the first version of this loop, 3x now (a pair table that
// resolves a whole match -- length and distance code -- in one lookup, two matches per refill, one shift per symbol
// with the extra bits cut from a copy of the buffer, a BMI2 clone for the variable shifts, CRC-32 by carry-less
// multiplication, per-thread scratch buffers).  It accepts
// exactly what RFC 1951/1952 allow; on ANYTHING unexpected (bad header, invalid code, distance too
// far, size or CRC-32 mismatch, truncated input) it returns false and the caller falls back to
// zlib, so the accepted language and the error behaviour stay zlib's.
#include "fast_inflate.h"

#include <immintrin.h>
#include <zlib.h>  // crc32() for the tail / CPUs without PCLMULQDQ

#include <cstring>

namespace fmio {
namespace {

struct Entry {
  uint16_t val;  // literal byte, length / distance base, or offset of a sub-table
  uint8_t bits;  // total code length in bits (root + sub for second-level entries); root bits for a link
  uint8_t op;    // kLiteral, or flags below
};
constexpr uint8_t kLiteral = 0;
constexpr uint8_t kBase = 0x10;     // length / distance: low nibble = number of extra bits
constexpr uint8_t kLink = 0x20;     // second-level table: low nibble = its index width
constexpr uint8_t kInvalid = 0x40;  // no code maps here
constexpr uint8_t kEnd = 0x80;      // end of block

constexpr int kLitRoot = 10, kDistRoot = 8, kLenRoot = 7;
constexpr int kLitSize = (1 << kLitRoot) + 288 * 32;  // root + worst-case second level (<= 2^5 entries per long code)
constexpr int kDistSize = (1 << kDistRoot) + 32 * 128;
constexpr int kLenSize = 1 << kLenRoot;

const uint16_t kLengthBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLengthExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

enum Kind { kLitLen, kDist, kCodeLen };

inline uint32_t reverse_bits(uint32_t v, int n) {
  uint32_t r = 0;
  for (int i = 0; i < n; i++) { r = (r << 1) | (v & 1); v >>= 1; }
  return r;
}

Entry symbol_entry(Kind kind, int sym) {
  Entry e{0, 0, kInvalid};
  if (kind == kCodeLen) { e.val = (uint16_t)sym; e.op = kLiteral; }
  else if (kind == kLitLen) {
    if (sym < 256) { e.val = (uint16_t)sym; e.op = kLiteral; }
    else if (sym == 256) { e.op = kEnd; }
    else if (sym < 286) { e.val = kLengthBase[sym - 257]; e.op = kBase | kLengthExtra[sym - 257]; }
  } else if (sym < 30) { e.val = kDistBase[sym]; e.op = kBase | kDistExtra[sym]; }
  return e;
}

// Canonical Huffman code (RFC 1951 3.2.2) -> two-level lookup table indexed by the next bits of the
// stream (LSB first).  Returns false for an over-subscribed code, or an incomplete one other than
// the single-code distance alphabet zlib also accepts.
bool build_table(const uint8_t* lens, int n, Kind kind, int root, Entry* table, int capacity) {
  int count[16] = {0};
  for (int s = 0; s < n; s++) count[lens[s]]++;
  count[0] = 0;
  int left = 1, total = 0, max_len = 0;
  for (int l = 1; l <= 15; l++) {
    left = (left << 1) - count[l];
    if (left < 0) return false;
    total += count[l];
    if (count[l]) max_len = l;
  }
  if (left > 0 && !(kind == kDist && total <= 1)) return false;
  const Entry invalid{0, (uint8_t)1, kInvalid};
  for (int i = 0; i < (1 << root); i++) table[i] = invalid;
  if (total == 0) return true;
  uint32_t next_code[16];
  uint32_t code = 0;
  for (int l = 1; l <= 15; l++) { code = (code + (uint32_t)count[l - 1]) << 1; next_code[l] = code; }
  // pass 1: longest code under every root prefix that needs a second level
  uint8_t sub_len[1 << kLitRoot];
  if (max_len > root) {
    memset(sub_len, 0, (size_t)1 << root);
    uint32_t nc[16];
    memcpy(nc, next_code, sizeof nc);
    for (int s = 0; s < n; s++) {
      const int l = lens[s];
      if (l <= root) { if (l) nc[l]++; continue; }
      const uint32_t c = nc[l]++;
      const uint32_t prefix = reverse_bits(c >> (l - root), root);
      if (sub_len[prefix] < l) sub_len[prefix] = (uint8_t)l;
    }
    int next = 1 << root;
    for (int p = 0; p < (1 << root); p++) {
      if (!sub_len[p]) continue;
      const int w = sub_len[p] - root;
      if (next + (1 << w) > capacity) return false;
      table[p] = Entry{(uint16_t)next, (uint8_t)root, (uint8_t)(kLink | w)};
      for (int i = 0; i < (1 << w); i++) table[next + i] = invalid;
      next += 1 << w;
    }
  }
  // pass 2: fill
  for (int s = 0; s < n; s++) {
    const int l = lens[s];
    if (!l) continue;
    const uint32_t c = next_code[l]++;
    Entry e = symbol_entry(kind, s);
    e.bits = (uint8_t)(l + ((e.op & kBase) ? (e.op & 15) : 0));  // length / distance entries: code + extra bits, taken in one shift
    if (l <= root) {
      const uint32_t r = reverse_bits(c, l);
      for (uint32_t i = r; i < (1u << root); i += 1u << l) table[i] = e;
    } else {
      const uint32_t prefix = reverse_bits(c >> (l - root), root);
      const Entry link = table[prefix];
      const int w = link.op & 0x0F, rest = l - root;
      const uint32_t r = reverse_bits(c & ((1u << rest) - 1), rest);
      for (uint32_t i = r; i < (1u << w); i += 1u << rest) table[link.val + i] = e;
    }
  }
  return true;
}

// A whole match in one lookup: when a length code (with its extra bits) and the distance code that follows fit in the
// next kPairBits bits of the stream, the entry holds the length, the distance base and how to cut the distance's extra
// bits.  Keypoint text is matches of length 3-6 at arbitrary distances: a 2-4 bit length code, a 4-6 bit distance code.
constexpr int kPairBits = 11;
struct Pair {
  uint16_t len;     // 0: no pair here, take the general path
  uint16_t dbase;
  uint8_t total;    // bits to consume: both codes and all extra bits
  uint8_t code;     // bits before the distance's extra bits
  uint8_t dextra;
  uint8_t pad;
};
struct Decoder {
  Entry lit[kLitSize];
  Entry dist[kDistSize];
  Entry clen[kLenSize];
  Pair pair[1 << kPairBits];
};

void build_pairs(Decoder& D) {
  for (uint32_t idx = 0; idx < (1u << kPairBits); idx++) {
    Pair p{0, 0, 0, 0, 0, 0};
    const Entry e = D.lit[idx & ((1u << kLitRoot) - 1)];
    if ((e.op & kBase) && !(e.op & (kLink | kEnd | kInvalid)) && e.bits <= kPairBits) {
      const int lextra = e.op & 15;
      const uint32_t len = e.val + ((idx >> (e.bits - lextra)) & ((1u << lextra) - 1));
      const int rem = kPairBits - e.bits;
      const Entry d = D.dist[(idx >> e.bits) & ((1u << kDistRoot) - 1)];
      const int dextra = d.op & 15;
      if ((d.op & kBase) && !(d.op & (kLink | kInvalid)) && d.bits - dextra <= rem) {
        p.len = (uint16_t)len;
        p.dbase = d.val;
        p.code = (uint8_t)(e.bits + d.bits - dextra);
        p.dextra = (uint8_t)dextra;
        p.total = (uint8_t)(e.bits + d.bits);
      }
    }
    D.pair[idx] = p;
  }
}

inline uint64_t load64(const uint8_t* p) {
  uint64_t v;
  memcpy(&v, p, 8);
  return v;  // little-endian hosts only (x86-64 / aarch64)
}

// CRC-32 (gzip polynomial, reflected) by carry-less multiplication: fold four 128-bit lanes over the buffer, reduce to
// 64 bits, Barrett reduction (Gopal et al., "Fast CRC computation for generic polynomials using PCLMULQDQ", Intel 2009).
// `crc` is the raw shift-register state (not inverted); len >= 64 and a multiple of 16.
__attribute__((target("pclmul,sse4.1")))
static uint32_t crc32_fold(uint32_t crc, const unsigned char* buf, size_t len) {
  const __m128i k1k2 = _mm_set_epi64x(0x01c6e41596, 0x0154442bd4);
  const __m128i k3k4 = _mm_set_epi64x(0x00ccaa009e, 0x01751997d0);
  const __m128i k5k0 = _mm_set_epi64x(0x0000000000, 0x0163cd6124);
  const __m128i poly = _mm_set_epi64x(0x01f7011641, 0x01db710641);
  __m128i x0, x1, x2, x3, x4, x5, x6, x7, x8, y5, y6, y7, y8;
  x1 = _mm_loadu_si128((const __m128i*)(buf + 0x00));
  x2 = _mm_loadu_si128((const __m128i*)(buf + 0x10));
  x3 = _mm_loadu_si128((const __m128i*)(buf + 0x20));
  x4 = _mm_loadu_si128((const __m128i*)(buf + 0x30));
  x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
  x0 = k1k2;
  buf += 64; len -= 64;
  while (len >= 64) {
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x6 = _mm_clmulepi64_si128(x2, x0, 0x00);
    x7 = _mm_clmulepi64_si128(x3, x0, 0x00); x8 = _mm_clmulepi64_si128(x4, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x2 = _mm_clmulepi64_si128(x2, x0, 0x11);
    x3 = _mm_clmulepi64_si128(x3, x0, 0x11); x4 = _mm_clmulepi64_si128(x4, x0, 0x11);
    y5 = _mm_loadu_si128((const __m128i*)(buf + 0x00)); y6 = _mm_loadu_si128((const __m128i*)(buf + 0x10));
    y7 = _mm_loadu_si128((const __m128i*)(buf + 0x20)); y8 = _mm_loadu_si128((const __m128i*)(buf + 0x30));
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x5), y5); x2 = _mm_xor_si128(_mm_xor_si128(x2, x6), y6);
    x3 = _mm_xor_si128(_mm_xor_si128(x3, x7), y7); x4 = _mm_xor_si128(_mm_xor_si128(x4, x8), y8);
    buf += 64; len -= 64;
  }
  x0 = k3k4;
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x3), x5);
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x4), x5);
  while (len >= 16) {
    x2 = _mm_loadu_si128((const __m128i*)buf);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
    buf += 16; len -= 16;
  }
  x2 = _mm_clmulepi64_si128(x1, x0, 0x10);
  x3 = _mm_setr_epi32(~0, 0, ~0, 0);
  x1 = _mm_srli_si128(x1, 8);
  x1 = _mm_xor_si128(x1, x2);
  x0 = k5k0;
  x2 = _mm_srli_si128(x1, 4);
  x1 = _mm_and_si128(x1, x3);
  x1 = _mm_clmulepi64_si128(x1, x0, 0x00);
  x1 = _mm_xor_si128(x1, x2);
  x0 = poly;
  x2 = _mm_and_si128(x1, x3);
  x2 = _mm_clmulepi64_si128(x2, x0, 0x10);
  x2 = _mm_and_si128(x2, x3);
  x2 = _mm_clmulepi64_si128(x2, x0, 0x00);
  x1 = _mm_xor_si128(x1, x2);
  return (uint32_t)_mm_extract_epi32(x1, 1);
}
// CRC-32 of a buffer: carry-less multiplication where the CPU has it (run-time check), zlib otherwise and for the tail.
uint32_t fast_crc32(const unsigned char* p, size_t n) {
  uint32_t c = 0;
  if (n >= 64 && __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1")) {
    const size_t body = n & ~(size_t)15;
    c = ~crc32_fold(~c, p, body);
    p += body; n -= body;
  }
  return (uint32_t)crc32(c, p, (uInt)n);
}

}  // namespace

// Compiled twice (GCC function multi-versioning, resolved once at load time): the BMI2 clone turns the variable shifts
// of the bit reader into single-uop SHRX / SHLX / BZHI, worth 25 % on the match loop.
__attribute__((target_clones("default", "bmi2")))
bool fast_inflate_gzip(const uint8_t* src, size_t n, std::vector<char>& out, size_t* produced) {
  *produced = 0;
  // ---- gzip member header (RFC 1952) ----
  if (n < 18 || src[0] != 0x1f || src[1] != 0x8b || src[2] != 8) return false;
  const uint8_t flg = src[3];
  if (flg & 0xE0) return false;  // reserved bits
  size_t pos = 10;
  if (flg & 4) {  // FEXTRA
    if (pos + 2 > n) return false;
    pos += 2 + (size_t)(src[pos] | (src[pos + 1] << 8));
  }
  for (int f = 0; f < 2; f++)  // FNAME, FCOMMENT: zero-terminated
    if (flg & (f ? 16 : 8)) {
      while (pos < n && src[pos]) pos++;
      pos++;
    }
  if (flg & 2) pos += 2;  // FHCRC
  if (pos + 8 > n) return false;

  // input copy with zero padding: the bit reader may load a few words past the end before a bounds check fires
  static thread_local std::vector<uint8_t> padded;  // reused from file to file: no page faults after the first
  padded.resize(n - pos + 64);
  memcpy(padded.data(), src + pos, n - pos);
  memset(padded.data() + (n - pos), 0, 64);
  const uint8_t* const in_begin = padded.data();
  const uint8_t* const in_end = in_begin + (n - pos);  // end of real data
  const uint8_t* in = in_begin;

  // The trailer of a single-member file gives the size; a multi-member file (or trailing bytes) puts
  // other data there, in which case the size check at the end fails and zlib takes over.
  const uint32_t isize = (uint32_t)src[n - 4] | ((uint32_t)src[n - 3] << 8) | ((uint32_t)src[n - 2] << 16) | ((uint32_t)src[n - 1] << 24);
  if ((size_t)isize > (n - pos) * 1032 + 1024) return false;  // beyond DEFLATE's maximum expansion: not a plain member
  constexpr size_t kSlack = 512;  // a match may be copied in 8-byte words past its end; the text gets a NUL appended
  if (out.size() < (size_t)isize + kSlack) out.resize((size_t)isize + kSlack);  // callers reuse `out`: grow only
  uint8_t* const out_begin = reinterpret_cast<uint8_t*>(out.data());
  uint8_t* const out_limit = out_begin + isize;
  uint8_t* o = out_begin;

  static thread_local std::vector<uint8_t> storage(sizeof(Decoder));
  Decoder& D = *reinterpret_cast<Decoder*>(storage.data());
  uint64_t bitbuf = 0;
  int bitcnt = 0;
#define FM_REFILL()                                  \
  do {                                               \
    bitbuf |= load64(in) << bitcnt;                  \
    in += (63 - bitcnt) >> 3;                        \
    bitcnt |= 56;                                    \
  } while (0)
#define FM_TAKE(nb) (bitbuf >>= (nb), bitcnt -= (nb))

  bool last = false;
  while (!last) {
    if (in > in_end) return false;
    FM_REFILL();
    last = bitbuf & 1;
    const int type = (int)((bitbuf >> 1) & 3);
    FM_TAKE(3);
    if (type == 0) {  // stored block: byte-align, LEN, NLEN, raw bytes
      FM_TAKE(bitcnt & 7);
      in -= bitcnt >> 3;  // give the whole bytes still in the buffer back
      bitbuf = 0;
      bitcnt = 0;
      if (in + 4 > in_end) return false;
      const uint32_t len = (uint32_t)in[0] | ((uint32_t)in[1] << 8), nlen = (uint32_t)in[2] | ((uint32_t)in[3] << 8);
      in += 4;
      if ((len ^ 0xFFFFu) != nlen || in + len > in_end || o + len > out_limit) return false;
      memcpy(o, in, len);
      o += len;
      in += len;
      continue;
    }
    if (type == 3) return false;
    if (type == 1) {  // fixed code (RFC 1951 3.2.6)
      uint8_t lens[288 + 32];
      for (int i = 0; i < 144; i++) lens[i] = 8;
      for (int i = 144; i < 256; i++) lens[i] = 9;
      for (int i = 256; i < 280; i++) lens[i] = 7;
      for (int i = 280; i < 288; i++) lens[i] = 8;
      for (int i = 0; i < 32; i++) lens[288 + i] = 5;
      if (!build_table(lens, 288, kLitLen, kLitRoot, D.lit, kLitSize)) return false;
      if (!build_table(lens + 288, 32, kDist, kDistRoot, D.dist, kDistSize)) return false;
      build_pairs(D);
    } else {  // dynamic code (3.2.7)
      const int hlit = (int)(bitbuf & 31) + 257, hdist = (int)((bitbuf >> 5) & 31) + 1, hclen = (int)((bitbuf >> 10) & 15) + 4;
      FM_TAKE(14);
      if (hlit > 286 || hdist > 30) return false;
      static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
      uint8_t cl[19] = {0};
      for (int i = 0; i < hclen; i++) {
        if (bitcnt < 3) FM_REFILL();
        cl[order[i]] = (uint8_t)(bitbuf & 7);
        FM_TAKE(3);
      }
      if (!build_table(cl, 19, kCodeLen, kLenRoot, D.clen, kLenSize)) return false;
      uint8_t lens[286 + 30 + 138];
      int i = 0;
      while (i < hlit + hdist) {
        if (in > in_end) return false;
        FM_REFILL();
        const Entry e = D.clen[bitbuf & ((1u << kLenRoot) - 1)];
        if (e.op != kLiteral) return false;
        FM_TAKE(e.bits);
        const int sym = e.val;
        if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
        int rep;
        uint8_t v = 0;
        if (sym == 16) {
          if (i == 0) return false;
          v = lens[i - 1];
          rep = 3 + (int)(bitbuf & 3);
          FM_TAKE(2);
        } else if (sym == 17) {
          rep = 3 + (int)(bitbuf & 7);
          FM_TAKE(3);
        } else {
          rep = 11 + (int)(bitbuf & 127);
          FM_TAKE(7);
        }
        if (i + rep > hlit + hdist) return false;
        while (rep--) lens[i++] = v;
      }
      if (lens[256] == 0) return false;  // no end-of-block code
      if (!build_table(lens, hlit, kLitLen, kLitRoot, D.lit, kLitSize)) return false;
      if (!build_table(lens + hlit, hdist, kDist, kDistRoot, D.dist, kDistSize)) return false;
      build_pairs(D);
    }

    // ---- the block's symbols ----
    // Keypoint text inflates almost entirely from matches (99.8 % of the bytes of a surf3d file, average length 4.2),
    // so the loop is built around the match.  Fast path: the pair table resolves length AND distance in one lookup
    // (two dependent table loads per match were the critical chain), two matches per refill.  General path: one shift
    // per symbol with the extra bits cut from a copy of the bit buffer (code + extra bits together, see build_table).
#define FM_LOOKUP(e)                                  \
  e = D.lit[bitbuf & ((1u << kLitRoot) - 1)];         \
  if (e.op & kLink) e = D.lit[e.val + ((bitbuf >> kLitRoot) & ((1u << (e.op & 15)) - 1))]
#define FM_COPY_MATCH(len, dist)                                                                          \
  do {                                                                                                    \
    if ((dist) > (size_t)(o - out_begin) || o + (len) > out_limit) return false;                          \
    const uint8_t* from = o - (dist);                                                                     \
    uint8_t* const stop = o + (len);                                                                      \
    if ((dist) >= 8) { /* words may run up to 7 bytes past `stop`: inside the slack, overwritten by what follows */ \
      memcpy(o, from, 8);                                                                                 \
      if ((len) > 8) {                                                                                    \
        o += 8; from += 8;                                                                                \
        do { memcpy(o, from, 8); o += 8; from += 8; } while (o < stop);                                   \
      }                                                                                                   \
    } else {                                                                                              \
      do { *o++ = *from++; } while (o < stop);                                                            \
    }                                                                                                     \
    o = stop;                                                                                             \
  } while (0)
    FM_REFILL();
    for (;;) {
      // here: at least 56 valid bits
      if (in > in_end + 8 || o > out_limit) return false;
      const Pair p = D.pair[bitbuf & ((1u << kPairBits) - 1)];
      if (p.len) {  // a whole match: at most kPairBits + 13 = 24 bits, so a second one fits before the refill
        const uint32_t dist = p.dbase + (uint32_t)((bitbuf >> p.code) & ((1u << p.dextra) - 1));
        FM_TAKE(p.total);
        const Pair p2 = D.pair[bitbuf & ((1u << kPairBits) - 1)];  // in flight during the copy
        FM_COPY_MATCH(p.len, dist);
        if (p2.len) {
          const uint32_t dist2 = p2.dbase + (uint32_t)((bitbuf >> p2.code) & ((1u << p2.dextra) - 1));
          FM_TAKE(p2.total);
          FM_COPY_MATCH(p2.len, dist2);
        }
        FM_REFILL();
        continue;
      }
      Entry e;
      FM_LOOKUP(e);
      if (e.op == kLiteral) {
        *o++ = (uint8_t)e.val;
        FM_TAKE(e.bits);
        FM_LOOKUP(e);
        if (e.op == kLiteral) {
          *o++ = (uint8_t)e.val;
          FM_TAKE(e.bits);
          FM_LOOKUP(e);
          if (e.op == kLiteral) {
            *o++ = (uint8_t)e.val;
            FM_TAKE(e.bits);
            FM_REFILL();
            continue;
          }
        }
        FM_REFILL();  // the entry in hand was looked up from bits that are still at the bottom of the buffer
      }
      if (e.op & kEnd) { FM_TAKE(e.bits); break; }
      if (!(e.op & kBase)) return false;  // invalid code
      // length (<= 15 + 5 bits) and distance (<= 15 + 13 bits): 48 of the >= 56 bits in the buffer
      const uint64_t lbits = bitbuf;
      const int lextra = e.op & 15;
      const uint32_t len = e.val + (uint32_t)((lbits >> (e.bits - lextra)) & ((1u << lextra) - 1));
      FM_TAKE(e.bits);
      Entry d = D.dist[bitbuf & ((1u << kDistRoot) - 1)];
      if (d.op & kLink) d = D.dist[d.val + ((bitbuf >> kDistRoot) & ((1u << (d.op & 15)) - 1))];
      if (!(d.op & kBase)) return false;
      const uint64_t dbits = bitbuf;
      const int dextra = d.op & 15;
      const uint32_t dist = d.val + (uint32_t)((dbits >> (d.bits - dextra)) & ((1u << dextra) - 1));
      FM_TAKE(d.bits);
      FM_REFILL();
      FM_COPY_MATCH(len, dist);
    }
#undef FM_COPY_MATCH
#undef FM_LOOKUP
  }
#undef FM_REFILL
#undef FM_TAKE
  // ---- trailer: CRC-32 and size of the uncompressed data ----
  in -= bitcnt >> 3;  // whole bytes still in the bit buffer belong to the trailer
  if (in + 8 != in_end) return false;  // the member must end where the file ends: further members / trailing bytes are zlib's business
  const uint32_t crc = (uint32_t)in[0] | ((uint32_t)in[1] << 8) | ((uint32_t)in[2] << 16) | ((uint32_t)in[3] << 24);
  const uint32_t size = (uint32_t)in[4] | ((uint32_t)in[5] << 8) | ((uint32_t)in[6] << 16) | ((uint32_t)in[7] << 24);
  const size_t got = (size_t)(o - out_begin);
  if (size != (uint32_t)got || got != (size_t)isize) return false;
  if (fast_crc32(out_begin, got) != crc) return false;
  *produced = got;
  return true;
}

}  // namespace fmio
