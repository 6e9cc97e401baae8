// Test-infrastructure shim: a std::istream that, once a gzip_decompressor tag and a source
// stream have been pushed (match.cpp:56-58), serves the zlib-inflated bytes of that source.
#pragma once
#include <istream>
#include <sstream>
#include <string>
#include <vector>
#include <zlib.h>
#include "filter/gzip.hpp"
namespace boost { namespace iostreams {
class filtering_istream : public std::istream {
  std::stringbuf buf_;
  bool gz_ = false;
 public:
  filtering_istream() : std::istream(nullptr) { rdbuf(&buf_); }
  void push(const gzip_decompressor&) { gz_ = true; }
  void push(std::istream& src) {
    std::string raw((std::istreambuf_iterator<char>(src)), std::istreambuf_iterator<char>());
    if (!gz_) { buf_.str(raw); return; }
    // Like boost::iostreams::gzip_decompressor, continue into the gzip members that follow the first one.  When
    // the stream ends in an error (truncated, corrupt, trailing non-gzip bytes) Boost throws inside the stream
    // buffer, std::getline sets badbit and the caller's loop (match.cpp:61) never sees the line that was being
    // read: the unterminated tail of what could be inflated is dropped here to the same effect.
    std::string out;
    std::vector<char> chunk(1 << 20);
    size_t pos = 0;
    bool clean = true;
    while (pos < raw.size() && clean) {
      z_stream zs{};
      if (inflateInit2(&zs, 15 + 16) != Z_OK) { setstate(std::ios::badbit); return; }
      zs.next_in = reinterpret_cast<Bytef*>(raw.data() + pos);
      zs.avail_in = static_cast<uInt>(raw.size() - pos);
      const size_t in0 = zs.avail_in;
      int rc = Z_OK;
      while (rc != Z_STREAM_END) {
        zs.next_out = reinterpret_cast<Bytef*>(chunk.data());
        zs.avail_out = static_cast<uInt>(chunk.size());
        rc = inflate(&zs, Z_NO_FLUSH);
        if (rc != Z_OK && rc != Z_STREAM_END) {
          out.append(chunk.data(), chunk.size() - zs.avail_out);
          break;
        }
        out.append(chunk.data(), chunk.size() - zs.avail_out);
        if (rc == Z_OK && zs.avail_in == 0 && zs.avail_out != 0) break;
      }
      pos += in0 - zs.avail_in;
      inflateEnd(&zs);
      if (rc != Z_STREAM_END) clean = false;
    }
    if (!clean) {
      size_t nl = out.find_last_of('\n');
      out.resize(nl == std::string::npos ? 0 : nl + 1);
    }
    buf_.str(out);
  }
};
}}
