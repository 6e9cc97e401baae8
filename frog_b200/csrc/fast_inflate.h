// fast_inflate.h -- table-driven gzip/DEFLATE decoder for in-memory `.csv.gz` keypoint files.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace fmio {

// Inflate the single gzip member src[0, n) into `out` (resized to the member's ISIZE plus slack; the
// first *produced bytes are the data).  Returns false -- leaving `out` unspecified -- for anything
// other than one well-formed member whose size and CRC-32 check out; callers then fall back to zlib.
bool fast_inflate_gzip(const uint8_t* src, size_t n, std::vector<char>& out, size_t* produced);

}  // namespace fmio
