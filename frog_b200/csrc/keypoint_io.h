// keypoint_io.h -- host side of the drop-in `match` executable: keypoint-file readers, filters,
// pruning and the pairs.bin writer.  Behaviour (including the reference's quirks) follows
// match/match.cpp; every function cites the lines it replaces.
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <vector>

namespace fmio {

// One image's keypoints as flat arrays (the reference keeps a vector<Point>, match.cpp:39-48).
struct KeypointSet {
  uint32_t n = 0;
  uint32_t d = 0;               // descriptor length (0 until the first row is seen)
  std::vector<float> head;      // n x 6: x, y, z, scale, laplacianSign, response
  std::vector<float> desc;      // n x d
  const float* row_head(uint32_t i) const { return head.data() + (size_t)i * 6; }
};

// Readers.  Return false and fill `err` on I/O failure or on input the reference itself would
// abort on (std::stof throwing, match.cpp:69/152).
bool read_csv(const std::string& path, KeypointSet& out, std::string& err);      // match.cpp:137-176
bool read_csv_gz(const std::string& path, KeypointSet& out, std::string& err);   // match.cpp:51-92
bool read_bin(const std::string& path, KeypointSet& out, std::string& err);      // match.cpp:179-208
// Parse CSV text already in memory (NUL-terminated); shared by the two text readers.
bool parse_csv_text(const char* text, size_t len, KeypointSet& out, std::string& err);

// Extension dispatch of match.cpp:514,528-536 ("csv" | "bin" | "gz").
bool read_keypoints(const std::string& path, KeypointSet& out, std::string& err);

// z-window filter, match.cpp:538-546: drop points with z + zT outside [zmin, zmax] (float math).
void filter_z(KeypointSet& k, float zT, float zmin, float zmax);
// Response threshold + top-np pruning, match.cpp:585-595 (same std::partial_sort call, same
// comparator, so survivors come out in the reference's order).
void prune(KeypointSet& k, float sp, int np);

struct PairBlock {
  int first, second;            // image ids (written as their low 16 bits, match.cpp:735-736)
  const uint32_t* pairs;        // 2 * count uint32: (first_idx, second_idx)
  uint32_t count;
};

// pairs.bin writer, match.cpp:675-742.  `blocks` must be in the order the reference's nested
// i/j loops visit them.  Returns false if the file cannot be opened.
bool write_pairs_bin(const std::string& path, const std::vector<std::string>& filenames,
                     const std::vector<std::array<double, 3>>& rigids, const std::vector<KeypointSet>& images,
                     const std::vector<PairBlock>& blocks);

// Test hook for the libc-free decimal parser (keypoint_io.cpp fast_strtof).
int debug_cell_to_float(const char* c, size_t len, float* v, int* endp_off);

}  // namespace fmio
