// fmio_capi.cpp -- C wrappers over the host-side readers / pruning / pairs.bin writer so the
// CPU test-suite (and Python callers) can exercise exactly the code bin/match runs.
#include <cstring>

#include "keypoint_io.h"

extern "C" {

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

}  // extern "C"
