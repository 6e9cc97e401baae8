// Test infrastructure only: in-process access to the UNMODIFIED reference hot path.
//
// This translation unit #includes the reference's match/match.cpp where it lies (through the
// symlink tree that oracle/Makefile creates under oracle/_ref/src) with its main() renamed, and
// exports a C ABI over the reference's own norm() (match.cpp:243-251) and ComputeMatches()
// (match.cpp:255-336).  Nothing here is shipped or measured as product code; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline leg may load the resulting library.
#define main ref_match_main
#include "match/match.cpp"
#undef main

#include <cstdint>

static void fill_points(Points& pts, const float* desc, const float* scale, const float* lap,
                        uint32_t n, uint32_t d) {
  pts.resize(n);
  for (uint32_t i = 0; i < n; i++) {
    pts[i].desc.assign(desc + (size_t)i * d, desc + (size_t)(i + 1) * d);
    pts[i].scale = scale[i];
    pts[i].laplacianSign = lap[i];
    pts[i].response = 0;
    for (int k = 0; k < 3; k++) pts[i].coordinates[k] = pts[i].transformedCoordinates[k] = 0;
  }
}

extern "C" {

float ref_norm(const float* a, const float* b, int n) {
  Descriptor da(a, a + n), db(b, b + n);
  return norm(da, db, n);
}

// ComputeMatches(points2 = image `first`, points1 = image `second`, ...) exactly as main() calls
// it at match.cpp:642 (sym=0) or :644 (sym=1, caller swaps the images).
// out_pairs must hold 2*n_second uint32; returns the number of matches written.
int64_t ref_compute_matches(const float* desc_first, const float* scale_first, const float* lap_first,
                            uint32_t n_first, const float* desc_second, const float* scale_second,
                            const float* lap_second, uint32_t n_second, uint32_t d, float threshold,
                            float dist2second, int sym, uint32_t* out_pairs) {
  Points p2, p1;
  fill_points(p2, desc_first, scale_first, lap_first, n_first, d);
  fill_points(p1, desc_second, scale_second, lap_second, n_second, d);
  MatchVect* m = ComputeMatches(p2, p1, threshold, dist2second, false, 0.0f, sym != 0);
  int64_t n = (int64_t)m->size();
  for (int64_t k = 0; k < n; k++) {
    out_pairs[2 * k] = (*m)[k].first;
    out_pairs[2 * k + 1] = (*m)[k].second;
  }
  delete m;
  return n;
}

// The same with matchAll = true (the `-all` key, match.cpp:417-418).  Writes at most `cap` pairs, returns the
// size of the reference's list.
int64_t ref_compute_matches_all(const float* desc_first, const float* scale_first, const float* lap_first,
                                uint32_t n_first, const float* desc_second, const float* scale_second,
                                const float* lap_second, uint32_t n_second, uint32_t d, float threshold,
                                int sym, uint32_t* out_pairs, int64_t cap) {
  Points p2, p1;
  fill_points(p2, desc_first, scale_first, lap_first, n_first, d);
  fill_points(p1, desc_second, scale_second, lap_second, n_second, d);
  MatchVect* m = ComputeMatches(p2, p1, threshold, 1.0f, true, 0.0f, sym != 0);
  int64_t n = (int64_t)m->size();
  for (int64_t k = 0; k < n && k < cap; k++) {
    out_pairs[2 * k] = (*m)[k].first;
    out_pairs[2 * k + 1] = (*m)[k].second;
  }
  delete m;
  return n;
}

// Reference distances for chosen (row of second image, column of first image) pairs.
void ref_distances(const float* desc_first, const float* desc_second, uint32_t d,
                   const uint32_t* first_idx, const uint32_t* second_idx, int64_t n, float* out) {
  for (int64_t k = 0; k < n; k++) {
    Descriptor a(desc_second + (size_t)second_idx[k] * d, desc_second + (size_t)(second_idx[k] + 1) * d);
    Descriptor b(desc_first + (size_t)first_idx[k] * d, desc_first + (size_t)(first_idx[k] + 1) * d);
    out[k] = norm(a, b, (int)d);
  }
}

int ref_main(int argc, char** argv) { return ref_match_main(argc, argv); }

}  // extern "C"
