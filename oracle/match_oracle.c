/*
 * match_oracle.c -- CPU restatement ("port") of the reference keypoint matcher's hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is product code: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and
 * only as the checker.  The product path (frog_b200/csrc, include/frogmatch.h) never links it.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors, so this
 * port is pinned against the reference itself: oracle/Makefile compiles the unmodified
 * /root/reference/match/match.cpp (oracle/_ref/match_ref, oracle/_ref/libmatch_ref.so) and
 * tests/test_oracle.py checks this file against it and against the committed outputs of that
 * binary under tests/golden/.
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off (no FMA contraction, no fast-math: the reference's
 * default x86-64 build evaluates r += (a-b)*(a-b) with separately rounded sub, mul, add).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* match/match.cpp:243-251 -- scalar `norm`: float accumulator, k ascending. */
float mo_norm(const float* a, const float* b, int size) {
  float result = 0.0f;
  for (int i = 0; i < size; i++) {
    float diff = a[i] - b[i];
    float sq = diff * diff;
    result = result + sq;
  }
  return result;
}

/*
 * match/match.cpp:255-336 -- ComputeMatches(points2 = image `first`, points1 = image `second`).
 * Outer loop over rows of `second` (:262), inner over rows of `first` (:267); Laplacian gate
 * (:270), scale-ratio gate against the double constant 1.3 (:273-275), top-2 with strict '<'
 * (:303-313), acceptance (:320-321), emit orientation (:323-327).  `match` is deliberately NOT
 * reset per row (:259).  -all is restated separately below (mo_compute_matches_all); -anat is out of
 * scope.
 * out_pairs: 2 * n_second uint32.  Returns the number of matches.
 */
int64_t mo_compute_matches(const float* desc_first, const float* scale_first, const float* lap_first,
                           uint32_t n_first, const float* desc_second, const float* scale_second,
                           const float* lap_second, uint32_t n_second, uint32_t d, float threshold,
                           float dist2second, int sym, uint32_t* out_pairs) {
  int64_t n_out = 0;
  int match = 0;
  int end1 = (int)n_second;
  for (int i = 0; i < end1; i++) {
    float d1 = FLT_MAX, d2 = FLT_MAX;
    int end2 = (int)n_first;
    const float* di = desc_second + (size_t)i * d;
    for (int j = 0; j < end2; j++) {
      if (lap_second[i] != lap_first[j]) continue;
      /* float division, result promoted to double for the comparison with 1.3 */
      if (((double)(scale_second[i] / scale_first[j]) > 1.3) ||
          ((double)(scale_first[j] / scale_second[i]) > 1.3))
        continue;
      float dist = mo_norm(di, desc_first + (size_t)j * d, (int)d);
      if (dist < d1) {
        d2 = d1;
        d1 = dist;
        match = j;
      } else if (dist < d2) {
        d2 = dist;
      }
    }
    /* std::sqrt(float) -> float (match.cpp:28 `using namespace std` + <cmath>) */
    if ((sqrtf(d1 / d2) < dist2second || d2 == FLT_MAX) && sqrtf(d1) < threshold) {
      if (sym) {
        out_pairs[2 * n_out] = (uint32_t)i;
        out_pairs[2 * n_out + 1] = (uint32_t)match;
      } else {
        out_pairs[2 * n_out] = (uint32_t)match;
        out_pairs[2 * n_out + 1] = (uint32_t)i;
      }
      n_out++;
    }
  }
  return n_out;
}

/*
 * match/match.cpp:255-336 with matchAll == true (the `-all` key, :417-418).  Every gated-in column
 * whose distance is under the threshold emits a pair (:295-300) -- but the pair names `match`, the
 * running nearest among the columns that were NOT under the threshold (:303-313 is the `else` of
 * :295), carried over from earlier rows when the current row has not seen such a column yet
 * (`match` is declared outside the row loop, :259).  Nothing is emitted at the end of a row (:319).
 * out_pairs may be NULL (count only); otherwise at most `cap` pairs are written.  Returns the count.
 */
int64_t mo_compute_matches_all(const float* desc_first, const float* scale_first, const float* lap_first,
                               uint32_t n_first, const float* desc_second, const float* scale_second,
                               const float* lap_second, uint32_t n_second, uint32_t d, float threshold,
                               int sym, uint32_t* out_pairs, int64_t cap) {
  int64_t n_out = 0;
  int match = 0;
  for (int i = 0; i < (int)n_second; i++) {
    float d1 = FLT_MAX, d2 = FLT_MAX;
    const float* di = desc_second + (size_t)i * d;
    for (int j = 0; j < (int)n_first; j++) {
      if (lap_second[i] != lap_first[j]) continue;
      if (((double)(scale_second[i] / scale_first[j]) > 1.3) ||
          ((double)(scale_first[j] / scale_second[i]) > 1.3))
        continue;
      float dist = mo_norm(di, desc_first + (size_t)j * d, (int)d);
      if (sqrtf(dist) < threshold) {
        if (out_pairs && n_out < cap) {
          out_pairs[2 * n_out] = sym ? (uint32_t)i : (uint32_t)match;
          out_pairs[2 * n_out + 1] = sym ? (uint32_t)match : (uint32_t)i;
        }
        n_out++;
      } else if (dist < d1) {
        d2 = d1;
        d1 = dist;
        match = j;
      } else if (dist < d2) {
        d2 = dist;
      }
    }
  }
  return n_out;
}

/* -all over a list of image pairs (:638-652): counts[p] = size of pair p's list (forward pass, then the
 * -sym reverse pass appended); with out_pairs != NULL the lists are written at out_pairs + 2*out_offsets[p]. */
void mo_match_pairs_all(const float* desc, const float* scale, const float* lap, const int64_t* offsets,
                        uint32_t d, const uint32_t* pair_first, const uint32_t* pair_second, int64_t n_pairs,
                        float threshold, int sym, const int64_t* out_offsets, uint32_t* out_pairs,
                        int64_t* counts) {
#pragma omp parallel for schedule(dynamic)
  for (int64_t p = 0; p < n_pairs; p++) {
    uint32_t a = pair_first[p], b = pair_second[p];
    uint32_t na = (uint32_t)(offsets[a + 1] - offsets[a]), nb = (uint32_t)(offsets[b + 1] - offsets[b]);
    uint32_t* out = out_pairs ? out_pairs + 2 * out_offsets[p] : NULL;
    int64_t cap = out_pairs ? out_offsets[p + 1] - out_offsets[p] : 0;
    int64_t n = mo_compute_matches_all(desc + offsets[a] * d, scale + offsets[a], lap + offsets[a], na,
                                       desc + offsets[b] * d, scale + offsets[b], lap + offsets[b], nb, d,
                                       threshold, 0, out, cap);
    if (sym)
      n += mo_compute_matches_all(desc + offsets[b] * d, scale + offsets[b], lap + offsets[b], nb,
                                  desc + offsets[a] * d, scale + offsets[a], lap + offsets[a], na, d,
                                  threshold, 1, out ? out + 2 * n : NULL, out ? cap - n : 0);
    counts[p] = n;
  }
}

/*
 * match/match.cpp:617-652 -- the pair scheduler: OpenMP dynamic over image pairs, optional -sym
 * reverse pass appended (:643-646).  Images are flat arrays; offsets[k] is the first point of
 * image k in the concatenated arrays, offsets[n_images] the total.  Results for pair p land in
 * out_pairs + 2*out_offsets[p] where out_offsets is the caller's exclusive prefix sum of the
 * per-pair capacity (n_second, or n_first + n_second with sym).  counts[p] receives the size.
 */
void mo_match_pairs(const float* desc, const float* scale, const float* lap, const int64_t* offsets,
                    uint32_t d, const uint32_t* pair_first, const uint32_t* pair_second, int64_t n_pairs,
                    float threshold, float dist2second, int sym, const int64_t* out_offsets,
                    uint32_t* out_pairs, int64_t* counts) {
#pragma omp parallel for schedule(dynamic)
  for (int64_t p = 0; p < n_pairs; p++) {
    uint32_t a = pair_first[p], b = pair_second[p];
    uint32_t na = (uint32_t)(offsets[a + 1] - offsets[a]), nb = (uint32_t)(offsets[b + 1] - offsets[b]);
    uint32_t* out = out_pairs + 2 * out_offsets[p];
    int64_t n = mo_compute_matches(desc + offsets[a] * d, scale + offsets[a], lap + offsets[a], na,
                                   desc + offsets[b] * d, scale + offsets[b], lap + offsets[b], nb, d,
                                   threshold, dist2second, 0, out);
    if (sym)
      n += mo_compute_matches(desc + offsets[b] * d, scale + offsets[b], lap + offsets[b], nb,
                              desc + offsets[a] * d, scale + offsets[a], lap + offsets[a], na, d,
                              threshold, dist2second, 1, out + 2 * n);
    counts[p] = n;
  }
}

/*
 * match/match.cpp:179-208 -- readBinary.  `while(!feof)` runs one extra iteration after the last
 * full record: every fread fails, so the six header fields all take the value left in valF (the
 * last record's response) and the descriptor stays 48 zeros -- a phantom record that is matched
 * and written like any other.  A truncated tail behaves the same way field by field.
 * rec_out: capacity rows x 54 floats.  Returns the number of records produced.
 */
int64_t mo_read_bin(const char* path, float* rec_out, int64_t capacity) {
  FILE* f = fopen(path, "rb");
  if (!f) return -1;
  int64_t n = 0;
  float valF = 0.0f;
  while (!feof(f) && n < capacity) {
    float* r = rec_out + n * 54;
    for (int k = 0; k < 6; k++) {
      size_t got = fread(&valF, sizeof(float), 1, f);
      (void)got;
      r[k] = valF;
    }
    memset(r + 6, 0, 48 * sizeof(float));
    size_t got = fread(r + 6, sizeof(float), 48, f);
    (void)got;
    n++;
  }
  fclose(f);
  return n;
}

/*
 * match/match.cpp:137-176 (and :51-92 after inflate) -- CSV rows.  Cells split on ','; a cell
 * whose first byte is CR ends the row (:150); each cell goes through std::stof (= strtof, which
 * skips leading blanks and ignores trailing junk); a row is kept iff it has more than 6 cells
 * (:170).  Text must be NUL-terminated.  Rows with a descriptor length other than `d` are an
 * error for this flat-array port (returns -2).  Returns rows parsed.
 */
int64_t mo_parse_csv(const char* text, float* rec_out, int64_t capacity, int d) {
  int64_t n = 0;
  const char* p = text;
  char cell[256];
  while (*p && n < capacity) {
    const char* eol = strchr(p, '\n');
    const char* end = eol ? eol : p + strlen(p);
    float* r = rec_out + n * (6 + d);
    int count = 0;
    const char* c = p;
    while (c <= end) {
      if (c == end) break; /* std::getline on an exhausted line stream fails: no trailing empty cell */
      const char* comma = memchr(c, ',', (size_t)(end - c));
      const char* ce = comma ? comma : end;
      if (*c == 13) break;
      size_t len = (size_t)(ce - c);
      if (len >= sizeof(cell)) len = sizeof(cell) - 1;
      memcpy(cell, c, len);
      cell[len] = 0;
      float v = strtof(cell, NULL);
      if (count < 6 + d) r[count] = v;
      count++;
      if (!comma) break;
      c = comma + 1;
    }
    if (count > 6) {
      if (count != 6 + d) return -2;
      n++;
    }
    if (!eol) break;
    p = eol + 1;
  }
  return n;
}

/*
 * match/match.cpp:684-742 -- pairs.bin writer.  names are the basenames after the last '/' or
 * '\\' (:690-691); rigid doubles (:697-708); nPoints as u32 (INT_PTIDS, tools/pointIdType.h:2-3);
 * six floats per point (:715-723); then one block per computed pair in (i, j) row-major order,
 * INCLUDING empty ones, with i and j truncated to their low 16 bits (:735-736).
 * pts: per image [n][6] floats concatenated; pair blocks must already be in (i, j) order.
 */
int mo_write_pairs_bin(const char* path, int n_images, const char* const* names, const double* rigids,
                       const int64_t* offsets, const float* pts6, int64_t n_pairs,
                       const uint32_t* pair_first, const uint32_t* pair_second, const int64_t* counts,
                       const int64_t* out_offsets, const uint32_t* pairs) {
  FILE* f = fopen(path, "wb");
  if (!f) return 1;
  unsigned short nb = (unsigned short)n_images;
  fwrite(&nb, sizeof nb, 1, f);
  for (int it = 0; it < n_images; it++) {
    unsigned short len = (unsigned short)strlen(names[it]);
    fwrite(&len, sizeof len, 1, f);
    fwrite(names[it], 1, strlen(names[it]), f);
    fwrite(rigids + 3 * it, sizeof(double), 3, f);
    uint32_t np = (uint32_t)(offsets[it + 1] - offsets[it]);
    fwrite(&np, sizeof np, 1, f);
    fwrite(pts6 + offsets[it] * 6, sizeof(float), (size_t)np * 6, f);
  }
  for (int64_t p = 0; p < n_pairs; p++) {
    int i = (int)pair_first[p], j = (int)pair_second[p];
    unsigned int size = (unsigned int)counts[p];
    fwrite(&i, sizeof(unsigned short), 1, f);
    fwrite(&j, sizeof(unsigned short), 1, f);
    fwrite(&size, sizeof size, 1, f);
    fwrite(pairs + 2 * out_offsets[p], 8, size, f);
  }
  fclose(f);
  return 0;
}
