/*
 * fast_oracle.c -- a FAST CPU restatement of ComputeMatches (match/match.cpp:255-336) for parity checks at
 * BASELINE scale (50k x 50k and 20k x 20k image pairs), SURVEY.md 8c "optional fast oracle".
 *
 * TEST INFRASTRUCTURE ONLY (see match_oracle.c).  Parity status: PINNED -- tests/test_oracle.py checks it against
 * the plain port (match_oracle.c) and against the verbatim reference build on inputs with ties, gate boundaries,
 * single survivors and duplicates before any large-scale test relies on it.
 *
 * How it can be fast and still bit-identical: the reference's `norm` (match.cpp:243-251) is a dependent chain over
 * k for ONE (row, column) pair, but different columns are independent.  Image `first` is transposed to [k][column],
 * so the compiler vectorises ACROSS COLUMNS: every SIMD lane runs the reference's exact sequence
 *     r = fl(r + fl(fl(b_k - a_k) * fl(b_k - a_k))),  k = 0 .. d-1
 * for its own column (built with -ffp-contract=off: no FMA; no reassociation is needed or allowed).  Gates
 * (match.cpp:270, :273-275), the strict-'<' top-2 scan in ascending column order (:303-313) and the acceptance rule
 * (:320-321) are the port's, applied to a tile of finished distances.  `match` is not reset per row (:259).
 * Rows are independent except for that carried `match`, which can only be observed when a row has no surviving
 * column AND is accepted -- impossible for thresholds below sqrt(FLT_MAX) (d1 stays FLT_MAX); the OpenMP split over
 * rows therefore cannot change the output for the thresholds the tests use (asserted: threshold < 1.8e19).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FO_TILE 512

/* distances of one row against FO_TILE transposed columns: lane-per-column, k sequential */
__attribute__((target_clones("avx512f", "avx2", "default")))
static void fo_tile_dist(const float* restrict row, const float* restrict at, size_t ld, uint32_t d, float* restrict acc) {
  /* blocks of 128 columns: the accumulators of a block stay in vector registers over the whole k loop */
  for (int jb = 0; jb < FO_TILE; jb += 128) {
    float a[128];
    for (int j = 0; j < 128; j++) a[j] = 0.0f;
    for (uint32_t k = 0; k < d; k++) {
      const float bk = row[k];
      const float* restrict col = at + (size_t)k * ld + jb;
      for (int j = 0; j < 128; j++) {
        const float diff = bk - col[j];
        const float sq = diff * diff;
        a[j] = a[j] + sq;
      }
    }
    for (int j = 0; j < 128; j++) acc[jb + j] = a[j];
  }
}

/* Gates of match.cpp:270 and :273-275 for one row against a tile of columns, also lane-per-column: a gated-out
 * column gets distance +inf (never below d1 / d2, exactly like `continue`).  Returns the smallest surviving distance. */
__attribute__((target_clones("avx512f", "avx2", "default")))
static float fo_tile_gate(float si, float li, const float* restrict sc, const float* restrict lp, float* restrict acc) {
  float mn[16];
  for (int l = 0; l < 16; l++) mn[l] = INFINITY;
  for (int j0 = 0; j0 < FO_TILE; j0 += 16)
    for (int l = 0; l < 16; l++) {
      const int j = j0 + l;
      const int out = (li != lp[j]) | ((double)(si / sc[j]) > 1.3) | ((double)(sc[j] / si) > 1.3);
      const float v = out ? INFINITY : acc[j];
      acc[j] = v;
      mn[l] = v < mn[l] ? v : mn[l];
    }
  float m = INFINITY;
  for (int l = 0; l < 16; l++) m = mn[l] < m ? mn[l] : m;
  return m;
}

int64_t fo_compute_matches(const float* desc_first, const float* scale_first, const float* lap_first,
                           uint32_t n_first, const float* desc_second, const float* scale_second,
                           const float* lap_second, uint32_t n_second, uint32_t d, float threshold,
                           float dist2second, int sym, uint32_t* out_pairs) {
  if (!(threshold < 1.8e19f)) return -1; /* see header: the carried `match` would become observable */
  const size_t ld = ((size_t)n_first + FO_TILE - 1) / FO_TILE * FO_TILE;
  const size_t ldp = ld ? ld : FO_TILE;
  float* at = (float*)aligned_alloc(64, ldp * (size_t)(d ? d : 1) * sizeof(float));
  float* sc = (float*)aligned_alloc(64, ldp * sizeof(float));  /* padded copies: columns past n_first never pass the gate */
  float* lp = (float*)aligned_alloc(64, ldp * sizeof(float));
  uint32_t* res = (uint32_t*)malloc(((size_t)n_second + 1) * sizeof(uint32_t));
  if (!at || !sc || !lp || !res) { free(at); free(sc); free(lp); free(res); return -2; }
  memset(at, 0, ldp * (size_t)(d ? d : 1) * sizeof(float));
  for (size_t j = 0; j < ldp; j++) {
    sc[j] = j < n_first ? scale_first[j] : 1.0f;
    lp[j] = j < n_first ? lap_first[j] : NAN; /* NaN != anything: the laplacian gate rejects the padding */
  }
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < (int64_t)n_first; j++)
    for (uint32_t k = 0; k < d; k++) at[(size_t)k * ld + j] = desc_first[(size_t)j * d + k];

#pragma omp parallel
  {
    float acc[FO_TILE] __attribute__((aligned(64)));
#pragma omp for schedule(dynamic, 16)
    for (int64_t i = 0; i < (int64_t)n_second; i++) {
      float d1 = FLT_MAX, d2 = FLT_MAX;
      uint32_t match = 0;
      const float si = scale_second[i], li = lap_second[i];
      for (size_t j0 = 0; j0 < n_first; j0 += FO_TILE) {
        fo_tile_dist(desc_second + (size_t)i * d, at + j0, ld, d, acc);
        const float tile_min = fo_tile_gate(si, li, sc + j0, lp + j0, acc);
        if (!(tile_min < d2)) continue; /* no column of the tile can change d1, d2 or match */
        for (size_t jj = 0; jj < FO_TILE; jj++) { /* match.cpp:303-313, ascending column order, strict '<' */
          const float dist = acc[jj];
          if (dist < d1) { d2 = d1; d1 = dist; match = (uint32_t)(j0 + jj); }
          else if (dist < d2) { d2 = dist; }
        }
      }
      res[i] = ((sqrtf(d1 / d2) < dist2second || d2 == FLT_MAX) && sqrtf(d1) < threshold) ? match : 0xFFFFFFFFu;
    }
  }
  int64_t n_out = 0;
  for (uint32_t i = 0; i < n_second; i++) {
    if (res[i] == 0xFFFFFFFFu) continue;
    if (sym) { out_pairs[2 * n_out] = i; out_pairs[2 * n_out + 1] = res[i]; }
    else { out_pairs[2 * n_out] = res[i]; out_pairs[2 * n_out + 1] = i; }
    n_out++;
  }
  free(at);
  free(sc);
  free(lp);
  free(res);
  return n_out;
}
