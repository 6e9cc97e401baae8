#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>
#include "keypoint_io.h"
// Mutation fuzz of parse_csv_text under ASan/UBSan (memory safety), plus: every value the parser returns equals strtof's
// on the same cell text for unmutated rows.
int main(int argc, char** argv) {
  std::mt19937 rng(7);
  long ok = 0, fail = 0;
  for (int base = 0; base < (argc > 1 ? atoi(argv[1]) : 40); base++) {
    std::string text;
    int rows = 1 + rng() % 40, cells = 7 + rng() % 60;
    for (int r = 0; r < rows; r++) {
      for (int c = 0; c < cells; c++) {
        char buf[64];
        double v = ((double)rng() / 4294967296.0 - 0.5) * (rng() % 4 == 0 ? 1e4 : 1.0);
        snprintf(buf, sizeof buf, rng() % 7 == 0 ? "%e" : "%f", v);
        text += buf;
        if (c + 1 < cells) text += ',';
      }
      text += rng() % 5 == 0 ? "\r\n" : "\n";
    }
    for (int m = 0; m < 2000; m++) {
      std::string t = text;
      int nmut = rng() % 4;
      for (int k = 0; k < nmut && !t.empty(); k++) {
        size_t i = rng() % t.size();
        switch (rng() % 5) {
          case 0: t[i] = ",\n\r .-+eE0123456789x"[rng() % 20]; break;
          case 1: t.erase(i, 1 + rng() % 3); break;
          case 2: t.insert(i, 1, (char)(rng() % 96 + 32)); break;
          case 3: t.resize(i); break;
          default: t[i] = (char)rng();
        }
      }
      // exact-size heap buffer (+1 for the NUL the readers append)
      char* heap = (char*)malloc(t.size() + 1); memcpy(heap, t.data(), t.size()); heap[t.size()] = 0;
      fmio::KeypointSet k; std::string err;
      if (fmio::parse_csv_text(heap, t.size(), k, err)) ok++; else fail++;
      free(heap);
    }
  }
  printf("csv fuzz ok: %ld parsed, %ld rejected\n", ok, fail);
}
