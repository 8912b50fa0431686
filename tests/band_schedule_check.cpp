// Host-only check of the band kernel's work list (csrc/band_schedule.h): compiled and run by tests/test_band_schedule.py.
// For a sweep of (rows, strips, groups, K, edge bands) it rebuilds the list exactly as launch_band_k does, decodes every
// piece with the function the kernel uses, and checks that the pieces tile [0, rows) x strips exactly once, that sizes
// never grow along the interior list (the one-piece-ahead fetch relies on it), that the edge bands are the last levels,
// and that nothing overflows BandSched::MAXLEV.  Prints "OK <cases>" or the first violation.
#include <cstdio>
#include <vector>
#include "../terrainwatersim_b200/csrc/band_schedule.h"

using namespace tws;

static int fail(const char* what, int rows, int nstrips, int groups, int K, int et, int eb) {
  std::printf("FAIL %s rows=%d nstrips=%d groups=%d K=%d e_top=%d e_bot=%d\n", what, rows, nstrips, groups, K, et, eb);
  return 1;
}

int main() {
  long cases = 0;
  const int rows_list[] = {1, 2, 7, 8, 9, 12, 16, 17, 40, 100, 255, 256, 1000, 1024, 4096, 8184, 8192, 16376, 32768, 65536, 1000003};
  const int strips_list[] = {1, 2, 3, 10, 74, 293, 586};
  const int groups_list[] = {2, 4, 74, 292, 296};
  for (int K = 1; K <= 4; ++K) {
    const int HP = 2 * K, BR = 12;
    for (int rows : rows_list)
      for (int nstrips : strips_list)
        for (int groups : groups_list)
          for (int mode = 0; mode < 4; ++mode) {             // none, top edge, bottom edge, both
            int et = (mode & 1) ? 8 : 0, eb = (mode & 2) ? 8 : 0;
            if (et + eb >= rows) { if (!mode) { et = eb = 0; } else { et = rows; eb = 0; } }
            int l0 = 0, nle = 0, esegs = 0;
            const BandSched s = band_build_schedule(rows, nstrips, groups, BR, HP, et, eb, &l0, &nle, &esegs);
            ++cases;
            if (s.nlev < 1 || s.nlev > BandSched::MAXLEV) return fail("level count", rows, nstrips, groups, K, et, eb);
            if (l0 + nle != s.nlev || nle != (et > 0) + (eb > 0) || esegs != nle) return fail("edge levels", rows, nstrips, groups, K, et, eb);
            if (s.npieces % nstrips) return fail("npieces", rows, nstrips, groups, K, et, eb);
            std::vector<int> cover((size_t)rows, 0);
            int prev_size = 1 << 30;
            for (int p = 0; p < s.npieces; p += nstrips) {   // strip 0 of every segment (all strips share the rows)
              int lev, strip, ya, yb;
              band_decode(s, p, nstrips, lev, strip, ya, yb);
              if (strip != 0 || lev < 0 || lev >= s.nlev || ya < 0 || yb > rows || yb <= ya) return fail("decode", rows, nstrips, groups, K, et, eb);
              int lev2, strip2, ya2, yb2;
              band_decode(s, p + nstrips - 1, nstrips, lev2, strip2, ya2, yb2);
              if (strip2 != nstrips - 1 || ya2 != ya || yb2 != yb || lev2 != lev) return fail("strip decode", rows, nstrips, groups, K, et, eb);
              for (int y = ya; y < yb; ++y) cover[(size_t)y]++;
              const bool is_edge = lev >= l0;
              if (is_edge != (yb <= et || ya >= rows - eb)) return fail("edge piece rows", rows, nstrips, groups, K, et, eb);
              if (!is_edge) {
                if (yb - ya > prev_size) return fail("interior sizes grow", rows, nstrips, groups, K, et, eb);
                prev_size = s.size[lev];
              }
            }
            for (int y = 0; y < rows; ++y)
              if (cover[(size_t)y] != 1) return fail("coverage", rows, nstrips, groups, K, et, eb);
          }
  }
  std::printf("OK %ld\n", cases);
  return 0;
}
