// band_schedule.h — the work list of one band-kernel launch (pure C++: shared by band_kernels.cu and the host-only
// schedule test, tests/band_schedule_check.cpp).
#pragma once
#include <algorithm>

#ifdef __CUDACC__
#define TWS_HD __host__ __device__
#else
#define TWS_HD
#endif

// tuning knobs of the guided schedule (defaults measured on 8192^2 / 32768^2, see profiles/)
#ifndef TWS_BAND_SMIN
#define TWS_BAND_SMIN 128        // shortest segment (rows) on large grids
#endif
#ifndef TWS_BAND_FIRST_DIV
#define TWS_BAND_FIRST_DIV 2     // first-level segment = even share / this
#endif

namespace tws {

// Work list of one launch.  A piece is (row segment, column strip); every warp group starts on piece
// `its index` and then takes pieces from a device-wide counter, so groups whose cells need the slow
// paths (the IEEE division of draining wet cells: lake and shore regions cost up to 1.4x the
// instructions of dry land) simply take fewer pieces.  Segments shrink geometrically down the strip
// ("guided" schedule: about half of the remaining rows per level), so the pieces handed out last are
// small and the groups finish together, while most rows are covered by long pieces whose 2K warm-up /
// feeder rows amortise.  Piece p = (segment p / nstrips, strip p % nstrips): the pieces in flight at any
// time are the same rows of neighbouring strips, which keeps their shared halo columns in L2.
struct BandSched {
  static constexpr int MAXLEV = 14;
  int nlev, npieces;
  int seg0[MAXLEV];    // first segment of level l
  int y0[MAXLEV];      // first row (relative to lr0) of level l
  int yend[MAXLEV];    // end row of level l
  int size[MAXLEV];    // rows per segment of level l (the last segment of a level may be shorter)
};

// Piece p of the list -> its level, column strip and row range [ya, yb) relative to the launch's first row.
TWS_HD inline void band_decode(const BandSched& sch, int piece, int nstrips, int& lev, int& strip, int& ya, int& yb) {
  const int seg = piece / nstrips;
  strip = piece - seg * nstrips;
  lev = 0;
  while (lev + 1 < sch.nlev && seg >= sch.seg0[lev + 1]) ++lev;
  ya = sch.y0[lev] + (seg - sch.seg0[lev]) * sch.size[lev];
  yb = ya + sch.size[lev] < sch.yend[lev] ? ya + sch.size[lev] : sch.yend[lev];
}

// Guided schedule (see BandSched): appends levels covering rows [y_begin, y_end) of `nstrips` strips for up to
// `max_groups` warp groups to `s`; `seg` is the running segment count.
inline void band_schedule_rows(BandSched& s, int& seg, int y_begin, int y_end, int nstrips, int max_groups, int BR, int HP) {
  const int rows = y_end - y_begin;
  if (rows <= 0) return;
  const long long total = (long long)rows * nstrips;
  const int even = (int)((total + max_groups - 1) / max_groups);            // rows per group if the work were split evenly
  // shortest segment: long enough that the 2K warm-up / feeder rows amortise (128 rows), but small grids are
  // latency bound and rather use every SM (at least ~4 bands per piece)
  const int smin = std::max(4 * BR, std::min(TWS_BAND_SMIN, even));
  // a piece computes rows + 2*HP rows in whole bands: sizes that make that a multiple of BR waste nothing
  auto whole_bands = [&](int sz) { return std::max(BR, (sz + 2 * HP + BR - 1) / BR * BR) - 2 * HP; };
  int size = whole_bands(std::max(smin, even / TWS_BAND_FIRST_DIV));
  int y = y_begin;
  while (y < y_end) {
    const int l = s.nlev;
    const int rem = y_end - y;
    const bool rest = size <= whole_bands(smin) || l >= BandSched::MAXLEV - 3;      // two levels stay free for a strip's edge bands
    const int n = rest ? (rem + size - 1) / size : std::max(1, rem / 2 / size);   // about half of what is left per level
    const int cover = (int)std::min<long long>(rem, (long long)n * size);
    s.seg0[l] = seg; s.y0[l] = y; s.yend[l] = y + cover; s.size[l] = size;
    seg += n;
    y += cover;
    s.nlev = l + 1;
    size = whole_bands(std::max(smin, size / 2));
  }
}
// one level holding a single segment [y_begin, y_end): an edge band of a strip
inline void band_schedule_one(BandSched& s, int& seg, int y_begin, int y_end) {
  if (y_end <= y_begin) return;
  const int l = s.nlev;
  s.seg0[l] = seg; s.y0[l] = y_begin; s.yend[l] = y_end; s.size[l] = y_end - y_begin;
  seg += 1;
  s.nlev = l + 1;
}

// The whole work list of a launch over `rows` rows: guided levels over the interior [e_top, rows - e_bot), then one
// single-segment level per edge band (strips only; e_top = e_bot = 0 otherwise).  Outputs where the edge levels sit.
inline BandSched band_build_schedule(int rows, int nstrips, int max_groups, int BR, int HP, int e_top, int e_bot, int* lev_edge0,
                                     int* nlev_edge, int* edge_segs) {
  BandSched s{};
  int seg = 0;
  band_schedule_rows(s, seg, e_top, rows - e_bot, nstrips, max_groups, BR, HP);
  const int l0 = s.nlev, seg_i = seg;
  band_schedule_one(s, seg, 0, e_top);
  band_schedule_one(s, seg, rows - e_bot, rows);
  if (lev_edge0) *lev_edge0 = l0;
  if (nlev_edge) *nlev_edge = s.nlev - l0;
  if (edge_segs) *edge_segs = seg - seg_i;
  s.npieces = seg * nstrips;
  return s;
}

}  // namespace tws
