// Device-resident state of the exact-median selection (svgd.cu) shared with the pipelined tensor-core kernels (svgd_tc2.cu).
#pragma once
#include "common.cuh"

namespace bode {

// prefix[0..1] = bit patterns of the two middle order statistics (equal when the count is odd), rank = remaining ranks.
// Window fields: the pipelined Gram kernel counts the entries below `win_lo` and histograms the raw bit patterns inside
// [win_lo, win_lo + WIN_SPAN] (a dense table: one counter per representable float).  When both middle ranks fall inside the
// window the median is read off the table (`hit`) and the three radix passes are skipped; the window for the next call is
// centred on the median just found (consecutive SVGD steps move the median by far less than the window's +-0.2 %).
struct SelState {
  unsigned int prefix[2];
  unsigned int maxbits;
  unsigned int hit;          // 1: the window resolved the median this call (radix passes return immediately)
  unsigned long long rank[2];
  unsigned int win_valid;    // a window is armed (set by gamma_kernel at the end of a call, read by the next Gram pass)
  unsigned int win_lo;       // lowest bit pattern covered
  unsigned int pad[2];
};

constexpr unsigned int WIN_HALF = 16384;                 // +- ulps around the previous median (2^-9 relative at most)
constexpr unsigned int WIN_SPAN = 2 * WIN_HALF;          // table covers win_lo .. win_lo + WIN_SPAN inclusive
constexpr unsigned int WIN_TABLE = WIN_SPAN + 1;         // counters; entry WIN_TABLE holds the count below the window

}  // namespace bode
