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

// ---------------------------------------------------------------------------------------------------------------------------
// Peer access for the distributed median (several ranks on one NVLink / NVSwitch node).  Every rank's SVGD workspace is one
// cudaMalloc'ed block mapped into all the others (cudaIpc), with identical layout, so rank r reads rank q's window table or
// radix histogram at `base[q] + offset` straight over NVLink and the selection needs no collective launch: the kernels meet at
// flag barriers instead.  flags[q] (in MY workspace) is written by rank q with a monotonically increasing epoch; the epoch
// counter itself is local and advances identically on every rank because all ranks run the same kernel sequence.
constexpr int MAX_PEERS = 8;
struct PeerInfo {
  unsigned char* base[MAX_PEERS];
  int rank, world;                       // world <= 1: single rank, nothing below is touched
  unsigned long long hist_off, table_off, flag_off;
};
struct PeerFlags {
  unsigned int arrived[MAX_PEERS];
  unsigned int epoch;
  unsigned int timed_out;                // a barrier gave up after ~2^24 polls (a peer never arrived): results are invalid
};

// One thread per rank calls this; everything the caller's GPU wrote before is visible to the peers after they pass.
__device__ __forceinline__ void peer_barrier(const PeerInfo& p) {
  PeerFlags* mine = reinterpret_cast<PeerFlags*>(p.base[p.rank] + p.flag_off);
  const unsigned int e = mine->epoch + 1u;
  mine->epoch = e;
  __threadfence_system();
  for (int q = 0; q < p.world; ++q) {
    unsigned int* f = &reinterpret_cast<PeerFlags*>(p.base[q] + p.flag_off)->arrived[p.rank];
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(e) : "memory");
  }
  for (int q = 0; q < p.world; ++q) {
    unsigned int v;
    long long spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(&mine->arrived[q]) : "memory");
    } while ((int)(v - e) < 0 && ++spins < (1ll << 24));
    if ((int)(v - e) < 0) mine->timed_out = 1u;
  }
}
__device__ __forceinline__ unsigned long long peer_ld_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

constexpr unsigned int WIN_HALF = 16384;                 // +- ulps around the previous median (2^-9 relative at most)
constexpr unsigned int WIN_SPAN = 2 * WIN_HALF;          // table covers win_lo .. win_lo + WIN_SPAN inclusive
constexpr unsigned int WIN_TABLE = WIN_SPAN + 1;         // counters; entry WIN_TABLE holds the count below the window

}  // namespace bode
