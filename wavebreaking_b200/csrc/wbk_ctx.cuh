// Detection context: device arenas of one batch of jobs (job = time step x contour level).
#pragma once
#include "wbk_common.cuh"

#define WBK_MAX_LEVELS 16
#define WBK_VERTEX_ID 0x80000000u  // point id of a contour vertex that coincides with a grid vertex
#ifndef WBK_CONTOUR_THREADS
#define WBK_CONTOUR_THREADS 512  // 2 CTAs per SM: all 296 jobs of a batch are resident at once (1024: 1.38 us, 512: 1.32, 256: 1.84 per step)
#endif

// Device-side view of the arenas (passed to kernels by value).  Per-job arrays have a fixed
// stride; `S` = seg_cap, `CC` = contour_cap, `R` = S + CC (raw contour points per job),
// `HC` = hash capacity per job (power of two >= 2 * R).
struct WbkDev {
  int nlat, nlon, add, W;  // W = nlon + add (extended width)
  int S, CC, R, HC;
  int sel_cap, pair_cap, event_cap, max_jobs;

  // marching-squares segments (written by ms_segments_kernel)
  u32 *rid, *fpid, *tpid, *fxy, *txy;  // [J][S]
  int* seg_count;                      // [J]
  int* status;                         // [J]
  // linking scratch
  u32 *nxt, *prv;           // [J][S]
  u64* w64;                 // [J][S]
  u32 *cminrid, *cnseg, *cidx, *cflag;  // [J][S]   (indexed by head segment)
  u64* hkeys;               // [J][HC]
  u32* hvals;               // [J][HC]
  u32 *raw, *rawc, *slot;   // [J][R]
  int* scan;                // [J][R]
  // per-contour (assembly order) scratch [J][CC]
  u64* sortkeys;            // [J][CCp2]
  int *craw_off, *craw_n, *cclosed, *cnuniq, *cdrop, *cymin, *cymax, *cout, *cand;
  int CCp2;
  // results of the contour stage
  u32* out_pts;             // [J][R]
  int* out_tab;             // [J][CC][4]: pt_off, npts, closed, nx
  int* out_sumy;            // [J][CC]
  int *out_nc, *out_np;     // [J]
  int* max_nx;              // [1]
  u32* planes;              // comparison bit planes written by the fused smoothing (wbk_smooth.cu), read by ms_planes_kernel
};

// Index-stage arenas.  PC = pair_cap rounded up to a power of two, EC = event_cap.
struct WbkIdx {
  int PC, EC, SC;            // SC = sel_cap
  int* sel;                  // [J][SC] packed-set contour index of the full-width contours
  int* nsel;                 // [J]
  u64* pairs;                // [J][SC][PC]  candidate pairs (i << 32 | j)
  int* pair_count;           // [J][SC]
  int* tile_off;             // [J*SC + 1]
  int TLC;                   // capacity of the active-tile list
  u64* tile_list;            // [TLC] active pair-scan tiles of the batch: slot << 40 | bi << 20 | bj
  double* blk_pf;            // [J][SC][NB][2] min / max along-contour prefix of every block of PT points
  int NB;                    // blocks of PT points per contour (capacity)
  int* blk_x;                // [J*SC][NB][4] column / row range of every block of a full-width contour
  u64* pairs_b;              // [J][PC]
  int* flag;                 // [J][SC][PC] keep flags of the touch kernel
  int *scanb, *label;        // [J][PC]
  u64 *hm1, *hm2;            // [J][SC][2*PC] smallest / second smallest pair key of a duplicate group
  int* cnt1;                 // [J][SC] pairs left after check_duplicates
  int* touch_off;            // [J*SC + 1] (unused since the touch kernel runs one CTA per contour)
  u64* pairs2;               // [J][SC][PC]  pairs left after check_duplicates (unordered)
  u64* hk;                   // [J][SC][2*PC]
  u32 *hv1, *hv2;            // [J][2*PC]
  int* ev_int;               // [3][J][EC][WBK_EV_INTS]
  double* ev_f64;            // [3][J][EC][WBK_EV_F64]
  int* ev_count;             // [3][J]
  int* ev_off;               // [3*J + 1]
  int* total;                // [4] scratch totals
  // meridian split (utils/index_utils.py:148-173) of the events that straddle the last meridian
  int SPV, SPR;              // vertex pool / ring list capacities
  int* split_xy;             // [SPV][2] vertex pool (chain scratch and final folded, truncated pieces)
  int* split_ring;           // [SPR][4] start, len, time index, kind
  int* split_count;          // [0] vertex cursor, [1] ring cursor, [2] overflow flag, [3] split_list cursor
  int* split_list;           // [SPR] events (index in ev_off order) that straddle the last meridian
  // near-threshold decisions of the streamer pair scan (accepted AND rejected pairs): job, contour, i, j | flags << 28
  int NRC;                   // capacity of the list
  int* near_rec;             // [NRC][4]
  int* near_cnt;             // [1] number of decisions seen (may exceed NRC)
};

struct wbk_ctx {
  wbk_caps caps;
  WbkDev d;
  WbkIdx x;
  int njobs;      // jobs of the last wbk_contours call
  int nlevels;
  void* owned;    // workspace allocated by the library (NULL if caller-owned)
  int split_clipped;  // the meridian-split work list of the last wbk_events_raster has been clipped into pieces
};
