// Marching-squares building blocks shared by the stand-alone segment kernel (wbk_contours.cu) and the fused
// smoothing + segment kernel (wbk_spatial.cu).  Reference: skimage.measure._find_contours_cy._get_contour_segments
// as called by wavebreaking/indices/contour_index.py:103.
#pragma once
#include "wbk_ctx.cuh"

struct LevelPack {
  double v[WBK_MAX_LEVELS];
};

#define MS_THREADS 256
#define MS_ROWS 16
#ifndef MS_MIN_CTAS
#define MS_MIN_CTAS 3
#endif

__device__ __forceinline__ double ms_fraction(double from_value, double to_value, double level) {
  if (to_value == from_value) return 0.0;
  return __ddiv_rn(__dsub_rn(level, from_value), __dsub_rn(to_value, from_value));
}

// edges: 0 top, 1 bottom, 2 left, 3 right;  code = from | to << 2 (first segment) | second << 4 | n << 8
__device__ __forceinline__ int ms_case_code(int c) {
  // (from, to) per case, fully_connected = 'low'
  switch (c) {
    case 1: return (0 | 2 << 2) | (1 << 8);
    case 2: return (3 | 0 << 2) | (1 << 8);
    case 3: return (3 | 2 << 2) | (1 << 8);
    case 4: return (2 | 1 << 2) | (1 << 8);
    case 5: return (0 | 1 << 2) | (1 << 8);
    case 6: return (3 | 0 << 2) | ((2 | 1 << 2) << 4) | (2 << 8);
    case 7: return (3 | 1 << 2) | (1 << 8);
    case 8: return (1 | 3 << 2) | (1 << 8);
    case 9: return (0 | 2 << 2) | ((1 | 3 << 2) << 4) | (2 << 8);
    case 10: return (1 | 0 << 2) | (1 << 8);
    case 11: return (1 | 2 << 2) | (1 << 8);
    case 12: return (2 | 3 << 2) | (1 << 8);
    case 13: return (0 | 3 << 2) | (1 << 8);
    case 14: return (2 | 0 << 2) | (1 << 8);
    default: return 0;
  }
}

// Warp-collective emission: every lane holds at most one square (r0, c0) of the BASE grid with its four corner
// values; the segments are built exactly as skimage does (float coordinates, np.round), the square of the
// periodic extension (c0 + nlon) is emitted alongside, slots come from one warp-aggregated atomic.
__device__ __forceinline__ void ms_emit_squares(const WbkDev& d, int job, bool active, int r0, int c0, double ul,
                                                double ur, double ll, double lr, double level) {
  const int lane = wbk_lane();
  const int W = d.W, nlon = d.nlon;
  int nemit = 0, code = 0, ncopy = 1;
  if (active) {
    const int sq = (ul > level ? 1 : 0) | (ur > level ? 2 : 0) | (ll > level ? 4 : 0) | (lr > level ? 8 : 0);
    code = ms_case_code(sq);
    ncopy = (c0 + nlon <= W - 2) ? 2 : 1;
    nemit = (code >> 8) * ncopy;
  }
  const int incl = wbk_warp_incl_scan(nemit);
  const int tot = __shfl_sync(WBK_FULL, incl, 31);
  int base = 0;
  if (lane == 31) base = atomicAdd(&d.seg_count[job], tot);
  base = __shfl_sync(WBK_FULL, base, 31);
  if (nemit == 0) return;
  int slot = base + incl - nemit;
  const int nseg = code >> 8;
  bool lattice = false;
  // only the edges this square uses are interpolated (identical expression from both adjacent squares)
  double frac[4];
  bool need[4] = {false, false, false, false};
  for (int s2 = 0; s2 < nseg; ++s2) {
    need[(code >> (4 * s2)) & 3] = true;
    need[(code >> (4 * s2 + 2)) & 3] = true;
  }
  frac[0] = need[0] ? ms_fraction(ul, ur, level) : 0.0;
  frac[1] = need[1] ? ms_fraction(ll, lr, level) : 0.0;
  frac[2] = need[2] ? ms_fraction(ul, ll, level) : 0.0;
  frac[3] = need[3] ? ms_fraction(ur, lr, level) : 0.0;
  for (int copy = 0; copy < ncopy; ++copy) {
    const int cc = c0 + copy * nlon;  // column of the square on the extended grid
    // float coordinates exactly as skimage builds them, then np.round (half to even)
    const double xt = __dadd_rn((double)cc, frac[0]), xb = __dadd_rn((double)cc, frac[1]);
    const double yl = __dadd_rn((double)r0, frac[2]), yr = __dadd_rn((double)r0, frac[3]);
    u32 pid[4], pxy[4];
    pid[0] = 2u * (u32)(r0 * W + cc);             // top: horizontal edge (r0, cc)
    pid[1] = 2u * (u32)((r0 + 1) * W + cc);       // bottom: horizontal edge (r0+1, cc)
    pid[2] = 2u * (u32)(r0 * W + cc) + 1u;        // left: vertical edge (r0, cc)
    pid[3] = 2u * (u32)(r0 * W + cc + 1) + 1u;    // right: vertical edge (r0, cc+1)
    pxy[0] = wbk_pack_xy((int)rint(xt), r0);
    pxy[1] = wbk_pack_xy((int)rint(xb), r0 + 1);
    pxy[2] = wbk_pack_xy(cc, (int)rint(yl));
    pxy[3] = wbk_pack_xy(cc + 1, (int)rint(yr));
    const bool von[4] = {xt == rint(xt), xb == rint(xb), yl == rint(yl), yr == rint(yr)};
    // a point that falls exactly on a grid vertex is identified by that vertex: skimage joins by float
    // equality, so all edges meeting there share the point (handled by the sequential linker)
    if (von[0]) pid[0] = WBK_VERTEX_ID | (u32)(r0 * W + (int)xt);
    if (von[1]) pid[1] = WBK_VERTEX_ID | (u32)((r0 + 1) * W + (int)xb);
    if (von[2]) pid[2] = WBK_VERTEX_ID | (u32)((int)yl * W + cc);
    if (von[3]) pid[3] = WBK_VERTEX_ID | (u32)((int)yr * W + cc + 1);
    for (int s2 = 0; s2 < nseg; ++s2) {
      const int fe = (code >> (4 * s2)) & 3, te = (code >> (4 * s2 + 2)) & 3;
      lattice = lattice || von[fe] || von[te];
      if (slot < d.S) {
        const size_t o = (size_t)job * d.S + slot;
        d.rid[o] = 2u * (u32)(r0 * (W - 1) + cc) + (u32)s2;
        d.fpid[o] = pid[fe];
        d.tpid[o] = pid[te];
        d.fxy[o] = pxy[fe];
        d.txy[o] = pxy[te];
      }
      ++slot;
    }
  }
  if (lattice) atomicOr(&d.status[job], (int)WBK_ST_LATTICE_VERTEX);
}

// hit h of a warp's compacted hit list: (mask index, lane) of the h-th set bit over masks[0..nmasks)
__device__ __forceinline__ bool ms_locate_hit(const u32* masks, int nmasks, int h, int& mask_idx, int& src_lane) {
  int cum = 0, i = 0;
  for (; i < nmasks; ++i) {
    const int c = __popc(masks[i]);
    if (h < cum + c) break;
    cum += c;
  }
  if (i >= nmasks) return false;
  u32 m = masks[i];
  for (int k = h - cum; k > 0; --k) m &= m - 1;
  mask_idx = i;
  src_lane = __ffs((int)m) - 1;
  return true;
}
