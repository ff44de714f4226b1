// Event tracking: exact polygon-overlap areas for the by_overlap method.
// Reference: wavebreaking/processing/events.py:205-214 (geopandas intersection(...).area of event pairs).
//
// area(A n B) is evaluated without clipping A against B: every polygon is the signed sum of the triangles
// (O, v_i, v_i+1) spanned by its edges and a common origin O, so
//     area(A n B) = sum_a sum_b sign_a sign_b area(T_a n T_b)
// and each triangle-triangle intersection is a tiny convex clip.  One CTA per event pair, threads over the
// (edge of A) x (edge of B) products, fp64, deterministic block reduction.
#include "wbk_common.cuh"

#define TRK_THREADS 256

struct TrkPoly {
  const double* xy;      // all vertices (x, y pairs)
  const int* ring_off;   // ring r = vertices [ring_off[r], ring_off[r+1])
  const int* poly_off;   // polygon p = rings [poly_off[p], poly_off[p+1])
};

// signed doubled area and CCW-ordered copy of triangle (o, a, b)
__device__ __forceinline__ double tri_ccw(double ox, double oy, double ax, double ay, double bx, double by, double* t) {
  const double c = (ax - ox) * (by - oy) - (ay - oy) * (bx - ox);
  t[0] = ox; t[1] = oy;
  if (c >= 0) { t[2] = ax; t[3] = ay; t[4] = bx; t[5] = by; }
  else        { t[2] = bx; t[3] = by; t[4] = ax; t[5] = ay; }
  return c;
}

// area of (convex CCW triangle p) n (convex CCW triangle q): Sutherland-Hodgman, at most 7 vertices
__device__ inline double tri_tri_area(const double* p, const double* q) {
  double poly[16], tmp[16];
  int n = 3;
  for (int i = 0; i < 6; ++i) poly[i] = q[i];
  for (int e = 0; e < 3 && n > 0; ++e) {
    const double ex0 = p[2 * e], ey0 = p[2 * e + 1];
    const double ex1 = p[2 * ((e + 1) % 3)], ey1 = p[2 * ((e + 1) % 3) + 1];
    const double dx = ex1 - ex0, dy = ey1 - ey0;
    int m = 0;
    for (int i = 0; i < n; ++i) {
      const double cx = poly[2 * i], cy = poly[2 * i + 1];
      const int j = i + 1 == n ? 0 : i + 1;
      const double nx = poly[2 * j], ny = poly[2 * j + 1];
      const double sc = dx * (cy - ey0) - dy * (cx - ex0);  // >= 0: inside (left of the edge)
      const double sn = dx * (ny - ey0) - dy * (nx - ex0);
      if (sc >= 0) {
        tmp[2 * m] = cx; tmp[2 * m + 1] = cy; ++m;
      }
      if ((sc > 0 && sn < 0) || (sc < 0 && sn > 0)) {
        const double t = sc / (sc - sn);
        tmp[2 * m] = cx + t * (nx - cx); tmp[2 * m + 1] = cy + t * (ny - cy); ++m;
      }
    }
    n = m;
    for (int i = 0; i < 2 * n; ++i) poly[i] = tmp[i];
  }
  if (n < 3) return 0.0;
  double a2 = 0.0;
  for (int i = 0; i < n; ++i) {
    const int j = i + 1 == n ? 0 : i + 1;
    a2 += poly[2 * i] * poly[2 * j + 1] - poly[2 * j] * poly[2 * i + 1];
  }
  return a2 > 0 ? 0.5 * a2 : 0.0;
}

// edge k (0-based over all rings of polygon p) -> its two end points
__device__ __forceinline__ void trk_edge(const TrkPoly& P, int p, int k, double& ax, double& ay, double& bx, double& by) {
  int r = P.poly_off[p];
  while (k >= P.ring_off[r + 1] - P.ring_off[r]) {
    k -= P.ring_off[r + 1] - P.ring_off[r];
    ++r;
  }
  const int s = P.ring_off[r], n = P.ring_off[r + 1] - s;
  const int k2 = k + 1 == n ? 0 : k + 1;
  ax = P.xy[2 * (s + k)]; ay = P.xy[2 * (s + k) + 1];
  bx = P.xy[2 * (s + k2)]; by = P.xy[2 * (s + k2) + 1];
}

__device__ __forceinline__ int trk_nedges(const TrkPoly& P, int p) {
  return P.ring_off[P.poly_off[p + 1]] - P.ring_off[P.poly_off[p]];
}

// out[3 * pair + {0,1,2}] = area(A), area(B), area(A n B)
__global__ void __launch_bounds__(TRK_THREADS) track_overlap_kernel(TrkPoly P, const int* __restrict__ pairs, int npairs,
                                                                    double* __restrict__ out) {
  __shared__ double red[40];
  for (int w = blockIdx.x; w < npairs; w += gridDim.x) {
    const int pa = pairs[2 * w], pb = pairs[2 * w + 1];
    const int na = trk_nedges(P, pa), nb = trk_nedges(P, pb);
    // common origin: first vertex of A (keeps the triangle areas small)
    const int sa = P.ring_off[P.poly_off[pa]];
    const double ox = na > 0 ? P.xy[2 * sa] : 0.0, oy = na > 0 ? P.xy[2 * sa + 1] : 0.0;
    double area_a = 0, area_b = 0, inter = 0;
    for (int k = threadIdx.x; k < na; k += TRK_THREADS) {
      double ax, ay, bx, by;
      trk_edge(P, pa, k, ax, ay, bx, by);
      area_a += (ax - ox) * (by - oy) - (ay - oy) * (bx - ox);
    }
    for (int k = threadIdx.x; k < nb; k += TRK_THREADS) {
      double ax, ay, bx, by;
      trk_edge(P, pb, k, ax, ay, bx, by);
      area_b += (ax - ox) * (by - oy) - (ay - oy) * (bx - ox);
    }
    const long long nprod = (long long)na * nb;
    for (long long q = threadIdx.x; q < nprod; q += TRK_THREADS) {
      const int ka = (int)(q / nb), kb = (int)(q - (long long)ka * nb);
      double a0x, a0y, a1x, a1y, b0x, b0y, b1x, b1y, ta[6], tb[6];
      trk_edge(P, pa, ka, a0x, a0y, a1x, a1y);
      trk_edge(P, pb, kb, b0x, b0y, b1x, b1y);
      const double ca = tri_ccw(ox, oy, a0x, a0y, a1x, a1y, ta);
      const double cb = tri_ccw(ox, oy, b0x, b0y, b1x, b1y, tb);
      if (ca == 0.0 || cb == 0.0) continue;
      // quick reject on bounding boxes
      const double axmin = fmin(ox, fmin(a0x, a1x)), axmax = fmax(ox, fmax(a0x, a1x));
      const double bxmin = fmin(ox, fmin(b0x, b1x)), bxmax = fmax(ox, fmax(b0x, b1x));
      const double aymin = fmin(oy, fmin(a0y, a1y)), aymax = fmax(oy, fmax(a0y, a1y));
      const double bymin = fmin(oy, fmin(b0y, b1y)), bymax = fmax(oy, fmax(b0y, b1y));
      if (axmax <= bxmin || bxmax <= axmin || aymax <= bymin || bymax <= aymin) continue;
      const double s = ((ca > 0) == (cb > 0)) ? 1.0 : -1.0;
      inter += s * tri_tri_area(ta, tb);
    }
    area_a = wbk_block_sum_f64(area_a, red);
    area_b = wbk_block_sum_f64(area_b, red);
    inter = wbk_block_sum_f64(inter, red);
    if (threadIdx.x == 0) {
      // orientation of the rings does not matter: |area|, and the product of winding signs for the overlap
      const double sgn = ((area_a > 0) == (area_b > 0)) ? 1.0 : -1.0;
      out[3 * w + 0] = fabs(0.5 * area_a);
      out[3 * w + 1] = fabs(0.5 * area_b);
      out[3 * w + 2] = sgn * inter;
    }
    __syncthreads();
  }
}

extern "C" int wbk_track_overlap(const double* d_xy, const int* d_ring_off, const int* d_poly_off, const int* d_pairs,
                                 int npairs, double* d_out, void* stream) {
  if (npairs < 0 || (npairs > 0 && (!d_xy || !d_ring_off || !d_poly_off || !d_pairs || !d_out))) {
    wbk_set_error("wbk_track_overlap: invalid argument");
    return WBK_ERR_INVALID;
  }
  if (npairs == 0) return WBK_OK;
  TrkPoly P{d_xy, d_ring_off, d_poly_off};
  const int grid = npairs < 148 * 8 ? npairs : 148 * 8;
  WBK_LAUNCH(KID_TRACK_OVERLAP, track_overlap_kernel, dim3(grid), dim3(TRK_THREADS), 0, (cudaStream_t)stream, P, d_pairs,
             npairs, d_out);
  WBK_LAUNCH_CHECK();
  return WBK_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Candidate pairs (events.py:160-181).  The events are sorted by date; for event i the host supplies the index window
// [lo[i], hi[i]) of the events j with 0 < date_j - date_i <= time_range (two vectorised searches).  One warp per
// event expands its window, drops the pairs whose bounding boxes are disjoint (their intersection is empty, so
// `inter.area / union.area` is 0) and appends the rest with one warp-aggregated atomic per 32 candidates.
// d_bbox may be NULL (by_distance: every pair of the window is a candidate).
__global__ void __launch_bounds__(256) track_candidates_kernel(const int* __restrict__ lo, const int* __restrict__ hi,
                                                               const int* __restrict__ bbox, int n, int* __restrict__ pairs,
                                                               int cap, int* __restrict__ count) {
  const int lane = wbk_lane();
  const int wpb = blockDim.x >> 5;
  for (int i = blockIdx.x * wpb + wbk_warp(); i < n; i += gridDim.x * wpb) {
    const int a = lo[i], b = hi[i];
    int bx0 = 0, by0 = 0, bx1 = 0, by1 = 0;
    if (bbox) {
      bx0 = bbox[4 * i]; by0 = bbox[4 * i + 1]; bx1 = bbox[4 * i + 2]; by1 = bbox[4 * i + 3];
    }
    for (int j0 = a; j0 < b; j0 += 32) {
      const int j = j0 + lane;
      bool keep = j < b;
      if (keep && bbox) {
        const int cx0 = bbox[4 * j], cy0 = bbox[4 * j + 1], cx1 = bbox[4 * j + 2], cy1 = bbox[4 * j + 3];
        keep = !(bx1 < cx0 || cx1 < bx0 || by1 < cy0 || cy1 < by0) && bx0 <= bx1 && cx0 <= cx1;  // empty boxes never match
      }
      const u32 m = __ballot_sync(WBK_FULL, keep);
      if (m == 0) continue;
      int base = 0;
      if (lane == 0) base = atomicAdd(count, __popc(m));
      base = __shfl_sync(WBK_FULL, base, 0);
      if (keep) {
        const int slot = base + __popc(m & ((1u << lane) - 1u));
        if (slot < cap) {
          pairs[2 * slot] = i;
          pairs[2 * slot + 1] = j;
        }
      }
    }
  }
}

extern "C" int wbk_track_candidates(const int* d_lo, const int* d_hi, const int* d_bbox, int n, int* d_pairs, int cap,
                                    int* d_count, void* stream) {
  if (n < 0 || cap < 0 || !d_count || (n > 0 && (!d_lo || !d_hi)) || (cap > 0 && !d_pairs)) {
    wbk_set_error("wbk_track_candidates: invalid argument");
    return WBK_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  WBK_CUDA_CHECK(cudaMemsetAsync(d_count, 0, sizeof(int), st));
  if (n == 0) return WBK_OK;
  const int blocks = (n + 7) / 8 < 148 * 8 ? (n + 7) / 8 : 148 * 8;
  WBK_LAUNCH(KID_TRACK_PAIRS, track_candidates_kernel, dim3(blocks), dim3(256), 0, st, d_lo, d_hi, d_bbox, n, d_pairs, cap,
             d_count);
  WBK_LAUNCH_CHECK();
  return WBK_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Exact "do the two lattice polygons overlap in a region of positive area?" (events.py:205-214 with overlap = 0:
// `inter.area / union.area > 0`; GEOS returns an empty / lower-dimensional intersection, area 0.0, for polygons
// that merely touch).  All vertices are integers below 2^12, so every predicate is exact in int64:
//   1. a proper (transversal, interior-interior) crossing of an edge of A with an edge of B  => overlap;
//   2. no edge of A meets an edge of B at all  => the boundaries are disjoint: overlap iff the first vertex of a ring
//      of one polygon lies inside the other (non-zero winding);
//   3. the boundaries touch but never cross properly: the touch points cut each boundary into arcs that lie
//      entirely inside, outside or on the other boundary.  Every edge that takes part in a touch is cut at the other
//      polygon's vertices on it; the midpoint of each piece is classified exactly (coordinates scaled by twice the
//      squared edge length): strictly inside => overlap; on the other boundary (collinear pieces) => overlap iff
//      both interiors lie on the same side of the shared piece (ring orientation signs).  Rings without a touched
//      edge are covered by rule 2.
// Polygon p = vertices [poly_voff[p], poly_voff[p+1]) of the edge table (x0, y0, x1, y1 per vertex: the edge from
// that vertex to the next one of its ring); vsign = orientation of the vertex's ring (+1 counter-clockwise).
// result: 0 / 1, | 2 when rule 3 decided.
#define TRX_THREADS 128
#define TRX_MAXTOUCH 96

__device__ __forceinline__ long long trx_cross(long long ax, long long ay, long long bx, long long by) { return ax * by - ay * bx; }

// winding number contribution of edge (x0,y0)->(x1,y1) for the point (px,py), all in the same (scaled) units;
// *on is set when the point lies on the closed edge
__device__ __forceinline__ int trx_wind(long long x0, long long y0, long long x1, long long y1, long long px, long long py, bool* on) {
  const long long cr = trx_cross(x1 - x0, y1 - y0, px - x0, py - y0);
  if (cr == 0 && px >= min(x0, x1) && px <= max(x0, x1) && py >= min(y0, y1) && py <= max(y0, y1)) *on = true;
  if (y0 <= py && py < y1) return cr > 0 ? 1 : 0;
  if (y1 <= py && py < y0) return cr < 0 ? -1 : 0;
  return 0;
}

__global__ void __launch_bounds__(TRX_THREADS) track_exact_kernel(const int4* __restrict__ edges, const int* __restrict__ vsign,
                                                                  const int* __restrict__ ring_off, const int* __restrict__ poly_off,
                                                                  const int* __restrict__ pairs, int npairs, int* __restrict__ result) {
  __shared__ int s_flag[4];                 // [0] proper crossing, [1] any touch, [2] touch list overflow
  __shared__ int s_ntouch[2];
  __shared__ int s_touch[2][TRX_MAXTOUCH];  // touched edges of A / of B (vertex indices; duplicates allowed)
  __shared__ int s_w, s_on;
  __shared__ long long s_next;
  const int tid = threadIdx.x;
  for (int w = blockIdx.x; w < npairs; w += gridDim.x) {
    const int pa = pairs[2 * w], pb = pairs[2 * w + 1];
    const int a0 = ring_off[poly_off[pa]], a1 = ring_off[poly_off[pa + 1]];
    const int b0 = ring_off[poly_off[pb]], b1 = ring_off[poly_off[pb + 1]];
    const int nA = a1 - a0, nB = b1 - b0;
    if (tid < 4) s_flag[tid] = 0;
    if (tid < 2) s_ntouch[tid] = 0;
    __syncthreads();
    // ---- rule 1: all edge pairs (thread: one edge of A against every edge of B, uniform broadcast reads)
    for (int at = 0; at < nA; at += TRX_THREADS) {
      const int ia = at + tid;
      if (ia < nA) {
        const int4 e = edges[a0 + ia];
        const int exmin = min(e.x, e.z), exmax = max(e.x, e.z), eymin = min(e.y, e.w), eymax = max(e.y, e.w);
        bool proper = false, touched = false;
        for (int ib = 0; ib < nB; ++ib) {
          const int4 f = edges[b0 + ib];
          if (exmax < min(f.x, f.z) || max(f.x, f.z) < exmin || eymax < min(f.y, f.w) || max(f.y, f.w) < eymin) continue;
          const long long d1 = trx_cross(f.z - f.x, f.w - f.y, e.x - f.x, e.y - f.y);
          const long long d2 = trx_cross(f.z - f.x, f.w - f.y, e.z - f.x, e.w - f.y);
          const long long d3 = trx_cross(e.z - e.x, e.w - e.y, f.x - e.x, f.y - e.y);
          const long long d4 = trx_cross(e.z - e.x, e.w - e.y, f.z - e.x, f.w - e.y);
          if (((d1 > 0 && d2 < 0) || (d1 < 0 && d2 > 0)) && ((d3 > 0 && d4 < 0) || (d3 < 0 && d4 > 0))) {
            proper = true;
            break;
          }
          const bool meet = !((d1 > 0 && d2 > 0) || (d1 < 0 && d2 < 0) || (d3 > 0 && d4 > 0) || (d3 < 0 && d4 < 0));
          if (meet) {  // closed segments share a point (collinear ones: their boxes intersect, checked above)
            touched = true;
            const int k = atomicAdd(&s_ntouch[1], 1);
            if (k < TRX_MAXTOUCH) s_touch[1][k] = b0 + ib;
            else s_flag[2] = 1;
          }
        }
        if (proper) s_flag[0] = 1;
        if (touched) {
          s_flag[1] = 1;
          const int k = atomicAdd(&s_ntouch[0], 1);
          if (k < TRX_MAXTOUCH) s_touch[0][k] = a0 + ia;
          else s_flag[2] = 1;
        }
      }
      __syncthreads();
      if (s_flag[0]) break;  // block-uniform
    }
    int res = s_flag[0] ? 1 : 0;
    const bool touch = s_flag[1] != 0, overflow = s_flag[2] != 0;
    __syncthreads();
    if (!res && nA > 0 && nB > 0) {
      // ---- rule 2: first vertex of every ring against the other polygon (skipped when it lies on the boundary)
      for (int side = 0; side < 2 && !res; ++side) {
        const int pr0 = side == 0 ? poly_off[pa] : poly_off[pb], pr1 = side == 0 ? poly_off[pa + 1] : poly_off[pb + 1];
        const int o0 = side == 0 ? b0 : a0, o1 = side == 0 ? b1 : a1;
        for (int r = pr0; r < pr1 && !res; ++r) {
          if (ring_off[r + 1] == ring_off[r]) continue;
          const int4 v = edges[ring_off[r]];
          if (tid == 0) { s_w = 0; s_on = 0; }
          __syncthreads();
          int wsum = 0;
          bool on = false;
          for (int i = o0 + tid; i < o1; i += TRX_THREADS) {
            const int4 f = edges[i];
            wsum += trx_wind(f.x, f.y, f.z, f.w, v.x, v.y, &on);
          }
          if (wsum) atomicAdd(&s_w, wsum);
          if (on) s_on = 1;
          __syncthreads();
          if (!s_on && s_w != 0) res = 1;  // block-uniform
          __syncthreads();
        }
      }
    }
    if (!res && touch) {
      // ---- rule 3: the pieces of the touched edges (all edges when the touch lists overflowed)
      res |= 2;
      for (int side = 0; side < 2 && !(res & 1); ++side) {
        const int s0 = side == 0 ? a0 : b0, s1 = side == 0 ? a1 : b1;  // subject polygon
        const int o0 = side == 0 ? b0 : a0, o1 = side == 0 ? b1 : a1;  // the other one
        const int nlist = overflow ? s1 - s0 : min(s_ntouch[side], TRX_MAXTOUCH);
        for (int li = 0; li < nlist && !(res & 1); ++li) {
          const int ei = overflow ? s0 + li : s_touch[side][li];
          const int4 e = edges[ei];
          const long long dx = e.z - e.x, dy = e.w - e.y;
          const long long L2 = dx * dx + dy * dy, S = 2 * L2;
          if (L2 == 0) continue;
          long long s = 0;
          while (s < L2 && !(res & 1)) {  // block-uniform loop
            if (tid == 0) { s_next = L2; s_w = 0; s_on = -1; }
            __syncthreads();
            // next cut: the smallest parameter > s of a vertex of the other polygon on the open edge
            long long best = L2;
            for (int i = o0 + tid; i < o1; i += TRX_THREADS) {
              const int4 f = edges[i];
              if (trx_cross(dx, dy, f.x - e.x, f.y - e.y) == 0) {
                const long long tv = (f.x - e.x) * dx + (f.y - e.y) * dy;
                if (tv > s && tv < best) best = tv;
              }
            }
            if (best < L2) atomicMin((unsigned long long*)&s_next, (unsigned long long)best);
            __syncthreads();
            const long long nxt = s_next;
            // midpoint of the piece (s, nxt), scaled by S = 2 L2
            const long long mx = (long long)e.x * S + dx * (s + nxt), my = (long long)e.y * S + dy * (s + nxt);
            int wsum = 0;
            for (int i = o0 + tid; i < o1; i += TRX_THREADS) {
              const int4 f = edges[i];
              bool on = false;
              wsum += trx_wind((long long)f.x * S, (long long)f.y * S, (long long)f.z * S, (long long)f.w * S, mx, my, &on);
              if (on) atomicMax(&s_on, i);
            }
            if (wsum) atomicAdd(&s_w, wsum);
            __syncthreads();
            if (s_on >= 0) {
              if (side == 0) {  // shared boundary piece: same side?
                const int4 f = edges[s_on];
                const long long dot = dx * (f.z - f.x) + dy * (f.w - f.y);
                const int sa = vsign[ei], sb = vsign[s_on] * (dot > 0 ? 1 : (dot < 0 ? -1 : 0));
                if (sa != 0 && sa == sb) res |= 1;
              }
            } else if (s_w != 0) {
              res |= 1;
            }
            __syncthreads();
            s = nxt;
          }
        }
      }
    }
    if (tid == 0) result[w] = res;
    __syncthreads();
  }
}

extern "C" int wbk_track_overlap_exact(const int* d_edges, const int* d_vsign, const int* d_ring_off, const int* d_poly_off,
                                       const int* d_pairs, int npairs, int* d_result, void* stream) {
  if (npairs < 0 || (npairs > 0 && (!d_edges || !d_vsign || !d_ring_off || !d_poly_off || !d_pairs || !d_result))) {
    wbk_set_error("wbk_track_overlap_exact: invalid argument");
    return WBK_ERR_INVALID;
  }
  if (npairs == 0) return WBK_OK;
  const int grid = npairs < 148 * 16 ? npairs : 148 * 16;
  WBK_LAUNCH(KID_TRACK_EXACT, track_exact_kernel, dim3(grid), dim3(TRX_THREADS), 0, (cudaStream_t)stream,
             (const int4*)d_edges, d_vsign, d_ring_off, d_poly_off, d_pairs, npairs, d_result);
  WBK_LAUNCH_CHECK();
  return WBK_OK;
}

// ------------------------------------------------------------------------------------------------------------
// by_distance (events.py:187-201): sklearn's haversine of the two centres of mass of every pair, evaluated in the
// operand order of sklearn/metrics/_dist_metrics.pyx.tp:2641-2656 (x1 = first event, x2 = second).  CUDA's sin /
// asin differ from glibc's by <= 2 ulp, so the caller re-evaluates with libm the pairs whose distance lies within
// 1e-9 (relative) of the threshold.  d_rad: [n][2] radians exactly as the reference feeds them.
__global__ void track_distance_kernel(const double* __restrict__ rad, const int* __restrict__ pairs, int npairs,
                                      double* __restrict__ out) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < npairs; k += gridDim.x * blockDim.x) {
    const int a = pairs[2 * k], b = pairs[2 * k + 1];
    const double a0 = rad[2 * a], a1 = rad[2 * a + 1], b0 = rad[2 * b], b1 = rad[2 * b + 1];
    const double s0 = sin(__dmul_rn(0.5, __dsub_rn(a0, b0))), s1 = sin(__dmul_rn(0.5, __dsub_rn(a1, b1)));
    const double h = __dadd_rn(__dmul_rn(s0, s0), __dmul_rn(__dmul_rn(__dmul_rn(cos(a0), cos(b0)), s1), s1));
    out[k] = __dmul_rn(2.0, asin(sqrt(h)));
  }
}

extern "C" int wbk_track_distance(const double* d_rad, const int* d_pairs, int npairs, double* d_out, void* stream) {
  if (npairs < 0 || (npairs > 0 && (!d_rad || !d_pairs || !d_out))) {
    wbk_set_error("wbk_track_distance: invalid argument");
    return WBK_ERR_INVALID;
  }
  if (npairs == 0) return WBK_OK;
  const int grid = (npairs + 255) / 256 < 148 * 8 ? (npairs + 255) / 256 : 148 * 8;
  WBK_LAUNCH(KID_TRACK_DIST, track_distance_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, d_rad, d_pairs, npairs,
             d_out);
  WBK_LAUNCH_CHECK();
  return WBK_OK;
}
