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
