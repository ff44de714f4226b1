// Polygon rasterisation on the lattice: event properties (area-weighted sums) and to_xarray flags.
// Reference: wavebreaking/utils/index_utils.py:35-126 (calculate_properties) and
// wavebreaking/processing/events.py:66-106 (to_xarray); buffer(r)+sjoin(contains) restated as
// "inside (non-zero winding) or on the boundary or closer than r to an edge" (SURVEY.md A.5).
#include "wbk_ctx.cuh"

struct RingView {
  const u32* packed;  // packed lattice points (x | y << 16), or
  const int* xy;      // int32 (x, y) pairs, or
  int bx0, by0, bx1, by1;  // a box (n == 4, both pointers NULL): shapely.box(minx, miny, maxx, maxy)
  int n;
  __device__ __forceinline__ void get(int k, int& x, int& y) const {
    if (packed) {
      const u32 p = packed[k];
      x = wbk_px(p);
      y = wbk_py(p);
    } else if (xy) {
      x = xy[2 * k];
      y = xy[2 * k + 1];
    } else {
      x = (k == 0 || k == 1) ? bx1 : bx0;
      y = (k == 1 || k == 2) ? by1 : by0;
    }
  }
};

__device__ __forceinline__ long long ceil_div_ll(long long n, long long d) {  // d != 0
  if (d < 0) {
    n = -n;
    d = -d;
  }
  return n >= 0 ? (n + d - 1) / d : -((-n) / d);
}

// double-double accumulation (error-free TwoSum): the area-weighted sums come out correctly rounded in
// practice, like the compensated (Kahan) group sums of pandas that the reference relies on
// (index_utils.py:94-102) -- the truncated centre-of-mass quotient is sensitive to the last bit.
struct DD {
  double hi, lo;
};
__device__ __forceinline__ void dd_add(DD& a, double v) {
  const double s = __dadd_rn(a.hi, v);
  const double bb = __dsub_rn(s, a.hi);
  const double err = __dadd_rn(__dsub_rn(a.hi, __dsub_rn(s, bb)), __dsub_rn(v, bb));
  a.hi = s;
  a.lo = __dadd_rn(a.lo, err);
}
__device__ __forceinline__ void dd_merge(DD& a, double hi, double lo) {
  dd_add(a, hi);
  a.lo = __dadd_rn(a.lo, lo);
}
// deterministic block-wide double-double sum; result (rounded to double) to every thread
__device__ inline double dd_block_sum(DD v, double* scratch /* >= 66 doubles */) {
  const int lane = wbk_lane(), warp = wbk_warp(), nwarps = wbk_nthreads() >> 5;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const double oh = __shfl_xor_sync(WBK_FULL, v.hi, d), ol = __shfl_xor_sync(WBK_FULL, v.lo, d);
    dd_merge(v, oh, ol);
  }
  __syncthreads();
  if (lane == 0) {
    scratch[2 * warp] = v.hi;
    scratch[2 * warp + 1] = v.lo;
  }
  __syncthreads();
  if (warp == 0) {
    DD w;
    w.hi = lane < nwarps ? scratch[2 * lane] : 0.0;
    w.lo = lane < nwarps ? scratch[2 * lane + 1] : 0.0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const double oh = __shfl_xor_sync(WBK_FULL, w.hi, d), ol = __shfl_xor_sync(WBK_FULL, w.lo, d);
      dd_merge(w, oh, ol);
    }
    if (lane == 0) scratch[64] = __dadd_rn(w.hi, w.lo);
  }
  __syncthreads();
  return scratch[64];
}

// Warp-cooperative scan of lattice row y over columns [bx0, bx0 + bw): on return acc[i] is the winding number
// of (bx0 + i, y) (for points not on the boundary) and flg[i] has bit 0 / bit 1 set if the point is on the
// boundary or closer than sqrt(r2a) / sqrt(r2b) to an edge.  R = floor(max radius).
__device__ inline void raster_scan_row(const RingView& rv, int y, int bx0, int bw, int* acc, u32* flg, double r2a,
                                       double r2b, double rmax, int R) {
  const int lane = wbk_lane();
  for (int i = lane; i < bw; i += 32) {
    acc[i] = 0;
    flg[i] = 0;
  }
  __syncwarp();
  const int n = rv.n;
  for (int e0 = 0; e0 < n; e0 += 32) {
    const int e = e0 + lane;
    if (e < n) {
      int xa, ya, xb, yb;
      rv.get(e, xa, ya);
      rv.get(e + 1 == n ? 0 : e + 1, xb, yb);
      const int dx = xb - xa, dy = yb - ya;
      if (dx != 0 || dy != 0) {
        const int lo = min(ya, yb), hi = max(ya, yb);
        // winding contribution: lattice points strictly left of the crossing get +-1
        if (dy != 0 && lo <= y && y < hi) {
          const int s = dy > 0 ? 1 : -1;
          const long long cx = (long long)xa + ceil_div_ll((long long)dx * (y - ya), (long long)dy);
          const long long idx = cx - bx0;
          if (idx > 0) {
            atomicAdd(&acc[0], s);
            if (idx < bw) atomicAdd(&acc[(int)idx], -s);
          }
        }
        // boundary / near-edge candidates of this row
        if (y >= lo - R && y <= hi + R) {
          const long long len2 = (long long)dx * dx + (long long)dy * dy;
          int cl, ch;
          if (dy == 0) {
            cl = min(xa, xb) - R;
            ch = max(xa, xb) + R;
          } else {
            const double xc = (double)xa + (double)dx * (double)(y - ya) / (double)dy;
            const double hw = rmax * sqrt((double)len2) / fabs((double)dy) + rmax + 1e-6;
            cl = max((int)ceil(xc - hw), min(xa, xb) - R - 1);
            ch = min((int)floor(xc + hw), max(xa, xb) + R + 1);
          }
          cl = max(cl, bx0);
          ch = min(ch, bx0 + bw - 1);
          for (int px = cl; px <= ch; ++px) {
            const long long ex = px - xa, ey = y - ya;
            const long long dot = ex * dx + ey * dy;
            u32 bits = 0;
            if (dot <= 0) {
              const double d2 = (double)(ex * ex + ey * ey);
              bits = (d2 < r2a ? 1u : 0u) | (d2 < r2b ? 2u : 0u);
              if (ex == 0 && ey == 0) bits = 3u;
            } else if (dot >= len2) {
              const long long fx = px - xb, fy = y - yb;
              const double d2 = (double)(fx * fx + fy * fy);
              bits = (d2 < r2a ? 1u : 0u) | (d2 < r2b ? 2u : 0u);
              if (fx == 0 && fy == 0) bits = 3u;
            } else {
              const long long cross = (long long)dx * ey - (long long)dy * ex;
              if (cross == 0) {
                bits = 3u;
              } else {
                const double c2 = (double)cross * (double)cross, l2 = (double)len2;
                bits = (c2 < r2a * l2 ? 1u : 0u) | (c2 < r2b * l2 ? 2u : 0u);
              }
            }
            if (bits) atomicOr(&flg[px - bx0], bits);
          }
        }
      }
    }
  }
  __syncwarp();
  // inclusive prefix sum of the winding differences along the row
  int carry = 0;
  for (int i0 = 0; i0 < bw; i0 += 32) {
    const int i = i0 + lane;
    int v = i < bw ? acc[i] : 0;
    int incl = wbk_warp_incl_scan(v);
    if (i < bw) acc[i] = incl + carry;
    carry += __shfl_sync(WBK_FULL, incl, 31);
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------ event list
__global__ void __launch_bounds__(1024) event_list_kernel(WbkIdx x, int J, int njobs) {
  __shared__ int sscan[40];
  const int n = 3 * njobs;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int kind = i / njobs, job = i - kind * njobs;
    x.ev_off[i] = x.ev_count[kind * J + job];
  }
  __syncthreads();
  const int total = wbk_block_excl_scan(x.ev_off, n, sscan);
  if (threadIdx.x == 0) {
    x.ev_off[n] = total;
    x.total[1] = total;
  }
}

#define RS_THREADS 256

template <typename T>
__global__ void __launch_bounds__(RS_THREADS)
events_raster_kernel(WbkDev d, WbkIdx x, const int* __restrict__ job_off, const int* __restrict__ pt_off,
                     const u32* __restrict__ pts, const double* __restrict__ area, const T* __restrict__ data,
                     const T* __restrict__ intensity, int8_t* __restrict__ flags, int ntime, int nlevels, int J,
                     int njobs, double r2_prop, double r2_flag, int rowcap) {
  WBK_DYN_SMEM(int, sm);
  __shared__ double red[72];
  __shared__ int s_box[5];
  const int tid = threadIdx.x, lane = wbk_lane(), warp = wbk_warp(), nwarps = blockDim.x >> 5;
  int* acc = sm + (size_t)warp * 2 * rowcap;
  u32* flg = reinterpret_cast<u32*>(acc + rowcap);
  const int nlat = d.nlat, nlon = d.nlon, W = d.W;
  const double rmax = sqrt(r2_prop > r2_flag ? r2_prop : r2_flag);
  const int R = (int)floor(rmax);
  const int nlist = 3 * njobs;
  const int total = x.ev_off[nlist];
  (void)job_off;
  for (int w = blockIdx.x; w < total; w += gridDim.x) {
    int lo = 0, hi = nlist - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (x.ev_off[mid] <= w) lo = mid; else hi = mid - 1;
    }
    const int kind = lo / njobs, job = lo - kind * njobs, e = w - x.ev_off[lo];
    int* ev = x.ev_int + (((size_t)kind * J + job) * x.EC + e) * WBK_EV_INTS;
    double* evf = x.ev_f64 + (((size_t)kind * J + job) * x.EC + e) * WBK_EV_F64;
    const int t = job / nlevels;
    RingView rv;
    rv.xy = nullptr;
    if (kind == WBK_EV_OVERTURNING) {
      rv.packed = nullptr;
      rv.bx0 = ev[3]; rv.by0 = ev[4]; rv.bx1 = ev[5]; rv.by1 = ev[6];
      rv.n = 4;
    } else {
      rv.packed = pts + pt_off[ev[0]] + ev[1];
      rv.n = ev[2] - ev[1] + 1;
      rv.bx0 = rv.by0 = rv.bx1 = rv.by1 = 0;
    }
    // bounding box of the vertices
    if (tid == 0) {
      s_box[0] = 0x7fffffff; s_box[1] = 0x7fffffff; s_box[2] = -1; s_box[3] = -1;
    }
    __syncthreads();
    {
      int x0 = 0x7fffffff, y0 = 0x7fffffff, x1 = -1, y1 = -1;
      for (int k = tid; k < rv.n; k += blockDim.x) {
        int vx, vy;
        rv.get(k, vx, vy);
        x0 = min(x0, vx); x1 = max(x1, vx); y0 = min(y0, vy); y1 = max(y1, vy);
      }
      x0 = wbk_warp_min(x0); y0 = wbk_warp_min(y0); x1 = wbk_warp_max(x1); y1 = wbk_warp_max(y1);
      if (lane == 0) {
        atomicMin(&s_box[0], x0); atomicMin(&s_box[1], y0); atomicMax(&s_box[2], x1); atomicMax(&s_box[3], y1);
      }
    }
    __syncthreads();
    const int vx0 = s_box[0], vy0 = s_box[1], vx1 = s_box[2], vy1 = s_box[3];
    // 0: all vertices on the real grid, 1: straddles the last meridian (host clips), 2: entirely in the extension
    const int split = vx1 < nlon ? 0 : (vx0 >= nlon ? 2 : 1);
    const int shift = split == 2 ? nlon : 0;
    const int bx0 = max(vx0 - R - 1, 0), bx1 = min(vx1 + R + 1, W - 1);
    const int by0 = max(vy0 - R - 1, 0), by1 = min(vy1 + R + 1, nlat - 1);
    const int bw = bx1 - bx0 + 1;
    DD s_a = {0, 0}, s_v = {0, 0}, s_i = {0, 0}, s_x = {0, 0}, s_y = {0, 0}, s_n = {0, 0};
    for (int y = by0 + warp; y <= by1; y += nwarps) {
      raster_scan_row(rv, y, bx0, bw, acc, flg, r2_prop, r2_flag, rmax, R);
      const double a = area[y];
      const size_t rowbase = ((size_t)t * nlat + y) * nlon;
      for (int i = lane; i < bw; i += 32) {
        const bool in = acc[i] != 0;
        const u32 f = flg[i];
        const int px = bx0 + i;
        if (in || (f & 1u)) {
          const int xf = px % nlon;
          dd_add(s_a, a);
          dd_add(s_v, __dmul_rn(a, (double)data[rowbase + xf]));
          if (intensity) dd_add(s_i, __dmul_rn(a, (double)intensity[rowbase + xf]));
          dd_add(s_x, __dmul_rn((double)px, a));
          dd_add(s_y, __dmul_rn((double)y, a));
          dd_add(s_n, 1.0);
        }
        if (flags && split != 1 && (in || (f & 2u))) {
          const int xf = px - shift;
          if (xf >= 0 && xf < nlon) flags[(((size_t)kind * ntime + t) * nlat + y) * nlon + xf] = 1;
        }
      }
      __syncwarp();
    }
    const double t_a = dd_block_sum(s_a, red), t_v = dd_block_sum(s_v, red), t_i = dd_block_sum(s_i, red);
    const double t_x = dd_block_sum(s_x, red), t_y = dd_block_sum(s_y, red), t_n = dd_block_sum(s_n, red);
    if (tid == 0) {
      evf[0] = t_a; evf[1] = t_v; evf[2] = t_i; evf[3] = t_x; evf[4] = t_y; evf[5] = t_n;
      if (kind != WBK_EV_OVERTURNING) {
        ev[3] = vx0; ev[4] = vy0; ev[5] = vx1; ev[6] = vy1;
      }
      ev[8] = split;
    }
    __syncthreads();
  }
}

static int raster_rowcap(int W) { return (W + 2 + 31) & ~31; }

extern "C" int wbk_events_raster(wbk_ctx* ctx, const int* d_job_off, const int* d_pt_off, const uint32_t* d_pts,
                                 const double* d_coords, const void* d_data, int dtype, const void* d_intensity,
                                 int ntime, int8_t* d_flags, const wbk_index_params* prm, void* stream) {
  if (!ctx || !d_job_off || !d_pt_off || !d_coords || !d_data || !prm || ntime < 0) {
    wbk_set_error("wbk_events_raster: invalid argument");
    return WBK_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  WbkDev& d = ctx->d;
  const int njobs = ctx->njobs, J = ctx->caps.max_jobs;
  if (d_flags && ntime > 0)
    WBK_CUDA_CHECK(cudaMemsetAsync(d_flags, 0, (size_t)3 * ntime * d.nlat * d.nlon, st));
  if (njobs == 0) return WBK_OK;
  if (d_flags && prm->dlon != prm->dlat) {
    wbk_set_error("wbk_events_raster: to_xarray flags need dlon == dlat (buffer radius is isotropic in degrees)");
    return WBK_ERR_INVALID;
  }
  WBK_LAUNCH(KID_EVENT_LIST, event_list_kernel, dim3(1), dim3(1024), 0, st, ctx->x, J, njobs);
  WBK_LAUNCH_CHECK();
  const double r_prop = (prm->dlon + prm->dlat) / 2.0 / 2.0;  // index units (index_utils.py:47-50)
  const double r_flag = d_flags ? ((prm->dlon + prm->dlat) / 2.0 / 2.0) / prm->dlon : r_prop;  // degrees -> cells
  const int rowcap = raster_rowcap(d.W);
  int nwarps = RS_THREADS / 32;
  const size_t smem = (size_t)nwarps * 2 * rowcap * sizeof(int);
  if (smem > 200 * 1024) {
    wbk_set_error("wbk_events_raster: extended width %d too large for the row buffers", d.W);
    return WBK_ERR_CAPACITY;
  }
  const double* area = d_coords + 3 * (size_t)d.nlat;
  const int grid = 148 * 4;
  if (dtype == WBK_F32) {
    WBK_CUDA_CHECK(cudaFuncSetAttribute(events_raster_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WBK_LAUNCH(KID_EVENTS_RASTER, events_raster_kernel<float>, dim3(grid), dim3(RS_THREADS), smem, st, d, ctx->x, d_job_off, d_pt_off,
               (const u32*)d_pts, area, (const float*)d_data, (const float*)d_intensity, d_flags, ntime, ctx->nlevels, J,
               njobs, r_prop * r_prop, r_flag * r_flag, rowcap);
  } else if (dtype == WBK_F64) {
    WBK_CUDA_CHECK(cudaFuncSetAttribute(events_raster_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WBK_LAUNCH(KID_EVENTS_RASTER, events_raster_kernel<double>, dim3(grid), dim3(RS_THREADS), smem, st, d, ctx->x, d_job_off, d_pt_off,
               (const u32*)d_pts, area, (const double*)d_data, (const double*)d_intensity, d_flags, ntime, ctx->nlevels,
               J, njobs, r_prop * r_prop, r_flag * r_flag, rowcap);
  } else {
    wbk_set_error("wbk_events_raster: unsupported dtype");
    return WBK_ERR_INVALID;
  }
  WBK_LAUNCH_CHECK();
  return WBK_OK;
}

// ------------------------------------------------------------------------------------------ generic rings
__global__ void __launch_bounds__(RS_THREADS)
rings_raster_kernel(const int* __restrict__ xy, const int* __restrict__ ring_off, const int* __restrict__ ring_t,
                    int nrings, int nlat, int nlon, int ntime, double r2, int8_t* out_i8, int* owner, int rowcap) {
  WBK_DYN_SMEM(int, sm);
  __shared__ int s_box[4];
  const int tid = threadIdx.x, lane = wbk_lane(), warp = wbk_warp(), nwarps = blockDim.x >> 5;
  int* acc = sm + (size_t)warp * 2 * rowcap;
  u32* flg = reinterpret_cast<u32*>(acc + rowcap);
  const double rmax = sqrt(r2);
  const int R = (int)floor(rmax);
  for (int r = blockIdx.x; r < nrings; r += gridDim.x) {
    RingView rv;
    rv.packed = nullptr;
    rv.xy = xy + 2 * (size_t)ring_off[r];
    rv.n = ring_off[r + 1] - ring_off[r];
    rv.bx0 = rv.by0 = rv.bx1 = rv.by1 = 0;
    const int t = ring_t[r];
    if (tid == 0) {
      s_box[0] = 0x7fffffff; s_box[1] = 0x7fffffff; s_box[2] = -0x7fffffff; s_box[3] = -0x7fffffff;
    }
    __syncthreads();
    {
      int x0 = 0x7fffffff, y0 = 0x7fffffff, x1 = -0x7fffffff, y1 = -0x7fffffff;
      for (int k = tid; k < rv.n; k += blockDim.x) {
        int vx, vy;
        rv.get(k, vx, vy);
        x0 = min(x0, vx); x1 = max(x1, vx); y0 = min(y0, vy); y1 = max(y1, vy);
      }
      x0 = wbk_warp_min(x0); y0 = wbk_warp_min(y0); x1 = wbk_warp_max(x1); y1 = wbk_warp_max(y1);
      if (lane == 0) {
        atomicMin(&s_box[0], x0); atomicMin(&s_box[1], y0); atomicMax(&s_box[2], x1); atomicMax(&s_box[3], y1);
      }
    }
    __syncthreads();
    const int bx0 = max(s_box[0] - R - 1, 0), bx1 = min(s_box[2] + R + 1, nlon - 1);
    const int by0 = max(s_box[1] - R - 1, 0), by1 = min(s_box[3] + R + 1, nlat - 1);
    const int bw = bx1 - bx0 + 1;
    if (rv.n > 0 && bw > 0 && t >= 0 && t < ntime) {
      for (int y = by0 + warp; y <= by1; y += nwarps) {
        raster_scan_row(rv, y, bx0, bw, acc, flg, r2, r2, rmax, R);
        for (int i = lane; i < bw; i += 32) {
          if (acc[i] != 0 || (flg[i] & 1u)) {
            const size_t o = ((size_t)t * nlat + y) * nlon + (bx0 + i);
            if (out_i8) out_i8[o] = 1;
            if (owner) atomicMax(&owner[o], r);
          }
        }
        __syncwarp();
      }
    }
    __syncthreads();
  }
}

__global__ void owner_fill_kernel(int* owner, size_t n, int v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) owner[i] = v;
}
__global__ void owner_apply_kernel(const int* __restrict__ owner, const double* __restrict__ val, double* __restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int r = owner[i];
    if (r >= 0) out[i] = val[r];
  }
}

extern "C" int wbk_rasterize_rings(const int* d_xy, const int* d_ring_off, const int* d_ring_t,
                                   const double* d_ring_val, int nrings, int nlat, int nlon, int ntime, double r2,
                                   int8_t* d_out_i8, double* d_out_f64, int* d_owner, void* stream) {
  if (nrings < 0 || nlat < 1 || nlon < 1 || ntime < 0 || !(r2 >= 0) || (!d_out_i8 && !d_out_f64) ||
      (d_out_f64 && (!d_owner || !d_ring_val)) || (nrings > 0 && (!d_xy || !d_ring_off || !d_ring_t))) {
    wbk_set_error("wbk_rasterize_rings: invalid argument");
    return WBK_ERR_INVALID;
  }
  if (nrings == 0 || ntime == 0) return WBK_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int rowcap = raster_rowcap(nlon);
  const size_t smem = (size_t)(RS_THREADS / 32) * 2 * rowcap * sizeof(int);
  if (smem > 200 * 1024) {
    wbk_set_error("wbk_rasterize_rings: grid width %d too large for the row buffers", nlon);
    return WBK_ERR_CAPACITY;
  }
  const size_t ncell = (size_t)ntime * nlat * nlon;
  if (d_out_f64) {
    WBK_LAUNCH(KID_OWNER, owner_fill_kernel, dim3(592), dim3(256), 0, st, d_owner, ncell, -1);
    WBK_LAUNCH_CHECK();
  }
  WBK_CUDA_CHECK(cudaFuncSetAttribute(rings_raster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = nrings < 148 * 4 ? nrings : 148 * 4;
  WBK_LAUNCH(KID_RINGS_RASTER, rings_raster_kernel, dim3(grid), dim3(RS_THREADS), smem, st, d_xy, d_ring_off, d_ring_t, nrings, nlat, nlon,
             ntime, r2, d_out_i8, d_out_f64 ? d_owner : (int*)nullptr, rowcap);
  WBK_LAUNCH_CHECK();
  if (d_out_f64) {
    WBK_LAUNCH(KID_OWNER, owner_apply_kernel, dim3(592), dim3(256), 0, st, (const int*)d_owner, d_ring_val, d_out_f64, ncell);
    WBK_LAUNCH_CHECK();
  }
  return WBK_OK;
}
