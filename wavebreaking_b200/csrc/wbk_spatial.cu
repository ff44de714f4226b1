// Spatial pre-processing kernels: fused multi-pass smoothing stencil, generic convolution,
// momentum flux, orientation flip, synthetic PV generator.
// Reference: wavebreaking/processing/spatial.py:27-128, utils/data_utils.py:196-213.
#include "wbk_common.cuh"
#include "wbk_ms.cuh"

// ------------------------------------------------------------------------------------------
// K1: fused `passes` x (5-point stencil / 6), periodic in longitude AND latitude
// (scipy mode="wrap" wraps both axes), accumulation in double in scipy's tap order
//   ((((N + W) + 2C) + E) + S)      [SURVEY.md A.7, probed against scipy 1.18]
// One CTA owns a TH x TW output tile and keeps the tile plus a `passes`-wide halo in shared
// memory (ping-pong), so the field is read once and written once whatever `passes` is.
// ------------------------------------------------------------------------------------------
#define SM_TW 64
#define SM_TH 32
#define SM_THREADS 256

__device__ __forceinline__ int wrap_idx(int i, int n) {
  i %= n;
  return i < 0 ? i + n : i;
}

// periodic index for values that are at most one period outside (falls back to the modulo otherwise)
__device__ __forceinline__ int wrap_near(int i, int n) {
  while (i < 0) i += n;
  while (i >= n) i -= n;
  return i;
}

// x / 6 correctly rounded without the generic division sequence: q = RN(x * RN(1/6)), exact remainder by FMA,
// one correction step (Markstein: with a correctly rounded reciprocal the corrected quotient is RN(x / 6)).
// Checked against 1.5e9 random operands on the host.  Valid for finite operands whose quotient is a normal
// number (or zero); a tile that holds anything else (NaN, Inf, |v| < 1e-150 or > 1e290) takes the plain division.
__device__ __forceinline__ double div6_fast(double x) {
  const double R6 = 0.16666666666666666;  // RN(1/6) = 0x3FC5555555555555
  const double q = __dmul_rn(x, R6);
  const double r = __fma_rn(-6.0, q, x);
  return __fma_rn(r, R6, q);
}

// rounding of one pass: 0 none (float64), 1 sum rounded to float32 then float64 division (NumPy >= 2, first pass
// of float32 data), 2 sum rounded to float32 and float32 division (NumPy 1.x, every pass)
template <int RND, bool SAFE>
__device__ __forceinline__ double smooth_finish(double v) {
  if (RND == 2) return (double)(__double2float_rn(v) / 6.0f);
  if (RND == 1) v = (double)__double2float_rn(v);
  return SAFE ? div6_fast(v) : __ddiv_rn(v, 6.0);
}

// ------------------------------------------------------------------------------------------
// Register-resident fused smoothing.  A warp owns a strip of 64 columns (lane l: columns 2l, 2l+1) and a band of
// SM_PER rows that stay in registers for all P passes; 8 warps stack their bands into a 64 x 64 tile of which
// the inner (64-2P) x (64-2P) cells are valid outputs.  Per pass the only traffic is two 64-bit shuffles per row
// (west / east neighbours) and one row of halo per band edge through shared memory; the arithmetic
// (4 DP adds/FMA + the 3-op exact division per cell) is what remains, so the kernel runs on the FP64 pipe.
// ------------------------------------------------------------------------------------------
#define SM_PER 8                       // rows per warp band
#define SM_TILE 64                     // tile edge = SM_NB bands x SM_PER rows = 32 lanes x 2 columns
#define SM_NB (SM_TILE / SM_PER)       // bands (= warps) per CTA
#define SM_STRIP_THREADS (32 * SM_NB)
#ifndef SM_MIN_CTAS
#define SM_MIN_CTAS 2                  // resident CTAs per SM the register budget is sized for
#endif

// EDGE: 1 = top band of the tile, 2 = bottom band: their outermost k rows are not needed any more after pass k
// (only rows [k, 64-k) of the tile feed the valid interior), so they are skipped; 0 = middle band.
template <int P, int RFIRST, int RREST, bool SAFE, int EDGE>
__device__ __forceinline__ void smooth_strip_passes(double (&vx)[SM_PER], double (&vy)[SM_PER],
                                                    double (*halo)[SM_NB][2][SM_TILE], int lane, int band, int r_base) {
#pragma unroll
  for (int k = 1; k <= P; ++k) {
    const int par = k & 1;
    // publish the band's edge rows (old values), fetch the neighbours' after the barrier
    *reinterpret_cast<double2*>(&halo[par][band][0][2 * lane]) = make_double2(vx[0], vy[0]);
    *reinterpret_cast<double2*>(&halo[par][band][1][2 * lane]) = make_double2(vx[SM_PER - 1], vy[SM_PER - 1]);
    __syncthreads();
    const double2 north = band > 0 ? *reinterpret_cast<const double2*>(&halo[par][band - 1][1][2 * lane]) : make_double2(0.0, 0.0);
    const double2 south = band < SM_NB - 1 ? *reinterpret_cast<const double2*>(&halo[par][band + 1][0][2 * lane]) : make_double2(0.0, 0.0);
    double nx = north.x, ny = north.y;
#pragma unroll
    for (int i = 0; i < SM_PER; ++i) {
      const double cx = vx[i], cy = vy[i];
      if (!((EDGE == 1 && i < k) || (EDGE == 2 && i >= SM_PER - k))) {
        const double sx = i + 1 < SM_PER ? vx[i + 1] : south.x;
        const double sy = i + 1 < SM_PER ? vy[i + 1] : south.y;
        const double wv = __shfl_up_sync(WBK_FULL, cy, 1);    // column 2l-1 (lane 0: unused halo garbage)
        const double ev = __shfl_down_sync(WBK_FULL, cx, 1);  // column 2l+2
        // scipy's tap order ((((N + W) + 2C) + E) + S); (N+W) + 2C in one FMA is the same single rounding
        double a0 = __dadd_rn(nx, wv);
        a0 = __fma_rn(2.0, cx, a0);
        a0 = __dadd_rn(a0, cy);
        a0 = __dadd_rn(a0, sx);
        double a1 = __dadd_rn(ny, cx);
        a1 = __fma_rn(2.0, cy, a1);
        a1 = __dadd_rn(a1, ev);
        a1 = __dadd_rn(a1, sy);
        if (k == 1) {
          vx[i] = smooth_finish<RFIRST, SAFE>(a0);
          vy[i] = smooth_finish<RFIRST, SAFE>(a1);
        } else {
          vx[i] = smooth_finish<RREST, SAFE>(a0);
          vy[i] = smooth_finish<RREST, SAFE>(a1);
        }
      }
      nx = cx;
      ny = cy;
    }
  }

}

// Persistent CTAs stride over the (time, tile row, tile column) list; the raw values of the NEXT tile are
// requested before the current tile is computed, so the DRAM latency hides behind the FP64 work.
// With FUSE the kernel also runs the marching-squares segment stage (contour_index.py:103) on the finished tile while
// it is still on chip (the tile is parked in shared memory), so the contour stage does not re-read the smoothed
// field: tiles then advance by one cell less (a square needs its right / lower neighbours in the same tile).
template <int P, typename TIn, typename TOut, int RMODE, bool FUSE>
__global__ void __launch_bounds__(SM_STRIP_THREADS, SM_MIN_CTAS)
smooth_fused_kernel(const TIn* __restrict__ in, TOut* __restrict__ out, int nlat, int nlon, int nan_border, int tiles_x,
                    int tiles_y, int ntime, const __grid_constant__ WbkDev dev, const __grid_constant__ LevelPack levels,
                    int nlevels) {
  constexpr int OUTW = SM_TILE - 2 * P - (FUSE ? 1 : 0);  // tile stride (valid outputs per tile edge: 64 - 2P)
  __shared__ double halo[2][SM_NB][2][SM_TILE];  // [parity][band][top/bottom][column]
  WBK_DYN_SMEM(double, tile);                    // FUSE: [64][64] finished tile, then SM_NB x 16 hit masks
  const int lane = wbk_lane(), band = wbk_warp();
  const int r_base = band * SM_PER;
  const size_t plane = (size_t)nlat * nlon;
  const unsigned ntiles = (unsigned)tiles_x * (unsigned)tiles_y * (unsigned)ntime;  // < 2^31 (chunked by the launcher)
  const unsigned utx = (unsigned)tiles_x, uty = (unsigned)tiles_y;

  // tile coordinates advance incrementally by the grid stride (no division in the loop)
  const unsigned g_q = gridDim.x / utx, g_bx = gridDim.x - g_q * utx;
  const unsigned g_bt = g_q / uty, g_by = g_q - g_bt * uty;
  unsigned nbx, nby, nbt;  // tile that is loaded next
  {
    const unsigned q = blockIdx.x / utx;
    nbx = blockIdx.x - q * utx;
    nbt = q / uty;
    nby = q - nbt * uty;
  }

  auto load_tile = [&](unsigned bx, unsigned by, unsigned bt, TIn (&tx)[SM_PER], TIn (&ty)[SM_PER]) {
    const int x0 = (int)bx * OUTW - P, y0 = (int)by * OUTW - P;
    const TIn* src = in + plane * bt;
    const int gx0 = wrap_near(x0 + 2 * lane, nlon), gx1 = wrap_near(x0 + 2 * lane + 1, nlon);
    const int gy0 = y0 + r_base;
    if (gy0 >= 0 && gy0 + SM_PER <= nlat) {  // warp-uniform: the band does not cross the latitude wrap
      const TIn* p0 = src + (size_t)gy0 * nlon + gx0;
      const int d1 = gx1 - gx0;
#pragma unroll
      for (int i = 0; i < SM_PER; ++i) {
        tx[i] = p0[0];
        ty[i] = p0[d1];
        p0 += nlon;
      }
    } else {
      int gy = wrap_near(gy0, nlat);
      const TIn* row = src + (size_t)gy * nlon;
#pragma unroll
      for (int i = 0; i < SM_PER; ++i) {
        tx[i] = row[gx0];
        ty[i] = row[gx1];
        ++gy;
        row += nlon;
        if (gy == nlat) {  // periodic in latitude too (scipy mode="wrap")
          gy = 0;
          row = src;
        }
      }
    }
  };

  TIn tx[SM_PER], ty[SM_PER];
  unsigned w = blockIdx.x;
  if (w < ntiles) load_tile(nbx, nby, nbt, tx, ty);
  for (; w < ntiles; w += gridDim.x) {
    const int x0 = (int)nbx * OUTW - P, y0 = (int)nby * OUTW - P, bt = (int)nbt;  // the tile computed now
    double vx[SM_PER], vy[SM_PER];
    int unsafe = 0;
    if (sizeof(TIn) == 4) {
      // float input: only NaN / Inf leave the range div6_fast is valid for (x * 0 is NaN exactly for those)
      float acc = 0.0f;
#pragma unroll
      for (int i = 0; i < SM_PER; ++i) {
        acc = fmaf((float)tx[i], 0.0f, acc);
        acc = fmaf((float)ty[i], 0.0f, acc);
        vx[i] = (double)tx[i];
        vy[i] = (double)ty[i];
      }
      unsafe = !(acc == 0.0f);
#ifdef __CUDA_ARCH__
      asm volatile("" : "+r"(nbx) : "f"(acc));  // keep the check ahead of the prefetch (no copies of the raw tile)
#endif
    } else {
#pragma unroll
      for (int i = 0; i < SM_PER; ++i) {
        vx[i] = (double)tx[i];
        vy[i] = (double)ty[i];
        const double ax = fabs(vx[i]), ay = fabs(vy[i]);
        unsafe |= !(ax <= 1e290) || (ax < 1e-150 && ax != 0.0) || !(ay <= 1e290) || (ay < 1e-150 && ay != 0.0);
      }
    }
    // advance to and prefetch the next tile of this CTA
    nbx += g_bx;
    if (nbx >= utx) {
      nbx -= utx;
      ++nby;
    }
    nby += g_by;
    if (nby >= uty) {
      nby -= uty;
      ++nbt;
    }
    nbt += g_bt;
    if (w + gridDim.x < ntiles) load_tile(nbx, nby, nbt, tx, ty);
    const int slow = __syncthreads_or(unsafe);

    constexpr int RFIRST = RMODE == WBK_ROUND_ALL ? 2 : (RMODE == WBK_ROUND_FIRST ? 1 : 0);
    constexpr int RREST = RMODE == WBK_ROUND_ALL ? 2 : 0;
    if (slow) smooth_strip_passes<P, RFIRST, RREST, false, 0>(vx, vy, halo, lane, band, r_base);
    else if (band == 0) smooth_strip_passes<P, RFIRST, RREST, true, 1>(vx, vy, halo, lane, band, r_base);
    else if (band == SM_NB - 1) smooth_strip_passes<P, RFIRST, RREST, true, 2>(vx, vy, halo, lane, band, r_base);
    else smooth_strip_passes<P, RFIRST, RREST, true, 0>(vx, vy, halo, lane, band, r_base);

    // write the valid interior: tile rows / columns [P, 64 - P)
    const int c_lo = 2 * lane;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    // the NaN border (spatial.py:106-107) only concerns bands at the top / bottom of the grid (warp-uniform)
    if (nan_border > 0 && (y0 + r_base < nan_border || y0 + r_base + SM_PER > nlat - nan_border)) {
#pragma unroll
      for (int i = 0; i < SM_PER; ++i) {
        const int gy = y0 + r_base + i;
        if (gy < nan_border || (gy >= nlat - nan_border && gy < nlat)) vx[i] = vy[i] = qnan;
      }
    }
    {
      const int ox0 = x0 + c_lo;
      const bool ok0 = c_lo >= P && c_lo < SM_TILE - P && ox0 < nlon;
      const bool ok1 = c_lo + 1 >= P && c_lo + 1 < SM_TILE - P && ox0 + 1 < nlon;
      TOut* dp = out + plane * bt + ((long long)(y0 + r_base) * nlon + ox0);
      int rows_left = nlat - (y0 + r_base);  // rows of this band that exist in the grid
#pragma unroll
      for (int i = 0; i < SM_PER; ++i) {
        const int r = r_base + i;
        if (r >= P && r < SM_TILE - P && i < rows_left) {  // warp-uniform
          if (ok0) dp[0] = (TOut)vx[i];
          if (ok1) dp[1] = (TOut)vy[i];
        }
        dp += nlon;
      }
    }
    if (FUSE) {
      // ---- marching squares on the finished tile (squares whose upper-left corner is in tile rows / columns
      //      [P, P + OUTW); all four corners are valid cells of this tile)
      u32* masks = reinterpret_cast<u32*>(tile + SM_TILE * SM_TILE) + band * 2 * SM_PER;
      unsigned char* bits = reinterpret_cast<unsigned char*>(tile + SM_TILE * SM_TILE) + SM_NB * 2 * SM_PER * sizeof(u32);
      double fx[SM_PER], fy[SM_PER];  // final values of this lane's cells (NaN border applied)
#pragma unroll
      for (int i = 0; i < SM_PER; ++i) {
        const int r = r_base + i;
        fx[i] = vx[i];
        fy[i] = vy[i];
        *reinterpret_cast<double2*>(&tile[r * SM_TILE + c_lo]) = make_double2(fx[i], fy[i]);
      }
      for (int l = 0; l < nlevels; ++l) {
        const double level = levels.v[l];
        // one comparison per cell: bit 0 = value > level, bit 2 = NaN; the squares are classified from these bytes
        __syncthreads();
#pragma unroll
        for (int i = 0; i < SM_PER; ++i) {
          const unsigned bxv = (fx[i] > level ? 1u : 0u) | (fx[i] != fx[i] ? 4u : 0u);
          const unsigned byv = (fy[i] > level ? 1u : 0u) | (fy[i] != fy[i] ? 4u : 0u);
          *reinterpret_cast<unsigned short*>(&bits[(r_base + i) * SM_TILE + c_lo]) = (unsigned short)(bxv | (byv << 8));
        }
        __syncthreads();
        u32 any_hits = 0;
        // bytes of columns c_lo .. c_lo+2 of a row (the third belongs to the next lane)
        auto row_bits = [&](int r) -> unsigned {
          const unsigned a2 = *reinterpret_cast<const unsigned short*>(&bits[r * SM_TILE + c_lo]);
          const unsigned b1 = c_lo + 2 < SM_TILE ? bits[r * SM_TILE + c_lo + 2] : 4u;
          return a2 | (b1 << 16);
        };
        unsigned up = row_bits(r_base);
#pragma unroll
        for (int i = 0; i < SM_PER; ++i) {
          const int r = r_base + i, gy = y0 + r;
          const unsigned dn = r + 1 < SM_TILE ? row_bits(r + 1) : 0x040404u;
          const bool row_ok = r >= P && r < P + OUTW && gy >= 0 && gy + 1 <= nlat - 1;
#pragma unroll
          for (int par = 0; par < 2; ++par) {
            const int c = c_lo + par, gx = x0 + c;
            const unsigned ul = (up >> (8 * par)) & 0xffu, ur = (up >> (8 * par + 8)) & 0xffu;
            const unsigned ll = (dn >> (8 * par)) & 0xffu, lr = (dn >> (8 * par + 8)) & 0xffu;
            int sq = (int)((ul & 1u) | ((ur & 1u) << 1) | ((ll & 1u) << 2) | ((lr & 1u) << 3));
            if (((ul | ur | ll | lr) & 4u) || sq == 15 || !row_ok || c < P || c >= P + OUTW || gx < 0 || gx > nlon - 1 ||
                gx > dev.W - 2)
              sq = 0;
            const u32 hits = __ballot_sync(WBK_FULL, sq != 0);
            if (lane == 0) masks[2 * i + par] = hits;
            any_hits |= hits;
          }
          up = dn;
        }
        if (any_hits) {  // warp-uniform
          __syncwarp();
          int total = 0;
          for (int m = 0; m < 2 * SM_PER; ++m) total += __popc(masks[m]);
          for (int h0 = 0; h0 < total; h0 += 32) {
            const int h = h0 + lane;
            int mi = 0, sl = 0, r0 = 0, c0 = 0;
            double ul = 0, ur = 0, ll = 0, lr = 0;
            const bool active = h < total && ms_locate_hit(masks, 2 * SM_PER, h, mi, sl);
            if (active) {
              const int r = r_base + (mi >> 1), c = 2 * sl + (mi & 1);
              r0 = y0 + r;
              c0 = x0 + c;
              ul = tile[r * SM_TILE + c]; ur = tile[r * SM_TILE + c + 1];
              ll = tile[(r + 1) * SM_TILE + c]; lr = tile[(r + 1) * SM_TILE + c + 1];
            }
            ms_emit_squares(dev, bt * nlevels + l, active, r0, c0, ul, ur, ll, lr, level);
          }
          __syncwarp();
        }
      }
    }
    // the halo buffers are reused by the next tile: its __syncthreads_or (before the first pass) already orders that;
    // only the parked tile of the fused variant needs a barrier of its own
    if (FUSE) __syncthreads();
  }
}

template <int P, typename TIn, typename TOut, int RMODE>
static int launch_smooth_p(const void* in, void* out, int ntime, int nlat, int nlon, int nan_border, cudaStream_t st) {
  constexpr int OUTW = SM_TILE - 2 * P;
  const int tiles_x = (nlon + OUTW - 1) / OUTW, tiles_y = (nlat + OUTW - 1) / OUTW;
  WbkDev none = {};
  LevelPack lv = {};
  // the kernel indexes tiles with 32 bits: very long series go in chunks of time steps
  const long long per_step = (long long)tiles_x * tiles_y;
  const int max_t = (int)((1LL << 30) / per_step) > 0 ? (int)((1LL << 30) / per_step) : 1;
  for (int t0 = 0; t0 < ntime; t0 += max_t) {
    const int nt = ntime - t0 < max_t ? ntime - t0 : max_t;
    const long long ntiles = per_step * nt;
    const int grid = (int)(ntiles < 148 * SM_MIN_CTAS ? ntiles : 148 * SM_MIN_CTAS);  // persistent CTAs
    const size_t off = (size_t)t0 * nlat * nlon;
    WBK_LAUNCH(KID_SMOOTH, (smooth_fused_kernel<P, TIn, TOut, RMODE, false>), dim3(grid), dim3(SM_STRIP_THREADS), 0, st,
               (const TIn*)in + off, (TOut*)out + off, nlat, nlon, nan_border, tiles_x, tiles_y, nt, none, lv, 0);
    WBK_LAUNCH_CHECK();
  }
  return WBK_OK;
}

// fused smoothing + marching squares (float64 output); the segment arenas of `dev` must have been reset
template <int P, typename TIn, int RMODE>
static int launch_smooth_ms_p(const void* in, double* out, int ntime, int nlat, int nlon, const WbkDev& dev,
                              const LevelPack& lv, int nlevels, cudaStream_t st) {
  constexpr int OUTW = SM_TILE - 2 * P - 1;
  const int tiles_x = (nlon + OUTW - 1) / OUTW, tiles_y = (nlat + OUTW - 1) / OUTW;
  const long long ntiles = (long long)tiles_x * tiles_y * ntime;
  if (ntiles > (1LL << 30)) {
    wbk_set_error("wbk_smooth_contours: batch too long (%lld tiles), split the time axis", ntiles);
    return WBK_ERR_INVALID;
  }
  const int grid = (int)(ntiles < 148 * SM_MIN_CTAS ? ntiles : 148 * SM_MIN_CTAS);
  const size_t smem = (size_t)SM_TILE * SM_TILE * sizeof(double) + (size_t)SM_NB * 2 * SM_PER * sizeof(u32) +
                      (size_t)SM_TILE * SM_TILE;  // parked tile, hit masks, comparison bytes
  WBK_CUDA_CHECK(cudaFuncSetAttribute(smooth_fused_kernel<P, TIn, double, RMODE, true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  WBK_LAUNCH(KID_SMOOTH_MS, (smooth_fused_kernel<P, TIn, double, RMODE, true>), dim3(grid), dim3(SM_STRIP_THREADS), smem, st,
             (const TIn*)in, out, nlat, nlon, 2, tiles_x, tiles_y, ntime, dev, lv, nlevels);
  WBK_LAUNCH_CHECK();
  return WBK_OK;
}

template <typename TIn, int RMODE>
static int launch_smooth_ms(const void* in, double* out, int ntime, int nlat, int nlon, int passes, const WbkDev& dev,
                            const LevelPack& lv, int nlevels, cudaStream_t st) {
  switch (passes) {
    case 1: return launch_smooth_ms_p<1, TIn, RMODE>(in, out, ntime, nlat, nlon, dev, lv, nlevels, st);
    case 2: return launch_smooth_ms_p<2, TIn, RMODE>(in, out, ntime, nlat, nlon, dev, lv, nlevels, st);
    case 3: return launch_smooth_ms_p<3, TIn, RMODE>(in, out, ntime, nlat, nlon, dev, lv, nlevels, st);
    case 4: return launch_smooth_ms_p<4, TIn, RMODE>(in, out, ntime, nlat, nlon, dev, lv, nlevels, st);
    case 5: return launch_smooth_ms_p<5, TIn, RMODE>(in, out, ntime, nlat, nlon, dev, lv, nlevels, st);
    case 6: return launch_smooth_ms_p<6, TIn, RMODE>(in, out, ntime, nlat, nlon, dev, lv, nlevels, st);
    case 7: return launch_smooth_ms_p<7, TIn, RMODE>(in, out, ntime, nlat, nlon, dev, lv, nlevels, st);
    case 8: return launch_smooth_ms_p<8, TIn, RMODE>(in, out, ntime, nlat, nlon, dev, lv, nlevels, st);
  }
  return WBK_ERR_INVALID;
}

// used by wbk_smooth_contours (wbk_contours.cu)
int wbk_launch_smooth_ms(const void* d_in, int in_dtype, double* d_out, int ntime, int nlat, int nlon, int passes,
                         const WbkDev& dev, const LevelPack& lv, int nlevels, cudaStream_t st) {
  if (in_dtype == WBK_F32) return launch_smooth_ms<float, WBK_ROUND_FIRST>(d_in, d_out, ntime, nlat, nlon, passes, dev, lv, nlevels, st);
  return launch_smooth_ms<double, WBK_ROUND_NONE>(d_in, d_out, ntime, nlat, nlon, passes, dev, lv, nlevels, st);
}

template <typename TIn, typename TOut, int RMODE>
static int launch_smooth_mode(const void* in, void* out, int ntime, int nlat, int nlon, int passes, int nan_border,
                              cudaStream_t st) {
  switch (passes) {
    case 1: return launch_smooth_p<1, TIn, TOut, RMODE>(in, out, ntime, nlat, nlon, nan_border, st);
    case 2: return launch_smooth_p<2, TIn, TOut, RMODE>(in, out, ntime, nlat, nlon, nan_border, st);
    case 3: return launch_smooth_p<3, TIn, TOut, RMODE>(in, out, ntime, nlat, nlon, nan_border, st);
    case 4: return launch_smooth_p<4, TIn, TOut, RMODE>(in, out, ntime, nlat, nlon, nan_border, st);
    case 5: return launch_smooth_p<5, TIn, TOut, RMODE>(in, out, ntime, nlat, nlon, nan_border, st);
    case 6: return launch_smooth_p<6, TIn, TOut, RMODE>(in, out, ntime, nlat, nlon, nan_border, st);
    case 7: return launch_smooth_p<7, TIn, TOut, RMODE>(in, out, ntime, nlat, nlon, nan_border, st);
    case 8: return launch_smooth_p<8, TIn, TOut, RMODE>(in, out, ntime, nlat, nlon, nan_border, st);
  }
  wbk_set_error("wbk_smooth: internal pass count %d", passes);
  return WBK_ERR_INVALID;
}

template <typename TIn, typename TOut>
static int launch_smooth(const void* in, void* out, int ntime, int nlat, int nlon, int passes, int round_first,
                         int round_all, int nan_border, cudaStream_t st) {
  if (round_all) return launch_smooth_mode<TIn, TOut, WBK_ROUND_ALL>(in, out, ntime, nlat, nlon, passes, nan_border, st);
  if (round_first) return launch_smooth_mode<TIn, TOut, WBK_ROUND_FIRST>(in, out, ntime, nlat, nlon, passes, nan_border, st);
  return launch_smooth_mode<TIn, TOut, WBK_ROUND_NONE>(in, out, ntime, nlat, nlon, passes, nan_border, st);
}

extern "C" int wbk_smooth(const void* d_in, int in_dtype, void* d_out, int out_dtype, void* d_tmp, int ntime,
                          int nlat, int nlon, int passes, int round_mode, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!d_in || !d_out || ntime < 0 || nlat < 4 || nlon < 1 || passes < 0) {
    wbk_set_error("wbk_smooth: invalid argument");
    return WBK_ERR_INVALID;
  }
  if (ntime == 0) return WBK_OK;
  const int border = 2;  // int(3 / 2 + 0.5), spatial.py:106
  if (passes == 0) {
    if (in_dtype != out_dtype) {
      wbk_set_error("wbk_smooth: passes == 0 keeps the dtype");
      return WBK_ERR_INVALID;
    }
    size_t bytes = (size_t)ntime * nlat * nlon * (in_dtype == WBK_F32 ? 4 : 8);
    if (d_in != d_out) WBK_CUDA_CHECK(cudaMemcpyAsync(d_out, d_in, bytes, cudaMemcpyDeviceToDevice, st));
    return wbk_nan_border(d_out, out_dtype, ntime, nlat, nlon, border, stream);
  }
  const bool f32_in = in_dtype == WBK_F32;
  if ((round_mode == WBK_ROUND_NONE && (f32_in || out_dtype != WBK_F64)) ||
      (round_mode == WBK_ROUND_FIRST && (!f32_in || out_dtype != WBK_F64)) ||
      (round_mode == WBK_ROUND_ALL && (!f32_in || out_dtype != WBK_F32))) {
    wbk_set_error("wbk_smooth: dtype / round_mode combination not supported");
    return WBK_ERR_INVALID;
  }
  if (passes > WBK_SMOOTH_MAX_FUSED && !d_tmp) {
    wbk_set_error("wbk_smooth: d_tmp required for passes > %d", WBK_SMOOTH_MAX_FUSED);
    return WBK_ERR_INVALID;
  }
  const int round_all = round_mode == WBK_ROUND_ALL;
  // chunk the passes; intermediates are stored in the output dtype (exact: f64, or f32 values under ROUND_ALL)
  int nchunks = (passes + WBK_SMOOTH_MAX_FUSED - 1) / WBK_SMOOTH_MAX_FUSED;
  const void* src = d_in;
  int done = 0;
  for (int ch = 0; ch < nchunks; ++ch) {
    int n = passes - done < WBK_SMOOTH_MAX_FUSED ? passes - done : WBK_SMOOTH_MAX_FUSED;
    bool last = ch == nchunks - 1;
    // ping-pong so that the last chunk lands in d_out
    void* dstp = ((nchunks - 1 - ch) % 2 == 0) ? d_out : d_tmp;
    int rf = (round_mode == WBK_ROUND_FIRST && ch == 0) ? 1 : 0;
    int nb = last ? border : 0;
    int rc;
    if (ch == 0) {
      if (f32_in && out_dtype == WBK_F64) rc = launch_smooth<float, double>(src, dstp, ntime, nlat, nlon, n, rf, round_all, nb, st);
      else if (f32_in) rc = launch_smooth<float, float>(src, dstp, ntime, nlat, nlon, n, rf, round_all, nb, st);
      else rc = launch_smooth<double, double>(src, dstp, ntime, nlat, nlon, n, rf, round_all, nb, st);
    } else {
      if (out_dtype == WBK_F64) rc = launch_smooth<double, double>(src, dstp, ntime, nlat, nlon, n, 0, round_all, nb, st);
      else rc = launch_smooth<float, float>(src, dstp, ntime, nlat, nlon, n, 0, round_all, nb, st);
    }
    if (rc != WBK_OK) return rc;
    src = dstp;
    done += n;
  }
  return WBK_OK;
}

// ------------------------------------------------------------------------------------------
// generic single-pass scipy.ndimage.convolve (any 2-D weights, five boundary modes)
// ------------------------------------------------------------------------------------------
#define CV_MAX_TAPS 225
struct ConvTaps {
  int n;
  int dy[CV_MAX_TAPS];
  int dx[CV_MAX_TAPS];
  double w[CV_MAX_TAPS];
};

__device__ __forceinline__ int extend_index(int i, int n, int mode, bool* outside) {
  *outside = false;
  if (i >= 0 && i < n) return i;
  switch (mode) {
    case 0:  // wrap
      return wrap_idx(i, n);
    case 1: {  // reflect: d c b a | a b c d | d c b a
      int period = 2 * n;
      int m = wrap_idx(i, period);
      return m < n ? m : period - 1 - m;
    }
    case 2: {  // mirror: d c b | a b c d | c b a
      if (n == 1) return 0;
      int period = 2 * n - 2;
      int m = wrap_idx(i, period);
      return m < n ? m : period - m;
    }
    case 3:  // nearest
      return i < 0 ? 0 : n - 1;
    default:  // constant
      *outside = true;
      return 0;
  }
}

template <typename TIn, typename TMid, typename TOut>
__global__ void convolve2d_cast_kernel(const TIn* __restrict__ in, TOut* __restrict__ out, int nlat, int nlon,
                                       ConvTaps taps, int mode, int divide, double divisor) {
  const size_t plane = (size_t)nlat * nlon;
  const TIn* src = in + plane * blockIdx.z;
  TOut* dst = out + plane * blockIdx.z;
  int gx = blockIdx.x * blockDim.x + threadIdx.x;
  int gy = blockIdx.y;
  if (gx >= nlon) return;
  double acc = 0.0;
  for (int t = 0; t < taps.n; ++t) {
    bool oy, ox;
    int yy = extend_index(gy + taps.dy[t], nlat, mode, &oy);
    int xx = extend_index(gx + taps.dx[t], nlon, mode, &ox);
    double v = (oy || ox) ? 0.0 : (double)src[(size_t)yy * nlon + xx];
    acc = __dadd_rn(acc, __dmul_rn(v, taps.w[t]));
  }
  TMid r = (TMid)acc;
  double rr = (double)r;
  if (divide) rr = __ddiv_rn(rr, divisor);
  dst[(size_t)gy * nlon + gx] = (TOut)rr;
}

extern "C" int wbk_convolve2d(const void* d_in, int in_dtype, void* d_out, int out_dtype, int ntime, int nlat,
                              int nlon, const double* h_weights, int kh, int kw, int mode, int divide, double divisor,
                              void* stream) {
  if (!d_in || !d_out || !h_weights || kh < 1 || kw < 1 || mode < 0 || mode > 4 || nlat < 1 || nlon < 1) {
    wbk_set_error("wbk_convolve2d: invalid argument");
    return WBK_ERR_INVALID;
  }
  if (ntime == 0) return WBK_OK;
  ConvTaps taps;
  taps.n = 0;
  // convolve(input, w) == correlate(input, flipped w) with origin shifted by -1 on even sizes
  const int cy = kh / 2 - ((kh % 2 == 0) ? 1 : 0);
  const int cx = kw / 2 - ((kw % 2 == 0) ? 1 : 0);
  for (int a = 0; a < kh; ++a)
    for (int b = 0; b < kw; ++b) {
      double w = h_weights[(kh - 1 - a) * kw + (kw - 1 - b)];
      if (w != 0.0) {
        if (taps.n >= CV_MAX_TAPS) {
          wbk_set_error("wbk_convolve2d: more than %d non-zero weights", CV_MAX_TAPS);
          return WBK_ERR_INVALID;
        }
        taps.dy[taps.n] = a - cy;
        taps.dx[taps.n] = b - cx;
        taps.w[taps.n] = w;
        ++taps.n;
      }
    }
  cudaStream_t st = (cudaStream_t)stream;
  dim3 block(128);
  dim3 grid((nlon + 127) / 128, nlat, ntime);
  if (in_dtype == WBK_F32 && out_dtype == WBK_F32) {
    WBK_LAUNCH(KID_CONVOLVE, (convolve2d_cast_kernel<float, float, float>), grid, block, 0, st, (const float*)d_in, (float*)d_out, nlat, nlon, taps, mode, divide, divisor);
  } else if (in_dtype == WBK_F32 && out_dtype == WBK_F64) {
    WBK_LAUNCH(KID_CONVOLVE, (convolve2d_cast_kernel<float, float, double>), grid, block, 0, st, (const float*)d_in, (double*)d_out, nlat, nlon, taps, mode, divide, divisor);
  } else if (in_dtype == WBK_F64 && out_dtype == WBK_F64) {
    WBK_LAUNCH(KID_CONVOLVE, (convolve2d_cast_kernel<double, double, double>), grid, block, 0, st, (const double*)d_in, (double*)d_out, nlat, nlon, taps, mode, divide, divisor);
  } else {
    wbk_set_error("wbk_convolve2d: unsupported dtype combination");
    return WBK_ERR_INVALID;
  }
  WBK_LAUNCH_CHECK();
  return WBK_OK;
}

// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void nan_border_kernel(T* f, int nlat, int nlon, int border) {
  int gx = blockIdx.x * blockDim.x + threadIdx.x;
  if (gx >= nlon) return;
  int k = blockIdx.y;  // 0 .. 2*border-1
  int gy = k < border ? k : nlat - 2 * border + k;
  if (gy < 0 || gy >= nlat) return;
  f[(size_t)blockIdx.z * nlat * nlon + (size_t)gy * nlon + gx] = (T)__longlong_as_double(0x7ff8000000000000LL);
}

extern "C" int wbk_nan_border(void* d_field, int dtype, int ntime, int nlat, int nlon, int border, void* stream) {
  if (!d_field || border < 0 || nlat < 1 || nlon < 1) {
    wbk_set_error("wbk_nan_border: invalid argument");
    return WBK_ERR_INVALID;
  }
  if (ntime == 0 || border == 0) return WBK_OK;
  dim3 grid((nlon + 255) / 256, 2 * border, ntime);
  if (dtype == WBK_F32) WBK_LAUNCH(KID_NAN_BORDER, nan_border_kernel<float>, grid, dim3(256), 0, (cudaStream_t)stream, (float*)d_field, nlat, nlon, border);
  else WBK_LAUNCH(KID_NAN_BORDER, nan_border_kernel<double>, grid, dim3(256), 0, (cudaStream_t)stream, (double*)d_field, nlat, nlon, border);
  WBK_LAUNCH_CHECK();
  return WBK_OK;
}

// ------------------------------------------------------------------------------------------
// K2: momentum flux.  One CTA per (time, lat) row: NaN-skipping zonal means of u and v in double,
// then (u - ubar) * (v - vbar) in the data dtype (xarray arithmetic keeps float32).
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void mflux_kernel(const T* __restrict__ u, const T* __restrict__ v, T* __restrict__ out, int nlon) {
  __shared__ double red[34];
  __shared__ double means[2];
  const size_t row = ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * nlon;
  double su = 0, sv = 0, cu = 0, cv = 0;
  for (int x = threadIdx.x; x < nlon; x += blockDim.x) {
    double a = (double)u[row + x], b = (double)v[row + x];
    if (!isnan(a)) { su += a; cu += 1; }
    if (!isnan(b)) { sv += b; cv += 1; }
  }
  su = wbk_block_sum_f64(su, red);
  cu = wbk_block_sum_f64(cu, red);
  sv = wbk_block_sum_f64(sv, red);
  cv = wbk_block_sum_f64(cv, red);
  if (threadIdx.x == 0) {
    means[0] = (double)(T)(su / cu);  // nanmean returns the data dtype
    means[1] = (double)(T)(sv / cv);
  }
  __syncthreads();
  const T mu = (T)means[0], mv = (T)means[1];
  for (int x = threadIdx.x; x < nlon; x += blockDim.x) {
    T up = u[row + x] - mu;
    T vp = v[row + x] - mv;
    out[row + x] = up * vp;
  }
}

extern "C" int wbk_mflux(const void* d_u, const void* d_v, void* d_out, int dtype, int ntime, int nlat, int nlon,
                         void* stream) {
  if (!d_u || !d_v || !d_out || nlat < 1 || nlon < 1) {
    wbk_set_error("wbk_mflux: invalid argument");
    return WBK_ERR_INVALID;
  }
  if (ntime == 0) return WBK_OK;
  dim3 grid(nlat, ntime);
  if (dtype == WBK_F32) WBK_LAUNCH(KID_MFLUX, mflux_kernel<float>, grid, dim3(256), 0, (cudaStream_t)stream, (const float*)d_u, (const float*)d_v, (float*)d_out, nlon);
  else WBK_LAUNCH(KID_MFLUX, mflux_kernel<double>, grid, dim3(256), 0, (cudaStream_t)stream, (const double*)d_u, (const double*)d_v, (double*)d_out, nlon);
  WBK_LAUNCH_CHECK();
  return WBK_OK;
}

// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void flip_kernel(const T* __restrict__ in, T* __restrict__ out, int nlat, int nlon, int flip_lat, int flip_lon) {
  int gx = blockIdx.x * blockDim.x + threadIdx.x;
  if (gx >= nlon) return;
  int gy = blockIdx.y;
  size_t base = (size_t)blockIdx.z * nlat * nlon;
  int sy = flip_lat ? nlat - 1 - gy : gy;
  int sx = flip_lon ? nlon - 1 - gx : gx;
  out[base + (size_t)gy * nlon + gx] = in[base + (size_t)sy * nlon + sx];
}

extern "C" int wbk_flip(const void* d_in, void* d_out, int dtype, int ntime, int nlat, int nlon, int flip_lat,
                        int flip_lon, void* stream) {
  if (!d_in || !d_out || d_in == d_out || nlat < 1 || nlon < 1) {
    wbk_set_error("wbk_flip: invalid argument (in-place is not supported)");
    return WBK_ERR_INVALID;
  }
  if (ntime == 0) return WBK_OK;
  dim3 grid((nlon + 255) / 256, nlat, ntime);
  if (dtype == WBK_F32) WBK_LAUNCH(KID_FLIP, flip_kernel<float>, grid, dim3(256), 0, (cudaStream_t)stream, (const float*)d_in, (float*)d_out, nlat, nlon, flip_lat, flip_lon);
  else WBK_LAUNCH(KID_FLIP, flip_kernel<double>, grid, dim3(256), 0, (cudaStream_t)stream, (const double*)d_in, (double*)d_out, nlat, nlon, flip_lat, flip_lon);
  WBK_LAUNCH_CHECK();
  return WBK_OK;
}

// ------------------------------------------------------------------------------------------
// synthetic PV (mirror of wavebreaking_b200/synthetic.py)
// ------------------------------------------------------------------------------------------
#define SYN_MAX_BLOB 32
struct SynthParams {
  double A, k, tilt, c, env0, env1, env_speed, blob_amp, sh_shift, A2, k2, c2, tilt2;
  int n_blob;
  double lat0[2][SYN_MAX_BLOB], lon0[2][SYN_MAX_BLOB], rad[2][SYN_MAX_BLOB];
};

__device__ __forceinline__ double synth_background(double alat, double lam_deg, double hours, const SynthParams& p) {
  const double d2r = 0.017453292519943295;
  double lam = lam_deg * d2r;
  double env = p.env0 + p.env1 * cos(lam - p.env_speed * hours * d2r);
  double phase = p.k * (lam - p.c * hours * d2r) + p.tilt * (alat - 45.0) / 10.0 + 0.7 * sin(2.0 * lam);
  double phase2 = p.k2 * (lam - p.c2 * hours * d2r) + p.tilt2 * (alat - 45.0) / 10.0;
  double phi = alat - p.A * env * sin(phase) - p.A2 * sin(phase2);
  double s = sin(phi * d2r) / sin(45.0 * d2r);
  double a = fabs(s);
  double r = 2.0 * a * a * a;
  return s < 0 ? -r : r;
}

template <typename T>
__global__ void synth_pv_kernel(T* out, int nlat, int nlon, double hour0, double hour_step, SynthParams p) {
  int gx = blockIdx.x * blockDim.x + threadIdx.x;
  if (gx >= nlon) return;
  int gy = blockIdx.y;
  int t = blockIdx.z;
  const double d2r = 0.017453292519943295;
  double hours = hour0 + hour_step * t;
  double lat = -90.0 + 180.0 * gy / (double)(nlat - 1);
  double lon = gx * (360.0 / nlon);
  bool south = lat < 0;
  double alat = fabs(lat);
  int hemi = south ? 1 : 0;
  double pv = synth_background(alat, south ? lon + p.sh_shift : lon, hours, p);
  for (int b = 0; b < p.n_blob; ++b) {
    double lonc = fmod(p.lon0[hemi][b] + p.c * hours, 360.0);
    double bg = synth_background(p.lat0[hemi][b], lonc + (hemi ? p.sh_shift : 0.0), hours, p);
    double sign = bg > 2.0 ? -1.0 : 1.0;
    double dl = fmod(lon - lonc + 180.0, 360.0);
    if (dl < 0) dl += 360.0;
    dl -= 180.0;
    double dy = alat - p.lat0[hemi][b];
    double dxs = dl * cos(p.lat0[hemi][b] * d2r);
    double d2 = dy * dy + dxs * dxs;
    double rad = p.rad[hemi][b];
    pv += sign * p.blob_amp * exp(-d2 / (2.0 * rad * rad));
  }
  out[((size_t)t * nlat + gy) * nlon + gx] = (T)(south ? -pv : pv);
}

extern "C" int wbk_synth_pv(void* d_out, int dtype, int ntime, int nlat, int nlon, double hour0, double hour_step,
                            const double* h_blobs, int n_blob, void* stream) {
  if (!d_out || nlat < 2 || nlon < 1 || n_blob < 0 || n_blob > SYN_MAX_BLOB || (n_blob > 0 && !h_blobs)) {
    wbk_set_error("wbk_synth_pv: invalid argument");
    return WBK_ERR_INVALID;
  }
  if (ntime == 0) return WBK_OK;
  SynthParams p;
  // frozen recipe: keep in sync with wavebreaking_b200/synthetic.py PARAMS
  p.A = 12.0; p.k = 7.0; p.tilt = 2.6; p.c = 0.5; p.env0 = 0.6; p.env1 = 0.4; p.env_speed = 1.5;
  p.blob_amp = 3.0; p.sh_shift = 37.0; p.A2 = 4.5; p.k2 = 19.0; p.c2 = 1.1; p.tilt2 = 2.0; p.n_blob = n_blob;
  for (int h = 0; h < 2; ++h)
    for (int b = 0; b < n_blob; ++b) {
      p.lat0[h][b] = h_blobs[(h * n_blob + b) * 3 + 0];
      p.lon0[h][b] = h_blobs[(h * n_blob + b) * 3 + 1];
      p.rad[h][b] = h_blobs[(h * n_blob + b) * 3 + 2];
    }
  dim3 grid((nlon + 127) / 128, nlat, ntime);
  if (dtype == WBK_F32) WBK_LAUNCH(KID_SYNTH, synth_pv_kernel<float>, grid, dim3(128), 0, (cudaStream_t)stream, (float*)d_out, nlat, nlon, hour0, hour_step, p);
  else WBK_LAUNCH(KID_SYNTH, synth_pv_kernel<double>, grid, dim3(128), 0, (cudaStream_t)stream, (double*)d_out, nlat, nlon, hour0, hour_step, p);
  WBK_LAUNCH_CHECK();
  return WBK_OK;
}
