// K1: fused multi-pass smoothing stencil, streamed along latitude in registers, with the marching-squares
// comparison bit planes as a by-product.
// Reference: wavebreaking/processing/spatial.py:60-128 (calculate_smoothed_field): `passes` x
//   [scipy.ndimage.convolve(weights=[[0,1,0],[1,2,1],[0,1,0]], mode="wrap") / np.sum(weights) = 6],
// periodic in longitude AND latitude (scipy wraps both axes), accumulation in double in scipy's tap order
//   ((((N + W) + 2C) + E) + S)      [SURVEY.md A.7, probed against scipy 1.18]
// then rows {0,1,nlat-2,nlat-1} := NaN (spatial.py:106-107).  The orientation fix of
// utils/data_utils.py:196-213 (descending latitude / longitude) is folded into the load addresses, and packed
// int16 input (NetCDF scale_factor / add_offset, decoded as xarray does: float64(v) * scale + offset) is decoded
// in the load.
//
// Design.  A warp owns a strip of 64 columns (lane l: columns 2l, 2l+1) of ONE time step and marches down the
// rows.  For every pass level p it keeps the last three rows of that level in registers; at step k the new
// input row k enters level 0 and level p emits its row k - 2p from the three rows of level p - 1 that were
// produced in EARLIER steps, so the P row updates of a step are independent of each other (P x 2 independent
// FP64 chains of 7 operations) and nothing is exchanged between warps: no shared memory, no block barrier.
// West / east neighbours come from two 64-bit shuffles per row update.  A strip yields 64 - 2P valid columns
// (halo recompute 64 / (64 - 2P) in x only; in y a warp marches over the whole chunk plus 3P warm-up rows).
// The field is read once and written once whatever `passes` is.
#pragma once
#include "wbk_common.cuh"
#include "wbk_ms.cuh"

#define SS_THREADS 128
#ifndef SS_MINCTA_SMALL
#define SS_MINCTA_SMALL 4  // resident CTAs per SM the register budget is sized for (P <= 5)
#endif
#ifndef SS_MAX_P_H2
#define SS_MAX_P_H2 5  // largest pass count that uses two 64-column halves per warp (register budget)
#endif
#ifndef SS_RING
#define SS_RING 4  // rows of the shared-memory input ring (power of two); 4 / 8 / 16 measured within 2 % on B200
#endif

__device__ __forceinline__ int ss_wrap(int i, int n) {
  i %= n;
  return i < 0 ? i + n : i;
}

// x / 6 correctly rounded without the generic division sequence: q = RN(x * RN(1/6)), exact remainder by FMA,
// one correction step (Markstein: with a correctly rounded reciprocal the corrected quotient is RN(x / 6)).
// Checked against 1.5e9 random operands on the host.  Valid for finite operands whose quotient and remainder are
// normal numbers (or zero); rows that may hold anything else take the plain division (see `slow_until`).
__device__ __forceinline__ double ss_div6_fast(double x) {
  const double R6 = 0.16666666666666666;  // RN(1/6) = 0x3FC5555555555555
  const double q = __dmul_rn(x, R6);
  const double r = __fma_rn(-6.0, q, x);
  return __fma_rn(r, R6, q);
}

// rounding of one pass: 0 none (float64), 1 sum rounded to float32 then float64 division (NumPy >= 2, first pass
// of float32 data), 2 sum rounded to float32 and float32 division (NumPy 1.x, every pass)
template <int RND, bool SAFE>
__device__ __forceinline__ double ss_finish(double v) {
  if (RND == 2) return (double)(__double2float_rn(v) / 6.0f);
  if (RND == 1) v = (double)__double2float_rn(v);
  return SAFE ? ss_div6_fast(v) : __ddiv_rn(v, 6.0);
}

struct SsParams {
  int nlat, nlon, ntime;
  int nan_border;        // rows set to NaN at either end after the passes (0: none)
  int nstrips, V;        // strips per row, valid columns per strip (V <= 64 - 2P)
  int nchunks, chunk_rows;  // latitude chunks per strip
  int flip_lat, flip_lon;
  // packed int16 input: value = double(v) * scale + offset; v == fill -> NaN (fill outside int16: none)
  double scale, offset;
  int fill;
  // comparison bit planes (NULL: not wanted): per (t, row, strip) PW = 2 + 2 * nlevels words:
  //   [0] NaN bits of the even columns (bit l = strip column 2l), [1] NaN bits of the odd columns,
  //   [2 + 2l], [3 + 2l] the same for "value > level l"
  u32* planes;
  int nlevels;
  LevelPack levels;
};

template <typename TIn>
__device__ __forceinline__ double ss_decode(TIn v, const SsParams& p) {
  return (double)v;
}
template <>
__device__ __forceinline__ double ss_decode<short>(short v, const SsParams& p) {
  if ((int)v == p.fill) return __longlong_as_double(0x7ff8000000000000LL);
  return __dadd_rn(__dmul_rn((double)v, p.scale), p.offset);
}

// is this raw input value outside the range ss_div6_fast is valid for?  Integer pipe only.  float32 data: finite
// and below 2^120 (the float32 rounding of the first pass overflows to Inf beyond 2^128); float64 data: exponent
// window [2^-480, 2^960], or zero; packed shorts: the fill value (decodes to NaN)
__device__ __forceinline__ bool ss_unsafe_raw(float a, float b, const SsParams&) {
  return max(__float_as_uint(a) & 0x7fffffffu, __float_as_uint(b) & 0x7fffffffu) > 0x7b800000u;
}
__device__ __forceinline__ bool ss_unsafe_raw(double a, double b, const SsParams&) {
  const u64 ba = (u64)__double_as_longlong(a) & 0x7fffffffffffffffULL, bb = (u64)__double_as_longlong(b) & 0x7fffffffffffffffULL;
  const u32 ha = (u32)(ba >> 32), hb = (u32)(bb >> 32);
  return (ba != 0 && (ha - 0x21f00000u) > (0x7bf00000u - 0x21f00000u)) ||
         (bb != 0 && (hb - 0x21f00000u) > (0x7bf00000u - 0x21f00000u));
}
__device__ __forceinline__ bool ss_unsafe_raw(short a, short b, const SsParams& p) {
  return (int)a == p.fill || (int)b == p.fill;  // finite scale / offset keep the decoded values in range
}

// One marching step of a strip of H halves of 64 columns (lane l of half h: strip columns 64h + 2l, 64h + 2l + 1).
// PH = k mod 3 selects which register row of every level is the oldest one (it is consumed for the last time in this
// step and then overwritten by the level's new row).  West / east neighbours: one rotation shuffle per half and
// direction; across the seam between two halves lane 0 / lane 31 take the value of the other half.
template <int P, int H, int RFIRST, int RREST, bool SAFE, int PH>
__device__ __forceinline__ void ss_step(double (&w)[P][3][2 * H], const double (&in)[2 * H], double (&o)[2 * H], int lane_up,
                                        int lane_dn, bool first_lane, bool last_lane) {
  constexpr int OLD = PH % 3, MID = (PH + 1) % 3, NEW = (PH + 2) % 3;
#pragma unroll
  for (int p = P; p >= 1; --p) {
    double wv[H], ev[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      if (H == 1) {
        // one half: shifts with an immediate distance (the strip's outer columns are halo, never valid)
        wv[h] = __shfl_up_sync(WBK_FULL, w[p - 1][MID][1], 1);
        ev[h] = __shfl_down_sync(WBK_FULL, w[p - 1][MID][0], 1);
      } else {
        wv[h] = __shfl_sync(WBK_FULL, w[p - 1][MID][2 * h + 1], lane_up);  // column 2l-1 of the half (lane 0: lane 31's)
        ev[h] = __shfl_sync(WBK_FULL, w[p - 1][MID][2 * h], lane_dn);      // column 2l+2 of the half (lane 31: lane 0's)
      }
    }
    if (H == 2) {
      // seam: column 63 is lane 31 of half 0, column 64 lane 0 of half 1 (the outer ends of the strip keep the
      // rotated garbage: halo columns, never valid)
      const double w1 = first_lane ? wv[0] : wv[1], e0 = last_lane ? ev[1] : ev[0];
      wv[H - 1] = w1;
      ev[0] = e0;
    }
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const double nx = w[p - 1][OLD][2 * h], ny = w[p - 1][OLD][2 * h + 1];
      const double cx = w[p - 1][MID][2 * h], cy = w[p - 1][MID][2 * h + 1];
      const double sx = w[p - 1][NEW][2 * h], sy = w[p - 1][NEW][2 * h + 1];
      // scipy's tap order ((((N + W) + 2C) + E) + S); (N+W) + 2C in one FMA is the same single rounding
      double a0 = __dadd_rn(nx, wv[h]);
      a0 = __fma_rn(2.0, cx, a0);
      a0 = __dadd_rn(a0, cy);
      a0 = __dadd_rn(a0, sx);
      double a1 = __dadd_rn(ny, cx);
      a1 = __fma_rn(2.0, cy, a1);
      a1 = __dadd_rn(a1, ev[h]);
      a1 = __dadd_rn(a1, sy);
      double r0, r1;
      if (p == 1) {
        r0 = ss_finish<RFIRST, SAFE>(a0);
        r1 = ss_finish<RFIRST, SAFE>(a1);
      } else {
        r0 = ss_finish<RREST, SAFE>(a0);
        r1 = ss_finish<RREST, SAFE>(a1);
      }
      if (p == P) {
        o[2 * h] = r0;
        o[2 * h + 1] = r1;
      } else {
        w[p][OLD][2 * h] = r0;  // level p + 1 has already consumed this row (levels run from P downwards)
        w[p][OLD][2 * h + 1] = r1;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 2 * H; ++c) w[0][OLD][c] = in[c];
}

// PL: 0 no bit planes, 1 planes for exactly one level (one 16-byte store per row and half), 2 planes for any level count
template <int P, int H, typename TIn, typename TOut, int RMODE, int PL>
__global__ void __launch_bounds__(SS_THREADS, (H == 2 ? 2 : (P <= 5 ? SS_MINCTA_SMALL : 3)))
smooth_stream_kernel(const TIn* __restrict__ in, TOut* __restrict__ out, const __grid_constant__ SsParams prm) {
  constexpr int RFIRST = RMODE == WBK_ROUND_ALL ? 2 : (RMODE == WBK_ROUND_FIRST ? 1 : 0);
  constexpr int RREST = RMODE == WBK_ROUND_ALL ? 2 : 0;
  constexpr int WS = 64 * H;  // strip width
  const int lane = wbk_lane();
  const int nlat = prm.nlat, nlon = prm.nlon;
  // the warp index through a shuffle: the compiler then treats everything derived from it as warp-uniform and
  // does not guard the collectives below with divergence checks
  const int warp_u = __shfl_sync(WBK_FULL, (int)(threadIdx.x >> 5), 0);
  const long long item = (long long)blockIdx.x * (SS_THREADS / 32) + warp_u;
  const long long nitems = (long long)prm.ntime * prm.nchunks * prm.nstrips;
  if (item >= nitems) return;  // whole warp
  const int s = (int)(item % prm.nstrips);
  const long long q = item / prm.nstrips;
  const int chunk = (int)(q % prm.nchunks);
  const int t = (int)(q / prm.nchunks);

  const int nb = prm.nan_border;
  const int row_lo = nb, row_hi = nlat - nb;  // rows that are computed (the rest is NaN)
  const int y0 = row_lo + chunk * prm.chunk_rows;
  const int y1 = min(y0 + prm.chunk_rows, row_hi);

  // columns of this lane: strip columns 64h + 2l, 64h + 2l + 1 = logical grid columns x0 + ..., periodic
  const int x0 = s * prm.V - P;
  const size_t plane = (size_t)nlat * nlon;
  const TIn* src = in + plane * t;
  TOut* dst = out + plane * t;
  // valid outputs of this lane: strip columns [P, P + V) that exist in the grid
  const int vcols = min(prm.V, nlon - s * prm.V);
  int pxa[H], d1[H];  // physical column of the lane's first column of half h, offset of its second column
  bool ok0[H], ok1[H];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    const int sc = 64 * h + 2 * lane;
    const int lx0 = ss_wrap(x0 + sc, nlon), lx1 = ss_wrap(x0 + sc + 1, nlon);
    const int p0 = prm.flip_lon ? nlon - 1 - lx0 : lx0, p1 = prm.flip_lon ? nlon - 1 - lx1 : lx1;
    pxa[h] = p0;
    d1[h] = p1 - p0;
    ok0[h] = sc >= P && sc < P + vcols;
    ok1[h] = sc + 1 >= P && sc + 1 < P + vcols;
  }
  const int ox = s * prm.V + 2 * lane - P;  // output column of strip column 2l (valid lanes only); half h: + 64h
  const int PW = 2 + 2 * prm.nlevels;
  // bit planes: one record of PW words per (row, 64-column half); half h of strip s is plane strip s * H + h
  const size_t pl_row = (size_t)prm.nstrips * H * (size_t)PW;  // plane words per grid row
  u32* pl_base = PL ? prm.planes + ((size_t)t * nlat * prm.nstrips * H + (size_t)s * H) * (size_t)PW : nullptr;
  const double level0 = prm.levels.v[0];
  const int lane_up = (lane + 31) & 31, lane_dn = (lane + 1) & 31;
  const bool first_lane = lane == 0, last_lane = lane == 31;

  // bit planes of one finished row: ballots of the values the lanes hold (NaN can only occur on the slow path)
  auto emit_planes = [&](u32* pl, const double (&v)[2 * H], bool may_nan) {
#pragma unroll
    for (int h = 0; h < H; ++h) {
      u32* ph = pl + h * PW;
      const double v0 = v[2 * h], v1 = v[2 * h + 1];
      u32 n0 = 0, n1 = 0;
      if (may_nan) {
        n0 = __ballot_sync(WBK_FULL, v0 != v0);
        n1 = __ballot_sync(WBK_FULL, v1 != v1);
      }
      if (PL == 1) {
        const u32 g0 = __ballot_sync(WBK_FULL, v0 > level0), g1 = __ballot_sync(WBK_FULL, v1 > level0);
        if (lane == 0) *reinterpret_cast<uint4*>(ph) = make_uint4(n0, n1, g0, g1);
      } else {
        if (lane == 0) {
          ph[0] = n0;
          ph[1] = n1;
        }
        for (int l = 0; l < prm.nlevels; ++l) {
          const double level = prm.levels.v[l];
          const u32 g0 = __ballot_sync(WBK_FULL, v0 > level), g1 = __ballot_sync(WBK_FULL, v1 > level);
          if (lane == 0) {
            ph[2 + 2 * l] = g0;
            ph[3 + 2 * l] = g1;
          }
        }
      }
    }
  };

  // ---- NaN border rows (spatial.py:106-107): written by the first / last chunk of the strip
  if (nb > 0) {
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    double vq[2 * H];
#pragma unroll
    for (int c = 0; c < 2 * H; ++c) vq[c] = qnan;
    for (int i = 0; i < 2 * nb; ++i) {
      const bool top = i < nb;
      const int row = top ? i : nlat - 2 * nb + i;
      if (top ? chunk != 0 : (chunk != prm.nchunks - 1 || row < nb)) continue;  // warp-uniform
#pragma unroll
      for (int h = 0; h < H; ++h) {
        if (ok0[h]) dst[(size_t)row * nlon + ox + 64 * h] = (TOut)qnan;
        if (ok1[h]) dst[(size_t)row * nlon + ox + 64 * h + 1] = (TOut)qnan;
      }
      if (PL) emit_planes(pl_base + (size_t)row * pl_row, vq, true);
    }
  }
  if (y0 >= y1) return;

  // ---- march: input rows k = y0 - P .. (rows past y1 + P - 1 are loaded too but never reach a valid output),
  //      level p emits row k - 2p, the output row of step k is k - 2P
  double w[P][3][2 * H];
#pragma unroll
  for (int p = 0; p < P; ++p)
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 2 * H; ++c) w[p][r][c] = 0.0;

  const int k_begin = y0 - P, k_end = y1 + 2 * P;  // steps k in [k_begin, k_end)
  // prefetch cursor: logical row gy_pf, physical row pointer of this lane's first column (half h: + dh[h])
  int gy_pf = ss_wrap(k_begin, nlat);
  const long long rstride = prm.flip_lat ? -(long long)nlon : (long long)nlon;
  const TIn* pfp = src + (long long)(prm.flip_lat ? nlat - 1 - gy_pf : gy_pf) * nlon + pxa[0];
  int dh[H];
#pragma unroll
  for (int h = 0; h < H; ++h) dh[h] = pxa[h] - pxa[0];
  const long long pf_wrap = rstride * nlat;  // back to logical row 0
  auto advance = [&]() {
    pfp += rstride;
    if (++gy_pf == nlat) {
      gy_pf = 0;
      pfp -= pf_wrap;
    }
  };
  // Input rows reach the registers through a per-warp ring of SS_RING rows in shared memory filled by asynchronous
  // copies (cp.async, 4 or 8 bytes per lane and column): SS_RING - 1 rows are in flight, no register is held for
  // them, and every lane reads back only what it copied itself (no barrier).  Packed shorts (2 bytes, below the
  // cp.async granularity) use plain loads two steps ahead instead.
  constexpr bool ASYNC = sizeof(TIn) >= 4;
  __shared__ __align__(16) TIn s_ring[ASYNC ? SS_THREADS / 32 : 1][ASYNC ? SS_RING : 1][WS];
  TIn* const my_ring = &s_ring[ASYNC ? warp_u : 0][0][ASYNC ? 2 * lane : 0];
  int slot_in = 0, slot_out = 0;  // ring slots of the next row to request / to consume
  TIn pf[3][2 * H];
  auto request_row = [&]() {
    if (ASYNC) {
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const TIn* g0 = pfp + dh[h];
#ifndef WBK_EMU
        const unsigned dst_s = (unsigned)__cvta_generic_to_shared(my_ring + slot_in * WS + 64 * h);
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst_s), "l"(g0), "n"(sizeof(TIn)) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst_s + (unsigned)sizeof(TIn)), "l"(g0 + d1[h]), "n"(sizeof(TIn)) : "memory");
#else
        my_ring[slot_in * WS + 64 * h] = g0[0];
        my_ring[slot_in * WS + 64 * h + 1] = g0[d1[h]];
#endif
      }
#ifndef WBK_EMU
      asm volatile("cp.async.commit_group;" ::: "memory");
#endif
      slot_in = (slot_in + 1) & (SS_RING - 1);
      advance();
    }
  };
  auto take_row = [&](TIn (&r)[2 * H]) {
#ifndef WBK_EMU
    asm volatile("cp.async.wait_group %0;" ::"n"(SS_RING - 1) : "memory");
#endif
#pragma unroll
    for (int h = 0; h < H; ++h) {
      r[2 * h] = my_ring[slot_out * WS + 64 * h];
      r[2 * h + 1] = my_ring[slot_out * WS + 64 * h + 1];
    }
    slot_out = (slot_out + 1) & (SS_RING - 1);
  };
  auto prefetch = [&](TIn (&slot)[2 * H]) {
#pragma unroll
    for (int h = 0; h < H; ++h) {
      slot[2 * h] = pfp[dh[h]];
      slot[2 * h + 1] = pfp[dh[h] + d1[h]];
    }
    advance();
  };
  if (ASYNC) {
#pragma unroll
    for (int i = 0; i < SS_RING - 1; ++i) request_row();
  } else {
    // two rows in flight: step k consumes slot k % 3 and refills slot (k + 2) % 3, the one the PREVIOUS step consumed
    // (refilling the slot that is being consumed makes the compiler copy the fresh value, i.e. wait for the load)
    prefetch(pf[0]);
    prefetch(pf[1]);
#pragma unroll
    for (int c = 0; c < 2 * H; ++c) pf[2][c] = (TIn)0;
  }

  int slow_until = k_begin;  // steps before this one use the plain division
  int k = k_begin;
  TOut* dp = dst + (long long)(k_begin - 2 * P) * nlon + ox;  // output row of step k (valid once k - 2P >= y0)
  u32* plp = PL ? pl_base + (long long)(k_begin - 2 * P) * (long long)pl_row : nullptr;
#define SS_ONE_STEP(PH)                                                                              \
  {                                                                                                  \
    TIn raw[2 * H];                                                                                  \
    if (ASYNC) {                                                                                     \
      request_row();                                                                                 \
      take_row(raw);                                                                                 \
    } else {                                                                                         \
      prefetch(pf[(PH + 2) % 3]);                                                                    \
      _Pragma("unroll") for (int c = 0; c < 2 * H; ++c) raw[c] = pf[PH][c];                          \
    }                                                                                                \
    double inv[2 * H], o[2 * H];                                                                     \
    bool bad = false;                                                                                \
    _Pragma("unroll") for (int h = 0; h < H; ++h) {                                                  \
      inv[2 * h] = ss_decode<TIn>(raw[2 * h], prm);                                                  \
      inv[2 * h + 1] = ss_decode<TIn>(raw[2 * h + 1], prm);                                          \
      bad = bad || ss_unsafe_raw(raw[2 * h], raw[2 * h + 1], prm);                                   \
    }                                                                                                \
    if (__any_sync(WBK_FULL, bad)) slow_until = k + 3 * P + 1;                                       \
    const bool slow = k < slow_until;                                                                \
    if (slow) ss_step<P, H, RFIRST, RREST, false, PH>(w, inv, o, lane_up, lane_dn, first_lane, last_lane); \
    else ss_step<P, H, RFIRST, RREST, true, PH>(w, inv, o, lane_up, lane_dn, first_lane, last_lane); \
    if (k - 2 * P >= y0) {                                                                           \
      _Pragma("unroll") for (int h = 0; h < H; ++h) {                                                \
        if (ok0[h]) dp[64 * h] = (TOut)o[2 * h];                                                     \
        if (ok1[h]) dp[64 * h + 1] = (TOut)o[2 * h + 1];                                             \
      }                                                                                              \
      if (PL) {                                                                                      \
        double ov[2 * H];                                                                            \
        _Pragma("unroll") for (int c = 0; c < 2 * H; ++c) ov[c] = (double)(TOut)o[c];                \
        emit_planes(plp, ov, slow);                                                                  \
      }                                                                                              \
    }                                                                                                \
    dp += nlon;                                                                                      \
    if (PL) plp += pl_row;                                                                           \
    ++k;                                                                                             \
  }

  // the register rotation has period 3: k_begin is mapped to phase 0
  while (k + 3 <= k_end) {
    SS_ONE_STEP(0)
    SS_ONE_STEP(1)
    SS_ONE_STEP(2)
  }
  if (k < k_end) {
    SS_ONE_STEP(0)
    if (k < k_end) SS_ONE_STEP(1)
  }
#ifndef WBK_EMU
  if (ASYNC) asm volatile("cp.async.wait_group 0;" ::: "memory");  // nothing of this warp is in flight at exit
#endif
#undef SS_ONE_STEP
}

// ------------------------------------------------------------------------------------------ launch
// Halves of 64 columns per warp strip.  Two halves cut the halo recompute from 64 / (64 - 2P) to 128 / (128 - 2P)
// and amortise the per-step overhead over twice the cells, but the 244 registers they need leave 8 warps per SM
// instead of 16: measured 1.43 ms against 1.31 ms per 296 steps of 721x1440 (tools/ab_halves.py), so
// one half is the default and two are opt-in through wbk_tune_smooth_halves().
int wbk_smooth_halves_override();  // wbk_smooth.cu: 0 = default
static int ss_halves(int nlon, int passes) {
  (void)nlon;
  return (passes <= SS_MAX_P_H2 && wbk_smooth_halves_override() == 2) ? 2 : 1;
}

static void ss_geometry(SsParams& p, int passes) {
  p.V = 64 * ss_halves(p.nlon, passes) - 2 * passes;
  p.nstrips = (p.nlon + p.V - 1) / p.V;
  const int rows = p.nlat - 2 * p.nan_border;
  // two latitude chunks per strip on tall grids: better balance over the SMs for 3P extra rows per chunk
  p.nchunks = rows >= 256 ? 2 : 1;
  p.chunk_rows = rows > 0 ? (rows + p.nchunks - 1) / p.nchunks : 1;
}

template <int P, int H, typename TIn, typename TOut, int RMODE>
static int ss_launch_ph(const void* in, void* out, SsParams& prm, cudaStream_t st) {
  ss_geometry(prm, P);
  // the kernel indexes items with 64 bits, the grid with 31: very long series go in chunks of time steps
  const long long per_step = (long long)prm.nchunks * prm.nstrips;
  const int wpb = SS_THREADS / 32;
  const int max_t = (int)(((1LL << 30) * wpb) / per_step) > 0 ? (int)(((1LL << 30) * wpb) / per_step) : 1;
  const int ntime = prm.ntime;
  const size_t plane = (size_t)prm.nlat * prm.nlon;
  for (int t0 = 0; t0 < ntime; t0 += max_t) {
    SsParams q = prm;
    q.ntime = ntime - t0 < max_t ? ntime - t0 : max_t;
    if (q.planes) q.planes += (size_t)t0 * prm.nlat * prm.nstrips * H * (size_t)(2 + 2 * prm.nlevels);
    const long long nitems = per_step * q.ntime;
    const int grid = (int)((nitems + wpb - 1) / wpb);
    bool launched = false;
    if constexpr (sizeof(TOut) == 8) {  // bit planes go with float64 output only
      if (prm.planes && prm.nlevels == 1) {
        WBK_LAUNCH(KID_SMOOTH, (smooth_stream_kernel<P, H, TIn, TOut, RMODE, 1>), dim3(grid), dim3(SS_THREADS), 0, st,
                   (const TIn*)in + plane * t0, (TOut*)out + plane * t0, q);
        launched = true;
      } else if (prm.planes) {
        WBK_LAUNCH(KID_SMOOTH, (smooth_stream_kernel<P, H, TIn, TOut, RMODE, 2>), dim3(grid), dim3(SS_THREADS), 0, st,
                   (const TIn*)in + plane * t0, (TOut*)out + plane * t0, q);
        launched = true;
      }
    } else if (prm.planes) {
      wbk_set_error("wbk_smooth: bit planes need float64 output");
      return WBK_ERR_INVALID;
    }
    if (!launched) {
      WBK_LAUNCH(KID_SMOOTH, (smooth_stream_kernel<P, H, TIn, TOut, RMODE, 0>), dim3(grid), dim3(SS_THREADS), 0, st,
                 (const TIn*)in + plane * t0, (TOut*)out + plane * t0, q);
    }
    WBK_LAUNCH_CHECK();
  }
  return WBK_OK;
}

template <int P, typename TIn, typename TOut, int RMODE>
static int ss_launch_p(const void* in, void* out, SsParams& prm, cudaStream_t st) {
  if constexpr (P <= SS_MAX_P_H2) {
    if (ss_halves(prm.nlon, P) == 2) return ss_launch_ph<P, 2, TIn, TOut, RMODE>(in, out, prm, st);
  }
  return ss_launch_ph<P, 1, TIn, TOut, RMODE>(in, out, prm, st);
}

template <typename TIn, typename TOut, int RMODE>
static int ss_launch(const void* in, void* out, int passes, SsParams& prm, cudaStream_t st) {
  switch (passes) {
    case 1: return ss_launch_p<1, TIn, TOut, RMODE>(in, out, prm, st);
    case 2: return ss_launch_p<2, TIn, TOut, RMODE>(in, out, prm, st);
    case 3: return ss_launch_p<3, TIn, TOut, RMODE>(in, out, prm, st);
    case 4: return ss_launch_p<4, TIn, TOut, RMODE>(in, out, prm, st);
    case 5: return ss_launch_p<5, TIn, TOut, RMODE>(in, out, prm, st);
    case 6: return ss_launch_p<6, TIn, TOut, RMODE>(in, out, prm, st);
    case 7: return ss_launch_p<7, TIn, TOut, RMODE>(in, out, prm, st);
    case 8: return ss_launch_p<8, TIn, TOut, RMODE>(in, out, prm, st);
  }
  wbk_set_error("wbk_smooth: internal pass count %d", passes);
  return WBK_ERR_INVALID;
}

