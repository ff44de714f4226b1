// Contour extraction: marching squares on the periodically extended grid, contour assembly in
// skimage's order, rounding + keep-first dedupe, length and periodic-duplicate filters.
// Reference: wavebreaking/indices/contour_index.py:86-194 (skimage.measure.find_contours inside).
#include "wbk_ctx.cuh"
#include "wbk_ms.cuh"

// ------------------------------------------------------------------------------------------ context
int wbk_smooth_plane_strips_max(int nlon);  // wbk_smooth.cu

namespace {
struct Bump {
  unsigned char* base;
  size_t off;
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

size_t layout(WbkDev& d, const wbk_caps& c, int nlat, int nlon, int add, unsigned char* base) {
  Bump b{base, 0};
  d.nlat = nlat; d.nlon = nlon; d.add = add; d.W = nlon + add;
  d.S = c.seg_cap; d.CC = c.contour_cap; d.R = c.seg_cap + c.contour_cap;
  d.HC = (int)wbk_pow2_ceil((u32)(2 * d.R));
  d.CCp2 = (int)wbk_pow2_ceil((u32)c.contour_cap);
  d.sel_cap = c.sel_cap; d.pair_cap = c.pair_cap; d.event_cap = c.event_cap; d.max_jobs = c.max_jobs;
  const size_t J = c.max_jobs, S = d.S, CC = d.CC, R = d.R, HC = d.HC;
  d.rid = b.take<u32>(J * S); d.fpid = b.take<u32>(J * S); d.tpid = b.take<u32>(J * S);
  d.fxy = b.take<u32>(J * S); d.txy = b.take<u32>(J * S);
  d.seg_count = b.take<int>(J); d.status = b.take<int>(J);
  d.nxt = b.take<u32>(J * S); d.prv = b.take<u32>(J * S); d.w64 = b.take<u64>(J * S);
  d.cminrid = b.take<u32>(J * S); d.cnseg = b.take<u32>(J * S); d.cidx = b.take<u32>(J * S); d.cflag = b.take<u32>(J * S);
  d.hkeys = b.take<u64>(J * HC); d.hvals = b.take<u32>(J * HC);
  d.raw = b.take<u32>(J * R); d.rawc = b.take<u32>(J * R); d.slot = b.take<u32>(J * R); d.scan = b.take<int>(J * R);
  d.sortkeys = b.take<u64>(J * d.CCp2);
  d.craw_off = b.take<int>(J * (CC + 1)); d.craw_n = b.take<int>(J * CC); d.cclosed = b.take<int>(J * CC);
  d.cnuniq = b.take<int>(J * CC); d.cdrop = b.take<int>(J * CC); d.cymin = b.take<int>(J * CC);
  d.cymax = b.take<int>(J * CC); d.cout = b.take<int>(J * CC); d.cand = b.take<int>(J * CC * 2);
  d.out_pts = b.take<u32>(J * R); d.out_tab = b.take<int>(J * CC * 4); d.out_sumy = b.take<int>(J * CC);
  d.out_nc = b.take<int>(J); d.out_np = b.take<int>(J); d.max_nx = b.take<int>(64);
  // comparison bit planes of the fused smoothing: <= 4 words per (job, row, plane strip)
  d.planes = b.take<u32>(J * 4 * (size_t)nlat * (size_t)wbk_smooth_plane_strips_max(nlon) + 8);
  return (b.off + 255) & ~(size_t)255;
}

bool caps_ok(const wbk_caps* c, int nlat, int nlon, int add) {
  return c && c->max_jobs >= 1 && c->seg_cap >= 16 && c->contour_cap >= 4 && c->sel_cap >= 1 && c->pair_cap >= 16 &&
         c->event_cap >= 1 && nlat >= 2 && nlon >= 2 && add >= 0 && nlon + add < 65536 && nlat < 65536 &&
         (long long)c->max_jobs * c->sel_cap <= 16384 && c->seg_cap + c->contour_cap <= 65536;
}
}  // namespace

size_t wbk_index_layout(struct wbk_ctx* ctx, unsigned char* base, size_t off);  // wbk_indices.cu

extern "C" size_t wbk_workspace_bytes(const wbk_caps* caps, int nlat, int nlon, int add) {
  if (!caps_ok(caps, nlat, nlon, add)) return 0;
  wbk_ctx tmp;
  tmp.caps = *caps;
  size_t n = layout(tmp.d, *caps, nlat, nlon, add, nullptr);
  return wbk_index_layout(&tmp, nullptr, n);
}

extern "C" int wbk_create(wbk_ctx** out, const wbk_caps* caps, int nlat, int nlon, int add, void* d_workspace,
                          size_t workspace_bytes) {
  if (!out || !caps_ok(caps, nlat, nlon, add)) {
    wbk_set_error("wbk_create: invalid capacities or grid");
    return WBK_ERR_INVALID;
  }
  size_t need = wbk_workspace_bytes(caps, nlat, nlon, add);
  wbk_ctx* ctx = new wbk_ctx();
  ctx->caps = *caps;
  ctx->owned = nullptr;
  ctx->njobs = 0;
  ctx->nlevels = 0;
  if (!d_workspace) {
    if (cudaMalloc(&ctx->owned, need) != cudaSuccess) {
      delete ctx;
      wbk_set_error("wbk_create: cudaMalloc of %zu bytes failed", need);
      return WBK_ERR_CUDA;
    }
    d_workspace = ctx->owned;
  } else if (workspace_bytes < need || ((uintptr_t)d_workspace & 255)) {
    delete ctx;
    wbk_set_error("wbk_create: workspace too small (%zu < %zu) or not 256-byte aligned", workspace_bytes, need);
    return WBK_ERR_INVALID;
  }
  size_t n = layout(ctx->d, *caps, nlat, nlon, add, (unsigned char*)d_workspace);
  wbk_index_layout(ctx, (unsigned char*)d_workspace, n);
  *out = ctx;
  return WBK_OK;
}

extern "C" int wbk_destroy(wbk_ctx* ctx) {
  if (!ctx) return WBK_OK;
  if (ctx->owned) cudaFree(ctx->owned);
  delete ctx;
  return WBK_OK;
}

// ------------------------------------------------------------------------------------------ K3
// One thread per base column, marching down a strip of rows.  Every grid value is loaded once per
// strip (the right neighbour comes from a warp shuffle); squares of the periodic extension
// (columns >= nlon) are emitted by the thread that owns the matching base column, so the extension
// costs no extra reads.  Segments are appended to the job's arena with warp-aggregated atomics.
// Emission of the stand-alone kernel: the hits of one warp strip are compacted, every lane then handles ONE hit;
// the four corner values are re-read from global memory (L1 / L2 hits: the strip has just been loaded).
template <typename T>
__device__ __noinline__ void ms_emit_hits(const WbkDev& d, const T* __restrict__ src, const u32* __restrict__ masks,
                                          int nrows, int job, int r_begin, int col_base, double level) {
  const int lane = wbk_lane();
  const int nlon = d.nlon;
  int total = 0;
  for (int i = 0; i < nrows; ++i) total += __popc(masks[i]);
  for (int h0 = 0; h0 < total; h0 += 32) {
    const int h = h0 + lane;
    int r0 = 0, c0 = 0, mi = 0, sl = 0;
    double ul = 0, ur = 0, ll = 0, lr = 0;
    const bool active = h < total && ms_locate_hit(masks, nrows, h, mi, sl);
    if (active) {
      r0 = r_begin + mi;
      c0 = col_base + sl;
      const int cr = c0 + 1 == nlon ? 0 : c0 + 1;
      ul = (double)src[(size_t)r0 * nlon + c0];
      ur = (double)src[(size_t)r0 * nlon + cr];
      ll = (double)src[(size_t)(r0 + 1) * nlon + c0];
      lr = (double)src[(size_t)(r0 + 1) * nlon + cr];
    }
    ms_emit_squares(d, job, active, r0, c0, ul, ur, ll, lr, level);
  }
}

template <typename T>
__global__ void __launch_bounds__(MS_THREADS, MS_MIN_CTAS)
ms_segments_kernel(const T* __restrict__ field, const __grid_constant__ WbkDev d, const __grid_constant__ LevelPack levels,
                   int nlevels) {
  __shared__ u32 smask[MS_THREADS / 32][WBK_MAX_LEVELS][MS_ROWS];
  const int nlat = d.nlat, nlon = d.nlon, W = d.W;
  // every warp covers 31 base columns; lane 31 only supplies the right neighbour of lane 30 (so no thread
  // needs a second, uncoalesced load) -- column nlon wraps to column 0 (periodic extension)
  const int lane = wbk_lane(), warp = wbk_warp();
  const int col_base = (blockIdx.x * (MS_THREADS / 32) + warp) * 31;
  const int c0 = col_base + lane;
  const bool loads = c0 <= nlon;
  const int csrc = c0 == nlon ? 0 : c0;
  const bool valid = c0 < nlon && lane < 31;
  const int r_begin = blockIdx.y * MS_ROWS;
  const int r_end = min(r_begin + MS_ROWS, nlat - 1);  // squares r0 in [r_begin, r_end)
  const int t = blockIdx.z;
  const T* src = field + (size_t)t * nlat * nlon;
  const bool own_base = valid && (c0 <= W - 2);  // base square (r0, c0) exists

  // all rows of the strip are requested before any is used (MS_ROWS + 1 independent loads in flight)
  T vals[MS_ROWS + 1];
  {
    const T* p = src + (size_t)r_begin * nlon + csrc;
    const int nrows_ld = r_end - r_begin + 1;  // block-uniform
#pragma unroll
    for (int i = 0; i <= MS_ROWS; ++i) {
      vals[i] = (loads && i < nrows_ld) ? *p : (T)0;
      p += nlon;
    }
  }
  // lanes whose base square (r0, c0)-(r0+1, c0+1) exists; bit c of a row mask belongs to column col_base + c
  const u32 sqmask = __ballot_sync(WBK_FULL, own_base);
  u32 level_hits = 0;  // bit l: this warp strip holds contour squares of level l (warp-uniform)
  for (int l = 0; l < nlevels; ++l) {
    const double level = levels.v[l];
    const T tl = (T)level;
    const bool exact_level = (double)tl == level;  // compare in T when the level is representable (always for f64)
    // one ballot per row turns the comparisons into 32-column bit rows (uniform across the warp); lane i keeps
    // the rows of strip row i, fetches row i + 1 from its neighbour and classifies the 31 squares of that row with a
    // handful of bit operations: a contour square has corner bits that are neither all set nor all clear, no NaN
    u32 myg = 0, myn = 0;
#pragma unroll
    for (int i = 0; i <= MS_ROWS; ++i) {
      const T v = vals[i];
      const u32 g = __ballot_sync(WBK_FULL, exact_level ? v > tl : (double)v > level);
      const u32 n = __ballot_sync(WBK_FULL, v != v);
      if (lane == i) {
        myg = g;
        myn = n;
      }
    }
    const u32 gl = __shfl_down_sync(WBK_FULL, myg, 1), nl = __shfl_down_sync(WBK_FULL, myn, 1);
    const u32 both = myg & gl, either = myg | gl, nn = myn | nl;
    u32 hits = ((either | (either >> 1)) & ~(both & (both >> 1))) & ~(nn | (nn >> 1)) & sqmask;
    if (lane >= MS_ROWS || r_begin + lane >= r_end) hits = 0;
    if (lane < MS_ROWS) smask[warp][l][lane] = hits;
    if (__ballot_sync(WBK_FULL, hits != 0)) level_hits |= 1u << l;
  }
  // emission after the scan (the strip values are dead by now: no registers live across the call)
  if (level_hits) {
    __syncwarp();
    for (int l = 0; l < nlevels; ++l)
      if ((level_hits >> l) & 1u)
        ms_emit_hits<T>(d, src, smask[warp][l], MS_ROWS, t * nlevels + l, r_begin, col_base, levels.v[l]);
  }
}

// ------------------------------------------------------------------------------------------ K4
// One CTA per job: link segments by edge identity, find the skimage contour order (R1: by the
// smallest raster id), start points (R2: open chains start where nothing comes in; R3: rings start
// at the `to` point of their largest raster id), emit rounded points, keep-first dedupe, >= 4 filter,
// periodic duplicate filter, per-contour nx / sum(y).
__global__ void __launch_bounds__(WBK_CONTOUR_THREADS) contour_link_kernel(WbkDev d, int njobs) {
  const int job = blockIdx.x;
  if (job >= njobs) return;
  const int tid = threadIdx.x, nt = blockDim.x;
  __shared__ int sscan[40];
  __shared__ int s_nheads, s_ncand, s_bad;

  const int S = d.S, CC = d.CC, R = d.R;
  const u32* rid = d.rid + (size_t)job * S;
  const u32* fpid = d.fpid + (size_t)job * S;
  const u32* tpid = d.tpid + (size_t)job * S;
  const u32* fxy = d.fxy + (size_t)job * S;
  const u32* txy = d.txy + (size_t)job * S;
  u32* nxt = d.nxt + (size_t)job * S;
  u32* prv = d.prv + (size_t)job * S;
  u64* w64 = d.w64 + (size_t)job * S;
  u32* cminrid = d.cminrid + (size_t)job * S;
  u32* cnseg = d.cnseg + (size_t)job * S;
  u32* cidx = d.cidx + (size_t)job * S;
  u32* cflag = d.cflag + (size_t)job * S;
  u64* hkeys = d.hkeys + (size_t)job * d.HC;
  u32* hvals = d.hvals + (size_t)job * d.HC;
  u32* raw = d.raw + (size_t)job * R;
  u32* rawc = d.rawc + (size_t)job * R;
  u32* slot = d.slot + (size_t)job * R;
  int* scan = d.scan + (size_t)job * R;
  u64* sortkeys = d.sortkeys + (size_t)job * d.CCp2;
  int* craw_off = d.craw_off + (size_t)job * (CC + 1);
  int* craw_n = d.craw_n + (size_t)job * CC;
  int* cclosed = d.cclosed + (size_t)job * CC;
  int* cnuniq = d.cnuniq + (size_t)job * CC;
  int* cdrop = d.cdrop + (size_t)job * CC;
  int* cymin = d.cymin + (size_t)job * CC;
  int* cymax = d.cymax + (size_t)job * CC;
  int* cout = d.cout + (size_t)job * CC;
  int* cand = d.cand + (size_t)job * CC * 2;
  u32* out_pts = d.out_pts + (size_t)job * R;
  int* out_tab = d.out_tab + (size_t)job * CC * 4;
  int* out_sumy = d.out_sumy + (size_t)job * CC;

  int n = d.seg_count[job];
  if (tid == 0) {
    s_nheads = 0;
    s_ncand = 0;
    s_bad = 0;
    d.out_nc[job] = 0;
    d.out_np[job] = 0;
  }
  __syncthreads();
  if (n > S) {
    if (tid == 0) atomicOr(&d.status[job], (int)WBK_ST_SEG_OVERFLOW);
    return;
  }
  if (n == 0) return;

  int nh = 0, nraw = 0;
  const bool degenerate = (d.status[job] & WBK_ST_LATTICE_VERTEX) != 0;
  if (!degenerate) {
    // ---- P1: hash  from-edge -> segment
    u32* k32 = reinterpret_cast<u32*>(hkeys);
    const u32 hcap = wbk_pow2_ceil((u32)(2 * n) < 64u ? 64u : (u32)(2 * n));
    u32* v32 = k32 + hcap;  // hkeys region holds 2*HC u32 >= 2*hcap
    for (u32 i = tid; i < hcap; i += nt) k32[i] = WBK_NONE;
    for (int s = tid; s < n; s += nt) {
      prv[s] = WBK_NONE;
      cminrid[s] = WBK_NONE;
      cnseg[s] = 0;
      cflag[s] = 0;
    }
    __syncthreads();
    for (int s = tid; s < n; s += nt) wbk_hash_insert32(k32, v32, hcap, fpid[s], (u32)s);
    __syncthreads();
    // ---- P2: links
    for (int s = tid; s < n; s += nt) {
      u32 nx = wbk_hash_find32(k32, v32, hcap, tpid[s]);
      nxt[s] = nx;
      if (nx != WBK_NONE) prv[nx] = (u32)s;
    }
    __syncthreads();
    // ---- P3: pointer doubling with max(rid): detects rings and their largest raster id
    for (int s = tid; s < n; s += nt) w64[s] = ((u64)rid[s] << 32) | (u64)nxt[s];
    __syncthreads();
    const int rounds = wbk_log2_ceil((u32)n) + 1;
    for (int it = 0; it < rounds; ++it) {
      for (int s = tid; s < n; s += nt) {
        u64 w = w64[s];
        u32 p = (u32)w;
        if (p != WBK_NONE) {
          u64 w2 = w64[p];
          u32 v = (u32)(w >> 32), v2 = (u32)(w2 >> 32);
          w64[s] = ((u64)(v > v2 ? v : v2) << 32) | (u64)(u32)w2;
        }
      }
      __syncthreads();
    }
    // ---- P4: cut every ring after its largest-id segment
    for (int s = tid; s < n; s += nt) {
      u64 w = w64[s];
      if ((u32)w != WBK_NONE && (u32)(w >> 32) == rid[s]) {
        u32 h = nxt[s];
        cflag[h] = 1;  // ring: closed contour starting at h
        prv[h] = WBK_NONE;
        // nxt[s] stays (not used below)
      }
    }
    __syncthreads();
    // ---- P5: rank from the head by pointer jumping along prv
    for (int s = tid; s < n; s += nt) {
      u32 p = prv[s];
      w64[s] = (p == WBK_NONE) ? (u64)(u32)s : (((u64)1 << 32) | (u64)p);
    }
    __syncthreads();
    while (true) {
      int changed = 0;
      for (int s = tid; s < n; s += nt) {
        u64 w = w64[s];
        u32 p = (u32)w;
        if (p != (u32)s) {
          u64 w2 = w64[p];
          u32 p2 = (u32)w2;
          if (p2 != p) {
            w64[s] = ((((w >> 32) + (w2 >> 32)) << 32) | (u64)p2);
            changed = 1;
          }
        }
      }
      if (!__syncthreads_or(changed)) break;
    }
    // ---- P6: per-contour statistics keyed by the head segment
    for (int s = tid; s < n; s += nt) {
      u32 h = (u32)w64[s];
      atomicMin(&cminrid[h], rid[s]);
      atomicAdd(&cnseg[h], 1u);
      if (h == (u32)s) {
        int k = atomicAdd(&s_nheads, 1);
        if (k < CC) sortkeys[k] = (u64)s;  // temporarily the head id
      }
    }
    __syncthreads();
    nh = s_nheads;
    if (nh > CC) {
      if (tid == 0) atomicOr(&d.status[job], (int)WBK_ST_CONTOUR_OVERFLOW);
      return;
    }
    // ---- P7: contour order = ascending smallest raster id
    const u32 np2 = wbk_pow2_ceil((u32)nh);
    for (u32 i = tid; i < np2; i += nt) {
      if ((int)i < nh) {
        u32 h = (u32)sortkeys[i];
        sortkeys[i] = ((u64)cminrid[h] << 32) | (u64)h;
      } else {
        sortkeys[i] = ~0ull;
      }
    }
    __syncthreads();
    wbk_block_bitonic_sort(sortkeys, np2);
    for (int k = tid; k < nh; k += nt) {
      u32 h = (u32)sortkeys[k];
      cidx[h] = (u32)k;
      craw_n[k] = (int)cnseg[h] + 1;
      craw_off[k] = (int)cnseg[h] + 1;
      cclosed[k] = (int)cflag[h];
      cnuniq[k] = 0;
      cdrop[k] = 0;
      cymin[k] = 0x7fffffff;
      cymax[k] = -1;
    }
    __syncthreads();
    nraw = wbk_block_excl_scan(craw_off, nh, sscan);  // == n + nh
    if (tid == 0) craw_off[nh] = nraw;
    // ---- P8: raw (rounded) points in contour order
    for (int s = tid; s < n; s += nt) {
      u64 w = w64[s];
      u32 h = (u32)w;
      int k = (int)cidx[h];
      int pos = craw_off[k] + (int)(w >> 32) + 1;
      raw[pos] = txy[s];
      rawc[pos] = (u32)k;
      if (h == (u32)s) {
        raw[craw_off[k]] = fxy[s];
        rawc[craw_off[k]] = (u32)k;
      }
    }

  } else {
    // ------------------------------------------------------------------------------------------------
    // Sequential linker for jobs in which a contour vertex coincides with a grid vertex (a field value equal
    // to the level): skimage's _assemble_contours joins segments through dicts keyed by the float point, so
    // several segments can meet in one point and the outcome depends on the raster order.  The same
    // dict / deque procedure is replayed here by one thread on canonical point ids (vertex id or edge id).
    // ------------------------------------------------------------------------------------------------
    const u32 np2s = wbk_pow2_ceil((u32)n);
    if (4 * n > S || (int)np2s > S) {  // tables are sized for n <= S / 4: ask for a larger arena
      if (tid == 0) atomicOr(&d.status[job], (int)WBK_ST_SEG_OVERFLOW);
      return;
    }
    for (u32 i = tid; i < np2s; i += nt) w64[i] = (int)i < n ? (((u64)rid[i] << 32) | (u64)i) : ~0ull;
    __syncthreads();
    wbk_block_bitonic_sort(w64, np2s);
    const u32 tcap = wbk_pow2_ceil((u32)(4 * n) < 64u ? 64u : (u32)(4 * n));  // <= HC / 2
    u32* st_k = reinterpret_cast<u32*>(hkeys);
    u32* st_v = st_k + tcap;
    u32* en_k = st_v + tcap;
    u32* en_v = en_k + tcap;
    for (u32 i = tid; i < tcap; i += nt) {
      st_k[i] = WBK_NONE;
      en_k[i] = WBK_NONE;
    }
    __syncthreads();
    u32* nd_next = nxt;      // nodes (points of the partial contours): <= 2n
    u32* nd_prev = prv;
    u32* nd_xy = cminrid;
    u32* nd_pid = cnseg;
    u32* c_head = cidx;      // partial contours in creation order: <= n
    u32* c_tail = cflag;
    u32* c_alive = slot;
    int* c_cnt = scan;
    if (tid == 0) {
      const u32 TOMB = 0xfffffffeu;
      auto tpop = [&](u32* keys, u32* vals, u32 key) -> u32 {
        u32 h = wbk_hash32(key) & (tcap - 1);
        while (true) {
          const u32 k = keys[h];
          if (k == key) {
            keys[h] = TOMB;
            return vals[h];
          }
          if (k == WBK_NONE) return WBK_NONE;
          h = (h + 1) & (tcap - 1);
        }
      };
      auto tput = [&](u32* keys, u32* vals, u32 key, u32 val) {
        u32 h = wbk_hash32(key) & (tcap - 1);
        u32 first_free = WBK_NONE;
        while (true) {
          const u32 k = keys[h];
          if (k == key) {
            vals[h] = val;
            return;
          }
          if (k == TOMB && first_free == WBK_NONE) first_free = h;
          if (k == WBK_NONE) {
            if (first_free == WBK_NONE) first_free = h;
            keys[first_free] = key;
            vals[first_free] = val;
            return;
          }
          h = (h + 1) & (tcap - 1);
        }
      };
      u32 nn = 0, nc = 0;
      auto new_node = [&](u32 pid, u32 xy) -> u32 {
        nd_pid[nn] = pid;
        nd_xy[nn] = xy;
        nd_next[nn] = WBK_NONE;
        nd_prev[nn] = WBK_NONE;
        return nn++;
      };
      for (int q = 0; q < n; ++q) {
        const int s2 = (int)(u32)w64[q];
        const u32 f = fpid[s2], t = tpid[s2];
        if (f == t) continue;  // degenerate segment
        const u32 tail = tpop(st_k, st_v, t);
        const u32 head = tpop(en_k, en_v, f);
        if (tail != WBK_NONE && head != WBK_NONE) {
          if (tail == head) {  // close the ring: add the end point
            const u32 nd = new_node(t, txy[s2]);
            nd_prev[nd] = c_tail[head];
            nd_next[c_tail[head]] = nd;
            c_tail[head] = nd;
            c_cnt[head] += 1;
          } else {
            // joined sequence = head ++ tail; the contour that was created first survives
            const u32 keep = tail > head ? head : tail, dead = tail > head ? tail : head;
            const u32 hh = c_head[head], ht = c_tail[head], th = c_head[tail], tt = c_tail[tail];
            nd_next[ht] = th;
            nd_prev[th] = ht;
            c_head[keep] = hh;
            c_tail[keep] = tt;
            c_cnt[keep] = c_cnt[head] + c_cnt[tail];
            c_alive[dead] = 0;
            if (keep == tail) tpop(st_k, st_v, nd_pid[hh]);
            tput(st_k, st_v, nd_pid[hh], keep);
            tput(en_k, en_v, nd_pid[tt], keep);
          }
        } else if (tail == WBK_NONE && head == WBK_NONE) {
          const u32 a0 = new_node(f, fxy[s2]), a1 = new_node(t, txy[s2]);
          nd_next[a0] = a1;
          nd_prev[a1] = a0;
          c_head[nc] = a0;
          c_tail[nc] = a1;
          c_alive[nc] = 1;
          c_cnt[nc] = 2;
          tput(st_k, st_v, f, nc);
          tput(en_k, en_v, t, nc);
          ++nc;
        } else if (head == WBK_NONE) {  // prepend the from-point to `tail`
          const u32 nd = new_node(f, fxy[s2]);
          nd_next[nd] = c_head[tail];
          nd_prev[c_head[tail]] = nd;
          c_head[tail] = nd;
          c_cnt[tail] += 1;
          tput(st_k, st_v, f, tail);
        } else {  // append the to-point to `head`
          const u32 nd = new_node(t, txy[s2]);
          nd_prev[nd] = c_tail[head];
          nd_next[c_tail[head]] = nd;
          c_tail[head] = nd;
          c_cnt[head] += 1;
          tput(en_k, en_v, t, head);
        }
      }
      // surviving contours in creation order
      int k = 0, off = 0;
      for (u32 c2 = 0; c2 < nc; ++c2) {
        if (!c_alive[c2]) continue;
        if (k < CC) {
          sortkeys[k] = (u64)c2;
          craw_off[k] = off;
          craw_n[k] = c_cnt[c2];
          cclosed[k] = nd_pid[c_head[c2]] == nd_pid[c_tail[c2]] ? 1 : 0;
          cnuniq[k] = 0;
          cdrop[k] = 0;
          cymin[k] = 0x7fffffff;
          cymax[k] = -1;
        }
        off += c_cnt[c2];
        ++k;
      }
      s_nheads = k;
      s_ncand = off;  // borrowed: total raw points (reset below)
    }
    __syncthreads();
    nh = s_nheads;
    nraw = s_ncand;
    __syncthreads();
    if (tid == 0) s_ncand = 0;
    if (nh > CC || nraw > R) {
      if (tid == 0) atomicOr(&d.status[job], (int)WBK_ST_CONTOUR_OVERFLOW);
      return;
    }
    if (tid == 0) craw_off[nh] = nraw;
    // one thread per contour walks its list
    for (int k = tid; k < nh; k += nt) {
      u32 nd = c_head[(u32)sortkeys[k]];
      int pos = craw_off[k];
      while (nd != WBK_NONE) {
        raw[pos] = nd_xy[nd];
        rawc[pos] = (u32)k;
        ++pos;
        nd = nd_next[nd];
      }
    }
    __syncthreads();
  }
  // ---- P9: keep-first dedupe of the rounded points inside every contour
  const u32 dcap = wbk_pow2_ceil((u32)(2 * nraw) < 64u ? 64u : (u32)(2 * nraw));
  for (u32 i = tid; i < dcap; i += nt) {
    hkeys[i] = ~0ull;
    hvals[i] = WBK_NONE;
  }
  __syncthreads();
  for (int i = tid; i < nraw; i += nt) {
    u64 key = ((u64)rawc[i] << 32) | (u64)raw[i];
    u32 sl = wbk_hash_slot64(hkeys, dcap, key, nullptr);
    slot[i] = sl;
    atomicMin(&hvals[sl], (u32)i);
  }
  __syncthreads();
  for (int i = tid; i < nraw; i += nt) {
    int keep = hvals[slot[i]] == (u32)i ? 1 : 0;
    scan[i] = keep;
    if (keep) {
      int k = (int)rawc[i];
      atomicAdd(&cnuniq[k], 1);
      int y = wbk_py(raw[i]);
      atomicMin(&cymin[k], y);
      atomicMax(&cymax[k], y);
    }
  }
  __syncthreads();
  // ---- P11: periodic duplicates (contour_index.py:119-140) among contours with >= 4 points:
  //      e1 (folded with x % nlon) subset of e2  ->  drop the shorter one (equal length: the later one)
  for (u32 i = tid; i < dcap; i += nt) hkeys[i] = ~0ull;
  __syncthreads();
  const int nlon = d.nlon;
  for (int i = tid; i < nraw; i += nt) {
    int k = (int)rawc[i];
    if (scan[i] && cnuniq[k] >= 4) {
      u32 p = raw[i];
      u64 key = ((u64)(u32)k << 32) | (u64)wbk_pack_xy(wbk_px(p) % nlon, wbk_py(p));
      wbk_hash_slot64(hkeys, dcap, key, nullptr);
    }
  }
  __syncthreads();
  for (long long q = tid; q < (long long)nh * nh; q += nt) {
    int k1 = (int)(q / nh), k2 = (int)(q % nh);
    if (k1 == k2 || cnuniq[k1] < 4 || cnuniq[k2] < 4) continue;
    if (cymin[k1] < cymin[k2] || cymax[k1] > cymax[k2]) continue;
    u32 p = raw[craw_off[k1]];  // first point of k1 (always kept)
    u64 key = ((u64)(u32)k2 << 32) | (u64)wbk_pack_xy(wbk_px(p) % nlon, wbk_py(p));
    if (wbk_hash_lookup64(hkeys, dcap, key) == WBK_NONE) continue;
    int c = atomicAdd(&s_ncand, 1);
    if (c < CC) {
      cand[2 * c] = k1;
      cand[2 * c + 1] = k2;
    }
  }
  __syncthreads();
  int ncand = s_ncand;
  if (ncand > CC) {
    if (tid == 0) atomicOr(&d.status[job], (int)WBK_ST_CONTOUR_OVERFLOW);
    return;
  }
  {
    const int lane = wbk_lane(), warp = wbk_warp(), nwarps = nt >> 5;
    for (int c = warp; c < ncand; c += nwarps) {
      const int k1 = cand[2 * c], k2 = cand[2 * c + 1];
      const int b = craw_off[k1], e = b + craw_n[k1];
      int ok = 1;
      for (int i0 = b; i0 < e; i0 += 32) {
        int i = i0 + lane;
        int good = 1;
        if (i < e && scan[i]) {
          u32 p = raw[i];
          u64 key = ((u64)(u32)k2 << 32) | (u64)wbk_pack_xy(wbk_px(p) % nlon, wbk_py(p));
          good = wbk_hash_lookup64(hkeys, dcap, key) != WBK_NONE;
        }
        if (!__all_sync(WBK_FULL, good)) {
          ok = 0;
          break;
        }
      }
      if (ok && lane == 0) {
        int l1 = cnuniq[k1], l2 = cnuniq[k2];
        int drop = (l1 == l2) ? (k1 > k2 ? k1 : k2) : (l1 < l2 ? k1 : k2);
        cdrop[drop] = 1;
      }
    }
  }
  __syncthreads();
  // ---- P12: compaction of contours and points
  for (int k = tid; k < nh; k += nt) cout[k] = (cnuniq[k] >= 4 && !cdrop[k]) ? 1 : 0;
  __syncthreads();
  for (int i = tid; i < nraw; i += nt) scan[i] = (scan[i] && cout[rawc[i]]) ? 1 : 0;
  __syncthreads();
  // remember the flags in `slot` (scan is overwritten by the prefix sum)
  for (int i = tid; i < nraw; i += nt) slot[i] = (u32)scan[i];
  __syncthreads();
  const int nout_pts = wbk_block_excl_scan(scan, nraw, sscan);
  // contour output index
  for (int k = tid; k < nh; k += nt) cymin[k] = cout[k];  // reuse cymin as scan buffer
  __syncthreads();
  const int nout_c = wbk_block_excl_scan(cymin, nh, sscan);
  for (int k = tid; k < nh; k += nt) {
    if (cout[k]) {
      int o = cymin[k];
      out_tab[4 * o + 0] = scan[craw_off[k]];
      out_tab[4 * o + 1] = cnuniq[k];
      out_tab[4 * o + 2] = cclosed[k];
      out_tab[4 * o + 3] = 0;
      out_sumy[o] = 0;
    }
  }
  // reset the hash for the distinct-column count
  for (u32 i = tid; i < dcap; i += nt) hkeys[i] = ~0ull;
  __syncthreads();
  for (int i = tid; i < nraw; i += nt) {
    if (slot[i]) {
      u32 p = raw[i];
      out_pts[scan[i]] = p;
      int o = cymin[rawc[i]];
      atomicAdd(&out_sumy[o], wbk_py(p));
      bool is_new;
      wbk_hash_slot64(hkeys, dcap, ((u64)(u32)o << 32) | (u64)(u32)wbk_px(p), &is_new);
      if (is_new) atomicAdd(&out_tab[4 * o + 3], 1);
    }
  }
  __syncthreads();
  int mx = 0;
  for (int o = tid; o < nout_c; o += nt) mx = max(mx, out_tab[4 * o + 3]);
  mx = wbk_warp_max(mx);
  if (wbk_lane() == 0 && mx > 0) atomicMax(d.max_nx, mx);
  if (tid == 0) {
    d.out_nc[job] = nout_c;
    d.out_np[job] = nout_pts;
  }
}

// ------------------------------------------------------------------------------------------ host API
extern "C" int wbk_contours(wbk_ctx* ctx, const void* d_field, int dtype, int ntime, const double* h_levels,
                            int nlevels, void* stream) {
  if (!ctx || !d_field || !h_levels || ntime < 0 || nlevels < 1 || nlevels > WBK_MAX_LEVELS) {
    wbk_set_error("wbk_contours: invalid argument (1..%d levels)", WBK_MAX_LEVELS);
    return WBK_ERR_INVALID;
  }
  const int njobs = ntime * nlevels;
  if (njobs > ctx->caps.max_jobs) {
    wbk_set_error("wbk_contours: %d jobs exceed max_jobs=%d", njobs, ctx->caps.max_jobs);
    return WBK_ERR_CAPACITY;
  }
  cudaStream_t st = (cudaStream_t)stream;
  ctx->njobs = njobs;
  ctx->nlevels = nlevels;
  if (njobs == 0) return WBK_OK;
  WbkDev& d = ctx->d;
  WBK_CUDA_CHECK(cudaMemsetAsync(d.seg_count, 0, sizeof(int) * njobs, st));
  WBK_CUDA_CHECK(cudaMemsetAsync(d.status, 0, sizeof(int) * njobs, st));
  WBK_CUDA_CHECK(cudaMemsetAsync(d.max_nx, 0, sizeof(int), st));
  LevelPack lv;
  for (int i = 0; i < WBK_MAX_LEVELS; ++i) lv.v[i] = i < nlevels ? h_levels[i] : 0.0;
  const int cols_per_block = (MS_THREADS / 32) * 31;
  dim3 grid((d.nlon + cols_per_block - 1) / cols_per_block, (d.nlat - 1 + MS_ROWS - 1) / MS_ROWS, ntime);
  if (dtype == WBK_F32) {
    WBK_LAUNCH(KID_MS_SEGMENTS, ms_segments_kernel<float>, grid, dim3(MS_THREADS), 0, st, (const float*)d_field, d, lv, nlevels);
  } else if (dtype == WBK_F64) {
    WBK_LAUNCH(KID_MS_SEGMENTS, ms_segments_kernel<double>, grid, dim3(MS_THREADS), 0, st, (const double*)d_field, d, lv, nlevels);
  } else {
    wbk_set_error("wbk_contours: unsupported dtype");
    return WBK_ERR_INVALID;
  }
  WBK_LAUNCH_CHECK();
  WBK_LAUNCH(KID_CONTOUR_LINK, contour_link_kernel, dim3(njobs), dim3(WBK_CONTOUR_THREADS), 0, st, d, njobs);
  WBK_LAUNCH_CHECK();
  return WBK_OK;
}

// ------------------------------------------------------------------------------------------ K3b
// Marching squares on the comparison bit planes the smoothing kernel leaves behind (wbk_smooth.cu): the smoothed
// field is not re-read; only the four corner values of the squares a contour actually crosses are fetched.
// One thread per (time step, chunk of MSP_ROWS rows, plane strip = 64-column half of a smoothing strip): it turns the
// even / odd column words of two consecutive rows into bit rows that start at the half's first valid column (the
// column right of its last valid one comes from the next half, or from the next strip's first valid column; the
// last strip wraps to column 0 = the periodic extension) and classifies all squares of the row with
// the same bit expression as ms_segments_kernel.  Hits are compacted per warp and emitted by ms_emit_squares.
#define MSP_THREADS 128
#define MSP_ROWS 4

__device__ __forceinline__ u64 msp_spread(u32 v) {
  u64 x = v;
  x = (x | (x << 16)) & 0x0000ffff0000ffffULL;
  x = (x | (x << 8)) & 0x00ff00ff00ff00ffULL;
  x = (x | (x << 4)) & 0x0f0f0f0f0f0f0f0fULL;
  x = (x | (x << 2)) & 0x3333333333333333ULL;
  x = (x | (x << 1)) & 0x5555555555555555ULL;
  return x;
}

struct MspGeom {
  int nstrips, V, P, PW;  // smoothing strips per row, valid columns per strip, left halo (= passes), words per (t, row, plane strip)
  int H, nps;             // 64-column halves per smoothing strip, plane strips per row (= nstrips * H)
  int ntime, nchunks;
};

// valid bits [lo, lo + n) of half h of smoothing strip s, and the grid column of bit lo (n <= 63 since P >= 1)
__device__ __forceinline__ void msp_half_range(const MspGeom& g, int nlon, int s, int h, int& lo, int& n, int& col0) {
  const int vcols = min(g.V, nlon - s * g.V);  // valid strip columns [P, P + vcols)
  lo = max(g.P - 64 * h, 0);
  n = max(min(g.P + vcols - 64 * h, 64) - lo, 0);
  col0 = s * g.V + 64 * h + lo - g.P;
}

// Bit row of a plane strip from its even / odd column words (e, o): bit i = strip bit lo + i; the bit right of the
// last valid column (i = n) is bit `nbit` of the words (ne, no) of the plane strip that follows
__device__ __forceinline__ u64 msp_bits(u32 e, u32 o, u32 ne, u32 no, int lo, int n, int nbit) {
  const u64 m = (msp_spread(e) | (msp_spread(o) << 1)) >> lo;
  const u32 nb = (((nbit & 1) ? no : ne) >> (nbit >> 1)) & 1u;
  return (m & ~(1ULL << n)) | ((u64)nb << n);
}

// Emission of the hits a warp found in its MSP_ROWS x 32 (row, strip) items: hit h goes to lane h % 32, so every
// ms_emit_squares call works on (up to) 32 squares whatever their distribution over rows and strips.
// masks: [MSP_ROWS][64] words of this warp (lane l: words 2l, 2l+1 of row i), prefix: [33] hits before lane l.
template <typename T>
__device__ __noinline__ void msp_emit(const WbkDev& d, const T* __restrict__ src, const u32* masks, const int* prefix,
                                      int job, int item0, const MspGeom& g, double level) {
  const int lane = wbk_lane(), nlon = d.nlon;
  const int total = prefix[32];
  for (int h0 = 0; h0 < total; h0 += 32) {
    const int h = h0 + lane;
    const bool active = h < total;
    int r0 = 0, c0 = 0;
    double ul = 0, ur = 0, ll = 0, lr = 0;
    if (active) {
      int sl = 0;  // the lane whose item holds hit h: largest sl with prefix[sl] <= h
#pragma unroll
      for (int step = 16; step > 0; step >>= 1)
        if (prefix[sl + step] <= h) sl += step;
      int k = h - prefix[sl], i = 0;
      u64 m = 0;
      for (; i < MSP_ROWS; ++i) {
        m = (u64)masks[i * 64 + 2 * sl] | ((u64)masks[i * 64 + 2 * sl + 1] << 32);
        const int c = __popcll(m);
        if (k < c) break;
        k -= c;
      }
      for (; k > 0; --k) m &= m - 1;
      const int bit = __ffsll((long long)m) - 1;
      const int hit_item = item0 + sl;
      const int hps = hit_item % g.nps, hchunk = hit_item / g.nps;
      int lo, n, col0;
      msp_half_range(g, nlon, hps / g.H, hps % g.H, lo, n, col0);
      r0 = hchunk * MSP_ROWS + i;
      c0 = col0 + bit;
      const int cr = c0 + 1 == nlon ? 0 : c0 + 1;
      ul = (double)src[(size_t)r0 * nlon + c0];
      ur = (double)src[(size_t)r0 * nlon + cr];
      ll = (double)src[(size_t)(r0 + 1) * nlon + c0];
      lr = (double)src[(size_t)(r0 + 1) * nlon + cr];
    }
    ms_emit_squares(d, job, active, r0, c0, ul, ur, ll, lr, level);
  }
}

// The plane rows a CTA needs are ONE contiguous block of global memory ([t][row][plane strip][PW] words, rows of
// nps * PW * 4 bytes): they are staged in shared memory with a single TMA bulk copy
// (cp.async.bulk.shared.global + mbarrier transaction count) instead of ~40 scattered loads per thread.
__device__ __forceinline__ void msp_stage_rows(u32* smem_dst, const u32* gsrc, unsigned bytes, u64* bar) {
#ifndef WBK_EMU
  const unsigned bar_s = (unsigned)__cvta_generic_to_shared(bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned dst_s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_s),
                 "l"(gsrc), "r"(bytes), "r"(bar_s)
                 : "memory");
  }
  // every thread waits for phase 0 of the barrier (the copy's bytes have landed)
  asm volatile(
      "{\n\t.reg .pred p;\n\tMSP_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra MSP_DONE;\n\tbra MSP_WAIT;\n\t"
      "MSP_DONE:\n\t}" ::"r"(bar_s)
      : "memory");
#else
  if (threadIdx.x == 0)
    for (unsigned i = 0; i < bytes / 4; ++i) smem_dst[i] = gsrc[i];
  __syncthreads();
#endif
}

// grid: (bands of MSP_THREADS items, time steps); item = (row chunk, strip) of the time step, strip fastest
template <typename T>
__global__ void __launch_bounds__(MSP_THREADS, 5)
ms_planes_kernel(const T* __restrict__ field, const u32* __restrict__ planes, const __grid_constant__ WbkDev d,
                 const __grid_constant__ LevelPack levels, int nlevels, const __grid_constant__ MspGeom g, int t_base) {
  __shared__ u32 smask[MSP_THREADS / 32][MSP_ROWS][64];
  __shared__ int sprefix[MSP_THREADS / 32][33];
  __shared__ __align__(8) u64 s_bar;
  WBK_DYN_SMEM(u32, srows);  // plane words of rows [row0, row0 + nrows) of this time step
  const int lane = wbk_lane(), warp = wbk_warp();
  const int nlat = d.nlat, nlon = d.nlon, W = d.W;
  const int per_t = g.nchunks * g.nps;
  const int t = t_base + (int)blockIdx.y;
  const int item_first = (int)blockIdx.x * MSP_THREADS;
  const int item_last = min(item_first + MSP_THREADS, per_t) - 1;  // the grid covers exactly ceil(per_t / MSP_THREADS) bands
  const size_t trow = (size_t)t * nlat;
  const size_t row_words = (size_t)g.nps * g.PW;
  const int row0 = (item_first / g.nps) * MSP_ROWS;
  const int row1 = min((item_last / g.nps + 1) * MSP_ROWS, nlat - 1);  // last row read
  // the bulk copy wants 16-byte aligned addresses and sizes: rows are only 8-byte multiples for an even number of
  // levels, so the block is widened to the enclosing 16-byte range (`delta` words precede the first row)
  const size_t gstart = (trow + row0) * row_words;
  const int delta = (int)(gstart & 3);
  const unsigned bytes = (unsigned)((((size_t)delta + (size_t)(row1 - row0 + 1) * row_words) * 4 + 15) & ~(size_t)15);
  msp_stage_rows(srows, planes + (gstart - delta), bytes, &s_bar);

  const int item0 = item_first + warp * 32;  // first item of this warp
  const bool live = item0 + lane < per_t;
  const int it = live ? item0 + lane : item_first;  // dead lanes shadow a staged item with an empty square mask
  const int ps = it % g.nps, chunk = it / g.nps;
  const int s = ps / g.H, h = ps - s * g.H;
  const int r_begin = chunk * MSP_ROWS;
  const int r_end = min(r_begin + MSP_ROWS, nlat - 1);  // squares r0 in [r_begin, r_end)
  int lo, n, col0;
  msp_half_range(g, nlon, s, h, lo, n, col0);
  // the column right of the last valid one: bit 0 of the next half when this half is valid up to its last bit,
  // else the first valid column of the strip to the right (the last strip wraps: periodic extension)
  const bool next_half = lo + n == 64;  // implies h + 1 < H (the last half ends at bit 64 - P)
  const int pn = next_half ? ps + 1 : (s + 1 == g.nstrips ? 0 : s + 1) * g.H;
  const int nbit = next_half ? 0 : g.P;
  // squares owned by this plane strip: its valid columns whose base square exists (c0 <= W - 2)
  int nsq = min(n, W - 1 - col0);
  if (nsq < 0 || !live) nsq = 0;
  const u64 sqmask = (1ULL << nsq) - 1ULL;  // nsq <= 63
  const T* src = field + trow * nlon;
  const u32* own = srows + delta + (size_t)(r_begin - row0) * row_words + (size_t)ps * g.PW;
  const u32* nxt = srows + delta + (size_t)(r_begin - row0) * row_words + (size_t)pn * g.PW;

  for (int l = 0; l < nlevels; ++l) {
    int cnt = 0;
    const bool ok0 = r_begin <= r_end;
    u64 g0 = ok0 ? msp_bits(own[2 + 2 * l], own[3 + 2 * l], nxt[2 + 2 * l], nxt[3 + 2 * l], lo, n, nbit) : 0;
    u64 n0 = ok0 ? msp_bits(own[0], own[1], nxt[0], nxt[1], lo, n, nbit) : 0;
#pragma unroll
    for (int i = 0; i < MSP_ROWS; ++i) {
      u64 hits = 0, g1 = 0, n1 = 0;
      if (r_begin + i < r_end) {
        const u32* po = own + (size_t)(i + 1) * row_words;
        const u32* px = nxt + (size_t)(i + 1) * row_words;
        g1 = msp_bits(po[2 + 2 * l], po[3 + 2 * l], px[2 + 2 * l], px[3 + 2 * l], lo, n, nbit);
        n1 = msp_bits(po[0], po[1], px[0], px[1], lo, n, nbit);
        const u64 both = g0 & g1, either = g0 | g1, nn = n0 | n1;
        hits = ((either | (either >> 1)) & ~(both & (both >> 1))) & ~(nn | (nn >> 1)) & sqmask;
      }
      smask[warp][i][2 * lane] = (u32)hits;
      smask[warp][i][2 * lane + 1] = (u32)(hits >> 32);
      cnt += __popcll(hits);
      g0 = g1;
      n0 = n1;
    }
    const int incl = wbk_warp_incl_scan(cnt);
    if (__shfl_sync(WBK_FULL, incl, 31) == 0) continue;  // warp-uniform: no contour crosses these items
    sprefix[warp][lane + 1] = incl;
    if (lane == 0) sprefix[warp][0] = 0;
    __syncwarp();
    msp_emit<T>(d, src, &smask[warp][0][0], sprefix[warp], t * nlevels + l, item0, g, levels.v[l]);
    __syncwarp();
  }
}

int wbk_launch_smooth(const void* d_in, int in_dtype, void* d_out, int out_dtype, int ntime, int nlat, int nlon, int passes,
                      int round_mode, int nan_border, const wbk_smooth_opts* opts, u32* d_planes, const double* h_levels,
                      int nlevels, cudaStream_t st);  // wbk_smooth.cu
void wbk_smooth_plane_geometry(int nlon, int passes, int* nstrips, int* V, int* halves);

extern "C" int wbk_smooth_contours(wbk_ctx* ctx, const void* d_in, int in_dtype, double* d_smoothed, int ntime,
                                   int passes, const double* h_levels, int nlevels, const wbk_smooth_opts* opts,
                                   void* stream) {
  if (!ctx || !d_in || !d_smoothed || !h_levels || ntime < 0 || nlevels < 1 || nlevels > WBK_MAX_LEVELS ||
      passes < 1 || passes > WBK_SMOOTH_MAX_FUSED || ctx->d.nlat < 4) {
    wbk_set_error("wbk_smooth_contours: invalid argument (1..%d passes, 1..%d levels)", WBK_SMOOTH_MAX_FUSED, WBK_MAX_LEVELS);
    return WBK_ERR_INVALID;
  }
  const int njobs = ntime * nlevels;
  if (njobs > ctx->caps.max_jobs) {
    wbk_set_error("wbk_smooth_contours: %d jobs exceed max_jobs=%d", njobs, ctx->caps.max_jobs);
    return WBK_ERR_CAPACITY;
  }
  cudaStream_t st = (cudaStream_t)stream;
  ctx->njobs = njobs;
  ctx->nlevels = nlevels;
  if (njobs == 0) return WBK_OK;
  WbkDev& d = ctx->d;
  WBK_CUDA_CHECK(cudaMemsetAsync(d.seg_count, 0, sizeof(int) * njobs, st));
  WBK_CUDA_CHECK(cudaMemsetAsync(d.status, 0, sizeof(int) * njobs, st));
  WBK_CUDA_CHECK(cudaMemsetAsync(d.max_nx, 0, sizeof(int), st));
  LevelPack lv;
  for (int i = 0; i < WBK_MAX_LEVELS; ++i) lv.v[i] = i < nlevels ? h_levels[i] : 0.0;
  const int rmode = in_dtype == WBK_F32 ? WBK_ROUND_FIRST : WBK_ROUND_NONE;
  int rc = wbk_launch_smooth(d_in, in_dtype, d_smoothed, WBK_F64, ntime, d.nlat, d.nlon, passes, rmode, 2, opts, d.planes,
                             h_levels, nlevels, st);
  if (rc != WBK_OK) return rc;
  MspGeom g;
  wbk_smooth_plane_geometry(d.nlon, passes, &g.nstrips, &g.V, &g.H);
  g.nps = g.nstrips * g.H;
  g.P = passes;
  g.PW = 2 + 2 * nlevels;
  g.ntime = ntime;
  g.nchunks = (d.nlat - 1 + MSP_ROWS - 1) / MSP_ROWS;
  const int per_t = g.nchunks * g.nps;
  const int bands = (per_t + MSP_THREADS - 1) / MSP_THREADS;
  // rows staged per CTA: the row chunks its MSP_THREADS items span, plus the row below the last one
  const int max_chunks = (MSP_THREADS + g.nps - 2) / g.nps + 1;
  const size_t smem = (size_t)(max_chunks * MSP_ROWS + 1) * g.nps * g.PW * sizeof(u32) + 32;
  if (smem > 160 * 1024) {
    wbk_set_error("wbk_smooth_contours: %d strips x %d levels do not fit the plane staging buffer", g.nstrips, nlevels);
    return WBK_ERR_CAPACITY;
  }
  WBK_CUDA_CHECK(cudaFuncSetAttribute(ms_planes_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int t0 = 0; t0 < ntime; t0 += 65535) {  // gridDim.y <= 65535
    const int nt = ntime - t0 < 65535 ? ntime - t0 : 65535;
    WBK_LAUNCH(KID_MS_SEGMENTS, ms_planes_kernel<double>, dim3(bands, nt), dim3(MSP_THREADS), smem, st,
               (const double*)d_smoothed, (const u32*)d.planes, d, lv, nlevels, g, t0);
    WBK_LAUNCH_CHECK();
  }
  WBK_LAUNCH(KID_CONTOUR_LINK, contour_link_kernel, dim3(njobs), dim3(WBK_CONTOUR_THREADS), 0, st, d, njobs);
  WBK_LAUNCH_CHECK();
  return WBK_OK;
}

extern "C" int wbk_contours_counts(wbk_ctx* ctx, int* h_ncontours, int* h_npoints, int* h_status, int* h_max_nx,
                                   void* stream) {
  if (!ctx || !h_ncontours || !h_npoints || !h_status || !h_max_nx) {
    wbk_set_error("wbk_contours_counts: invalid argument");
    return WBK_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int J = ctx->njobs;
  *h_max_nx = 0;
  if (J == 0) return WBK_OK;
  WBK_CUDA_CHECK(cudaMemcpyAsync(h_ncontours, ctx->d.out_nc, sizeof(int) * J, cudaMemcpyDeviceToHost, st));
  WBK_CUDA_CHECK(cudaMemcpyAsync(h_npoints, ctx->d.out_np, sizeof(int) * J, cudaMemcpyDeviceToHost, st));
  WBK_CUDA_CHECK(cudaMemcpyAsync(h_status, ctx->d.status, sizeof(int) * J, cudaMemcpyDeviceToHost, st));
  WBK_CUDA_CHECK(cudaMemcpyAsync(h_max_nx, ctx->d.max_nx, sizeof(int), cudaMemcpyDeviceToHost, st));
  WBK_CUDA_CHECK(cudaStreamSynchronize(st));
  for (int j = 0; j < J; ++j) {
    if (h_status[j] & (WBK_ST_SEG_OVERFLOW | WBK_ST_CONTOUR_OVERFLOW)) {
      wbk_set_error("wbk_contours: job %d overflowed its arena (status %d); raise seg_cap / contour_cap", j, h_status[j]);
      return WBK_ERR_CAPACITY;
    }
  }
  return WBK_OK;
}

__global__ void contours_pack_kernel(WbkDev d, int njobs, const int* __restrict__ job_off, const int* __restrict__ pt_job_off,
                                     int* __restrict__ pt_off, int* __restrict__ meta, u32* __restrict__ pts) {
  const int job = blockIdx.x;
  if (job >= njobs) return;
  const int nc = d.out_nc[job], np = d.out_np[job];
  const int c0 = job_off[job], p0 = pt_job_off[job];
  const int* tab = d.out_tab + (size_t)job * d.CC * 4;
  const int* sumy = d.out_sumy + (size_t)job * d.CC;
  const u32* src = d.out_pts + (size_t)job * d.R;
  for (int c = threadIdx.x; c < nc; c += blockDim.x) {
    pt_off[c0 + c] = p0 + tab[4 * c + 0];
    meta[4 * (c0 + c) + 0] = tab[4 * c + 2];
    meta[4 * (c0 + c) + 1] = tab[4 * c + 3];
    meta[4 * (c0 + c) + 2] = sumy[c];
    meta[4 * (c0 + c) + 3] = job;
  }
  for (int i = threadIdx.x; i < np; i += blockDim.x) pts[p0 + i] = src[i];
  if (job == njobs - 1 && threadIdx.x == 0) pt_off[c0 + nc] = p0 + np;
}

extern "C" int wbk_contours_pack(wbk_ctx* ctx, const int* h_ncontours, const int* h_npoints, int* d_job_off,
                                 int* d_pt_off, int* d_meta, uint32_t* d_pts, void* stream) {
  if (!ctx || !h_ncontours || !h_npoints || !d_job_off || !d_pt_off || !d_meta || !d_pts) {
    wbk_set_error("wbk_contours_pack: invalid argument");
    return WBK_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int J = ctx->njobs;
  // exclusive prefix sums on the host (J is small), staged through the scan arena
  int* h = new int[2 * (J + 1)];
  int c = 0, p = 0;
  for (int j = 0; j < J; ++j) {
    h[j] = c;
    h[J + 1 + j] = p;
    c += h_ncontours[j];
    p += h_npoints[j];
  }
  h[J] = c;
  h[2 * J + 1] = p;
  int* d_pt_job_off = ctx->d.scan;  // scratch, free after wbk_contours
  cudaError_t e1 = cudaMemcpyAsync(d_job_off, h, sizeof(int) * (J + 1), cudaMemcpyHostToDevice, st);
  cudaError_t e2 = cudaMemcpyAsync(d_pt_job_off, h + J + 1, sizeof(int) * (J + 1), cudaMemcpyHostToDevice, st);
  cudaError_t e3 = cudaStreamSynchronize(st);  // h is pageable: make sure it was consumed
  delete[] h;
  WBK_CUDA_CHECK(e1);
  WBK_CUDA_CHECK(e2);
  WBK_CUDA_CHECK(e3);
  if (J == 0) return WBK_OK;
  if (c == 0) {
    int zero = 0;
    WBK_CUDA_CHECK(cudaMemcpyAsync(d_pt_off, &zero, sizeof(int), cudaMemcpyHostToDevice, st));
    WBK_CUDA_CHECK(cudaStreamSynchronize(st));
    return WBK_OK;
  }
  WBK_LAUNCH(KID_CONTOUR_PACK, contours_pack_kernel, dim3(J), dim3(256), 0, st, ctx->d, J, (const int*)d_job_off, (const int*)d_pt_job_off, d_pt_off, d_meta, (u32*)d_pts);
  WBK_LAUNCH_CHECK();
  return WBK_OK;
}
