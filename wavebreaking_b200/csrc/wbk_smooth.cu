// Dispatch of the fused smoothing (kernel: wbk_smooth_impl.cuh; one translation unit per dtype combination so that
// the 8 pass counts x 3 plane variants compile in parallel).
#include "wbk_common.cuh"
#include "wbk_ms.cuh"
#include "wbk_smooth_impl.cuh"

int wbk_ss_launch_f32_f64(const void* in, void* out, int passes, SsParams& prm, cudaStream_t st);  // ROUND_FIRST
int wbk_ss_launch_f32_f32(const void* in, void* out, int passes, SsParams& prm, cudaStream_t st);  // ROUND_ALL
int wbk_ss_launch_f64_f64(const void* in, void* out, int passes, SsParams& prm, cudaStream_t st);  // ROUND_NONE
int wbk_ss_launch_i16_f64(const void* in, void* out, int passes, SsParams& prm, cudaStream_t st);  // ROUND_NONE

// geometry of the bit planes a fused smoothing of `passes` (<= WBK_SMOOTH_MAX_FUSED) passes writes
static int g_halves_override = 0;
int wbk_smooth_halves_override() { return g_halves_override; }
extern "C" void wbk_tune_smooth_halves(int halves) { g_halves_override = halves == 1 || halves == 2 ? halves : 0; }

// upper bound of the plane strips per grid row over every pass count and strip width the smoothing kernel may use
int wbk_smooth_plane_strips_max(int nlon) {
  const int one = (nlon + (64 - 2 * WBK_SMOOTH_MAX_FUSED) - 1) / (64 - 2 * WBK_SMOOTH_MAX_FUSED);
  const int two = 2 * ((nlon + (128 - 2 * SS_MAX_P_H2) - 1) / (128 - 2 * SS_MAX_P_H2));
  return one > two ? one : two;
}

void wbk_smooth_plane_geometry(int nlon, int passes, int* nstrips, int* V, int* halves) {
  *halves = ss_halves(nlon, passes);
  *V = 64 * *halves - 2 * passes;
  *nstrips = (nlon + *V - 1) / *V;
}

// One fused launch of 1..WBK_SMOOTH_MAX_FUSED passes (used by wbk_smooth and wbk_smooth_contours).
int wbk_launch_smooth(const void* d_in, int in_dtype, void* d_out, int out_dtype, int ntime, int nlat, int nlon, int passes,
                      int round_mode, int nan_border, const wbk_smooth_opts* opts, u32* d_planes, const double* h_levels,
                      int nlevels, cudaStream_t st) {
  SsParams prm = {};
  prm.nlat = nlat; prm.nlon = nlon; prm.ntime = ntime; prm.nan_border = nan_border;
  prm.flip_lat = opts ? opts->flip_lat : 0;
  prm.flip_lon = opts ? opts->flip_lon : 0;
  prm.scale = opts ? opts->scale : 1.0;
  prm.offset = opts ? opts->offset : 0.0;
  prm.fill = (opts && opts->has_fill) ? opts->fill : 0x7fffffff;
  prm.planes = d_planes;
  prm.nlevels = d_planes ? nlevels : 0;
  for (int i = 0; i < WBK_MAX_LEVELS; ++i) prm.levels.v[i] = (d_planes && i < nlevels) ? h_levels[i] : 0.0;
  if (in_dtype == WBK_I16) {
    if (round_mode != WBK_ROUND_NONE || out_dtype != WBK_F64) {
      wbk_set_error("wbk_smooth: packed int16 input decodes to float64 (round_mode NONE, float64 output)");
      return WBK_ERR_INVALID;
    }
    return wbk_ss_launch_i16_f64(d_in, d_out, passes, prm, st);
  }
  if (in_dtype == WBK_F32 && out_dtype == WBK_F64 && round_mode == WBK_ROUND_FIRST)
    return wbk_ss_launch_f32_f64(d_in, d_out, passes, prm, st);
  if (in_dtype == WBK_F32 && out_dtype == WBK_F32 && round_mode == WBK_ROUND_ALL)
    return wbk_ss_launch_f32_f32(d_in, d_out, passes, prm, st);
  if (in_dtype == WBK_F64 && out_dtype == WBK_F64 && round_mode == WBK_ROUND_NONE)
    return wbk_ss_launch_f64_f64(d_in, d_out, passes, prm, st);
  wbk_set_error("wbk_smooth: dtype / round_mode combination not supported");
  return WBK_ERR_INVALID;
}

extern "C" int wbk_smooth(const void* d_in, int in_dtype, void* d_out, int out_dtype, void* d_tmp, int ntime,
                          int nlat, int nlon, int passes, int round_mode, const wbk_smooth_opts* opts, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!d_in || !d_out || d_in == d_out || ntime < 0 || nlat < 4 || nlon < 1 || passes < 0) {
    wbk_set_error("wbk_smooth: invalid argument (in-place is not supported)");
    return WBK_ERR_INVALID;
  }
  if (ntime == 0) return WBK_OK;
  const int border = 2;  // int(3 / 2 + 0.5), spatial.py:106
  if (passes == 0) {
    // no pass: orientation / decode only, then the NaN border (the reference still writes it, spatial.py:106-107)
    if (in_dtype == WBK_I16 ? out_dtype != WBK_F64 : in_dtype != out_dtype) {
      wbk_set_error("wbk_smooth: passes == 0 keeps the dtype");
      return WBK_ERR_INVALID;
    }
    const bool plain = in_dtype != WBK_I16 && !(opts && (opts->flip_lat || opts->flip_lon));
    if (plain) {
      const size_t bytes = (size_t)ntime * nlat * nlon * (in_dtype == WBK_F32 ? 4 : 8);
      WBK_CUDA_CHECK(cudaMemcpyAsync(d_out, d_in, bytes, cudaMemcpyDeviceToDevice, st));
    } else {
      const int rc = wbk_orient(d_in, in_dtype, d_out, ntime, nlat, nlon, opts, stream);
      if (rc != WBK_OK) return rc;
    }
    return wbk_nan_border(d_out, out_dtype, ntime, nlat, nlon, border, stream);
  }
  if (passes > WBK_SMOOTH_MAX_FUSED && !d_tmp) {
    wbk_set_error("wbk_smooth: d_tmp required for passes > %d", WBK_SMOOTH_MAX_FUSED);
    return WBK_ERR_INVALID;
  }
  // chunk the passes; intermediates are stored in the output dtype (exact: f64, or f32 values under ROUND_ALL)
  const int nchunks = (passes + WBK_SMOOTH_MAX_FUSED - 1) / WBK_SMOOTH_MAX_FUSED;
  const void* src = d_in;
  int src_dtype = in_dtype, done = 0;
  for (int ch = 0; ch < nchunks; ++ch) {
    const int n = passes - done < WBK_SMOOTH_MAX_FUSED ? passes - done : WBK_SMOOTH_MAX_FUSED;
    const bool last = ch == nchunks - 1;
    void* dstp = ((nchunks - 1 - ch) % 2 == 0) ? d_out : d_tmp;  // ping-pong so that the last chunk lands in d_out
    int rm = round_mode;
    if (ch > 0 && round_mode == WBK_ROUND_FIRST) rm = WBK_ROUND_NONE;  // later chunks: plain float64
    // a later chunk under ROUND_ALL reads float32 intermediates and rounds every pass, like the first
    const int rc = wbk_launch_smooth(src, src_dtype, dstp, out_dtype, ntime, nlat, nlon, n, rm, last ? border : 0,
                                     ch == 0 ? opts : nullptr, nullptr, nullptr, 0, st);
    if (rc != WBK_OK) return rc;
    src = dstp;
    src_dtype = out_dtype;
    done += n;
  }
  return WBK_OK;
}
