// fused smoothing, short input -> double output (WBK_ROUND_NONE): instantiations of wbk_smooth_impl.cuh
#include "wbk_smooth_impl.cuh"

int wbk_ss_launch_i16_f64(const void* in, void* out, int passes, SsParams& prm, cudaStream_t st) {
  return ss_launch<short, double, WBK_ROUND_NONE>(in, out, passes, prm, st);
}
