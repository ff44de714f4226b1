// fused smoothing, float input -> double output (WBK_ROUND_FIRST): instantiations of wbk_smooth_impl.cuh
#include "wbk_smooth_impl.cuh"

int wbk_ss_launch_f32_f64(const void* in, void* out, int passes, SsParams& prm, cudaStream_t st) {
  return ss_launch<float, double, WBK_ROUND_FIRST>(in, out, passes, prm, st);
}
