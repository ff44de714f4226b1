// fused smoothing, float input -> float output (WBK_ROUND_ALL): instantiations of wbk_smooth_impl.cuh
#include "wbk_smooth_impl.cuh"

int wbk_ss_launch_f32_f32(const void* in, void* out, int passes, SsParams& prm, cudaStream_t st) {
  return ss_launch<float, float, WBK_ROUND_ALL>(in, out, passes, prm, st);
}
