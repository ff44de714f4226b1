// Streamer / overturning / cutoff indices on the device-resident contour tables.
#include "wbk_ctx.cuh"

size_t wbk_index_layout(struct wbk_ctx* ctx, unsigned char* base, size_t off) {
  (void)ctx; (void)base;
  return off;
}
