// Fiber scheduler of the SIMT emulation shim (test infrastructure only; see simt_emu.h).
#include "simt_emu.h"

namespace simt {
Fiber* cur = nullptr;
ucontext_t sched_ctx;
uint3_ g_blockIdx;
dim3 g_blockDim, g_gridDim;
unsigned char* g_dyn_smem = nullptr;
std::vector<Fiber> fibers;

static void (*g_entry)(void*);
static void* g_args;
static const size_t kStack = 64 * 1024;

void yield_wait(int state) {
  cur->state = state;
  swapcontext(&cur->ctx, &sched_ctx);
}

void warp_sync() { yield_wait(2); }

static void trampoline() {
  g_entry(g_args);
  cur->state = 3;
  swapcontext(&cur->ctx, &sched_ctx);
}

static void run_cta(int nthreads) {
  for (int i = 0; i < nthreads; ++i) {
    Fiber& f = fibers[i];
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = kStack;
    f.ctx.uc_link = &sched_ctx;
    f.state = 0;
    f.xchg = 0;
    makecontext(&f.ctx, trampoline, 0);
  }
  int done = 0;
  while (done < nthreads) {
    // run every runnable fiber until it blocks or finishes
    bool progressed = false;
    for (int i = 0; i < nthreads; ++i) {
      Fiber& f = fibers[i];
      if (f.state == 0) {
        cur = &f;
        swapcontext(&sched_ctx, &f.ctx);
        progressed = true;
        if (f.state == 3) ++done;
      }
    }
    // release warp barriers whose live lanes have all arrived
    for (int b = 0; b < nthreads; b += 32) {
      int e = std::min(b + 32, nthreads);
      bool all = true, any = false;
      for (int i = b; i < e; ++i) {
        if (fibers[i].state == 2) any = true;
        else if (fibers[i].state != 3) all = false;
      }
      if (any && all) {
        for (int i = b; i < e; ++i)
          if (fibers[i].state == 2) fibers[i].state = 0;
        progressed = true;
      }
    }
    // release the CTA barrier when every live thread waits on it
    bool all = true, any = false;
    for (int i = 0; i < nthreads; ++i) {
      if (fibers[i].state == 1) any = true;
      else if (fibers[i].state != 3) all = false;
    }
    if (any && all) {
      for (int i = 0; i < nthreads; ++i)
        if (fibers[i].state == 1) fibers[i].state = 0;
      progressed = true;
    }
    if (!progressed && done < nthreads) {
      std::fprintf(stderr, "simt_emu: deadlock (divergent barrier) in block (%u,%u,%u)\n", g_blockIdx.x,
                   g_blockIdx.y, g_blockIdx.z);
      std::abort();
    }
  }
}

void run_grid(void (*entry)(void*), void* args, dim3 grid, dim3 block, size_t smem) {
  g_entry = entry;
  g_args = args;
  g_blockDim = block;
  g_gridDim = grid;
  int nthreads = (int)(block.x * block.y * block.z);
  if ((int)fibers.size() < nthreads) {
    size_t old = fibers.size();
    fibers.resize(nthreads);
    for (size_t i = old; i < fibers.size(); ++i) fibers[i].stack = (char*)std::malloc(kStack);
  }
  for (int i = 0; i < nthreads; ++i) {
    fibers[i].lin = i;
    fibers[i].tid.x = i % block.x;
    fibers[i].tid.y = (i / block.x) % block.y;
    fibers[i].tid.z = i / (block.x * block.y);
  }
  std::vector<unsigned char> dyn(smem + 16);
  g_dyn_smem = dyn.data();
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        g_blockIdx.x = bx;
        g_blockIdx.y = by;
        g_blockIdx.z = bz;
        run_cta(nthreads);
      }
  g_dyn_smem = nullptr;
}
}  // namespace simt
