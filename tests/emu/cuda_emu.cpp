// cuda_emu.cpp — scheduler of the host-side CUDA emulator (TEST INFRASTRUCTURE ONLY, see cuda_emu.h).
#include "cuda_emu.h"

namespace nbemu {

State g;
uint3 g_threadIdx, g_blockIdx;
dim3 g_blockDim, g_gridDim;

static const size_t kStack = 256 * 1024;

static void trampoline() {
  g.body();
  g.done[g.cur] = 1;
  g.nlive--;
  // a thread that exits releases barriers other threads may be waiting on
  if (g.nlive > 0 && g.bar_count >= g.nlive) {
    g.bar_count = 0;
    g.bar_gen++;
  }
  swapcontext(&g.ctx[g.cur], &g.main_ctx);
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  int nt = (int)(block.x * block.y * block.z);
  if ((int)g.stacks.size() < nt) {
    size_t old = g.stacks.size();
    g.stacks.resize(nt);
    for (size_t i = old; i < (size_t)nt; ++i) g.stacks[i] = (char*)malloc(kStack);
  }
  g.ctx.resize(nt);
  g.done.assign(nt, 0);
  g.slots.assign(nt, 0);
  int nw = (nt + 31) / 32;
  std::vector<unsigned char> smem_buf(smem + 64);
  g.dyn_smem = (unsigned char*)(((uintptr_t)smem_buf.data() + 15) & ~(uintptr_t)15);
  g.body = body;
  g_blockDim = block;
  g_gridDim = grid;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        g_blockIdx = uint3{bx, by, bz};
        g.nthreads = nt;
        g.nlive = nt;
        g.bar_count = 0;
        g.wbar_count.assign(nw, 0);
        g.wbar_gen.assign(nw, 0);
        memset(g.dyn_smem, 0xCD, smem);  // poison: uninitialised shared memory reads show up as garbage
        for (int t = 0; t < nt; ++t) {
          g.done[t] = 0;
          getcontext(&g.ctx[t]);
          g.ctx[t].uc_stack.ss_sp = g.stacks[t];
          g.ctx[t].uc_stack.ss_size = kStack;
          g.ctx[t].uc_link = &g.main_ctx;
          makecontext(&g.ctx[t], (void (*)())trampoline, 0);
        }
        long spins = 0;
        while (g.nlive > 0) {
          for (int t = 0; t < nt; ++t) {
            if (g.done[t]) continue;
            g.cur = t;
            unsigned tx = t % block.x, ty = (t / block.x) % block.y, tz = t / (block.x * block.y);
            g_threadIdx = uint3{tx, ty, tz};
            swapcontext(&g.main_ctx, &g.ctx[t]);
          }
          if (++spins > 50000000L) {
            fprintf(stderr, "nbemu: block (%u,%u,%u) appears deadlocked (divergent barrier?)\n", bx, by, bz);
            abort();
          }
        }
      }
  g.dyn_smem = nullptr;
}

}  // namespace nbemu
