// cpu_simt.h — a tiny SIMT emulator so the product's CUDA kernels can be exercised on a machine
// without a GPU.  TEST INFRASTRUCTURE ONLY: nothing under agimus_controller_b200/ includes it.
//
// Each CUDA thread of a block becomes a ucontext fiber; __syncthreads / __syncwarp(mask) /
// __shfl_sync yield round-robin until every participant has arrived.  Blocks run one after the
// other, so `extern __shared__` memory is one static buffer.  Deterministic and single-threaded.
#ifndef AGX_CPU_SIMT_H_
#define AGX_CPU_SIMT_H_

#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define AGX_EMULATE 1

struct dim3 {
  unsigned x = 1, y = 1, z = 1;
  dim3() {}
  dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

namespace simt {

constexpr int MAX_THREADS = 256;
constexpr size_t STACK = 256 * 1024;

struct Fiber {
  ucontext_t ctx;
  std::vector<char> stack;
  bool done = false;
  // barrier bookkeeping: how many times this fiber has arrived at a barrier with the given key
  uint64_t arrivals_block = 0;
  uint64_t arrivals_warp[33] = {0};  // keyed by lowest set bit of the mask (+32 for full-warp)
  double shfl_slot = 0;
};

struct Block {
  std::vector<Fiber> fibers;
  ucontext_t sched;
  int current = 0;
  int nthreads = 0;
  std::function<void()> body;
};

inline Block*& cur_block() {
  static Block* b = nullptr;
  return b;
}
inline dim3& tIdx() { static dim3 v; return v; }
inline dim3& bIdx() { static dim3 v; return v; }
inline dim3& bDim() { static dim3 v; return v; }
inline dim3& gDim() { static dim3 v; return v; }

alignas(16) inline char g_smem[232448];

inline void yield_fiber() {
  Block* b = cur_block();
  Fiber& f = b->fibers[b->current];
  swapcontext(&f.ctx, &b->sched);
}

inline void trampoline() {
  Block* b = cur_block();
  b->body();
  b->fibers[b->current].done = true;
  swapcontext(&b->fibers[b->current].ctx, &b->sched);
}

inline void run_block(int nthreads, const std::function<void()>& body) {
  Block blk;
  blk.nthreads = nthreads;
  blk.body = body;
  blk.fibers.resize(nthreads);
  cur_block() = &blk;
  for (int i = 0; i < nthreads; ++i) {
    Fiber& f = blk.fibers[i];
    f.stack.resize(STACK);
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack.data();
    f.ctx.uc_stack.ss_size = STACK;
    f.ctx.uc_link = nullptr;
    makecontext(&f.ctx, (void (*)())trampoline, 0);
  }
  int live = nthreads;
  uint64_t spins = 0;
  while (live > 0) {
    for (int i = 0; i < nthreads; ++i) {
      Fiber& f = blk.fibers[i];
      if (f.done) continue;
      blk.current = i;
      tIdx() = dim3(i, 0, 0);
      swapcontext(&blk.sched, &f.ctx);
      if (f.done) --live;
    }
    if (++spins > (1ull << 34)) { std::fprintf(stderr, "simt: deadlock?\n"); std::abort(); }
  }
  cur_block() = nullptr;
}

template <class F>
inline void launch(dim3 grid, dim3 block, F&& kernel_call) {
  gDim() = grid;
  bDim() = block;
  for (unsigned bx = 0; bx < grid.x; ++bx) {
    bIdx() = dim3(bx, 0, 0);
    run_block((int)block.x, kernel_call);
  }
}

// barrier among the fibers selected by `sel(i)`; `counter(f)` is the arrival counter to use
template <class Sel, class Cnt>
inline void barrier(Sel sel, Cnt counter) {
  Block* b = cur_block();
  const int me = b->current;
  const uint64_t mine = ++counter(b->fibers[me]);
  for (;;) {
    bool all = true;
    for (int i = 0; i < b->nthreads; ++i)
      if (sel(i) && !b->fibers[i].done && counter(b->fibers[i]) < mine) { all = false; break; }
    if (all) return;
    yield_fiber();
    tIdx() = dim3(me, 0, 0);
  }
}

}  // namespace simt

#define threadIdx (simt::tIdx())
#define blockIdx (simt::bIdx())
#define blockDim (simt::bDim())
#define gridDim (simt::gDim())
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)

inline void __syncthreads() {
  simt::barrier([](int) { return true; }, [](simt::Fiber& f) -> uint64_t& { return f.arrivals_block; });
}
inline void __syncwarp(unsigned mask = 0xffffffffu) {
  const int me = simt::cur_block()->current;
  const int warp = me >> 5;
  const int key = (mask == 0xffffffffu) ? 32 : __builtin_ctz(mask);
  simt::barrier([=](int i) { return (i >> 5) == warp && ((mask >> (i & 31)) & 1u); },
                [=](simt::Fiber& f) -> uint64_t& { return f.arrivals_warp[key]; });
}
inline double __shfl_sync(unsigned mask, double v, int src, int width = 32) {
  simt::Block* b = simt::cur_block();
  const int me = b->current;
  b->fibers[me].shfl_slot = v;
  __syncwarp(mask);
  const int base = me & ~(width - 1);
  const double r = b->fibers[base + (src & (width - 1))].shfl_slot;
  __syncwarp(mask);
  return r;
}
inline double __shfl_xor_sync(unsigned mask, double v, int lanemask, int width = 32) {
  const int me = simt::cur_block()->current;
  return __shfl_sync(mask, v, (me & (width - 1)) ^ lanemask, width);
}
inline double __shfl_down_sync(unsigned mask, double v, int delta, int width = 32) {
  const int me = simt::cur_block()->current;
  const int l = me & (width - 1);
  return __shfl_sync(mask, v, (l + delta < width) ? l + delta : l, width);
}
inline double __shfl_up_sync(unsigned mask, double v, int delta, int width = 32) {
  const int me = simt::cur_block()->current;
  const int l = me & (width - 1);
  return __shfl_sync(mask, v, (l - delta >= 0) ? l - delta : l, width);
}
inline unsigned __activemask() {
  simt::Block* b = simt::cur_block();
  const int warp = b->current >> 5;
  unsigned m = 0;
  for (int i = 0; i < 32; ++i) {
    const int t = warp * 32 + i;
    if (t < b->nthreads && !b->fibers[t].done) m |= 1u << i;
  }
  return m;
}
inline int __any_sync(unsigned mask, int pred) {
  simt::Block* b = simt::cur_block();
  const int me = b->current;
  b->fibers[me].shfl_slot = pred ? 1.0 : 0.0;
  __syncwarp(mask);
  int any = 0;
  const int base = me & ~31;
  for (int i = 0; i < 32; ++i)
    if (((mask >> i) & 1u) && base + i < b->nthreads && !b->fibers[base + i].done && b->fibers[base + i].shfl_slot != 0.0) any = 1;
  __syncwarp(mask);
  return any;
}
// mma.sync.aligned.m8n8k4.row.col.f64: D(8x8) = A(8x4) B(4x8) + C.  Lane T holds a = A[T/4][T%4], b = B[T%4][T/4],
// c/d = C[T/4][2(T%4) + {0,1}].  All 32 lanes of the warp take part.
struct simt_mma_slots { double a[32], b[32]; };
inline simt_mma_slots& simt_mma(int warp) { static simt_mma_slots s[8]; return s[warp]; }
inline void agx_emul_dmma(double& d0, double& d1, double a, double b, double c0, double c1) {
  simt::Block* blk = simt::cur_block();
  const int me = blk->current, lane = me & 31, warp = me >> 5;
  simt_mma_slots& s = simt_mma(warp);
  s.a[lane] = a; s.b[lane] = b;
  __syncwarp();
  const int g = lane >> 2, q = lane & 3;
  double r0 = c0, r1 = c1;
  for (int k = 0; k < 4; ++k) {
    r0 = std::fma(s.a[g * 4 + k], s.b[(2 * q) * 4 + k], r0);
    r1 = std::fma(s.a[g * 4 + k], s.b[(2 * q + 1) * 4 + k], r1);
  }
  __syncwarp();
  d0 = r0; d1 = r1;
}
inline int atomicAdd(int* p, int v) { const int o = *p; *p = o + v; return o; }
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
inline void sincos(double x, double* s, double* c) { *s = std::sin(x); *c = std::cos(x); }
inline double __ldg(const double* p) { return *p; }
inline int __ldg(const int* p) { return *p; }

#endif  // AGX_CPU_SIMT_H_
