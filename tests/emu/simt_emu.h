// simt_emu.h -- TEST INFRASTRUCTURE ONLY: a host-side SIMT emulator so the warp-level device code of
// zpaqfranz_b200/csrc/*.cuh can be compiled by g++ and checked against the oracle without a GPU.
// Every CUDA thread of one CTA is a ucontext coroutine; warp collectives (__shfl_sync, __ballot_sync,
// __reduce_*_sync, __syncwarp) and CTA barriers rendezvous the participating coroutines, so divergent
// per-lane control flow runs exactly as written.  Spin waits (__nanosleep) yield to the scheduler.
// Nothing on the product path includes this file; it is pulled in through tests/emu/shim/cuda_runtime.h.
#pragma once
#include <ucontext.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define ZQ_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static
#define __align__(x) alignas(x)
#define __restrict__

struct uint3 { unsigned x, y, z; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
struct uint2 { unsigned x, y; };
inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r; r.x = x; r.y = y; return r; }
struct alignas(16) ulonglong2 { unsigned long long x, y; };
inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { ulonglong2 r; r.x = x; r.y = y; return r; }
struct int2 { int x, y; };
inline int2 make_int2(int x, int y) { int2 r; r.x = x; r.y = y; return r; }

inline uint3 threadIdx, blockIdx, blockDim, gridDim;

namespace emu {
struct Thread { ucontext_t ctx; char* stack = nullptr; bool done = false; unsigned tid = 0; };
struct Warp { uint64_t xchg[32], res[32]; int arrived = 0; uint64_t gen = 0; };
struct NamedBar { unsigned arrived = 0; uint64_t gen = 0; };
inline std::vector<Thread> threads;
inline std::vector<Warp> warps;
inline NamedBar bars[17];   // 16 = __syncthreads
inline ucontext_t sched;
inline Thread* cur = nullptr;
inline unsigned char* dyn_smem = nullptr;
inline std::function<void()> body;
inline uint64_t collectives = 0;

inline void yield() { swapcontext(&cur->ctx, &sched); }
inline void entry() { body(); cur->done = true; swapcontext(&cur->ctx, &sched); }

// all 32 lanes of the calling warp meet; returns the snapshot of everyone's value
inline const uint64_t* rendezvous(uint64_t v) {
  Warp& w = warps[cur->tid >> 5];
  w.xchg[cur->tid & 31] = v;
  ++collectives;
  if (++w.arrived == 32) { memcpy(w.res, w.xchg, sizeof w.res); w.arrived = 0; ++w.gen; }
  else { const uint64_t g = w.gen; while (w.gen == g) yield(); }
  return w.res;
}
inline void barrier(int id, unsigned count) {
  NamedBar& b = bars[id];
  if (++b.arrived == count) { b.arrived = 0; ++b.gen; }
  else { const uint64_t g = b.gen; while (b.gen == g) yield(); }
}

// run `fn` as a kernel: grid x block threads, one CTA at a time
inline void launch(unsigned grid, unsigned block, size_t smem_bytes, std::function<void()> fn) {
  if (block % 32) { fprintf(stderr, "emu: block size must be a multiple of 32\n"); abort(); }
  body = fn;
  const size_t STACK = 512 << 10;
  std::vector<unsigned char> smem(smem_bytes + 64);
  for (unsigned b = 0; b < grid; ++b) {
    dyn_smem = (unsigned char*)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
    memset(smem.data(), 0xA5, smem.size());
    threads.assign(block, Thread());
    warps.assign(block / 32, Warp());
    for (auto& nb : bars) nb = NamedBar();
    for (unsigned t = 0; t < block; ++t) {
      Thread& th = threads[t];
      th.tid = t; th.stack = (char*)malloc(STACK);
      getcontext(&th.ctx);
      th.ctx.uc_stack.ss_sp = th.stack; th.ctx.uc_stack.ss_size = STACK; th.ctx.uc_link = &sched;
      makecontext(&th.ctx, (void (*)())entry, 0);
    }
    blockIdx = {b, 0, 0}; blockDim = {block, 1, 1}; gridDim = {grid, 1, 1};
    for (;;) {
      bool alive = false;
      for (unsigned t = 0; t < block; ++t) {
        if (threads[t].done) continue;
        alive = true;
        cur = &threads[t];
        threadIdx = {t, 0, 0};
        swapcontext(&sched, &cur->ctx);
      }
      if (!alive) break;
    }
    for (auto& th : threads) free(th.stack);
  }
  cur = nullptr;
}
}  // namespace emu

#define ZQ_EMU_FULLMASK(m) do { if ((m) != 0xffffffffu) { fprintf(stderr, "emu: partial mask\n"); abort(); } } while (0)
template <class T> inline uint64_t emu_pack(T v) { uint64_t u = 0; memcpy(&u, &v, sizeof(T)); return u; }
template <class T> inline T emu_unpack(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }

template <class T> inline T __shfl_sync(unsigned m, T v, int src, int width = 32) {
  ZQ_EMU_FULLMASK(m);
  const unsigned lane = emu::cur->tid & 31;
  const uint64_t* r = emu::rendezvous(emu_pack(v));
  const unsigned s = (lane & ~(unsigned)(width - 1)) | ((unsigned)src & (unsigned)(width - 1));
  return emu_unpack<T>(r[s]);
}
template <class T> inline T __shfl_up_sync(unsigned m, T v, unsigned delta, int width = 32) {
  ZQ_EMU_FULLMASK(m);
  const unsigned lane = emu::cur->tid & 31;
  const uint64_t* r = emu::rendezvous(emu_pack(v));
  return ((lane & (unsigned)(width - 1)) >= delta) ? emu_unpack<T>(r[lane - delta]) : v;
}
template <class T> inline T __shfl_down_sync(unsigned m, T v, unsigned delta, int width = 32) {
  ZQ_EMU_FULLMASK(m);
  const unsigned lane = emu::cur->tid & 31;
  const uint64_t* r = emu::rendezvous(emu_pack(v));
  return ((lane & (unsigned)(width - 1)) + delta < (unsigned)width) ? emu_unpack<T>(r[lane + delta]) : v;
}
template <class T> inline T __shfl_xor_sync(unsigned m, T v, int x, int width = 32) {
  ZQ_EMU_FULLMASK(m);
  const unsigned lane = emu::cur->tid & 31;
  const uint64_t* r = emu::rendezvous(emu_pack(v));
  (void)width;
  return emu_unpack<T>(r[lane ^ (unsigned)x]);
}
inline unsigned __ballot_sync(unsigned m, int pred) {
  ZQ_EMU_FULLMASK(m);
  const uint64_t* r = emu::rendezvous(pred ? 1 : 0);
  unsigned b = 0;
  for (int i = 0; i < 32; ++i) b |= (unsigned)(r[i] & 1) << i;
  return b;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
inline void __syncwarp(unsigned m = 0xffffffffu) { ZQ_EMU_FULLMASK(m); emu::rendezvous(0); }
inline void __syncthreads() { emu::barrier(16, blockDim.x); }
inline int __reduce_add_sync(unsigned m, int v) { ZQ_EMU_FULLMASK(m); const uint64_t* r = emu::rendezvous(emu_pack(v)); unsigned s = 0; for (int i = 0; i < 32; ++i) s += emu_unpack<unsigned>(r[i]); return (int)s; }
inline unsigned __reduce_add_sync(unsigned m, unsigned v) { return (unsigned)__reduce_add_sync(m, (int)v); }
inline unsigned __reduce_or_sync(unsigned m, unsigned v) { ZQ_EMU_FULLMASK(m); const uint64_t* r = emu::rendezvous(v); unsigned s = 0; for (int i = 0; i < 32; ++i) s |= (unsigned)r[i]; return s; }
inline unsigned __reduce_and_sync(unsigned m, unsigned v) { ZQ_EMU_FULLMASK(m); const uint64_t* r = emu::rendezvous(v); unsigned s = ~0u; for (int i = 0; i < 32; ++i) s &= (unsigned)r[i]; return s; }
inline unsigned __reduce_min_sync(unsigned m, unsigned v) { ZQ_EMU_FULLMASK(m); const uint64_t* r = emu::rendezvous(v); unsigned s = ~0u; for (int i = 0; i < 32; ++i) s = (unsigned)r[i] < s ? (unsigned)r[i] : s; return s; }
inline unsigned __reduce_max_sync(unsigned m, unsigned v) { ZQ_EMU_FULLMASK(m); const uint64_t* r = emu::rendezvous(v); unsigned s = 0; for (int i = 0; i < 32; ++i) s = (unsigned)r[i] > s ? (unsigned)r[i] : s; return s; }
inline int __reduce_min_sync(unsigned m, int v) { ZQ_EMU_FULLMASK(m); const uint64_t* r = emu::rendezvous(emu_pack(v)); int s = 0x7fffffff; for (int i = 0; i < 32; ++i) { const int x = emu_unpack<int>(r[i]); s = x < s ? x : s; } return s; }
inline int __reduce_max_sync(unsigned m, int v) { ZQ_EMU_FULLMASK(m); const uint64_t* r = emu::rendezvous(emu_pack(v)); int s = -0x7fffffff - 1; for (int i = 0; i < 32; ++i) { const int x = emu_unpack<int>(r[i]); s = x > s ? x : s; } return s; }

// single host thread runs every coroutine: plain read-modify-write is atomic
template <class T, class V> inline T atomicAdd(T* p, V v) { const T o = *p; *p = (T)(o + (T)v); return o; }
template <class T, class V> inline T atomicOr(T* p, V v) { const T o = *p; *p = (T)(o | (T)v); return o; }
template <class T, class V> inline T atomicAnd(T* p, V v) { const T o = *p; *p = (T)(o & (T)v); return o; }
template <class T, class V> inline T atomicMax(T* p, V v) { const T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <class T, class V> inline T atomicMin(T* p, V v) { const T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <class T, class A, class V> inline T atomicCAS(T* p, A cmp, V v) { const T o = *p; if (o == (T)cmp) *p = (T)v; return o; }
template <class T, class V> inline T atomicExch(T* p, V v) { const T o = *p; *p = (T)v; return o; }
inline void __threadfence_block() {}
inline void __threadfence() {}
inline void __nanosleep(unsigned) { emu::yield(); }
inline int __ffs(unsigned x) { return x ? __builtin_ctz(x) + 1 : 0; }
inline int __ffs(int x) { return __ffs((unsigned)x); }
inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
inline int __clz(int x) { return __clz((unsigned)x); }
inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline unsigned __brev(unsigned x) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i); return r; }
inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
  const uint64_t v = (uint64_t)b << 32 | a; unsigned r = 0;
  for (int i = 0; i < 4; ++i) { const unsigned sel = (s >> (4 * i)) & 7; r |= (unsigned)((v >> (8 * sel)) & 255) << (8 * i); }
  return r;
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) { const uint64_t v = (uint64_t)hi << 32 | lo; return (unsigned)(v >> (sh & 31)); }
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) { const uint64_t v = (uint64_t)hi << 32 | lo; return (unsigned)((v << (sh & 31)) >> 32); }
inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
template <class T> inline T __ldg(const T* p) { return *p; }
#ifndef INT_MIN
#include <climits>
#endif
template <class A, class B> inline auto min(A a, B b) -> decltype(a + b) { typedef decltype(a + b) R; return (R)a < (R)b ? (R)a : (R)b; }
template <class A, class B> inline auto max(A a, B b) -> decltype(a + b) { typedef decltype(a + b) R; return (R)a > (R)b ? (R)a : (R)b; }
