// zq_cm.cuh -- the ZPAQ context-mixing compressor (methods -m3/-m4/-m5 and any explicit model of
// <= 32 components): HCOMP virtual machine, component chain, logistic mixing and the binary
// arithmetic coder.  One WARP per block, LANE i = COMPONENT i.
//
// Replaces (bit-exactly): ZPAQL::run0/execute Z:14232-14467, Predictor::predict0 Z:15041,
// update0 Z:15139, find Z:15254, train Z:13161, Encoder::encode/compress Z:15557-15589.
//
// Why this mapping: coding is a strict per-bit recurrence (the next probability needs the previous
// bit's updates), so a block cannot be split; but inside one bit every component does an independent
// dependent-gather (hash row -> bit history -> adaptive probability) into its own tables.  On the
// CPU those are n serial cache misses per bit; here the n lanes issue them at once and only the
// final arithmetic (ISSE chains, MIX dot products, SSE interpolation) is resolved in dependency
// levels with warp shuffles / one REDUX per mixer.  Component tables live in HBM (up to ~85 MB per
// block for -m5); the 78 KB of model-independent tables (stretch, squash, dt, state table) are
// staged into shared memory once per CTA.  ICM/ISSE hash rows (16 B) are held in registers for the
// four bits of a nibble and written back on the next row switch.
#pragma once
#include "zq_cm_types.h"
#include "zq_common.cuh"

namespace zqdev {

struct CmTablesDev {      // device mirror of zq::CmTables (same field order)
  int16_t stretch[32768];
  u16 squash[4096];
  int dt[1024];
  int dt2k[256];
  u8 ns[1024];
  u32 icm_init[256];
  u32 isse_init[512];
};
struct CmSmem {           // what the coder needs per bit
  int16_t stretch[32768];
  u16 squash[4096];
  int dt[1024];
  int dt2k[256];
  u8 ns[1024];
};

// ---- model initialisation ------------------------------------------------------------------------
// One CTA writes one fill job (a table's initial values, Z:14968-15032) of one unit's model region.
__device__ __forceinline__ void cm_fill_job(const ZqCmFill f, const CmTablesDev* __restrict__ tab, u8* __restrict__ region) {
  u8* __restrict__ dst = region + f.off;
  const u64 nq = f.bytes >> 4;  // regions are 256 B aligned; tails shorter than 16 B are written bytewise
  uint4* __restrict__ d4 = (uint4*)dst;
  if (f.kind == ZQ_FILL_ZERO || f.kind == ZQ_FILL_U32 || f.kind == ZQ_FILL_U16 || f.kind == ZQ_FILL_MATCHBUF) {
    u32 v = f.kind == ZQ_FILL_U32 ? f.value : f.kind == ZQ_FILL_U16 ? (f.value | f.value << 16) : 0u;
    const uint4 q = make_uint4(v, v, v, v);
    for (u64 k = threadIdx.x; k < nq; k += blockDim.x) d4[k] = q;
    for (u64 k = (nq << 4) + threadIdx.x; k < f.bytes; k += blockDim.x) dst[k] = (u8)(v >> (8 * (k & 3)));
    if (f.kind == ZQ_FILL_MATCHBUF) { __syncthreads(); if (threadIdx.x == 0) dst[0] = 1; }
  } else if (f.kind == ZQ_FILL_SSE) {
    // cm[j] = squash((j&31)*64-992)<<17 | start : period of 32 entries = 8 uint4
    for (u64 k = threadIdx.x; k < nq; k += blockDim.x) {
      const u32 e = (u32)(k & 7) * 4;
      uint4 q;
      q.x = (u32)tab->squash[(e + 0) * 64 - 992 + 2048] << 17 | f.value;
      q.y = (u32)tab->squash[(e + 1) * 64 - 992 + 2048] << 17 | f.value;
      q.z = (u32)tab->squash[(e + 2) * 64 - 992 + 2048] << 17 | f.value;
      q.w = (u32)tab->squash[(e + 3) * 64 - 992 + 2048] << 17 | f.value;
      d4[k] = q;
    }
  } else {
    const u32* src = f.kind == ZQ_FILL_ICM ? tab->icm_init : tab->isse_init;
    u32* d = (u32*)dst;
    for (u32 k = threadIdx.x; k < f.bytes / 4; k += blockDim.x) d[k] = src[k];
  }
}

// grid.x = units in the wave * fill jobs of the largest plan
__global__ void __launch_bounds__(256)
k_cm_init(const ZqUnit* __restrict__ units, const ZqPlan* __restrict__ plans, const ZqCmPlan* __restrict__ cmplans,
          const ZqCmFill* __restrict__ fills, const int* __restrict__ todo, int ntodo, int maxjobs,
          const CmTablesDev* __restrict__ tab, u8* __restrict__ model_base) {
  const int t = blockIdx.x / maxjobs, j = blockIdx.x % maxjobs;
  if (t >= ntodo) return;
  const ZqUnit u = units[todo[t]];
  const ZqCmPlan& cp = cmplans[plans[u.plan].cm_plan];
  if (j >= (int)cp.fill_count) return;
  cm_fill_job(fills[cp.fill_first + j], tab, model_base + u.model_off);
}
// same, for explicit (model offset, plan) pairs (decoder)
__global__ void __launch_bounds__(256)
k_cm_init_pairs(const u64* __restrict__ model_off, const u32* __restrict__ plan_of, const ZqCmPlan* __restrict__ cmplans,
                const ZqCmFill* __restrict__ fills, int n, int maxjobs, const CmTablesDev* __restrict__ tab, u8* __restrict__ model_base) {
  const int t = blockIdx.x / maxjobs, j = blockIdx.x % maxjobs;
  if (t >= n) return;
  const ZqCmPlan& cp = cmplans[plan_of[t]];
  if (j >= (int)cp.fill_count) return;
  cm_fill_job(fills[cp.fill_first + j], tab, model_base + model_off[t]);
}

// ---- per-warp coder state ------------------------------------------------------------------------
struct CmCoder {
  u32 low, high;
  u8* out; u8* end; u32 overflow;
  __device__ __forceinline__ void init(u8* o, u32 cap) { low = 1; high = 0xffffffffu; out = o; end = o + cap; overflow = 0; }
  __device__ __forceinline__ void encode(int y, u32 p16) {   // Encoder::encode, Z:15557
    const u32 mid = low + (u32)(((u64)(high - low) * p16) >> 16);
    if (y) high = mid; else low = mid + 1;
    while ((high ^ low) < 0x1000000u) {
      if (out < end) { if (lane_id() == 0) *out = (u8)(high >> 24); } else overflow = 1;
      ++out;
      high = high << 8 | 255;
      low = low << 8;
      low += (low == 0);
    }
  }
};

struct CmVm {   // HCOMP machine; registers are warp-uniform
  u32 a, b, c, d; int f;
  u8* m; u32* h; u32* r;
  u32 mmask, hmask;
  const u8* code; int len;
  int error;
};

// One HCOMP run (Z:14232). Every lane executes the same instruction stream; stores are issued by all
// lanes with identical address and value, so each lane later reads back what it wrote itself.
__device__ void cm_vm_run(CmVm& v, u32 input) {
  const u8* __restrict__ P = v.code;
  int pc = 0;
  u32 a = input, b = v.b, c = v.c, d = v.d; int f = v.f;
#define ZQ_MB v.m[b & v.mmask]
#define ZQ_MC v.m[c & v.mmask]
#define ZQ_HD v.h[d & v.hmask]
  for (;;) {
    if (pc >= v.len) { v.error = 1; break; }
    const int op = P[pc++];
    if (op == 56) break;
    if (op >= 64) {
      if (op == 255) { pc = P[pc] + 256 * P[pc + 1]; continue; }
      const int src = op & 7, grp = op >> 3;
      u32 x;
      switch (src) {
        case 0: x = a; break; case 1: x = b; break; case 2: x = c; break; case 3: x = d; break;
        case 4: x = ZQ_MB; break; case 5: x = ZQ_MC; break; case 6: x = ZQ_HD; break;
        default: x = P[pc++];
      }
      switch (grp) {
        case 8: a = x; break; case 9: b = x; break; case 10: c = x; break; case 11: d = x; break;
        case 12: ZQ_MB = (u8)x; break; case 13: ZQ_MC = (u8)x; break; case 14: ZQ_HD = x; break;
        case 16: a += x; break; case 17: a -= x; break; case 18: a *= x; break;
        case 19: a = x ? a / x : 0; break; case 20: a = x ? a % x : 0; break;
        case 21: a &= x; break; case 22: a &= ~x; break; case 23: a |= x; break; case 24: a ^= x; break;
        case 25: a <<= (x & 31); break; case 26: a >>= (x & 31); break;
        case 27: f = a == x; break; case 28: f = a < x; break; case 29: f = a > x; break;
        default: v.error = 1;
      }
      if (v.error) break;
      continue;
    }
    switch (op) {
      case 1: ++a; break; case 2: --a; break; case 3: a = ~a; break; case 4: a = 0; break;
      case 7: a = v.r[P[pc++]]; break;
      case 8: { const u32 t = a; a = b; b = t; } break;
      case 9: ++b; break; case 10: --b; break; case 11: b = ~b; break; case 12: b = 0; break;
      case 15: b = v.r[P[pc++]]; break;
      case 16: { const u32 t = a; a = c; c = t; } break;
      case 17: ++c; break; case 18: --c; break; case 19: c = ~c; break; case 20: c = 0; break;
      case 23: c = v.r[P[pc++]]; break;
      case 24: { const u32 t = a; a = d; d = t; } break;
      case 25: ++d; break; case 26: --d; break; case 27: d = ~d; break; case 28: d = 0; break;
      case 31: d = v.r[P[pc++]]; break;
      case 32: { const u8 t = ZQ_MB; ZQ_MB = (u8)a; a = (a & ~255u) | t; } break;
      case 33: ZQ_MB = ZQ_MB + 1; break; case 34: ZQ_MB = ZQ_MB - 1; break; case 35: ZQ_MB = ~ZQ_MB; break; case 36: ZQ_MB = 0; break;
      case 39: if (f) pc += ((P[pc] + 128) & 255) - 127; else ++pc; break;
      case 40: { const u8 t = ZQ_MC; ZQ_MC = (u8)a; a = (a & ~255u) | t; } break;
      case 41: ZQ_MC = ZQ_MC + 1; break; case 42: ZQ_MC = ZQ_MC - 1; break; case 43: ZQ_MC = ~ZQ_MC; break; case 44: ZQ_MC = 0; break;
      case 47: if (!f) pc += ((P[pc] + 128) & 255) - 127; else ++pc; break;
      case 48: { const u32 t = ZQ_HD; ZQ_HD = a; a = t; } break;
      case 49: ZQ_HD = ZQ_HD + 1; break; case 50: ZQ_HD = ZQ_HD - 1; break; case 51: ZQ_HD = ~ZQ_HD; break; case 52: ZQ_HD = 0; break;
      case 55: v.r[P[pc++]] = a; break;
      case 57: break;   // OUT: HCOMP has no output
      case 59: a = (a + ZQ_MB + 512) * 773; break;
      case 60: ZQ_HD = (ZQ_HD + a + 512) * 773; break;
      case 63: pc += ((P[pc] + 128) & 255) - 127; break;
      default: v.error = 1;
    }
    if (v.error || pc < 0) { v.error = 1; break; }
  }
  v.a = a; v.b = b; v.c = c; v.d = d; v.f = f;
#undef ZQ_MB
#undef ZQ_MC
#undef ZQ_HD
}

__device__ __forceinline__ int cm_clamp2k(int x) { return min(max(x, -2048), 2047); }
__device__ __forceinline__ int cm_clamp512k(int x) { return min(max(x, -(1 << 19)), (1 << 19) - 1); }

__device__ __forceinline__ u32 row_get(const uint4& r, u32 idx) {
  const u32 w = idx < 8 ? (idx < 4 ? r.x : r.y) : (idx < 12 ? r.z : r.w);
  return (w >> ((idx & 3) * 8)) & 255u;
}
__device__ __forceinline__ void row_set(uint4& r, u32 idx, u32 v) {
  const u32 sh = (idx & 3) * 8, msk = ~(255u << sh), val = v << sh;
  if (idx < 4) r.x = (r.x & msk) | val;
  else if (idx < 8) r.y = (r.y & msk) | val;
  else if (idx < 12) r.z = (r.z & msk) | val;
  else r.w = (r.w & msk) | val;
}

// Everything one lane knows about its component.
struct CmLane {
  u32 type, a1, a2, a3, a4, a5, level;
  u32 in1, in2;          // lanes of the inputs (AVG j,k / MIX2 j,k / ISSE j / SSE j)
  u32* cm; u8* ht; u32 cm_mask, ht_mask;
  u32 h;                 // context hash H[i]
  int p;                 // stretched prediction of this component
  u32 cxt, limit, ca, cb, cc;
  u32 pn;                // table entry fetched for predict (trained in update)
  int w0, w1, pj, pk;    // ISSE weights / MIX2 weight in w0; inputs seen at predict time
  uint4 row; u32 rowpos; bool rowok;
};

// Predictor::find (Z:15254) on this lane's hash table + switch of the register-cached row
__device__ __forceinline__ void cm_row_switch(CmLane& L, u32 cxt) {
  if (L.rowok) *(uint4*)(L.ht + L.rowpos) = L.row;
  const int sizebits = (int)L.a1 + 2;
  const u32 chk = (cxt >> sizebits) & 255u;
  const u32 h0 = (cxt * 16u) & (L.ht_mask - 15u), h1 = h0 ^ 16u, h2 = h0 ^ 32u;
  const u32 w0 = *(const u32*)(L.ht + h0), w1 = *(const u32*)(L.ht + h1), w2 = *(const u32*)(L.ht + h2);
  u32 r; bool fresh = false;
  if ((w0 & 255u) == chk) r = h0;
  else if ((w1 & 255u) == chk) r = h1;
  else if ((w2 & 255u) == chk) r = h2;
  else {
    const u32 p0 = (w0 >> 8) & 255u, p1 = (w1 >> 8) & 255u, p2 = (w2 >> 8) & 255u;
    if (p0 <= p1 && p0 <= p2) r = h0; else if (p1 < p2) r = h1; else r = h2;
    fresh = true;
  }
  L.row = fresh ? make_uint4(chk, 0, 0, 0) : *(const uint4*)(L.ht + r);
  L.rowpos = r; L.rowok = true; L.cc = r;
}

struct CmCtx {             // warp-uniform per-block state
  int n, nlevels, c8, hmap4;
  u32 mix_mask;            // lanes that are MIX components
  u32 mix_levels;          // bit L: some MIX sits at dependency level L
};

// p for the next bit: Predictor::predict0 (Z:15041)
__device__ int cm_predict(CmLane& L, const CmCtx& X, const CmSmem& T) {
  const u32 lane = lane_id();
  const int c8 = X.c8, hmap4 = X.hmap4;
  const bool nib = c8 == 1 || (c8 & 0xf0) == 16;
  switch (L.type) {
    case ZQ_CM:
      L.cxt = (L.h ^ (u32)hmap4) & L.cm_mask;
      L.pn = L.cm[L.cxt];
      L.p = T.stretch[L.pn >> 17];
      break;
    case ZQ_ICM:
      if (nib) cm_row_switch(L, L.h + 16u * (u32)c8);
      L.cxt = row_get(L.row, hmap4 & 15);
      L.pn = L.cm[L.cxt];
      L.p = T.stretch[L.pn >> 8];
      break;
    case ZQ_ISSE:
      if (nib) cm_row_switch(L, L.h + 16u * (u32)c8);
      L.cxt = row_get(L.row, hmap4 & 15);
      L.w0 = (int)L.cm[L.cxt * 2]; L.w1 = (int)L.cm[L.cxt * 2 + 1];
      break;
    case ZQ_MATCH:
      if (L.ca == 0) L.p = 0;
      else {
        L.cc = (L.ht[(L.limit - L.cb) & L.ht_mask] >> (7 - L.cxt)) & 1u;
        L.p = T.stretch[(T.dt2k[L.ca] * (L.cc ? -1 : 1)) & 32767];
      }
      break;
    case ZQ_MIX2:
      L.cxt = (L.h + ((u32)c8 & L.a5)) & L.cm_mask;
      L.w0 = ((const u16*)L.cm)[L.cxt];
      break;
    default: break;
  }
  for (int lev = 1; lev < X.nlevels; ++lev) {
    const int pj = __shfl_sync(ZQ_FULL, L.p, L.in1), pk = __shfl_sync(ZQ_FULL, L.p, L.in2);
    if ((int)L.level == lev) {
      switch (L.type) {
        case ZQ_AVG: L.p = (pj * (int)L.a3 + pk * (256 - (int)L.a3)) >> 8; break;
        case ZQ_MIX2: L.pj = pj; L.pk = pk; L.p = (L.w0 * pj + (65536 - L.w0) * pk) >> 16; break;
        case ZQ_ISSE: L.pj = pj; L.p = cm_clamp2k((L.w0 * pj + L.w1 * 64) >> 16); break;
        case ZQ_SSE: {
          u32 cx = (L.h + (u32)c8) * 32u;
          int pq = min(max(pj + 992, 0), 1983);
          const int wt = pq & 63;
          pq >>= 6;
          cx += (u32)pq;
          const u32 e0 = L.cm[cx & L.cm_mask], e1 = L.cm[(cx + 1) & L.cm_mask];
          L.p = T.stretch[((e0 >> 10) * (u32)(64 - wt) + (e1 >> 10) * (u32)wt) >> 13];
          cx += (u32)(wt >> 5);
          L.cxt = cx & L.cm_mask;
          L.pn = (wt >> 5) ? e1 : e0;
          break;
        }
        default: break;
      }
    }
    // mixers of this level: lanes j0..j0+m-1 each multiply their own p by their weight, one REDUX sums
    u32 mm = ((X.mix_levels >> lev) & 1u) ? __ballot_sync(ZQ_FULL, L.type == ZQ_MIX && (int)L.level == lev) : 0u;
    while (mm) {
      const int xl = __ffs(mm) - 1;
      mm &= mm - 1;
      const u32 hX = __shfl_sync(ZQ_FULL, L.h, xl), j0 = __shfl_sync(ZQ_FULL, L.a2, xl), m = __shfl_sync(ZQ_FULL, L.a3, xl);
      const u32 msk = __shfl_sync(ZQ_FULL, L.a5, xl), rmask = __shfl_sync(ZQ_FULL, L.cm_mask, xl);
      const u64 base = __shfl_sync(ZQ_FULL, (u64)(uintptr_t)L.cm, xl);
      const u32 row = ((hX + ((u32)c8 & msk)) & rmask) * m;
      const bool mine = lane >= j0 && lane < j0 + m;
      int term = 0;
      if (mine) term = (((const int*)(uintptr_t)base)[row + lane - j0] >> 8) * L.p;
      const int sum = __reduce_add_sync(ZQ_FULL, term);
      if ((int)lane == xl) { L.p = cm_clamp2k(sum >> 8); L.cxt = row; }
    }
  }
  return T.squash[__shfl_sync(ZQ_FULL, L.p, X.n - 1) + 2048];
}

// train all components on bit y, advance the bit context; Predictor::update0 (Z:15139).
// Returns true when a byte was completed (HCOMP must run).
__device__ void cm_update(CmLane& L, CmCtx& X, const CmSmem& T, int y, CmVm& vm) {
  const u32 lane = lane_id();
  // mixers first need every lane's p as seen at predict time: p is unchanged until the next predict
  u32 mm = X.mix_mask;
  while (mm) {
    const int xl = __ffs(mm) - 1;
    mm &= mm - 1;
    const int pX = __shfl_sync(ZQ_FULL, L.p, xl);
    const u32 j0 = __shfl_sync(ZQ_FULL, L.a2, xl), m = __shfl_sync(ZQ_FULL, L.a3, xl), rate = __shfl_sync(ZQ_FULL, L.a4, xl);
    const u32 row = __shfl_sync(ZQ_FULL, L.cxt, xl);
    const u64 base = __shfl_sync(ZQ_FULL, (u64)(uintptr_t)L.cm, xl);
    const int err = ((y * 32767 - (int)T.squash[pX + 2048]) * (int)rate) >> 4;
    if (lane >= j0 && lane < j0 + m) {
      int* wp = (int*)(uintptr_t)base + row + lane - j0;
      *wp = cm_clamp512k(*wp + ((err * L.p + (1 << 12)) >> 13));
    }
  }
  switch (L.type) {
    case ZQ_CM: case ZQ_SSE: {
      const u32 count = L.pn & 0x3ffu;
      const int error = y * 32767 - (int)(L.pn >> 17);
      L.pn += (u32)((error * T.dt[count]) & -1024) + (count < L.limit ? 1u : 0u);
      L.cm[L.cxt] = L.pn;
      break;
    }
    case ZQ_ICM: {
      row_set(L.row, X.hmap4 & 15, T.ns[L.cxt * 4 + y]);
      L.pn += (u32)(((int)(y * 32767 - (int)(L.pn >> 8))) >> 2);
      L.cm[L.cxt] = L.pn;
      break;
    }
    case ZQ_ISSE: {
      const int err = y * 32767 - (int)T.squash[L.p + 2048];
      L.cm[L.cxt * 2] = (u32)cm_clamp512k(L.w0 + ((err * L.pj + (1 << 12)) >> 13));
      L.cm[L.cxt * 2 + 1] = (u32)cm_clamp512k(L.w1 + ((err + 16) >> 5));
      row_set(L.row, X.hmap4 & 15, T.ns[L.cxt * 4 + y]);
      break;
    }
    case ZQ_MIX2: {
      const int err = ((y * 32767 - (int)T.squash[L.p + 2048]) * (int)L.a4) >> 5;
      int w = L.w0 + ((err * (L.pj - L.pk) + (1 << 12)) >> 13);
      w = min(max(w, 0), 65535);
      ((u16*)L.cm)[L.cxt] = (u16)w;
      break;
    }
    case ZQ_MATCH: {
      const u32 hm = L.ht_mask;
      if ((int)L.cc != y) L.ca = 0;
      L.ht[L.limit & hm] = (u8)(L.ht[L.limit & hm] * 2 + y);
      if (++L.cxt == 8) {
        L.cxt = 0;
        L.limit = (L.limit + 1) & hm;
        if (L.ca == 0) {
          L.cb = L.limit - L.cm[L.h & L.cm_mask];
          if (L.cb & hm)
            while (L.ca < 255 && L.ht[(L.limit - L.ca - 1) & hm] == L.ht[(L.limit - L.ca - L.cb - 1) & hm]) ++L.ca;
        } else L.ca += L.ca < 255;
        L.cm[L.h & L.cm_mask] = L.limit;
      }
      break;
    }
    default: break;
  }
  X.c8 += X.c8 + y;
  if (X.c8 >= 256) {
    if (lane == 0) cm_vm_run(vm, (u32)(X.c8 - 256));   // one lane owns the machine: no same-address races
    __syncwarp();
    X.hmap4 = 1; X.c8 = 1;
    if ((int)lane < X.n) L.h = vm.h[lane & vm.hmask];
  } else if (X.c8 >= 16 && X.c8 < 32)
    X.hmap4 = (X.hmap4 & 0xf) << 5 | y << 4 | 1;
  else
    X.hmap4 = (X.hmap4 & 0x1f0) | (((X.hmap4 & 0xf) * 2 + y) & 0xf);
}

__device__ __forceinline__ void cm_code_byte(CmCoder& E, CmLane& L, CmCtx& X, const CmSmem& T, CmVm& vm, u32 c) {
  E.encode(0, 0);
  for (int i = 7; i >= 0; --i) {
    const u32 p16 = (u32)cm_predict(L, X, T) * 2 + 1;
    const int y = (c >> i) & 1;
    E.encode(y, p16);
    cm_update(L, X, T, y, vm);
  }
}

// grid of persistent 16-warp CTAs; warps pull modeled units from a counter
template <int MINB>
__global__ void __launch_bounds__(512, MINB)
k_cm_encode(const u8* __restrict__ in_base, const ZqUnit* __restrict__ units, const ZqPlan* __restrict__ plans,
            const ZqCmPlan* __restrict__ cmplans, const int* __restrict__ todo, int ntodo,
            const CmTablesDev* __restrict__ tab, const u8* __restrict__ blob, const u8* __restrict__ lz_base,
            const u32* __restrict__ lz_len, u8* __restrict__ model_base, u8* __restrict__ coded_base,
            u32* __restrict__ coded_len, u32* __restrict__ err_flag, u32* __restrict__ next_unit) {
  ZQ_DYN_SMEM(smem_raw);
  CmSmem& T = *reinterpret_cast<CmSmem*>(smem_raw);
  {
    const uint4* src = (const uint4*)tab;
    uint4* dst = (uint4*)smem_raw;
    for (u32 k = threadIdx.x; k < sizeof(CmSmem) / 16; k += blockDim.x) dst[k] = src[k];
  }
  __syncthreads();
  const u32 lane = lane_id();
  for (;;) {
    int t = 0;
    if (lane == 0) t = (int)atomicAdd(next_unit, 1u);
    t = __shfl_sync(ZQ_FULL, t, 0);
    if (t >= ntodo) break;
    const int ui = todo[t];
    const ZqUnit u = units[ui];
    const ZqPlan pl = plans[u.plan];
    const ZqCmPlan& cp = cmplans[pl.cm_plan];
    u8* model = model_base + u.model_off;
    CmCtx X; X.n = cp.n; X.nlevels = cp.nlevels; X.c8 = 1; X.hmap4 = 1; X.mix_mask = cp.mix_mask;
    CmLane L;
    {
      const ZqCmComp c = cp.comp[lane < (u32)cp.n ? lane : 0];
      const bool act = lane < (u32)cp.n;
      L.type = act ? c.type : 0; L.a1 = c.a1; L.a2 = c.a2; L.a3 = c.a3; L.a4 = c.a4; L.a5 = c.a5; L.level = act ? c.level : 255;
      X.mix_levels = __reduce_or_sync(ZQ_FULL, (act && c.type == ZQ_MIX) ? (1u << c.level) : 0u);
      L.cm = (u32*)(model + c.cm_off); L.ht = model + c.ht_off; L.cm_mask = c.cm_mask; L.ht_mask = c.ht_mask;
      L.in1 = 0; L.in2 = 0;
      if (L.type == ZQ_AVG) { L.in1 = c.a1; L.in2 = c.a2; }
      else if (L.type == ZQ_MIX2) { L.in1 = c.a2; L.in2 = c.a3; }
      else if (L.type == ZQ_ISSE || L.type == ZQ_SSE) L.in1 = c.a2;
      L.h = 0; L.p = L.type == ZQ_CONS ? ((int)c.a1 - 128) * 4 : 0;
      L.cxt = 0; L.ca = L.cb = L.cc = 0; L.pn = 0; L.w0 = L.w1 = L.pj = L.pk = 0;
      L.limit = L.type == ZQ_CM ? c.a2 * 4u : L.type == ZQ_SSE ? c.a4 * 4u : L.type == ZQ_ICM ? 1023u : 0u;
      L.row = make_uint4(0, 0, 0, 0); L.rowpos = 0; L.rowok = false;
    }
    CmVm vm;
    vm.a = vm.b = vm.c = vm.d = 0; vm.f = 0; vm.error = 0;
    vm.m = model + cp.m_off; vm.h = (u32*)(model + cp.h_off); vm.r = (u32*)(model + cp.r_off);
    vm.mmask = (1u << cp.hm) - 1; vm.hmask = (1u << cp.hh) - 1;
    vm.code = blob + cp.hcomp_off; vm.len = (int)cp.hcomp_len;
    CmCoder E; E.init(coded_base + u.coded_off, u.coded_cap);
    // the coded stream: post-processor selector (+ PCOMP bytes), then the (pre-processed) data
    for (u32 k = 0; k < pl.payload_len; ++k) cm_code_byte(E, L, X, T, vm, blob[pl.payload_off + k]);
    const u8* __restrict__ stream; u32 slen;
    if (pl.lz_level) { stream = lz_base + u.lz_off; slen = lz_len[ui]; }
    else { stream = in_base + u.in_off; slen = u.n; }
    for (u32 k = 0; k < slen; ++k) cm_code_byte(E, L, X, T, vm, stream[k]);
    E.encode(1, 0);   // end of segment
    if (lane == 0) {
      coded_len[ui] = (u32)(E.out - (coded_base + u.coded_off));
      if (E.overflow) atomicOr(err_flag, 1u);
      if (vm.error) atomicOr(err_flag, 2u);
    }
  }
}

}  // namespace zqdev
