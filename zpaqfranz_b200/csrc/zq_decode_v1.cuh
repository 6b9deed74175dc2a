// zq_decode.cuh -- block decompression on the device: arithmetic decoder + the same predictor as the
// encoder, the unmodeled chunk reader, and the post-processor (PASS or a PCOMP program run by the
// ZPAQL VM with OUT).  One warp per block.
//
// Replaces Decoder::decode/decompress (Z:15282-15332), PostProcessor::write (Z:15368-15414) and the
// data part of Decompresser::decompress (Z:15481-15508); the host parses the block/segment framing
// (Decompresser::findBlock/findFilename/readComment/readSegmentEnd, Z:15418-15534).
#pragma once
#include "zq_cm_v1.cuh"

namespace zqdev {

struct ZqDecUnit {
  u64 data_off;     // first byte of the coded data in the input arena
  u64 data_len;     // bytes available from there to the end of the block
  u64 out_off;      // where the restored bytes go
  u64 model_off;
  u32 out_cap;
  u32 plan;         // index into the ZqCmPlan table (n may be 0 = unmodeled)
};
struct ZqDecResult {
  u32 out_len;      // restored bytes
  u32 consumed;     // coded bytes read, including the 4 end-of-stream zeros
  u32 error;        // 0 ok; 1 corrupted; 2 unexpected end; 3 output overflow; 4 ZPAQL error; 5 bad post-processing type
  u32 pad;
};

struct DecIn {
  const u8* p; u64 len, pos; u32 error;
  __device__ __forceinline__ int get() { if (pos < len) return p[pos++]; error = 2; return 0; }
};

// The post-processor state machine (PostProcessor::write, Z:15368)
struct DecPost {
  u32 state, hsize, loaded;
  CmVm vm;
  u8* code;
  u8* out; u32 outlen, outcap, error;
};

// PCOMP run with OUT (same interpreter as HCOMP, plus the output sink)
__device__ void dec_vm_run(DecPost& pp, u32 input) {
  CmVm& v = pp.vm;
  const u8* __restrict__ P = v.code;
  int pc = 0;
  u32 a = input, b = v.b, c = v.c, d = v.d; int f = v.f;
#define ZQ_MB v.m[b & v.mmask]
#define ZQ_MC v.m[c & v.mmask]
#define ZQ_HD v.h[d & v.hmask]
  for (;;) {
    if (pc < 0 || pc >= v.len) { v.error = 1; break; }
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
      case 57:
        if (pp.outlen < pp.outcap) { if (lane_id() == 0) pp.out[pp.outlen] = (u8)a; } else pp.error = 3;
        ++pp.outlen;
        break;
      case 59: a = (a + ZQ_MB + 512) * 773; break;
      case 60: ZQ_HD = (ZQ_HD + a + 512) * 773; break;
      case 63: pc += ((P[pc] + 128) & 255) - 127; break;
      default: v.error = 1;
    }
    if (v.error) break;
  }
  v.a = a; v.b = b; v.c = c; v.d = d; v.f = f;
#undef ZQ_MB
#undef ZQ_MC
#undef ZQ_HD
}

__device__ __forceinline__ void dec_post_write(DecPost& pp, int c) {
  switch (pp.state) {
    case 0:
      if (c < 0) { pp.error = 2; break; }
      pp.state = (u32)c + 1;
      if (pp.state > 2) pp.error = 5;
      break;
    case 1:
      if (c >= 0) {
        if (pp.outlen < pp.outcap) { if (lane_id() == 0) pp.out[pp.outlen] = (u8)c; } else pp.error = 3;
        ++pp.outlen;
      }
      break;
    case 2: if (c < 0) { pp.error = 2; break; } pp.hsize = (u32)c; pp.state = 3; break;
    case 3:
      if (c < 0) { pp.error = 2; break; }
      pp.hsize += (u32)c * 256u;
      if (pp.hsize < 1) { pp.error = 1; break; }
      pp.loaded = 0; pp.state = 4;
      break;
    case 4:
      if (c < 0) { pp.error = 2; break; }
      pp.code[pp.loaded++] = (u8)c;       // every lane stores the same byte: each reads back its own
      if (pp.loaded == pp.hsize) { pp.vm.code = pp.code; pp.vm.len = (int)pp.hsize; pp.state = 5; }
      break;
    default:
      dec_vm_run(pp, (u32)c);             // c == -1 at end of segment -> a = 0xFFFFFFFF
      break;
  }
}

__global__ void __launch_bounds__(512, 1)
k_cm_decode(const u8* __restrict__ in_base, const ZqDecUnit* __restrict__ units, const ZqCmPlan* __restrict__ cmplans, int nunits,
            const CmTablesDev* __restrict__ tab, const u8* __restrict__ blob, u8* __restrict__ model_base,
            u8* __restrict__ out_base, ZqDecResult* __restrict__ results, u32* __restrict__ next_unit) {
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
    if (t >= nunits) break;
    const ZqDecUnit u = units[t];
    const ZqCmPlan& cp = cmplans[u.plan];
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
    DecPost pp;
    pp.state = 0; pp.hsize = 0; pp.loaded = 0; pp.error = 0;
    pp.out = out_base + u.out_off; pp.outlen = 0; pp.outcap = u.out_cap;
    pp.code = model + cp.pcode_off;
    pp.vm.a = pp.vm.b = pp.vm.c = pp.vm.d = 0; pp.vm.f = 0; pp.vm.error = 0;
    pp.vm.m = model + cp.pm_off; pp.vm.h = (u32*)(model + cp.ph_off); pp.vm.r = (u32*)(model + cp.pr_off);
    pp.vm.mmask = (1u << cp.pm) - 1; pp.vm.hmask = (1u << cp.ph) - 1; pp.vm.code = pp.code; pp.vm.len = 0;
    DecIn in; in.p = in_base + u.data_off; in.len = u.data_len; in.pos = 0; in.error = 0;
    u32 err = 0;
    if (cp.n > 0) {
      u32 low = 1, high = 0xffffffffu, curr = 0;
      for (int k = 0; k < 4; ++k) curr = curr << 8 | (u32)in.get();
      auto decode = [&](u32 p16) -> int {   // Decoder::decode, Z:15282
        if (curr < low || curr > high) { err = 1; return 0; }
        const u32 mid = low + (u32)(((u64)(high - low) * p16) >> 16);
        int y;
        if (curr <= mid) { y = 1; high = mid; } else { y = 0; low = mid + 1; }
        while ((high ^ low) < 0x1000000u) {
          high = high << 8 | 255; low = low << 8; low += (low == 0);
          curr = curr << 8 | (u32)in.get();
        }
        return y;
      };
      for (;;) {
        if (decode(0)) { if (curr != 0 && !err) err = 1; break; }
        int c = 1;
        while (c < 256) {
          const u32 p16 = (u32)cm_predict(L, X, T) * 2 + 1;
          const int y = decode(p16);
          c += c + y;
          cm_update(L, X, T, y, vm);
        }
        dec_post_write(pp, c - 256);
        if (err || in.error || pp.error || vm.error || pp.vm.error) break;
      }
    } else {
      for (;;) {   // unmodeled: big-endian u32 length, that many bytes, ... , length 0
        u32 cl = 0;
        for (int k = 0; k < 4; ++k) cl = cl << 8 | (u32)in.get();
        if (cl == 0 || in.error) break;
        for (u32 k = 0; k < cl; ++k) dec_post_write(pp, in.get());
        if (in.error || pp.error || pp.vm.error) break;
      }
    }
    if (!err && !in.error && !pp.error && !vm.error && !pp.vm.error) dec_post_write(pp, -1);
    if (lane == 0) {
      ZqDecResult r;
      r.out_len = pp.outlen; r.consumed = (u32)in.pos; r.pad = 0;
      r.error = err ? err : in.error ? in.error : pp.error ? pp.error : (vm.error || pp.vm.error) ? 4u : 0u;
      results[t] = r;
    }
  }
}

}  // namespace zqdev
