// zq_decode.cuh -- block decompression on the device: arithmetic decoder + the same predictor as the
// encoder (zq_cm.cuh; here one warp runs coder and context machine, the machine on lane 0), the
// unmodeled chunk reader, and the post-processor (PASS or a PCOMP program run by the ZPAQL VM with OUT).
// One warp per block.
//
// Replaces Decoder::decode/decompress (Z:15282-15332), PostProcessor::write (Z:15368-15414) and the
// data part of Decompresser::decompress (Z:15481-15508); the host parses the block/segment framing
// (Decompresser::findBlock/findFilename/readComment/readSegmentEnd, Z:15418-15534).
#pragma once
#include "zq_cm.cuh"

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
  u32 error;        // 0 ok; 1 corrupted; 2 unexpected end; 3 output overflow; 4 ZPAQL error; 5 bad post-processing type; 6 segment table full
  u32 pad;
};

struct ZqDecSeg {     // one record per segment that is FOLLOWED by another one in its block (the last one is in ZqDecResult)
  u32 unit;         // index of the block in the launch
  u32 out_end;      // restored bytes of the block up to the end of this segment
  u32 trailer;      // offset (from data_off) of this segment's 253 / 254 byte
  u32 pad;
};

struct DecIn {
  const u8* p; u64 len, pos; u32 error;
  __device__ __forceinline__ int get() { if (pos < len) return p[pos++]; error = 2; return 0; }
};

// After an end of stream: does another segment follow in this block (Decompresser::readSegmentEnd + findFilename +
// readComment, Z:15509-15534, Z:15440-15475)?  The trailer is 254, or 253 and a 20-byte SHA-1; a following segment
// opens with 1, filename, 0, comment, 0, reserved 0.  Leaves `in` at the trailer when the block ends here (the host
// reads it), else at the next segment's first coded byte and returns the trailer's position.
__device__ __forceinline__ bool dec_next_segment(DecIn& in, u32& trailer) {
  if (in.pos >= in.len) return false;
  const u32 c = in.p[in.pos];
  const u64 nxt = in.pos + (c == 253 ? 21 : 1);
  if ((c != 253 && c != 254) || nxt >= in.len || in.p[nxt] != 1) return false;
  trailer = (u32)in.pos;
  in.pos = nxt + 1;
  while (in.get() != 0 && !in.error) {}
  while (in.get() != 0 && !in.error) {}
  if (in.get() != 0 && !in.error) in.error = 1;
  return !in.error;
}
__device__ __forceinline__ void dec_record_segment(ZqDecSeg* segs, u32* nseg, u32 segcap, u32 unit, u32 out_end, u32 trailer, u32& err) {
  const u32 k = atomicAdd(nseg, 1u);
  if (k < segcap) { ZqDecSeg r; r.unit = unit; r.out_end = out_end; r.trailer = trailer; r.pad = 0; segs[k] = r; }
  else err = 6;
}

// The post-processor state machine (PostProcessor::write, Z:15368)
struct DecPost {
  u32 state, hsize, loaded;
  CmVm vm;          // PCOMP machine, owned by lane 0
  u8* code;
  VmOut o;          // output sink: lane 0 writes, the count is re-broadcast after every run
  u32 error;
};

// PostProcessor::write (Z:15368). Machine state and the output count live on lane 0; `outlen`/errors are
// re-broadcast after every step so the warp's control flow stays uniform.
// SOLO: the whole warp's work is being done by the calling lane alone (chain fast path): no lane tests, no shuffles.
template <int VM, bool SOLO>
__device__ __forceinline__ void dec_post_write(DecPost& pp, int c) {
  const u32 lane = SOLO ? 0u : lane_id();
  switch (pp.state) {
    case 0:
      if (c < 0) { pp.error = 2; break; }
      pp.state = (u32)c + 1;
      if (pp.state > 2) pp.error = 5;
      break;
    case 1:
      if (c >= 0) {
        if (pp.o.len < pp.o.cap) { if (lane == 0) pp.o.out[pp.o.len] = (u8)c; } else pp.error = 3;
        ++pp.o.len;
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
      if (lane == 0) pp.code[pp.loaded] = (u8)c;   // read back by lane 0 only (it runs the machine)
      ++pp.loaded;
      if (pp.loaded == pp.hsize) { pp.vm.code = pp.code; pp.vm.len = (int)pp.hsize; pp.state = 5; }
      break;
    default:
      if (lane == 0) cm_vm_run<VM, true>(pp.vm, (u32)c, &pp.o);   // c == -1 at end of segment -> a = 0xFFFFFFFF
      if (!SOLO) {
        pp.o.len = __shfl_sync(ZQ_FULL, pp.o.len, 0);
        pp.o.error = __shfl_sync(ZQ_FULL, pp.o.error, 0);
        pp.vm.error = __shfl_sync(ZQ_FULL, pp.vm.error, 0);
      }
      if (pp.o.error) pp.error = pp.o.error;
      break;
  }
}

// ---- decoder fast path for the chain ICM -> ISSE (see zq_cm.cuh): one lane, straight-line per bit ----------
struct ChainDec {
  u32 low, high, curr, err;
  DecIn in;
  __device__ __forceinline__ int bit(u32 p16) {   // Decoder::decode, Z:15282
    if (curr < low || curr > high) { err = 1; return 0; }
    const u32 mid = low + (u32)(((u64)(high - low) * p16) >> 16);
    int y;
    if (curr <= mid) { y = 1; high = mid; } else { y = 0; low = mid + 1; }
    while ((high ^ low) < 0x1000000u) {
      high = high << 8 | 255; low = low << 8; low += (low == 0);
      curr = curr << 8 | (u32)in.get();
    }
    return y;
  }
};
// one bit: node slot `k` of row words (wa of the ICM, wb of the ISSE); returns y and updates tables and rows
__device__ __forceinline__ u32 chain1_dec_bit(ChainComp& A, ChainComp& B, u32& wa, u32& wb, u32 k, ChainDec& D, const CmSmem& T) {
  const u32 sa = chain_get(wa, k), sb = chain_get(wb, k);
  u32 pn = A.cm[sa];
  int2 w = *(const int2*)(B.cm + sb * 2);
  const int pa = T.stretch[pn >> 8];
  const int pb = cm_clamp2k((w.x * pa + w.y * 64) >> 16);
  const int pr = T.squash[pb + 2048];
  const u32 y = (u32)D.bit((u32)pr * 2 + 1);
  pn += (u32)(((int)(y * 32767u) - (int)(pn >> 8)) >> 2);
  A.cm[sa] = pn;
  const int err = (int)(y * 32767u) - pr;
  w.x = cm_clamp512k(w.x + ((err * pa + (1 << 12)) >> 13));
  w.y = cm_clamp512k(w.y + ((err + 16) >> 5));
  *(int2*)(B.cm + sb * 2) = w;
  wa = chain_put(wa, k, T.ns[sa * 4 + y]);
  wb = chain_put(wb, k, T.ns[sb * 4 + y]);
  return y;
}
__device__ __forceinline__ u32 chain1_dec_nibble(ChainComp& A, ChainComp& B, ChainDec& D, const CmSmem& T) {
  const u32 y0 = chain1_dec_bit(A, B, A.row.x, B.row.x, 1u, D, T);
  const u32 y1 = chain1_dec_bit(A, B, A.row.x, B.row.x, 2u + y0, D, T);
  const u32 y2 = chain1_dec_bit(A, B, A.row.y, B.row.y, 2u * y0 + y1, D, T);
  u32 y3;
  if (y0 == 0) y3 = chain1_dec_bit(A, B, A.row.z, B.row.z, 2u * y1 + y2, D, T);
  else y3 = chain1_dec_bit(A, B, A.row.w, B.row.w, 2u * y1 + y2, D, T);
  return y0 << 3 | y1 << 2 | y2 << 1 | y3;
}

// 16 warps per CTA, one block per warp; dynamic shared memory: CmSmem + one CmUnitSmem per warp.
template <int VM>
__global__ void __launch_bounds__(512, 1)
k_cm_decode(const u8* __restrict__ in_base, const ZqDecUnit* __restrict__ units, const ZqCmPlan* __restrict__ cmplans, int nunits,
            const CmTablesDev* __restrict__ tab, const u8* __restrict__ blob, u8* __restrict__ model_base,
            u8* __restrict__ out_base, ZqDecResult* __restrict__ results, u32* __restrict__ next_unit, int fast,
            ZqDecSeg* __restrict__ segs, u32* __restrict__ nseg, u32 segcap, u32 unit_base) {
  ZQ_DYN_SMEM(smem_raw);
  CmSmem& T = *reinterpret_cast<CmSmem*>(smem_raw);
  {
    const uint4* src = (const uint4*)tab;
    uint4* dst = (uint4*)smem_raw;
    for (u32 k = threadIdx.x; k < sizeof(CmSmem) / 16; k += blockDim.x) dst[k] = src[k];
  }
  __syncthreads();
  const u32 lane = lane_id();
  CmUnitSmem& S = reinterpret_cast<CmUnitSmem*>(smem_raw + sizeof(CmSmem))[threadIdx.x >> 5];
  for (;;) {
    int t = 0;
    if (lane == 0) t = (int)atomicAdd(next_unit, 1u);
    t = __shfl_sync(ZQ_FULL, t, 0);
    if (t >= nunits) break;
    const ZqDecUnit u = units[t];
    const ZqCmPlan& cp = cmplans[u.plan];
    u8* model = model_base + u.model_off;
    CmCtx X; CmLane L;
    cm_setup(L, X, cp, model, S);
    CmVm vm;
    cm_vm_setup(vm, cp, model, blob, S);
    DecPost pp;
    pp.state = 0; pp.hsize = 0; pp.loaded = 0; pp.error = 0;
    pp.o.out = out_base + u.out_off; pp.o.len = 0; pp.o.cap = u.out_cap; pp.o.error = 0;
    pp.code = model + cp.pcode_off;
    pp.vm.a = pp.vm.b = pp.vm.c = pp.vm.d = 0; pp.vm.f = 0; pp.vm.error = 0;
    pp.vm.m = model + cp.pm_off; pp.vm.h = (u32*)(model + cp.ph_off); pp.vm.r = (u32*)(model + cp.pr_off);
    pp.vm.mmask = (1u << cp.pm) - 1; pp.vm.hmask = (1u << cp.ph) - 1; pp.vm.code = pp.code; pp.vm.len = 0;
    DecIn in; in.p = in_base + u.data_off; in.len = u.data_len; in.pos = 0; in.error = 0;
    u32 err = 0, vmerr = 0;
    if (cp.chain == 1 && fast) {
      if (lane == 0) {
        ChainComp A, B;
        chain_setup(A, cp.comp[0], model);
        chain_setup(B, cp.comp[1], model);
        ChainDec D; D.low = 1; D.high = 0xffffffffu; D.curr = 0; D.err = 0; D.in = in;
        u32 ha = 0, hb = 0;
        for (;;) {   // segments: the model, the coder's range and the post-processor carry on (Z:15481-15508)
          for (int k = 0; k < 4; ++k) D.curr = D.curr << 8 | (u32)D.in.get();
          for (;;) {
            if (D.bit(0)) { if (D.curr != 0 && !D.err) D.err = 1; break; }
            chain_row_switch(A, ha + 16u); chain_row_switch(B, hb + 16u);
            const u32 hi = chain1_dec_nibble(A, B, D, T);
            chain_row_switch(A, ha + 16u * (16u + hi)); chain_row_switch(B, hb + 16u * (16u + hi));
            const u32 c = hi << 4 | chain1_dec_nibble(A, B, D, T);
            cm_vm_run<VM, false>(vm, c, nullptr);
            ha = vm.h[0]; hb = vm.h[1 & vm.hmask];
            dec_post_write<VM, true>(pp, (int)c);
            if (D.err || D.in.error || pp.error || vm.error || pp.vm.error) break;
          }
          if (D.err || D.in.error || pp.error || vm.error || pp.vm.error) break;
          dec_post_write<VM, true>(pp, -1);
          u32 trailer = 0;
          if (pp.error || pp.vm.error || !dec_next_segment(D.in, trailer)) break;
          dec_record_segment(segs, nseg, segcap, unit_base + (u32)t, pp.o.len, trailer, D.err);
          if (D.err) break;
        }
        ZqDecResult r;
        r.out_len = pp.o.len; r.consumed = (u32)D.in.pos; r.pad = 0;
        r.error = D.err ? D.err : D.in.error ? D.in.error : pp.error ? pp.error : (vm.error || pp.vm.error) ? 4u : 0u;
        results[t] = r;
      }
      __syncwarp();
      continue;
    }
    if (cp.n > 0) {
      const bool wide = cp.n > ZQ_CM_LANES;   // lane 0 evaluates every component (zq_cm_wide.cuh); the warp still shares the coder
      CmWide W;
      if (wide && lane == 0) cmw_setup(W, cp, model);
      u32 low = 1, high = 0xffffffffu, curr = 0;
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
      for (;;) {   // segments: the model, the coder's range and the post-processor carry on (Z:15481-15508)
        for (int k = 0; k < 4; ++k) curr = curr << 8 | (u32)in.get();
        for (;;) {
          if (decode(0)) { if (curr != 0 && !err) err = 1; break; }
          int c = 1;
          while (c < 256) {
            int pr = 0;
            if (!wide) pr = cm_predict(L, X, T);
            else { if (lane == 0) pr = cmw_predict(W, vm.h, vm.hmask, T); pr = __shfl_sync(ZQ_FULL, pr, 0); }
            const u32 p16 = (u32)pr * 2 + 1;
            const int y = decode(p16);
            c += c + y;
            bool done;
            if (!wide) done = cm_update(L, X, T, y);
            else { int d = 0; if (lane == 0) d = cmw_update(W, vm.h, vm.hmask, T, y) ? 1 : 0; done = __shfl_sync(ZQ_FULL, d, 0) != 0; }
            if (done) {   // byte complete: contexts of the next one
              if (lane == 0) cm_vm_run<VM, false>(vm, (u32)(c - 256), nullptr);
              __syncwarp();
              if (lane < (u32)X.n) L.h = vm.h[lane & vm.hmask];
              vmerr = (u32)__shfl_sync(ZQ_FULL, vm.error, 0);
            }
          }
          dec_post_write<VM, false>(pp, c - 256);
          if (err || in.error || pp.error || vmerr || pp.vm.error) break;
        }
        if (err || in.error || pp.error || vmerr || pp.vm.error) break;
        dec_post_write<VM, false>(pp, -1);
        u32 trailer = 0;
        if (pp.error || pp.vm.error || !dec_next_segment(in, trailer)) break;
        if (lane == 0) dec_record_segment(segs, nseg, segcap, unit_base + (u32)t, pp.o.len, trailer, err);
        err = __shfl_sync(ZQ_FULL, err, 0);
        if (err) break;
      }
    } else {
      for (;;) {   // segments
        for (;;) {   // unmodeled: big-endian u32 length, that many bytes, ... , length 0
          u32 cl = 0;
          for (int k = 0; k < 4; ++k) cl = cl << 8 | (u32)in.get();
          if (cl == 0 || in.error) break;
          // the length comes from the archive: never loop past the bytes that are there, stop at the first error
          if ((u64)cl > (u64)in.len - in.pos) { in.error = 1; break; }
          for (u32 k = 0; k < cl; ++k) {
            dec_post_write<VM, false>(pp, in.get());
            if (in.error || pp.error || pp.vm.error) break;
          }
          if (in.error || pp.error || pp.vm.error) break;
        }
        if (in.error || pp.error || pp.vm.error) break;
        dec_post_write<VM, false>(pp, -1);
        u32 trailer = 0;
        if (pp.error || pp.vm.error || !dec_next_segment(in, trailer)) break;
        if (lane == 0) dec_record_segment(segs, nseg, segcap, unit_base + (u32)t, pp.o.len, trailer, err);
        err = __shfl_sync(ZQ_FULL, err, 0);
        if (err) break;
      }
    }
    if (lane == 0) {
      ZqDecResult r;
      r.out_len = pp.o.len; r.consumed = (u32)in.pos; r.pad = 0;
      r.error = err ? err : in.error ? in.error : pp.error ? pp.error : (vmerr || pp.vm.error) ? 4u : 0u;
      results[t] = r;
    }
  }
}

}  // namespace zqdev
