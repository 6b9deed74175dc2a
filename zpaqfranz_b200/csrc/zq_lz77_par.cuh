// zq_lz77_par.cuh -- the suffix-array LZ77 parse (LZBuffer::fill, Z:19395-19430, Z:19472-19527) split so
// that the heavy part is embarrassingly parallel:
//
//   A  k_lz_candidates : ONE THREAD PER INPUT POSITION evaluates the reference's candidate scan at that
//      position (rows isa[i+h] +-bucket, lengths by running minimum over the capped LCP array) and
//      stores the decision it leads to, for both values of the only parse state that enters the
//      scores: "no literal pending" (lit == 0) and "literals pending".  No dependence between
//      positions, every lane does useful work (the warp-per-block parser keeps ~1 lane in 8 busy).
//   B  k_lz_chain      : one thread per block walks i -> i + len through those decisions (the only
//      sequential part: ~15 instructions per step) and records the tokens.  Positions whose scan
//      met a neighbour with LCP >= 256 (exact length needed, Z:19419) are evaluated here, exactly,
//      only if the walk actually visits them.
//   C  k_lz_emit       : tokens -> code lengths -> prefix sum -> every token writes its own bits
//      (level 1: LSB-first bit codes; level 2: byte codes) with word-wise atomicOr into a zeroed stream.
// Bit-exact by construction: A evaluates literally the reference's loop per position; B applies the
// same accept / advance / flush rules (Z:19472-19519).
#pragma once
#include "zq_lz77.cuh"

namespace zqdev {

// decision at one position for one literal state
//   bit 63 deferred (needs the exact evaluator), bit 62 match, 48..55 blit, 32..47 blen, 0..31 offset
constexpr u64 LZD_DEFER = 1ull << 63, LZD_MATCH = 1ull << 62;

// The reference's candidate scan at position i (Z:19404-19430) for both literal states at once.
// EXACT=false: stops with LZD_DEFER as soon as a usable neighbour has a capped LCP.
template <typename IdxT, bool EXACT>
__device__ void lz_decide_at(const u8* __restrict__ in, u32 n, const IdxT* __restrict__ sa, const IdxT* __restrict__ isa,
                             const u16* __restrict__ lcp, const u8* __restrict__ bwt, const LzParams& P, u32 i, u64 (&out)[2]) {
  const u32 maxMatch = 3u << 14;
  const u32 lmax = min(maxMatch, n - i);
  LzBest b[2];
  bool alive[2] = {true, true};
#pragma unroll
  for (int st = 0; st < 2; ++st) { b[st].blen = P.minMatch - 1; b[st].bp = 0; b[st].blit = 0; b[st].bscore = 0; }
  for (u32 h = 0; h <= P.lookahead && (alive[0] || alive[1]); ++h) {
    const u32 pos = i + h;
    if (pos >= n || (pos >> P.checkbits) != (i >> P.checkbits)) continue;
    const u32 q = isa[pos];
    const u32 ci = h > 0 ? (u32)in[i + h - 1] : 0u;
    for (int dir = 0; dir < 2; ++dir) {
      bool stop0 = !alive[0], stop1 = !alive[1];
      u32 runmin = 0xffffffffu;
      for (u32 k = 1; k <= P.bucket && !(stop0 && stop1); ++k) {
        if (dir == 0 ? q < k : (u64)q + k >= n) break;      // off the array: so is every later k
        const u32 x = dir == 0 ? q - k : q + k;
        runmin = min(runmin, (u32)lcp[dir == 0 ? x + 1 : x]);
        // Pruning (exact): every remaining neighbour of this scan matches at most h+runmin bytes, so its
        // score is at most (h+runmin)*8 - lg(1) - 11, scaled; once that cannot beat a state's best score
        // the rest of the reference's scan can no longer change that state.
        if (runmin < ZQ_LCP_CAP) {   // (a capped LCP only says ">= 256": no bound then)
          int ub = (int)(h + runmin) * 8 - 12;
          for (u32 a = 0; a < h; ++a) ub = ub * 5 / 8;
          if (ub <= b[0].bscore) stop0 = true;
          if (ub <= b[1].bscore) stop1 = true;
          if (stop0 && stop1) break;
        }
        const u32 p = (u32)sa[x] - h;
        if (!(p < i)) continue;
        u32 l = h + runmin;
        if (runmin >= ZQ_LCP_CAP) {
          if (!EXACT) { out[0] = out[1] = LZD_DEFER; return; }
          while (l < lmax && in[p + l] == in[i + l]) ++l;
        }
        l = min(l, lmax);
        u32 l1 = h;
        if (h > 0 && (u32)bwt[x] == ci) { --l1; while (l1 > 0 && in[p + l1 - 1] == in[i + l1 - 1]) --l1; }
        const int base = (int)(l - l1) * 8 - zq_bitlen(i - p) - 11;
#pragma unroll
        for (int st = 0; st < 2; ++st) {
          bool& stop = st ? stop1 : stop0;
          if (stop) continue;
          int sc = base - ((st == 0 && l1 > 0) ? 4 : 0);
          for (u32 a = 0; a < h; ++a) sc = sc * 5 / 8;
          if (sc > b[st].bscore) { b[st].blen = l; b[st].bp = p; b[st].blit = l1; b[st].bscore = sc; }
          if (l < b[st].blen || l < P.minMatch || l > 255) stop = true;
        }
      }
    }
#pragma unroll
    for (int st = 0; st < 2; ++st) if (alive[st] && (b[st].bscore <= 0 || b[st].blen < P.minMatch)) alive[st] = false;
  }
#pragma unroll
  for (int st = 0; st < 2; ++st) {
    const u32 off = i - b[st].bp;
    const bool match = off > 0 && b[st].bscore > 0 &&
                       b[st].blen - b[st].blit >= P.minMatch + (P.level == 2 ? (u32)(off >= (1u << 16)) + (u32)(off >= (1u << 24)) : 0u);
    out[st] = match ? (LZD_MATCH | ((u64)b[st].blit << 48) | ((u64)b[st].blen << 32) | off) : 0ull;
  }
}

struct LzUnitView {
  const u8* in; u32 n; const void* sa; const void* isa; const u16* lcp; const u8* bwt; bool idx16; LzParams P;
};
__device__ __forceinline__ LzUnitView lz_unit_view(const u8* in_base, const u8* work_base, const ZqUnit& u, const ZqPlan& pl) {
  LzUnitView v;
  v.P.level = pl.lz_level; v.P.minMatch = pl.args[2]; v.P.lookahead = pl.args[6];
  v.P.bucket = (1u << pl.args[4]) - 1; v.P.rb = pl.args[0] > 4 ? pl.args[0] - 4 : 0; v.P.checkbits = 17 + pl.args[0];
  const u8* w = work_base + u.work_off;
  const u64 stride = zq_work_stride(u.n, u.idx16 ? 2 : 4);
  v.in = in_base + u.in_off; v.n = u.n; v.sa = w; v.isa = w + stride; v.lcp = (const u16*)(w + 2 * stride);
  v.bwt = w + 2 * stride + zq_work_stride(u.n, 2); v.idx16 = u.idx16 != 0;
  return v;
}

// A: every lane runs the reference's candidate scan of one position at a time as a small state machine
// (one SA neighbour per loop iteration) and pulls the next position of its CTA's 2048-position chunk when
// it finishes, so lanes with short scans do not wait for lanes with long ones.
// grid (chunks, units); dec = 2 x u64 per position at dec_base + 16*(dec_off[unit] + i)
constexpr u32 LZC_CHUNK = 2048;
template <typename IdxT>
__device__ void lz_candidates_chunk(const u8* __restrict__ in, u32 n, const IdxT* __restrict__ sa, const IdxT* __restrict__ isa,
                                    const u16* __restrict__ lcp, const u8* __restrict__ bwt, const LzParams P,
                                    u32 chunk_begin, u32 chunk_end, u32* next_pos, u64* __restrict__ dec) {
  const u32 maxMatch = 3u << 14;
  u32 i = 0, h = 0, dir = 0, k = 0, q = 0, runmin = 0, ci = 0, lmax = 0;
  LzBest b0, b1; b0.blen = b0.bp = b0.blit = 0; b0.bscore = 0; b1 = b0;
  bool alive0 = false, alive1 = false, stop0 = true, stop1 = true;
  bool have = false, deferred = false;
  int phase = 0;   // 0: need a position, 1: start look-ahead h, 2: scanning
  for (;;) {
    if (phase == 0) {
      i = chunk_begin + atomicAdd(next_pos, 1u);
      if (i >= chunk_end) break;
      lmax = min(maxMatch, n - i);
      b0.blen = P.minMatch - 1; b0.bp = 0; b0.blit = 0; b0.bscore = 0; b1 = b0;
      alive0 = alive1 = true; deferred = false; h = 0; phase = 1; have = true;
    }
    if (phase == 1) {
      // start look-ahead h (or finish the position)
      bool fin = deferred || h > P.lookahead || !(alive0 || alive1);
      if (!fin) {
        const u32 pos = i + h;
        if (pos >= n || (pos >> P.checkbits) != (i >> P.checkbits)) { ++h; continue; }
        q = isa[pos];
        ci = h > 0 ? (u32)in[i + h - 1] : 0u;
        dir = 0; k = 1; runmin = 0xffffffffu; stop0 = !alive0; stop1 = !alive1;
        phase = 2;
      } else {
        u64 o0, o1;
        if (deferred) o0 = o1 = LZD_DEFER;
        else {
          const u32 off0 = i - b0.bp, off1 = i - b1.bp;
          const bool m0 = off0 > 0 && b0.bscore > 0 &&
                          b0.blen - b0.blit >= P.minMatch + (P.level == 2 ? (u32)(off0 >= (1u << 16)) + (u32)(off0 >= (1u << 24)) : 0u);
          const bool m1 = off1 > 0 && b1.bscore > 0 &&
                          b1.blen - b1.blit >= P.minMatch + (P.level == 2 ? (u32)(off1 >= (1u << 16)) + (u32)(off1 >= (1u << 24)) : 0u);
          o0 = m0 ? (LZD_MATCH | ((u64)b0.blit << 48) | ((u64)b0.blen << 32) | off0) : 0ull;
          o1 = m1 ? (LZD_MATCH | ((u64)b1.blit << 48) | ((u64)b1.blen << 32) | off1) : 0ull;
        }
        *(ulonglong2*)(dec + 2 * (u64)i) = make_ulonglong2(o0, o1);
        phase = 0;
        continue;
      }
    }
    // phase 2: one neighbour of row q in direction dir
    const bool off_array = dir == 0 ? q < k : (u64)q + k >= n;
    if (k > P.bucket || (stop0 && stop1) || off_array) {
      if (dir == 0) { dir = 1; k = 1; runmin = 0xffffffffu; stop0 = !alive0; stop1 = !alive1; }
      else {
        if (alive0 && (b0.bscore <= 0 || b0.blen < P.minMatch)) alive0 = false;
        if (alive1 && (b1.bscore <= 0 || b1.blen < P.minMatch)) alive1 = false;
        ++h; phase = 1;
      }
      continue;
    }
    const u32 x = dir == 0 ? q - k : q + k;
    ++k;
    runmin = min(runmin, (u32)lcp[dir == 0 ? x + 1 : x]);
    if (runmin < ZQ_LCP_CAP) {   // exact pruning, see lz_decide_at
      int ub = (int)(h + runmin) * 8 - 12;
      if (h == 1) ub = ub * 5 / 8; else for (u32 a = 0; a < h; ++a) ub = ub * 5 / 8;
      if (ub <= b0.bscore) stop0 = true;
      if (ub <= b1.bscore) stop1 = true;
      if (stop0 && stop1) continue;
    }
    const u32 p = (u32)sa[x] - h;
    if (!(p < i)) continue;
    if (runmin >= ZQ_LCP_CAP) { deferred = true; phase = 1; continue; }
    const u32 l = min(h + runmin, lmax);
    u32 l1 = h;
    if (h > 0 && (u32)bwt[x] == ci) { --l1; while (l1 > 0 && in[p + l1 - 1] == in[i + l1 - 1]) --l1; }
    const int base = (int)(l - l1) * 8 - zq_bitlen(i - p) - 11;
    const bool brk_len = l < P.minMatch || l > 255;
    if (!stop0) {
      int sc = base - (l1 > 0 ? 4 : 0);
      if (h == 1) sc = sc * 5 / 8; else for (u32 a = 0; a < h; ++a) sc = sc * 5 / 8;
      if (sc > b0.bscore) { b0.blen = l; b0.bp = p; b0.blit = l1; b0.bscore = sc; }
      if (l < b0.blen || brk_len) stop0 = true;
    }
    if (!stop1) {
      int sc = base;
      if (h == 1) sc = sc * 5 / 8; else for (u32 a = 0; a < h; ++a) sc = sc * 5 / 8;
      if (sc > b1.bscore) { b1.blen = l; b1.bp = p; b1.blit = l1; b1.bscore = sc; }
      if (l < b1.blen || brk_len) stop1 = true;
    }
  }
  (void)have;
}

__global__ void __launch_bounds__(256)
k_lz_candidates(const u8* __restrict__ in_base, const ZqUnit* __restrict__ units, const ZqPlan* __restrict__ plans,
                const int* __restrict__ todo, const u8* __restrict__ work_base, const u64* __restrict__ dec_off,
                u64* __restrict__ dec_base) {
  __shared__ u32 next_pos;
  const int t = blockIdx.y;
  const int ui = todo[t];
  const ZqUnit u = units[ui];
  const u32 cb = blockIdx.x * LZC_CHUNK;
  if (cb >= u.n) return;
  if (threadIdx.x == 0) next_pos = 0;
  __syncthreads();
  const LzUnitView v = lz_unit_view(in_base, work_base, u, plans[u.plan]);
  const u32 ce = min(u.n, cb + LZC_CHUNK);
  u64* dec = dec_base + 2 * dec_off[t];
  if (v.idx16) lz_candidates_chunk<u16>(v.in, v.n, (const u16*)v.sa, (const u16*)v.isa, v.lcp, v.bwt, v.P, cb, ce, &next_pos, dec);
  else lz_candidates_chunk<u32>(v.in, v.n, (const u32*)v.sa, (const u32*)v.isa, v.lcp, v.bwt, v.P, cb, ce, &next_pos, dec);
}

struct LzToken { u32 pos, lit, mlen, off; };   // lit literals starting at pos, then (mlen > 0) a match

// B: one thread per unit
__global__ void __launch_bounds__(64)
k_lz_chain(const u8* __restrict__ in_base, const ZqUnit* __restrict__ units, const ZqPlan* __restrict__ plans,
           const int* __restrict__ todo, int ntodo, const u8* __restrict__ work_base, const u64* __restrict__ dec_off,
           const u64* __restrict__ dec_base, const u64* __restrict__ tok_off, LzToken* __restrict__ tok_base,
           u32* __restrict__ ntok) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntodo) return;
  const int ui = todo[t];
  const ZqUnit u = units[ui];
  const LzUnitView v = lz_unit_view(in_base, work_base, u, plans[u.plan]);
  const u64* dec = dec_base + 2 * dec_off[t];
  LzToken* tok = tok_base + tok_off[t];
  const u32 maxLiteral = 1u << 12;
  u32 i = 0, lit = 0, nt = 0;
  while (i < v.n) {
    u64 d = dec[2 * i + (lit ? 1 : 0)];
    if (d & LZD_DEFER) {
      u64 o[2];
      if (v.idx16) lz_decide_at<u16, true>(v.in, v.n, (const u16*)v.sa, (const u16*)v.isa, v.lcp, v.bwt, v.P, i, o);
      else lz_decide_at<u32, true>(v.in, v.n, (const u32*)v.sa, (const u32*)v.isa, v.lcp, v.bwt, v.P, i, o);
      d = o[lit ? 1 : 0];
    }
    if (d & LZD_MATCH) {
      const u32 blit = (u32)(d >> 48) & 255u, blen = (u32)(d >> 32) & 0xffffu, off = (u32)d;
      lit += blit;
      LzToken k; k.pos = i + blit - lit; k.lit = lit; k.mlen = blen - blit; k.off = off;
      tok[nt++] = k;
      lit = 0;
      i += blen;
    } else {
      ++lit; ++i;
      if (lit >= maxLiteral) { LzToken k; k.pos = i - lit; k.lit = lit; k.mlen = 0; k.off = 0; tok[nt++] = k; lit = 0; }
    }
  }
  if (lit) { LzToken k; k.pos = v.n - lit; k.lit = lit; k.mlen = 0; k.off = 0; tok[nt++] = k; }
  ntok[t] = nt;
}

// bits a token occupies (level 1) / 8 x bytes (level 2)
__device__ __forceinline__ u64 lz_token_bits(const LzToken& k, const LzParams& P) {
  u64 bits = 0;
  if (P.level == 1) {
    if (k.lit) bits += 2 + (2 * (zq_bitlen(k.lit) - 1) + 1) + 8ull * k.lit;
    if (k.mlen) {
      const u32 off = k.off + (1u << P.rb) - 1;
      const u32 lo = (u32)zq_bitlen(off) - 1 - P.rb;
      bits += 5 + (2 * (zq_bitlen(k.mlen >> 2) - 1) + 1) + 2 + P.rb + lo;
    }
  } else {
    bits += 8ull * (k.lit + (k.lit + 63) / 64);
    const u32 ob = k.off - 1 < (1u << 16) ? 3 : k.off - 1 < (1u << 24) ? 4 : 5;
    u32 len = k.mlen; const u32 mm = P.minMatch;
    while (len > 0) { const u32 l1 = len > mm * 2 + 63 ? mm + 63 : len > mm + 63 ? len - mm : len; bits += 8ull * ob; len -= l1; }
  }
  return bits;
}

struct BitOut {   // writes bit strings at absolute bit positions of a zeroed, 4-byte aligned buffer
  u32* w;
  __device__ __forceinline__ void put(u64 bitpos, u64 code, u32 nbits) {   // nbits <= 57
    if (!nbits) return;
    code &= nbits < 64 ? ((1ull << nbits) - 1) : ~0ull;
    const u64 wi = bitpos >> 5; const u32 sh = (u32)bitpos & 31u;
    const u64 lo = code << sh;                       // bits for words wi, wi+1
    atomicOr(&w[wi], (u32)lo);
    if (sh + nbits > 32) atomicOr(&w[wi + 1], (u32)(lo >> 32));
    if (sh + nbits > 64) atomicOr(&w[wi + 2], (u32)(code >> (64 - sh)));
  }
};

// C: one CTA per unit; tokens -> offsets (scan) -> bits
__global__ void __launch_bounds__(256)
k_lz_emit(const u8* __restrict__ in_base, const ZqUnit* __restrict__ units, const ZqPlan* __restrict__ plans,
          const int* __restrict__ todo, int ntodo, const u64* __restrict__ tok_off, const LzToken* __restrict__ tok_base,
          const u32* __restrict__ ntok, u64* __restrict__ bitpos_base, u8* __restrict__ lz_base, u32* __restrict__ lz_len,
          u32* __restrict__ err_flag) {
  __shared__ u64 wsum[8];
  __shared__ u64 carry_s;
  for (int t = blockIdx.x; t < ntodo; t += gridDim.x) {
    const int ui = todo[t];
    const ZqUnit u = units[ui];
    const ZqPlan pl = plans[u.plan];
    LzParams P;
    P.level = pl.lz_level; P.minMatch = pl.args[2]; P.lookahead = pl.args[6];
    P.bucket = (1u << pl.args[4]) - 1; P.rb = pl.args[0] > 4 ? pl.args[0] - 4 : 0; P.checkbits = 17 + pl.args[0];
    const u8* __restrict__ in = in_base + u.in_off;
    const LzToken* tok = tok_base + tok_off[t];
    u64* bp = bitpos_base + tok_off[t];
    const u32 nt = ntok[t];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    // exclusive scan of the token bit lengths
    for (u32 b0 = 0; b0 < nt; b0 += blockDim.x) {
      const u32 k = b0 + threadIdx.x;
      const u64 mine = k < nt ? lz_token_bits(tok[k], P) : 0;
      u64 inc = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const u64 x = __shfl_up_sync(ZQ_FULL, inc, o); if (lane >= (u32)o) inc += x; }
      if (lane == 31) wsum[warp] = inc;
      __syncthreads();
      u64 pre = carry_s;
      for (u32 w = 0; w < warp; ++w) pre += wsum[w];
      if (k < nt) bp[k] = pre + inc - mine;
      __syncthreads();
      if (threadIdx.x == blockDim.x - 1) carry_s = pre + inc;
      __syncthreads();
    }
    const u64 total_bits = carry_s;
    const u64 total_bytes = (total_bits + 7) >> 3;
    if (threadIdx.x == 0) { lz_len[ui] = (u32)total_bytes; if (total_bytes > u.lz_cap) atomicOr(err_flag, 1u); }
    if (total_bytes > u.lz_cap) { __syncthreads(); continue; }
    BitOut bo; bo.w = (u32*)(lz_base + u.lz_off);
    for (u32 k = threadIdx.x; k < nt; k += blockDim.x) {
      const LzToken tk = tok[k];
      u64 pos = bp[k];
      if (P.level == 1) {
        if (tk.lit) {
          u32 nb; const u32 g = gamma_code(tk.lit, &nb);
          bo.put(pos, (u64)g << 2, nb + 2); pos += nb + 2;
          const u8* src = in + tk.pos;
          u32 j = 0;
          for (; j + 4 <= tk.lit; j += 4) {
            const u32 v4 = (u32)src[j] | (u32)src[j + 1] << 8 | (u32)src[j + 2] << 16 | (u32)src[j + 3] << 24;
            bo.put(pos, v4, 32); pos += 32;
          }
          for (; j < tk.lit; ++j) { bo.put(pos, src[j], 8); pos += 8; }
        }
        if (tk.mlen) {
          const u32 off = tk.off + (1u << P.rb) - 1;
          const u32 lo = (u32)zq_bitlen(off) - 1 - P.rb;
          u32 nb; const u32 g = gamma_code(tk.mlen >> 2, &nb);
          const u64 head = (u64)((lo + 8) >> 3) | ((u64)(lo & 7) << 2) | ((u64)g << 5) | ((u64)(tk.mlen & 3) << (5 + nb));
          bo.put(pos, head, 7 + nb); pos += 7 + nb;
          bo.put(pos, ((u64)off & ((1u << P.rb) - 1)) | ((u64)((off >> P.rb) & ((1u << lo) - 1)) << P.rb), P.rb + lo);
        }
      } else {
        u32 lit = tk.lit; const u8* src = in + tk.pos;
        while (lit > 0) {
          const u32 l1 = lit > 64 ? 64 : lit;
          bo.put(pos, l1 - 1, 8); pos += 8;
          for (u32 j = 0; j < l1; ++j) { bo.put(pos, src[j], 8); pos += 8; }
          src += l1; lit -= l1;
        }
        u32 len = tk.mlen; const u32 mm = P.minMatch; const u32 off = tk.off - 1;
        while (len > 0) {
          const u32 l1 = len > mm * 2 + 63 ? mm + 63 : len > mm + 63 ? len - mm : len;
          if (off < (1u << 16)) { bo.put(pos, (64 + l1 - mm) | ((off >> 8) << 8) | ((off & 255) << 16), 24); pos += 24; }
          else if (off < (1u << 24)) { bo.put(pos, (u64)(128 + l1 - mm) | ((u64)(off >> 16) << 8) | ((u64)((off >> 8) & 255) << 16) | ((u64)(off & 255) << 24), 32); pos += 32; }
          else { bo.put(pos, (u64)(192 + l1 - mm) | ((u64)(off >> 24) << 8) | ((u64)((off >> 16) & 255) << 16) | ((u64)((off >> 8) & 255) << 24) | ((u64)(off & 255) << 32), 40); pos += 40; }
          len -= l1;
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace zqdev
