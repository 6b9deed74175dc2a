// zq_lz77_half.cuh -- the suffix-array LZ77 parse with TWO blocks per warp (16 lanes each).
//
// Same algorithm and same exactness argument as lz77_sa_parse_pipe (zq_lz77.cuh); what changes is the
// mapping.  The warp-per-block parser is bound by the latency of its own dependent chain (isa -> row
// neighbourhood -> decision -> next position) with only 24 blocks resident per SM; most scans need just
// the first few suffix-array neighbours.  Here a block owns a half-warp: lanes 0-7 of the half look at
// the 8 rows below isa[i+h], lanes 8-15 at the 8 rows above (16 more per extra round), so twice as many
// independent parses are in flight per SM for the same registers.  All warp primitives run on the
// half's own lane mask; the two halves of a warp are fully independent (each pulls its own blocks).
#pragma once
#include "zq_lz77.cuh"

namespace zqdev {

struct Half {                      // lane geometry of one 16-lane group
  u32 mask, hl, base;
  __device__ __forceinline__ void init() { const u32 l = lane_id(); base = l & 16u; hl = l & 15u; mask = 0xffffu << base; }
  __device__ __forceinline__ u32 ballot(bool p) const { return (__ballot_sync(mask, p) >> base) & 0xffffu; }
  template <class T> __device__ __forceinline__ T shfl(T v, int src) const { return __shfl_sync(mask, v, src, 16); }
  template <class T> __device__ __forceinline__ T shfl_up(T v, int d) const { return __shfl_up_sync(mask, v, d, 16); }
  __device__ __forceinline__ u32 redmin(u32 v) const { return __reduce_min_sync(mask, v); }
};

struct HalfSink {                  // same contract as WarpSink, 16 lanes
  u8* out; u8* end; u64 bits; u32 nbits; u32 overflow;
  __device__ __forceinline__ void init(u8* o, u32 cap) { out = o; end = o + cap; bits = 0; nbits = overflow = 0; }
  __device__ __forceinline__ void byte(const Half& H, u32 c) {
    if (out < end) { if (H.hl == 0) *out = (u8)c; } else overflow = 1;
    ++out;
  }
  __device__ __forceinline__ void putb(const Half& H, u64 x, u32 k) {
    bits |= (x & ((1ull << k) - 1)) << nbits;
    nbits += k;
    const u32 nb = nbits >> 3;
    if (nb) {
      if (H.hl < nb) { if (out + H.hl < end) out[H.hl] = (u8)(bits >> (8 * H.hl)); else overflow = 1; }
      if (out + nb > end) overflow = 1;
      out += nb;
      bits = nb < 8 ? bits >> (8 * nb) : 0;
      nbits &= 7;
    }
  }
  __device__ __forceinline__ void flush(const Half& H) { if (nbits > 0) byte(H, (u32)bits & 255); bits = 0; nbits = 0; }
  __device__ __forceinline__ void bytes(const Half& H, const u8* __restrict__ src, u32 cnt) {
    const u32 s = nbits;
    u32 carry = (u32)bits;
    for (u32 b = 0; b < cnt; b += 16) {
      const u32 idx = b + H.hl;
      const u32 c = idx < cnt ? (u32)src[idx] : 0u;
      const u32 prev = H.shfl_up(c, 1);
      const u32 low = H.hl == 0 ? carry : (prev >> (8 - s));
      const u32 o = (low | (c << s)) & 255u;
      if (idx < cnt) { if (out + idx < end) out[idx] = (u8)o; else overflow = 1; }
      const u32 last = min(cnt - b, 16u) - 1;
      carry = H.shfl(c, (int)last) >> (8 - s);
    }
    overflow = H.ballot(overflow != 0) ? 1u : 0u;
    out += cnt;
    bits = carry & ((1u << s) - 1);
  }
};

__device__ __forceinline__ void hlz_write_literal(const Half& H, HalfSink& sk, const LzParams& P, const u8* __restrict__ in, u32 i, u32& lit) {
  if (P.level == 1) {
    if (lit < 1) return;
    u32 nb;
    const u32 g = gamma_code(lit, &nb);
    sk.putb(H, (u64)g << 2, nb + 2);
    sk.bytes(H, in + (i - lit), lit);
    lit = 0;
  } else {
    while (lit > 0) {
      const u32 l1 = lit > 64 ? 64 : lit;
      sk.byte(H, l1 - 1);
      sk.bytes(H, in + (i - lit), l1);
      lit -= l1;
    }
  }
}
__device__ __forceinline__ void hlz_write_match(const Half& H, HalfSink& sk, const LzParams& P, u32 len, u32 off) {
  if (P.level == 1) {
    off += (1u << P.rb) - 1;
    const u32 lo = (u32)zq_bitlen(off) - 1 - P.rb;
    u32 nb;
    const u32 g = gamma_code(len >> 2, &nb);
    const u64 head = (u64)((lo + 8) >> 3) | ((u64)(lo & 7) << 2) | ((u64)g << 5) | ((u64)(len & 3) << (5 + nb));
    sk.putb(H, head, 7 + nb);
    sk.putb(H, ((u64)off & ((1u << P.rb) - 1)) | ((u64)((off >> P.rb) & ((1u << lo) - 1)) << P.rb), P.rb + lo);
  } else {
    const u32 mm = P.minMatch;
    --off;
    while (len > 0) {
      const u32 l1 = len > mm * 2 + 63 ? mm + 63 : len > mm + 63 ? len - mm : len;
      if (off < (1u << 16)) { sk.byte(H, 64 + l1 - mm); sk.byte(H, off >> 8); sk.byte(H, off); }
      else if (off < (1u << 24)) { sk.byte(H, 128 + l1 - mm); sk.byte(H, off >> 16); sk.byte(H, off >> 8); sk.byte(H, off); }
      else { sk.byte(H, 192 + l1 - mm); sk.byte(H, off >> 24); sk.byte(H, off >> 16); sk.byte(H, off >> 8); sk.byte(H, off); }
      len -= l1;
    }
  }
}
__device__ __forceinline__ u32 half_match_len(const Half& H, const u8* __restrict__ a, const u8* __restrict__ b, u32 start, u32 limit) {
  for (u32 base = start; base < limit; base += 16) {
    const u32 idx = base + H.hl;
    const bool diff = idx < limit ? (a[idx] != b[idx]) : true;
    const u32 mk = H.ballot(diff);
    if (mk) return min(limit, base + (u32)(__ffs(mk) - 1));
  }
  return limit;
}

// rows around q: relative lanes 0-7 -> q-1..q-8, lanes 8-15 -> q+1..q+8
template <typename IdxT>
__device__ __forceinline__ LzChunk hlz_chunk_issue(const Half& H, const IdxT* __restrict__ sa, const u16* __restrict__ lcp,
                                                   const u8* __restrict__ bwt, u32 n, u32 q, u32 bucket, bool enable) {
  const u32 up = H.hl >> 3, kk = (H.hl & 7) + 1;
  LzChunk c;
  c.inr = enable && kk <= bucket && (up == 0 ? q >= kk : (u64)q + kk < n);
  const u32 x = up == 0 ? q - kk : q + kk;
  c.s = 0; c.e = 0; c.bw = 0;
  if (c.inr) { c.s = sa[x]; c.e = lcp[up == 0 ? x + 1 : x]; c.bw = bwt[x]; }
  return c;
}

// sequential accept/break rules over the lanes of `segmask` (16-bit, relative) -- see lz_resolve
__device__ __forceinline__ bool hlz_resolve(const Half& H, u32 segmask, const LzCand& c, int exmax, LzBest& b, u32 minMatch,
                                            const u8* __restrict__ in, u32 i, u32 h, u32 lit, u32 lmax) {
  const u32 vm = H.ballot(c.valid) & segmask;
  if (!vm) return false;
  const int f = __ffs(vm) - 1;
  if (H.shfl((int)c.capped, f)) {
    const u32 cp = H.shfl(c.p, f);
    u32 l = half_match_len(H, in + cp, in + i, h + ZQ_LCP_CAP, lmax);
    l = min(l, lmax);
    const u32 l1 = H.shfl(c.l1, f);
    const int sc = lz_score(l, l1, i - cp, lit, h);
    if (sc > b.bscore) { b.blen = l; b.bp = cp; b.blit = l1; b.bscore = sc; }
    return true;
  }
  const int M = max(b.bscore, exmax);
  const bool acc = c.valid && c.score > M;
  const u32 am = H.ballot(acc) & segmask;
  const u32 la = am & ((2u << H.hl) - 1);
  const u32 lsrc = H.shfl(c.l, la ? 31 - __clz(la) : (int)H.hl);
  const u32 bl = la ? lsrc : b.blen;
  const bool brk = c.valid && (c.l < bl || c.l < minMatch || c.l > 255);
  const u32 bm = H.ballot(brk) & segmask;
  const u32 upto = bm ? ((2u << (__ffs(bm) - 1)) - 1) : 0xffffu;
  const u32 fa = am & upto;
  if (fa) {
    const int w = 31 - __clz(fa);
    b.blen = H.shfl(c.l, w); b.bp = H.shfl(c.p, w); b.blit = H.shfl(c.l1, w); b.bscore = H.shfl(c.score, w);
  }
  return bm != 0;
}

template <typename IdxT>
__device__ __forceinline__ void hlz_scan_pos(const Half& H, const u8* __restrict__ in, u32 n, const IdxT* __restrict__ sa,
                                             const u16* __restrict__ lcp, const u8* __restrict__ bwt, const LzParams& P,
                                             u32 i, u32 h, u32 lit, u32 lmax, u32 q, const LzChunk& ch, LzBest& b) {
  const u32 ci = h > 0 ? (u32)in[i + h - 1] : 0u;
  const u32 p = ch.s - h;
  const u32 vm_all = H.ballot(ch.inr && p < i);
  const u32 outm = H.ballot(!ch.inr);
  bool have = false;
  LzCand c; int exmax = INT_MIN; u32 pm = 0;
  c.valid = false; c.capped = false; c.p = c.l = c.l1 = 0; c.score = 0;
  for (u32 dir = 0; dir < 2; ++dir) {
    const u32 segmask = dir ? 0xff00u : 0x00ffu;
    const u32 vm = vm_all & segmask;
    bool stop = false, general = false;
    if (vm) {
      u32 rest = vm;
      for (int tries = 0; rest; ++tries) {
        if (tries == 3) { general = true; break; }
        const int f = __ffs(rest) - 1;
        rest &= rest - 1;
        const u32 upto = segmask & ((2u << f) - 1);
        const u32 pmf = H.redmin(((upto >> H.hl) & 1u) ? ch.e : 0xffffffffu);
        if (pmf >= ZQ_LCP_CAP) { general = true; break; }
        const u32 cp = H.shfl(p, f);
        const u32 lf = h + pmf;
        u32 l1 = h;
        if (h > 0 && H.shfl(ch.bw, f) == ci) { --l1; while (l1 > 0 && in[cp + l1 - 1] == in[i + l1 - 1]) --l1; }
        const int sc = lz_score(lf, l1, i - cp, lit, h);
        if (sc > b.bscore) { b.blen = lf; b.bp = cp; b.blit = l1; b.bscore = sc; }
        if (lf < b.blen || lf < P.minMatch || lf > 255) { stop = true; break; }
      }
      if (general) {
        if (!have) {
          pm = ch.e;
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) { const u32 t = __shfl_up_sync(H.mask, pm, o, 8); if ((H.hl & 7) >= (u32)o) pm = min(pm, t); }
          c = lz_eval(in, i, h, lit, ch.inr, ch.s, pm, ch.bw, ci);
          int im = c.valid ? c.score : INT_MIN;
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) { const int t = __shfl_up_sync(H.mask, im, o, 8); if ((H.hl & 7) >= (u32)o) im = max(im, t); }
          exmax = __shfl_up_sync(H.mask, im, 1, 8);
          if ((H.hl & 7) == 0) exmax = INT_MIN;
          have = true;
        }
        stop = hlz_resolve(H, segmask, c, exmax, b, P.minMatch, in, i, h, lit, lmax);
      }
    }
    if (stop || (outm & segmask)) continue;
    // more than 8 rows needed in this direction: 16 per round from k = 9
    u32 run_min = H.redmin(((segmask >> H.hl) & 1u) ? ch.e : 0xffffffffu);
    for (u32 k0 = 8; k0 < P.bucket && !stop; k0 += 16) {
      const u32 k = k0 + H.hl + 1;
      const bool inr2 = k <= P.bucket && (dir == 0 ? q >= k : (u64)q + k < n);
      const u32 x2 = dir == 0 ? q - k : q + k;
      u32 s2 = 0, pm2 = 0, bw2 = 0;
      if (inr2) { s2 = sa[x2]; pm2 = lcp[dir == 0 ? x2 + 1 : x2]; bw2 = bwt[x2]; }
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) { const u32 t = H.shfl_up(pm2, o); if (H.hl >= (u32)o) pm2 = min(pm2, t); }
      pm2 = min(pm2, run_min);
      run_min = H.shfl(pm2, 15);
      const LzCand c2 = lz_eval(in, i, h, lit, inr2, s2, pm2, bw2, ci);
      int im2 = c2.valid ? c2.score : INT_MIN;
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) { const int t = H.shfl_up(im2, o); if (H.hl >= (u32)o) im2 = max(im2, t); }
      int ex2 = H.shfl_up(im2, 1);
      if (H.hl == 0) ex2 = INT_MIN;
      stop = hlz_resolve(H, 0xffffu, c2, ex2, b, P.minMatch, in, i, h, lit, lmax);
      if (H.ballot(!inr2)) break;
    }
  }
}

template <typename IdxT>
__device__ void hlz_parse(const Half& H, const u8* __restrict__ in, u32 n, const IdxT* __restrict__ sa, const IdxT* __restrict__ isa,
                          const u16* __restrict__ lcp, const u8* __restrict__ bwt, const LzParams P, HalfSink& sk) {
  const u32 maxMatch = 3u << 14;
  u32 i = 0, lit = 0;
  if (n == 0) { sk.flush(H); return; }
  u32 qv = (H.hl < 2 && i + H.hl < n) ? (u32)isa[i + H.hl] : 0u;
  u32 qA = H.shfl(qv, 0), qB = H.shfl(qv, 1);
  LzChunk A = hlz_chunk_issue(H, sa, lcp, bwt, n, qA, P.bucket, true);
  while (i < n) {
    const bool haveB = i + 1 < n;
    LzChunk B = hlz_chunk_issue(H, sa, lcp, bwt, n, qB, P.bucket, haveB);
    const u32 qC = i + 2 < n ? (u32)isa[i + 2] : 0u;
    LzBest b; b.blen = P.minMatch - 1; b.bp = 0; b.blit = 0; b.bscore = 0;
    const u32 lmax = min(maxMatch, n - i);
    const u32 adj = max(H.shfl(A.inr ? A.e : 0u, 0), H.shfl(A.inr ? A.e : 0u, 8));
    if (adj >= P.minMatch) hlz_scan_pos(H, in, n, sa, lcp, bwt, P, i, 0, lit, lmax, qA, A, b);
    if (P.lookahead >= 1 && !(b.bscore <= 0 || b.blen < P.minMatch) && haveB && ((i + 1) >> P.checkbits) == (i >> P.checkbits))
      hlz_scan_pos(H, in, n, sa, lcp, bwt, P, i, 1, lit, lmax, qB, B, b);
    const u32 off = i - b.bp;
    u32 adv;
    if (off > 0 && b.bscore > 0 &&
        b.blen - b.blit >= P.minMatch + (P.level == 2 ? (u32)(off >= (1u << 16)) + (u32)(off >= (1u << 24)) : 0u)) {
      lit += b.blit;
      hlz_write_literal(H, sk, P, in, i + b.blit, lit);
      hlz_write_match(H, sk, P, b.blen - b.blit, off);
      adv = b.blen;
    } else { adv = 1; ++lit; }
    if (lit >= (1u << 12)) hlz_write_literal(H, sk, P, in, i + adv, lit);
    i += adv;
    if (adv == 1) { A = B; qA = qB; qB = qC; }
    else if (i < n) {
      qv = (H.hl < 2 && i + H.hl < n) ? (u32)isa[i + H.hl] : 0u;
      qA = H.shfl(qv, 0); qB = H.shfl(qv, 1);
      A = hlz_chunk_issue(H, sa, lcp, bwt, n, qA, P.bucket, true);
    }
  }
  hlz_write_literal(H, sk, P, in, n, lit);
  sk.flush(H);
}

// persistent half-warps pull blocks (look-ahead <= 1 only) from the queue
template <typename IdxT, int MINB>
__global__ void __launch_bounds__(128, MINB)
k_lz77_sa_half(const u8* __restrict__ in_base, const ZqUnit* __restrict__ units, const ZqPlan* __restrict__ plans,
               const int* __restrict__ todo, int ntodo, const u8* __restrict__ work_base,
               u8* __restrict__ lz_base, u32* __restrict__ lz_len, u32* __restrict__ err_flag, u32* __restrict__ next_unit) {
  Half H; H.init();
  for (;;) {
    int t = 0;
    if (H.hl == 0) t = (int)atomicAdd(next_unit, 1u);
    t = H.shfl(t, 0);
    if (t >= ntodo) break;
    const int ui = todo[t];
    const ZqUnit u = units[ui];
    const ZqPlan pl = plans[u.plan];
    LzParams P;
    P.level = pl.lz_level; P.minMatch = pl.args[2]; P.lookahead = pl.args[6];
    P.bucket = (1u << pl.args[4]) - 1; P.rb = pl.args[0] > 4 ? pl.args[0] - 4 : 0; P.checkbits = 17 + pl.args[0];
    HalfSink sk; sk.init(lz_base + u.lz_off, u.lz_cap);
    const u8* w = work_base + u.work_off;
    const u64 stride = zq_work_stride(u.n, sizeof(IdxT));
    const u16* lcp = (const u16*)(w + 2 * stride);
    const u8* bwt = w + 2 * stride + zq_work_stride(u.n, 2);
    hlz_parse<IdxT>(H, in_base + u.in_off, u.n, (const IdxT*)w, (const IdxT*)(w + stride), lcp, bwt, P, sk);
    if (H.hl == 0) {
      lz_len[ui] = (u32)(sk.out - (lz_base + u.lz_off));
      if (sk.overflow) atomicOr(err_flag, 1u);
    }
  }
}

}  // namespace zqdev
