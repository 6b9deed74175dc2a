// zq_lz77.cuh -- the exact greedy LZ77 parse of LZBuffer::fill (suffix-array search variant) and
// its two code formats, one WARP per block.
//
// Replaces LZBuffer::fill / write_literal / write_match (Z:19381-19612) for methods whose args[5]-
// args[0] >= 21 (-m2: "x0,1,4,0,7,21,1"; -m3 binary: "x0,2,12,0,7,21,1...").  The parse is a strict
// sequential recurrence over the position i (each decision moves i by the chosen match length and the
// pending-literal count feeds the next score), so a block is owned by one warp and the parallel
// work is the candidate scan: the +-bucket suffix-array neighbours of isa[i+h] are examined 32 at a
// time by the lanes.  Their match lengths come from a running minimum over the capped LCP array
// (zq_sufsort.cuh) instead of byte compares; only a neighbour whose LCP hits the cap (>= 256 bytes,
// after which the reference's scan stops, Z:19424) is measured exactly, cooperatively by the warp.
// The reference's scan order, strict '>' tie-break, early exits and the lit==0 penalty are replayed
// literally, because any deviation changes the output bytes.
#pragma once
#include "zq_common.cuh"
#include "zq_sufsort.cuh"

namespace zqdev {

// Sequential byte/bit sink owned by one warp. State is warp-uniform; stores are spread over lanes.
struct WarpSink {
  u8* out;      // next byte to write
  u8* end;      // capacity guard (overflow -> flag, bytes dropped)
  u64 bits;     // pending bits, LSB first (level-1 codes); always < 8 of them between calls
  u32 nbits;
  u32 overflow;
  __device__ __forceinline__ void init(u8* o, u32 cap) { out = o; end = o + cap; bits = 0; nbits = overflow = 0; }
  __device__ __forceinline__ void byte(u32 c) {
    if (out < end) { if (lane_id() == 0) *out = (u8)c; } else overflow = 1;
    ++out;
  }
  // append the k (<= 56) low bits of x; whole bytes are stored by lanes 0..7 in one instruction
  __device__ __forceinline__ void putb(u64 x, u32 k) {
    bits |= (x & ((1ull << k) - 1)) << nbits;
    nbits += k;
    const u32 nb = nbits >> 3;
    if (nb) {
      const u32 lane = lane_id();
      if (lane < nb) { if (out + lane < end) out[lane] = (u8)(bits >> (8 * lane)); else overflow = 1; }
      if (out + nb > end) overflow = 1;
      out += nb;
      bits = nb < 8 ? bits >> (8 * nb) : 0;
      nbits &= 7;
    }
  }
  __device__ __forceinline__ void flush() { if (nbits > 0) byte((u32)bits & 255); bits = 0; nbits = 0; }
  // append cnt bytes src[0..cnt) at the current bit phase (== cnt x putb(byte, 8)); all lanes help
  __device__ __forceinline__ void bytes(const u8* __restrict__ src, u32 cnt) {
    const u32 lane = lane_id();
    const u32 s = nbits;  // 0..7 pending bits
    u32 carry = (u32)bits;
    for (u32 b = 0; b < cnt; b += 32) {
      const u32 idx = b + lane;
      const u32 c = idx < cnt ? (u32)src[idx] : 0u;
      const u32 prev = __shfl_up_sync(ZQ_FULL, c, 1);
      const u32 low = lane == 0 ? carry : (prev >> (8 - s));
      const u32 o = (low | (c << s)) & 255u;
      if (idx < cnt) { if (out + idx < end) out[idx] = (u8)o; else overflow = 1; }
      const u32 last = min(cnt - b, 32u) - 1;
      carry = __shfl_sync(ZQ_FULL, c, last) >> (8 - s);
    }
    overflow = __any_sync(ZQ_FULL, overflow) ? 1u : 0u;
    out += cnt;
    bits = carry & ((1u << s) - 1);
  }
};

struct LzParams {
  u32 level;       // 1 = variable-length codes, 2 = byte codes
  u32 minMatch, lookahead, bucket, rb, checkbits;
};

// Interleaved Elias-gamma body of v >= 1 as the reference emits it LSB first: for every bit of v
// below the leading one, MSB first: a 1 then the bit; then a terminating 0 (Z:19535-19543,
// Z:19572-19576).  Returns the bits; *nb = 2*(bitlen(v)-1)+1.
__device__ __forceinline__ u32 gamma_code(u32 v, u32* nb) {
  const u32 L = (u32)zq_bitlen(v);
  *nb = 2 * (L - 1) + 1;
  if (L <= 1) return 0;
  u32 r = __brev(v) >> (33 - L);       // bit j = bit (L-2-j) of v
  r = (r | (r << 8)) & 0x00FF00FFu;
  r = (r | (r << 4)) & 0x0F0F0F0Fu;
  r = (r | (r << 2)) & 0x33333333u;
  r = (r | (r << 1)) & 0x55555555u;    // bit j -> position 2j
  return ((r << 1) | 0x55555555u) & ((1u << (2 * (L - 1))) - 1);
}

__device__ __forceinline__ void lz_write_literal(WarpSink& sk, const LzParams& P, const u8* __restrict__ in, u32 i, u32& lit) {
  if (P.level == 1) {
    if (lit < 1) return;
    u32 nb;
    const u32 g = gamma_code(lit, &nb);
    sk.putb((u64)g << 2, nb + 2);      // "00" then the length
    sk.bytes(in + (i - lit), lit);
    lit = 0;
  } else {
    while (lit > 0) {
      const u32 l1 = lit > 64 ? 64 : lit;
      sk.byte(l1 - 1);
      sk.bytes(in + (i - lit), l1);   // bit phase is always 0 for byte codes
      lit -= l1;
    }
  }
}

__device__ __forceinline__ void lz_write_match(WarpSink& sk, const LzParams& P, u32 len, u32 off) {
  if (P.level == 1) {
    off += (1u << P.rb) - 1;
    const u32 lo = (u32)zq_bitlen(off) - 1 - P.rb;
    u32 nb;
    const u32 g = gamma_code(len >> 2, &nb);
    // mm, mmm, gamma(len/4), ll
    const u64 head = (u64)((lo + 8) >> 3) | ((u64)(lo & 7) << 2) | ((u64)g << 5) | ((u64)(len & 3) << (5 + nb));
    sk.putb(head, 7 + nb);
    // r (rb low offset bits), q (lo bits below the leading one)
    sk.putb(((u64)off & ((1u << P.rb) - 1)) | ((u64)((off >> P.rb) & ((1u << lo) - 1)) << P.rb), P.rb + lo);
  } else {
    const u32 mm = P.minMatch;
    --off;
    while (len > 0) {
      const u32 l1 = len > mm * 2 + 63 ? mm + 63 : len > mm + 63 ? len - mm : len;
      if (off < (1u << 16)) { sk.byte(64 + l1 - mm); sk.byte(off >> 8); sk.byte(off); }
      else if (off < (1u << 24)) { sk.byte(128 + l1 - mm); sk.byte(off >> 16); sk.byte(off >> 8); sk.byte(off); }
      else { sk.byte(192 + l1 - mm); sk.byte(off >> 24); sk.byte(off >> 16); sk.byte(off >> 8); sk.byte(off); }
      len -= l1;
    }
  }
}

// first l in [start, limit] with a[l] != b[l] (limit if none); warp cooperative
__device__ __forceinline__ u32 warp_match_len(const u8* __restrict__ a, const u8* __restrict__ b, u32 start, u32 limit) {
  const u32 lane = lane_id();
  for (u32 base = start; base < limit; base += 32) {
    const u32 idx = base + lane;
    const bool diff = idx < limit ? (a[idx] != b[idx]) : true;
    const u32 mk = __ballot_sync(ZQ_FULL, diff);
    if (mk) return min(limit, base + (u32)(__ffs(mk) - 1));
  }
  return limit;
}

struct LzBest { u32 blen, bp, blit; int bscore; };

// One candidate per lane, in scan order = ascending lane inside the segment `segmask`.
struct LzCand {
  bool valid, capped;
  u32 p, l, l1;
  int score;
};

// The 16 suffix-array neighbours below (lanes 0-15) and above (lanes 16-31) row q: suffix start,
// LCP with the adjacent row towards q (raw, not yet a running minimum), preceding text byte.
struct LzChunk { u32 s, e, bw; bool inr; };

template <typename IdxT>
__device__ __forceinline__ LzChunk lz_chunk_issue(const IdxT* __restrict__ sa, const u16* __restrict__ lcp,
                                                  const u8* __restrict__ bwt, u32 n, u32 q, u32 bucket, bool enable) {
  const u32 lane = lane_id(), half = lane >> 4, kk = (lane & 15) + 1;
  LzChunk c;
  c.inr = enable && kk <= bucket && (half == 0 ? q >= kk : (u64)q + kk < n);
  const u32 x = half == 0 ? q - kk : q + kk;
  c.s = 0; c.e = 0; c.bw = 0;
  if (c.inr) { c.s = sa[x]; c.e = lcp[half == 0 ? x + 1 : x]; c.bw = bwt[x]; }
  return c;
}

__device__ __forceinline__ int lz_score(u32 l, u32 l1, u32 dist, u32 lit, u32 h) {
  int sc = (int)(l - l1) * 8 - zq_bitlen(dist) - ((lit == 0 && l1 > 0) ? 4 : 0) - 11;
  for (u32 a = 0; a < h; ++a) sc = sc * 5 / 8;
  return sc;
}

// Per-lane evaluation of the neighbour s (= sa[x]) whose LCP with the suffix at i+h is pm.
// bw = in[s-1] (the row's BWT byte), ci = in[i+h-1]: first step of the backward extension.
__device__ __forceinline__ LzCand lz_eval(const u8* __restrict__ in, u32 i, u32 h, u32 lit, bool inr, u32 s, u32 pm, u32 bw, u32 ci) {
  LzCand c;
  c.p = s - h;                       // wraps for s < h; rejected by p < i exactly like the reference
  c.valid = inr && c.p < i;
  c.capped = pm >= ZQ_LCP_CAP;
  c.l = h + pm;
  u32 l1 = h;
  if (h > 0 && c.valid && bw == ci) { --l1; while (l1 > 0 && in[c.p + l1 - 1] == in[i + l1 - 1]) --l1; }
  c.l1 = l1;
  c.score = lz_score(c.l, l1, i - c.p, lit, h);
  return c;
}

// Replays the reference's sequential candidate loop (Z:19414-19426) over the lanes of one segment:
// take a candidate iff its score beats every earlier one (strict), stop after the first candidate
// whose length is below the current best / below minMatch / above 255.  `exmax` is the exclusive
// running maximum of the valid scores inside the segment.  Returns true if the scan ended.
__device__ __forceinline__ bool lz_resolve(u32 segmask, const LzCand& c, int exmax, LzBest& b, u32 minMatch,
                                           const u8* __restrict__ in, u32 i, u32 h, u32 lit, u32 lmax) {
  const u32 lane = lane_id();
  const u32 vm = __ballot_sync(ZQ_FULL, c.valid) & segmask;
  if (!vm) return false;
  const int f = __ffs(vm) - 1;
  if (__shfl_sync(ZQ_FULL, (int)c.capped, f)) {
    // LCP >= 256: the only candidate the reference evaluates from here on; measure it exactly
    const u32 cp = __shfl_sync(ZQ_FULL, c.p, f);
    u32 l = warp_match_len(in + cp, in + i, h + ZQ_LCP_CAP, lmax);
    l = min(l, lmax);
    const u32 l1 = __shfl_sync(ZQ_FULL, c.l1, f);
    const int sc = lz_score(l, l1, i - cp, lit, h);
    if (sc > b.bscore) { b.blen = l; b.bp = cp; b.blit = l1; b.bscore = sc; }
    return true;   // l > 255 always ends the scan
  }
  const int M = max(b.bscore, exmax);
  const bool acc = c.valid && c.score > M;
  const u32 am = __ballot_sync(ZQ_FULL, acc) & segmask;
  const u32 la = am & (lanemask_lt() | (1u << lane));
  const u32 lsrc = __shfl_sync(ZQ_FULL, c.l, la ? 31 - __clz(la) : lane);
  const u32 bl = la ? lsrc : b.blen;
  const bool brk = c.valid && (c.l < bl || c.l < minMatch || c.l > 255);
  const u32 bm = __ballot_sync(ZQ_FULL, brk) & segmask;
  const u32 upto = bm ? (0xffffffffu >> (31 - (__ffs(bm) - 1))) : 0xffffffffu;
  const u32 fa = am & upto;
  if (fa) {
    const int w = 31 - __clz(fa);
    b.blen = __shfl_sync(ZQ_FULL, c.l, w);
    b.bp = __shfl_sync(ZQ_FULL, c.p, w);
    b.blit = __shfl_sync(ZQ_FULL, c.l1, w);
    b.bscore = __shfl_sync(ZQ_FULL, c.score, w);
  }
  return bm != 0;
}

// All candidates of look-ahead h at position i (row q = isa[i+h], first 16 neighbours per
// direction already in `ch`): both directions, in the reference's order, updating b.
// Fast path: in most scans the first usable neighbour already ends the scan (its match is shorter
// than minMatch or than the best so far), so it alone is evaluated, with warp-uniform scalars; its
// LCP with row q is one REDUX min over the lanes between it and q.  Only otherwise are all lanes
// scored and the sequential semantics resolved with scans (lz_resolve).
template <typename IdxT>
__device__ __forceinline__ void lz_scan_pos(const u8* __restrict__ in, u32 n, const IdxT* __restrict__ sa,
                                            const u16* __restrict__ lcp, const u8* __restrict__ bwt, const LzParams& P,
                                            u32 i, u32 h, u32 lit, u32 lmax, u32 q, const LzChunk& ch, LzBest& b) {
  const u32 lane = lane_id();
  const u32 ci = h > 0 ? (u32)in[i + h - 1] : 0u;
  const u32 p = ch.s - h;
  const u32 vm_all = __ballot_sync(ZQ_FULL, ch.inr && p < i);
  const u32 outm = __ballot_sync(ZQ_FULL, !ch.inr);
  bool have = false;
  LzCand c; int exmax = INT_MIN; u32 pm = 0;
  for (u32 dir = 0; dir < 2; ++dir) {
    const u32 segmask = dir ? 0xffff0000u : 0x0000ffffu;
    const u32 vm = vm_all & segmask;
    bool stop = false;
    bool general = false;
    if (vm) {
      // serial fast path: up to three usable neighbours, each with warp-uniform scalars
      u32 rest = vm;
      for (int tries = 0; rest; ++tries) {
        if (tries == 3) { general = true; break; }
        const int f = __ffs(rest) - 1;
        rest &= rest - 1;
        const u32 upto = segmask & (0xffffffffu >> (31 - f));
        const u32 pmf = __reduce_min_sync(ZQ_FULL, ((upto >> lane) & 1u) ? ch.e : 0xffffffffu);
        if (pmf >= ZQ_LCP_CAP) { general = true; break; }   // only ever the first one (LCPs shrink away from q)
        const u32 cp = __shfl_sync(ZQ_FULL, p, f);
        const u32 lf = h + pmf;
        u32 l1 = h;
        if (h > 0 && __shfl_sync(ZQ_FULL, ch.bw, f) == ci) { --l1; while (l1 > 0 && in[cp + l1 - 1] == in[i + l1 - 1]) --l1; }
        const int sc = lz_score(lf, l1, i - cp, lit, h);
        if (sc > b.bscore) { b.blen = lf; b.bp = cp; b.blit = l1; b.bscore = sc; }
        if (lf < b.blen || lf < P.minMatch || lf > 255) { stop = true; break; }
      }
      // (re-running the neighbours already taken above through the full resolve is idempotent: their
      //  scores no longer beat the best and their lengths are not below it)
      if (general) {
        if (!have) {
          pm = ch.e;
#pragma unroll
          for (int o = 1; o < 16; o <<= 1) { const u32 t = __shfl_up_sync(ZQ_FULL, pm, o, 16); if ((lane & 15) >= (u32)o) pm = min(pm, t); }
          c = lz_eval(in, i, h, lit, ch.inr, ch.s, pm, ch.bw, ci);
          int im = c.valid ? c.score : INT_MIN;
#pragma unroll
          for (int o = 1; o < 16; o <<= 1) { const int t = __shfl_up_sync(ZQ_FULL, im, o, 16); if ((lane & 15) >= (u32)o) im = max(im, t); }
          exmax = __shfl_up_sync(ZQ_FULL, im, 1, 16);
          if ((lane & 15) == 0) exmax = INT_MIN;
          have = true;
        }
        stop = lz_resolve(segmask, c, exmax, b, P.minMatch, in, i, h, lit, lmax);
      }
    }
    if (stop || (outm & segmask)) continue;
    // ---- rare: more than 16 neighbours needed in this direction; 32 per round from k = 17
    u32 run_min = __reduce_min_sync(ZQ_FULL, ((segmask >> lane) & 1u) ? ch.e : 0xffffffffu);
    for (u32 k0 = 16; k0 < P.bucket && !stop; k0 += 32) {
      const u32 k = k0 + lane + 1;
      const bool inr2 = k <= P.bucket && (dir == 0 ? q >= k : (u64)q + k < n);
      const u32 x2 = dir == 0 ? q - k : q + k;
      u32 s2 = 0, pm2 = 0, bw2 = 0;
      if (inr2) { s2 = sa[x2]; pm2 = lcp[dir == 0 ? x2 + 1 : x2]; bw2 = bwt[x2]; }
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(ZQ_FULL, pm2, o); if (lane >= (u32)o) pm2 = min(pm2, t); }
      pm2 = min(pm2, run_min);
      run_min = __shfl_sync(ZQ_FULL, pm2, 31);
      const LzCand c2 = lz_eval(in, i, h, lit, inr2, s2, pm2, bw2, ci);
      int im2 = c2.valid ? c2.score : INT_MIN;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(ZQ_FULL, im2, o); if (lane >= (u32)o) im2 = max(im2, t); }
      int ex2 = __shfl_up_sync(ZQ_FULL, im2, 1);
      if (lane == 0) ex2 = INT_MIN;
      stop = lz_resolve(0xffffffffu, c2, ex2, b, P.minMatch, in, i, h, lit, lmax);
      if (__ballot_sync(ZQ_FULL, !inr2)) break;   // ran off the suffix array: nothing further
    }
  }
}

// token decision + output (Z:19472-19519); returns how far i advances
__device__ __forceinline__ u32 lz_emit_step(WarpSink& sk, const LzParams& P, const u8* __restrict__ in, u32 i,
                                            const LzBest& b, u32& lit) {
  const u32 off = i - b.bp;
  u32 adv;
  if (off > 0 && b.bscore > 0 &&
      b.blen - b.blit >= P.minMatch + (P.level == 2 ? (u32)(off >= (1u << 16)) + (u32)(off >= (1u << 24)) : 0u)) {
    lit += b.blit;
    lz_write_literal(sk, P, in, i + b.blit, lit);
    lz_write_match(sk, P, b.blen - b.blit, off);
    adv = b.blen;
  } else {
    adv = 1; ++lit;
  }
  if (lit >= (1u << 12)) lz_write_literal(sk, P, in, i + adv, lit);
  return adv;
}

// Parse block in[0..n) with its suffix array / inverse / capped LCP / BWT bytes (index type u16 for
// n <= 65536).  Generic form: any look-ahead.
template <typename IdxT>
__device__ void lz77_sa_parse(const u8* __restrict__ in, u32 n, const IdxT* __restrict__ sa, const IdxT* __restrict__ isa,
                              const u16* __restrict__ lcp, const u8* __restrict__ bwt, const LzParams P, WarpSink& sk) {
  const u32 maxMatch = 3u << 14;
  u32 i = 0, lit = 0;
  while (i < n) {
    LzBest b; b.blen = P.minMatch - 1; b.bp = 0; b.blit = 0; b.bscore = 0;
    const u32 lmax = min(maxMatch, n - i);
    for (u32 h = 0; h <= P.lookahead; ++h) {
      const u32 pos = i + h;
      // the reference's windowed ISA only resolves positions in i's 2^checkbits window (Z:19405-19412)
      if (pos >= n || (pos >> P.checkbits) != (i >> P.checkbits)) continue;
      const u32 q = isa[pos];
      const LzChunk ch = lz_chunk_issue(sa, lcp, bwt, n, q, P.bucket, true);
      lz_scan_pos(in, n, sa, lcp, bwt, P, i, h, lit, lmax, q, ch, b);
      if (b.bscore <= 0 || b.blen < P.minMatch) break;
    }
    i += lz_emit_step(sk, P, in, i, b, lit);
  }
  lz_write_literal(sk, P, in, n, lit);
  sk.flush();
}

// ---------------------------------------------------------------------------------------------------
// Hash-table search variant of the same parse (LZBuffer::fill, Z:19434-19471, index update
// Z:19490-19515): methods with args[5]-args[0] < 21, i.e. -m1 and the low-redundancy forms of -m2..-m4.
// One warp per block.  The reference's rolling context hashes
//     h1 <- ((h1*5 << shift1) + (in[i+minMatch]+1)*123456791) mod 2^bits
// forget a byte after ceil(bits/shift1) steps (the multiplier is 5*2^shift1), so the hash in force at
// position j is a closed form of the next few input bytes: every lane derives it for its own position
// and 32 positions are indexed per round, the highest position winning a slot like the sequential loop.
// The bucket probes ht[h1^k] are fetched by the lanes at once and resolved in k order.
struct LzHashParams {
  u32 minMatch2, htbits, shift1, shift2, checkbits, frozen_from;  // updates stop at i + minMatchBoth >= n
};

__device__ __forceinline__ u32 lz_roll_hash(const u8* __restrict__ in, u32 j, u32 upto, u32 bits, u32 shift, u32 mult, u32 cmul, u32 ahead) {
  // hash state before position j (after updates of positions 0..j-1), j <= upto
  (void)upto;
  const u32 T = (bits + shift - 1) / shift;       // updates that still influence the low `bits` bits
  u32 h = 0;
  const u32 first = j > T ? j - T : 0;
  for (u32 t = first; t < j; ++t) h = (h * mult << shift) + ((u32)in[t + ahead] + 1u) * cmul;
  return h & ((1u << bits) - 1u);
}

__device__ void lz77_hash_parse(const u8* __restrict__ in, u32 n, u32* __restrict__ ht, const LzParams P, const LzHashParams Q,
                                WarpSink& sk) {
  const u32 lane = lane_id();
  const u32 maxMatch = 3u << 14;
  const u32 mask = (1u << Q.checkbits) - 1;
  const u32 htmask = (1u << Q.htbits) - 1;
  const u32 stop = Q.frozen_from;   // positions >= stop no longer update the index (i + minMatchBoth >= n)
  u32 i = 0, lit = 0;
  while (i < n) {
    LzBest b; b.blen = P.minMatch - 1; b.bp = 0; b.blit = 0; b.bscore = 0;
    const u32 lmax = min(maxMatch, n - i);
    const u32 hpos = min(i, stop);
    if (P.level == 1 || P.minMatch <= 64) {
      for (int pass = (Q.minMatch2 > 0 ? 0 : 1); pass < 2; ++pass) {
        if (pass == 1 && Q.minMatch2 && b.blen >= Q.minMatch2) break;
        const bool hi = pass == 0;   // high-order context first
        const u32 hh = hi ? lz_roll_hash(in, hpos, stop, Q.htbits, Q.shift2, 9u, 23456789u, Q.minMatch2 + P.lookahead)
                          : lz_roll_hash(in, hpos, stop, Q.htbits, Q.shift1, 5u, 123456791u, P.minMatch);
        const u32 start = hi ? P.lookahead : 0u;
        const u32 ci3 = i + 3 < n ? (u32)in[i + 3] : 0u;
        bool done = false;
        for (u32 k0 = 0; k0 <= P.bucket && !done; k0 += 32) {
          const u32 k = k0 + lane;
          u32 p = 0; bool ok = false; u32 l = 0;
          if (k <= P.bucket) {
            p = ht[(hh ^ k) & htmask];
            ok = p != 0 && (hi || i + 3 < n) && (p & mask) == (ci3 & mask);
            p >>= Q.checkbits;
            ok = ok && p < i;
            if (ok) { l = start; const u32 cap = min(lmax, start + 64u); while (l < cap && in[p + l] == in[i + l]) ++l; }
          }
          u32 vm = __ballot_sync(ZQ_FULL, k <= P.bucket);
          const u32 okm = __ballot_sync(ZQ_FULL, ok);
          while (vm) {   // k order; state-dependent pre-check replayed literally
            const int src = __ffs(vm) - 1;
            vm &= vm - 1;
            if ((okm >> src) & 1u) {
              const u32 cp = __shfl_sync(ZQ_FULL, p, src);
              u32 cl = __shfl_sync(ZQ_FULL, l, src);
              if (i + b.blen <= n && in[cp + b.blen - 1] == in[i + b.blen - 1]) {
                if (cl >= start + 64u) cl = warp_match_len(in + cp, in + i, cl, lmax);
                if (hi) {
                  if (cl >= Q.minMatch2 + P.lookahead) {
                    u32 l1 = P.lookahead;
                    while (l1 > 0 && in[cp + l1 - 1] == in[i + l1 - 1]) --l1;
                    const int sc = (int)(cl - l1) * 8 - zq_bitlen(i - cp) - ((lit == 0 && l1 > 0) ? 8 : 0) - 11;
                    if (sc > b.bscore) { b.blen = cl; b.bp = cp; b.blit = l1; b.bscore = sc; }
                  }
                } else {
                  const int sc = (int)cl * 8 - zq_bitlen(i - cp) - (lit > 0 ? 2 : 0) - 11;
                  if (sc > b.bscore) { b.blen = cl; b.bp = cp; b.blit = 0; b.bscore = sc; }
                }
              }
            }
            if (b.blen >= 128) { done = true; break; }
          }
        }
      }
    }
    const u32 adv = lz_emit_step(sk, P, in, i, b, lit);
    // index the adv consumed positions
    const u32 jend = min(i + adv, stop);
    if (Q.minMatch2 == 0) {
      for (u32 j0 = i; j0 < jend; j0 += 32) {
        const u32 j = j0 + lane;
        const bool act = j < jend;
        u32 slot = 0xffffffffu, val = 0;
        if (act) {
          const u32 h1 = lz_roll_hash(in, j, stop, Q.htbits, Q.shift1, 5u, 123456791u, P.minMatch);
          const u32 ih = ((j * 1234547u) >> 19) & P.bucket;
          slot = (h1 ^ ih) & htmask;
          val = (j << Q.checkbits) | ((u32)in[j + 3] & mask);
        }
        // the later position wins a shared slot: a lane stores unless a higher lane targets the same slot
        bool loses = false;
#pragma unroll
        for (int o = 1; o < 32; ++o) { const u32 other = __shfl_down_sync(ZQ_FULL, slot, o); if (lane + o < 32 && other == slot) loses = true; }
        if (act && !loses) ht[slot] = val;
      }
    } else if (lane == 0) {
      // second context order present (never produced by the digit methods): literal sequential update
      for (u32 j = i; j < jend; ++j) {
        const u32 ih = ((j * 1234547u) >> 19) & P.bucket;
        const u32 val = (j << Q.checkbits) | ((u32)in[j + 3] & mask);
        const u32 h2 = lz_roll_hash(in, j, stop, Q.htbits, Q.shift2, 9u, 23456789u, Q.minMatch2 + P.lookahead);
        const u32 h1 = lz_roll_hash(in, j, stop, Q.htbits, Q.shift1, 5u, 123456791u, P.minMatch);
        ht[(h2 ^ ih) & htmask] = val;
        ht[(h1 ^ ih) & htmask] = val;
      }
    }
    __syncwarp();
    i += adv;
  }
  lz_write_literal(sk, P, in, n, lit);
  sk.flush();
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB)
k_lz77_hash(const u8* __restrict__ in_base, const ZqUnit* __restrict__ units, const ZqPlan* __restrict__ plans,
            const int* __restrict__ todo, int ntodo, u8* __restrict__ work_base,
            u8* __restrict__ lz_base, u32* __restrict__ lz_len, u32* __restrict__ err_flag, u32* __restrict__ next_unit) {
  for (;;) {
    int t = 0;
    if (lane_id() == 0) t = (int)atomicAdd(next_unit, 1u);
    t = __shfl_sync(ZQ_FULL, t, 0);
    if (t >= ntodo) break;
    const int ui = todo[t];
    const ZqUnit u = units[ui];
    const ZqPlan pl = plans[u.plan];
    LzParams P;
    P.level = pl.lz_level; P.minMatch = pl.args[2]; P.lookahead = pl.args[6];
    P.bucket = (1u << pl.args[4]) - 1; P.rb = pl.args[0] > 4 ? pl.args[0] - 4 : 0; P.checkbits = 12 - pl.args[0];
    LzHashParams Q;
    Q.minMatch2 = pl.args[3]; Q.htbits = pl.args[5]; Q.checkbits = 12 - pl.args[0];
    Q.shift1 = P.minMatch > 0 ? (pl.args[5] - 1) / P.minMatch + 1 : 1;
    Q.shift2 = Q.minMatch2 > 0 ? (pl.args[5] - 1) / Q.minMatch2 + 1 : 0;
    const u32 mmb = max(P.minMatch, Q.minMatch2 + P.lookahead) + 4;
    Q.frozen_from = u.n > mmb ? u.n - mmb : 0;
    WarpSink sk; sk.init(lz_base + u.lz_off, u.lz_cap);
    lz77_hash_parse(in_base + u.in_off, u.n, (u32*)(work_base + u.work_off), P, Q, sk);
    if (lane_id() == 0) {
      lz_len[ui] = (u32)(sk.out - (lz_base + u.lz_off));
      if (sk.overflow) atomicOr(err_flag, 1u);
    }
  }
}

// BWT pre-pass (LZBuffer::fill level 3, Z:19383-19393): last column with the end-of-string row
// coded as 255, then that row's index, 4 bytes LSB first (n+5 bytes). One CTA per block.
template <typename IdxT>
__device__ void bwt_emit(const u8* __restrict__ in, u32 n, const IdxT* __restrict__ sa, u8* __restrict__ out) {
  __shared__ u32 idx_s;
  if (threadIdx.x == 0) idx_s = 0;
  __syncthreads();
  for (u32 i = threadIdx.x; i <= n; i += blockDim.x) {
    if (i == 0) out[0] = n > 0 ? in[n - 1] : 255;
    else {
      const u32 s = sa[i - 1];
      if (s == 0) { out[i] = 255; idx_s = i; }
      else out[i] = in[s - 1];
    }
  }
  __syncthreads();
  if (threadIdx.x < 4) out[n + 1 + threadIdx.x] = (u8)(idx_s >> (8 * threadIdx.x));
  __syncthreads();
}

__global__ void __launch_bounds__(256)
k_bwt_stream(const u8* __restrict__ in_base, const ZqUnit* __restrict__ units, const int* __restrict__ todo, int ntodo,
             const u8* __restrict__ work_base, u8* __restrict__ lz_base, u32* __restrict__ lz_len) {
  for (int t = blockIdx.x; t < ntodo; t += gridDim.x) {
    const int ui = todo[t];
    const ZqUnit u = units[ui];
    const u8* w = work_base + u.work_off;
    if (u.idx16) bwt_emit<u16>(in_base + u.in_off, u.n, (const u16*)w, lz_base + u.lz_off);
    else bwt_emit<u32>(in_base + u.in_off, u.n, (const u32*)w, lz_base + u.lz_off);
    if (threadIdx.x == 0) lz_len[ui] = u.n + 5;
  }
}

// One warp per unit (persistent warps pull units from a counter).  The general form of the parse: any look-ahead
// and bucket size.  The built-in methods go through the position-parallel pipeline of zq_lz77_scan.cuh instead;
// this kernel serves what that one does not cover, and its per-position evaluator (lz_scan_pos) is the exact
// slow path of the pipeline's walk.
// work layout per unit (bytes from work_base + work_off): sa | isa | lcp | bwt, index width 2 B when
// n <= 65536 (ZqUnit::idx16) else 4 B, each array padded to 128 B.
template <typename IdxT>
__global__ void __launch_bounds__(128, 6)
k_lz77_sa(const u8* __restrict__ in_base, const ZqUnit* __restrict__ units, const ZqPlan* __restrict__ plans,
          const int* __restrict__ todo, int ntodo, const u8* __restrict__ work_base,
          u8* __restrict__ lz_base, u32* __restrict__ lz_len, u32* __restrict__ err_flag, u32* __restrict__ next_unit) {
  // persistent warps pull units from a shared counter: unit costs vary a lot, a static split leaves a tail
  for (;;) {
    int t = 0;
    if (lane_id() == 0) t = (int)atomicAdd(next_unit, 1u);
    t = __shfl_sync(ZQ_FULL, t, 0);
    if (t >= ntodo) break;
    const int ui = todo[t];
    const ZqUnit u = units[ui];
    const ZqPlan pl = plans[u.plan];
    LzParams P;
    P.level = pl.lz_level; P.minMatch = pl.args[2]; P.lookahead = pl.args[6];
    P.bucket = (1u << pl.args[4]) - 1; P.rb = pl.args[0] > 4 ? pl.args[0] - 4 : 0; P.checkbits = 17 + pl.args[0];
    WarpSink sk; sk.init(lz_base + u.lz_off, u.lz_cap);
    const u8* w = work_base + u.work_off;
    const u64 stride = zq_work_stride(u.n, sizeof(IdxT));
    const u16* lcp = (const u16*)(w + 2 * stride);
    const u8* bwt = w + 2 * stride + zq_work_stride(u.n, 2);
    const u8* in = in_base + u.in_off;
    lz77_sa_parse<IdxT>(in, u.n, (const IdxT*)w, (const IdxT*)(w + stride), lcp, bwt, P, sk);
    if (lane_id() == 0) {
      lz_len[ui] = (u32)(sk.out - (lz_base + u.lz_off));
      if (sk.overflow) atomicOr(err_flag, 1u);
    }
  }
}

}  // namespace zqdev
