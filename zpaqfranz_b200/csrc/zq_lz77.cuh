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

// Sequential byte/bit sink owned by one warp. State is warp-uniform; lane 0 performs scalar stores.
struct WarpSink {
  u8* out;      // next byte to write
  u8* end;      // capacity guard (overflow -> flag, bytes dropped)
  u32 bits;     // pending bits (level-1 codes are LSB first)
  u32 nbits;
  u32 overflow;
  __device__ __forceinline__ void init(u8* o, u32 cap) { out = o; end = o + cap; bits = nbits = overflow = 0; }
  __device__ __forceinline__ void byte(u32 c) {
    if (out < end) { if (lane_id() == 0) *out = (u8)c; } else overflow = 1;
    ++out;
  }
  __device__ __forceinline__ void putb(u32 x, int k) {  // k <= 24
    x &= (1u << k) - 1;
    bits |= x << nbits; nbits += k;
    while (nbits > 7) { byte(bits & 255); bits >>= 8; nbits -= 8; }
  }
  __device__ __forceinline__ void flush() { if (nbits > 0) byte(bits & 255); bits = nbits = 0; }
  // append cnt bytes src[0..cnt) at the current bit phase (== cnt x putb(byte, 8)); all lanes help
  __device__ __forceinline__ void bytes(const u8* __restrict__ src, u32 cnt) {
    const u32 lane = lane_id();
    const u32 s = nbits;  // 0..7 pending bits
    u32 carry = bits;
    for (u32 b = 0; b < cnt; b += 32) {
      const u32 idx = b + lane;
      const u32 c = idx < cnt ? (u32)src[idx] : 0u;
      u32 prev = __shfl_up_sync(ZQ_FULL, c, 1);
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

__device__ __forceinline__ void lz_write_literal(WarpSink& sk, const LzParams& P, const u8* __restrict__ in, u32 i, u32& lit) {
  if (P.level == 1) {
    if (lit < 1) return;
    int ll = zq_bitlen(lit);
    sk.putb(0, 2);
    --ll;
    while (--ll >= 0) { sk.putb(1, 1); sk.putb((lit >> ll) & 1, 1); }
    sk.putb(0, 1);
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
    int ll = zq_bitlen(len) - 1;
    off += (1u << P.rb) - 1;
    const int lo = zq_bitlen(off) - 1 - (int)P.rb;
    sk.putb((lo + 8) >> 3, 2);
    sk.putb(lo & 7, 3);
    while (--ll >= 2) { sk.putb(1, 1); sk.putb((len >> ll) & 1, 1); }
    sk.putb(0, 1);
    sk.putb(len & 3, 2);
    sk.putb(off, P.rb);
    sk.putb(off >> P.rb, lo);
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

// Parse block in[0..n) with its suffix array; returns stream length in *out_len (lane-uniform).
__device__ void lz77_sa_parse(const u8* __restrict__ in, u32 n, const u32* __restrict__ sa, const u32* __restrict__ isa,
                              const u16* __restrict__ lcp, const LzParams P, WarpSink& sk) {
  const u32 lane = lane_id();
  const u32 maxMatch = 3u << 14, maxLiteral = 1u << 12;
  const u32 minMatch = P.minMatch;
  u32 i = 0, lit = 0;
  while (i < n) {
    u32 blen = minMatch - 1, bp = 0, blit = 0; int bscore = 0;
    for (u32 h = 0; h <= P.lookahead; ++h) {
      const u32 pos = i + h;
      // the reference's windowed ISA only resolves positions in i's 2^checkbits window (Z:19405-19412)
      if (pos >= n || (pos >> P.checkbits) != (i >> P.checkbits)) continue;
      const u32 q = isa[pos];
      for (int dir = 0; dir < 2; ++dir) {           // 0: towards smaller suffixes, 1: larger
        u32 run_min = 0xffffffffu;
        bool stop = false;
        for (u32 k0 = 0; k0 < P.bucket && !stop; k0 += 32) {
          const u32 k = k0 + lane + 1;
          const bool inr = k <= P.bucket && (dir == 0 ? q >= k : (u64)q + k < n);
          const u32 x = dir == 0 ? q - k : q + k;
          u32 s = 0, e = 0;
          if (inr) { s = sa[x]; e = lcp[dir == 0 ? x + 1 : x]; }
          // running minimum of adjacent LCPs = LCP(sa[x], sa[q])
          u32 pm = e;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(ZQ_FULL, pm, o); if (lane >= (u32)o) pm = min(pm, t); }
          pm = min(pm, run_min);
          run_min = __shfl_sync(ZQ_FULL, pm, 31);
          const u32 p = s - h;                       // wraps for s < h, rejected by p < i as in the reference
          u32 vmask = __ballot_sync(ZQ_FULL, inr && p < i);
          while (vmask) {
            const int src = __ffs(vmask) - 1;
            vmask &= vmask - 1;
            const u32 cp = __shfl_sync(ZQ_FULL, p, src);
            const u32 cm = __shfl_sync(ZQ_FULL, pm, src);
            u32 l = h + cm;
            const u32 lmax = min(maxMatch, n - i);
            if (cm >= ZQ_LCP_CAP) l = warp_match_len(in + cp, in + i, l, lmax);
            l = min(l, lmax);
            u32 l1 = h;
            while (l1 > 0 && in[cp + l1 - 1] == in[i + l1 - 1]) --l1;
            int score = (int)(l - l1) * 8 - zq_bitlen(i - cp) - 4 * (lit == 0 && l1 > 0) - 11;
            for (u32 a = 0; a < h; ++a) score = score * 5 / 8;
            if (score > bscore) { blen = l; bp = cp; blit = l1; bscore = score; }
            if (l < blen || l < minMatch || l > 255) { stop = true; break; }
          }
          if (!stop && __ballot_sync(ZQ_FULL, !inr)) break;  // ran off the suffix array: nothing further
        }
      }
      if (bscore <= 0 || blen < minMatch) break;
    }
    const u32 off = i - bp;
    if (off > 0 && bscore > 0 &&
        blen - blit >= minMatch + (P.level == 2 ? (u32)(off >= (1u << 16)) + (u32)(off >= (1u << 24)) : 0u)) {
      lit += blit;
      lz_write_literal(sk, P, in, i + blit, lit);
      lz_write_match(sk, P, blen - blit, off);
    } else {
      blen = 1; ++lit;
    }
    i += blen;
    if (lit >= maxLiteral) lz_write_literal(sk, P, in, i, lit);
  }
  lz_write_literal(sk, P, in, n, lit);
  sk.flush();
}

// BWT pre-pass (LZBuffer::fill level 3, Z:19383-19393): last column with the end-of-string row
// coded as 255, then that row's index, 4 bytes LSB first. Whole CTA.
__device__ void bwt_emit(const u8* __restrict__ in, u32 n, const u32* __restrict__ sa, u8* __restrict__ out, u32* idx_slot) {
  // out[0] = in[n-1] (255 if empty); out[i] for i=1..n: sa[i-1]==0 ? 255 (idx=i) : in[sa[i-1]-1]
  for (u32 i = threadIdx.x; i <= n; i += blockDim.x) {
    if (i == 0) out[0] = n > 0 ? in[n - 1] : 255;
    else {
      const u32 s = sa[i - 1];
      if (s == 0) { out[i] = 255; *idx_slot = i; }
      else out[i] = in[s - 1];
    }
  }
  if (n == 0 && threadIdx.x == 0) *idx_slot = 0;
  __syncthreads();
  if (threadIdx.x < 4) out[n + 1 + threadIdx.x] = (u8)(*idx_slot >> (8 * threadIdx.x));
}

// One warp per unit (grid-stride over the wave's units that use the SA parse).
__global__ void __launch_bounds__(128)
k_lz77_sa(const u8* __restrict__ in_base, const ZqUnit* __restrict__ units, const ZqPlan* __restrict__ plans,
          const int* __restrict__ todo, int ntodo,
          const u32* __restrict__ sa_all, const u32* __restrict__ isa_all, const u16* __restrict__ lcp_all,
          u8* __restrict__ lz_base, u32* __restrict__ lz_len, u32* __restrict__ err_flag) {
  const int warps_per_cta = blockDim.x >> 5;
  const int gw = blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
  const int nw = gridDim.x * warps_per_cta;
  for (int t = gw; t < ntodo; t += nw) {
    const int ui = todo[t];
    const ZqUnit u = units[ui];
    const ZqPlan pl = plans[u.plan];
    LzParams P;
    P.level = pl.lz_level; P.minMatch = pl.args[2]; P.lookahead = pl.args[6];
    P.bucket = (1u << pl.args[4]) - 1; P.rb = pl.args[0] > 4 ? pl.args[0] - 4 : 0; P.checkbits = 17 + pl.args[0];
    WarpSink sk; sk.init(lz_base + u.lz_off, u.lz_cap);
    lz77_sa_parse(in_base + u.in_off, u.n, sa_all + u.work_off, isa_all + u.work_off, lcp_all + u.work_off, P, sk);
    if (lane_id() == 0) {
      lz_len[ui] = (u32)(sk.out - (lz_base + u.lz_off));
      if (sk.overflow) atomicOr(err_flag, 1u);
    }
  }
}

}  // namespace zqdev
