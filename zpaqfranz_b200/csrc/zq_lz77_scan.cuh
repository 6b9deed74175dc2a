// zq_lz77_scan.cuh -- the suffix-array LZ77 parse of LZBuffer::fill (Z:19395-19430 candidate scan, Z:19472-19527
// token decision) as a position-parallel pipeline.  The parse is sequential only through two things: which
// positions are visited (i -> i + len) and whether literals are pending (the lit==0 penalty, Z:19421).  Everything
// else -- the whole candidate scan -- is a pure function of the position, so it is evaluated for EVERY position and
// both literal states up front, by one thread per suffix-array row, and the sequential part shrinks to a table walk.
//
//   k_lz_scan0  rows in SUFFIX-ARRAY order, one thread per row q (suffix s = sa[q]).  The look-ahead-0 scan of
//               position s only touches rows q-127..q+127, so a tile of rows plus a 128-row halo is brought into
//               shared memory with bulk async copies (cp.async.bulk + mbarrier, double buffered) and the scan
//               runs out of shared memory: no gathers from HBM at all.  Result (best length / pointer) is
//               scattered to r0[s].  Lengths come from the running minimum of the capped LCP array.
//   k_lz_scan1  same tiles again: row q also is the look-ahead-1 row of position s-1 (Z:19412: q = isa[i+1]).
//               It picks up r0[s-1], continues the reference's scan with h = 1 for both literal states and
//               scatters the two final decisions to f[s-1].
//   k_lz_walk   one warp per block walks i -> i + len through f[] in 32-position windows (ballots pick the next
//               match, literal runs are skipped in one hop) and writes the token list.  A position whose scan
//               met a capped LCP (>= 256 bytes: the reference measures it exactly, Z:19419) is marked DEFER by
//               the scans and evaluated here, exactly and warp-cooperatively (lz_scan_pos), only if visited.
//   k_lz_emit   tokens -> code lengths -> prefix sum -> every token writes its own bits (level 1: LSB-first bit
//               codes; level 2: byte codes) into a stream assembled in shared memory, stored with one bulk copy.
//
// Lanes finish their rows at very different times (scan lengths 1..254), so each lane pulls the next row of the
// tile when it is done, at round boundaries every LZS_ROUND steps: ~65-70 % of the lanes do useful steps instead
// of 13 % with one row per lane per loop.
// Bit-exact by construction: the per-row state machine is the reference's loop with one exact pruning rule (no
// remaining neighbour can beat the best score once 8*(h+runmin)-12, scaled, is not above it).
#pragma once
#include "zq_lz77.cuh"

namespace zqdev {

constexpr u32 LZS_HALO = 128;    // rows of context on either side of a tile (bucket <= 127)
constexpr u32 LZS_TILE = 4096;   // rows per tile
constexpr int LZS_NT = 512;      // threads per CTA of the scan kernels
constexpr int LZS_ROUND = 6;     // scan steps between two refills
constexpr u32 LZS_IDLE = 0xffffffffu;

// per-unit arrays behind sa | isa | lcp | bwt in the work region: r0[n] then f[n][2]
template <typename IdxT> struct LzsFmt;
template <> struct LzsFmt<u16> {
  typedef u32 R0T; typedef u32 FT;
  static constexpr u32 R0_DEFER = 0x80000000u;
  static constexpr u32 F_DEFER = 0x80000000u, F_MATCH = 0x40000000u;
  static __device__ __forceinline__ u32 r0_pack(u32 blen, u32 bp) { return (blen << 16) | bp; }
  static __device__ __forceinline__ u32 r0_blen(u32 r) { return (r >> 16) & 0x7fffu; }
  static __device__ __forceinline__ u32 r0_bp(u32 r) { return r & 0xffffu; }
  static __device__ __forceinline__ u32 f_pack(u32 off, u32 blen, u32 blit) { return F_MATCH | off | (blen << 16) | (blit << 25); }
  static __device__ __forceinline__ u32 f_off(u32 d) { return d & 0xffffu; }
  static __device__ __forceinline__ u32 f_blen(u32 d) { return (d >> 16) & 0x1ffu; }
  static __device__ __forceinline__ u32 f_blit(u32 d) { return (d >> 25) & 1u; }
};
template <> struct LzsFmt<u32> {
  typedef u64 R0T; typedef u64 FT;
  static constexpr u64 R0_DEFER = 1ull << 63;
  static constexpr u64 F_DEFER = 1ull << 63, F_MATCH = 1ull << 62;
  static __device__ __forceinline__ u64 r0_pack(u32 blen, u32 bp) { return ((u64)blen << 32) | bp; }
  static __device__ __forceinline__ u32 r0_blen(u64 r) { return (u32)(r >> 32) & 0xffffu; }
  static __device__ __forceinline__ u32 r0_bp(u64 r) { return (u32)r; }
  static __device__ __forceinline__ u64 f_pack(u32 off, u32 blen, u32 blit) { return F_MATCH | off | ((u64)blen << 32) | ((u64)blit << 48); }
  static __device__ __forceinline__ u32 f_off(u64 d) { return (u32)d; }
  static __device__ __forceinline__ u32 f_blen(u64 d) { return (u32)(d >> 32) & 0xffffu; }
  static __device__ __forceinline__ u32 f_blit(u64 d) { return (u32)(d >> 48) & 1u; }
};

__host__ __device__ inline u64 zq_align128(u64 x) { return (x + 127) & ~(u64)127; }
// bytes of the extended work region (sa | isa | lcp | bwt | r0 | f) of a block of n bytes with index width w
__host__ __device__ inline u64 zq_work_bytes_scan(u32 n, u32 w) {
  return zq_work_bytes(n, w) + zq_align128((u64)n * (w == 2 ? 4 : 8)) + zq_align128((u64)n * (w == 2 ? 8 : 16));
}
__host__ __device__ inline u32 lzs_tiles(u32 n) { return (n + LZS_TILE - 1) / LZS_TILE; }

template <typename IdxT>
struct LzsView {
  const IdxT* sa; const IdxT* isa; const u16* lcp; const u8* bwt;
  typename LzsFmt<IdxT>::R0T* r0; typename LzsFmt<IdxT>::FT* f;
};
template <typename IdxT>
__device__ __forceinline__ LzsView<IdxT> lzs_view(u8* work_base, u64 work_off, u32 n) {
  LzsView<IdxT> v;
  u8* w = work_base + work_off;
  const u64 stride = zq_work_stride(n, sizeof(IdxT));
  v.sa = (const IdxT*)w; v.isa = (const IdxT*)(w + stride); v.lcp = (const u16*)(w + 2 * stride);
  v.bwt = w + 2 * stride + zq_work_stride(n, 2);
  u8* r = w + zq_work_bytes(n, sizeof(IdxT));
  v.r0 = (typename LzsFmt<IdxT>::R0T*)r;
  v.f = (typename LzsFmt<IdxT>::FT*)(r + zq_align128((u64)n * sizeof(typename LzsFmt<IdxT>::R0T)));
  return v;
}

struct LzsDesc {   // one tile in flight
  u32 valid, n, lo, t0, t1, plan;
  u64 work_off;
};

template <typename IdxT>
struct LzsSmem {
  ZqMbar bar[2];
  LzsDesc d[2];
  u32 next_row[2];
  u32 pad[2];
  alignas(128) IdxT sa[2][LZS_TILE + 2 * LZS_HALO];
  alignas(128) u16 lcp[2][LZS_TILE + 2 * LZS_HALO];
  alignas(128) u8 bwt[2][LZS_TILE + 2 * LZS_HALO];
};

// thread 0: claim the next tile of the launch, describe it in slot `sl` and start its bulk copies
template <typename IdxT, bool WITH_BWT>
__device__ void lzs_fetch(LzsSmem<IdxT>& sm, int sl, const ZqUnit* __restrict__ units, const int* __restrict__ todo,
                          const u32* __restrict__ tile_first, int ntodo, u8* work_base, u32* tile_ctr) {
  LzsDesc& d = sm.d[sl];
  const u32 tile = atomicAdd(tile_ctr, 1u);
  if (tile >= tile_first[ntodo]) { d.valid = 0; return; }
  int lo = 0, hi = ntodo - 1;      // last t with tile_first[t] <= tile
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (tile_first[mid] <= tile) lo = mid; else hi = mid - 1; }
  const ZqUnit u = units[todo[lo]];
  const u32 n = u.n;
  d.valid = 1; d.n = n; d.plan = u.plan; d.work_off = u.work_off;
  d.t0 = (tile - tile_first[lo]) * LZS_TILE;
  d.t1 = min(n, d.t0 + LZS_TILE);
  d.lo = d.t0 >= LZS_HALO ? d.t0 - LZS_HALO : 0u;
  const u32 rhi = min(d.t1 + LZS_HALO, (n + 63u) & ~63u);
  const u32 rows = rhi - d.lo;
  sm.next_row[sl] = d.t0;
  const LzsView<IdxT> v = lzs_view<IdxT>(work_base, u.work_off, n);
  zq_mbar_expect_tx(&sm.bar[sl], rows * (u32)(sizeof(IdxT) + 2 + (WITH_BWT ? 1 : 0)));
  zq_bulk_g2s(sm.sa[sl], v.sa + d.lo, rows * (u32)sizeof(IdxT), &sm.bar[sl]);
  zq_bulk_g2s(sm.lcp[sl], v.lcp + d.lo, rows * 2u, &sm.bar[sl]);
  if (WITH_BWT) zq_bulk_g2s(sm.bwt[sl], v.bwt + d.lo, rows, &sm.bar[sl]);
}

struct LzsParams { u32 minMatch, bucket, lookahead, checkbits, level; };
__device__ __forceinline__ LzsParams lzs_params(const ZqPlan& pl) {
  LzsParams P;
  P.minMatch = pl.args[2]; P.bucket = (1u << pl.args[4]) - 1; P.lookahead = pl.args[6];
  P.checkbits = 17 + pl.args[0]; P.level = pl.lz_level;
  return P;
}

// final decision (Z:19474-19477) from a best candidate; off = i - bp
template <typename IdxT>
__device__ __forceinline__ typename LzsFmt<IdxT>::FT lzs_decide(const LzsParams& P, u32 off, u32 blen, u32 blit, int bscore) {
  typedef LzsFmt<IdxT> F;
  const bool match = off > 0 && bscore > 0 &&
                     blen - blit >= P.minMatch + (P.level == 2 ? (u32)(off >= (1u << 16)) + (u32)(off >= (1u << 24)) : 0u);
  return match ? F::f_pack(off, blen, blit) : (typename F::FT)0;
}

// ---- pass 1: look-ahead 0 -----------------------------------------------------------------------------------
template <typename IdxT>
__device__ void lzs_scan0_tile(const IdxT* __restrict__ s_sa, const u16* __restrict__ s_lcp, const LzsDesc& d, const LzsParams P,
                               u32* next_row, typename LzsFmt<IdxT>::R0T* __restrict__ r0) {
  typedef LzsFmt<IdxT> F;
  const u32 lane = lane_id(), n = d.n, lo = d.lo, t1 = d.t1;
  u32 q = LZS_IDLE, s = 0, k = 0, runmin = 0, blen = 0, bp = 0, dir = 0;
  int bscore = 0;
  for (;;) {
    const u32 need = __ballot_sync(ZQ_FULL, q == LZS_IDLE);
    if (need) {
      u32 base = 0;
      if (lane == 0) base = atomicAdd(next_row, (u32)__popc(need));
      base = __shfl_sync(ZQ_FULL, base, 0);
      if (q == LZS_IDLE) {
        const u32 r = base + (u32)__popc(need & lanemask_lt());
        if (r < t1) { q = r; s = s_sa[q - lo]; dir = 0; k = 1; runmin = 0xffffffffu; blen = P.minMatch - 1; bp = 0; bscore = 0; }
      }
      if (__all_sync(ZQ_FULL, q == LZS_IDLE)) break;
    }
#pragma unroll
    for (int r = 0; r < LZS_ROUND; ++r) {
      if (q != LZS_IDLE) {
        bool end_dir = false, fin = false;
        typename F::R0T res = 0;
        const bool inr = k <= P.bucket && (dir == 0 ? q >= k : q + k < n);
        if (!inr) end_dir = true;
        else {
          const u32 x = dir == 0 ? q - k : q + k;
          runmin = min(runmin, (u32)s_lcp[(dir == 0 ? x + 1 : x) - lo]);
          if ((int)(runmin * 8u) - 12 <= bscore) end_dir = true;   // exact pruning (a capped LCP never prunes: 8*256-12 > any score)
          else {
            const u32 p = s_sa[x - lo];
            ++k;
            if (p < s) {
              if (runmin >= ZQ_LCP_CAP) { fin = true; res = F::R0_DEFER; }   // exact length needed (Z:19419): the walk does it
              else {
                const int sc = (int)(runmin * 8u) - zq_bitlen(s - p) - 11;
                if (sc > bscore) { blen = runmin; bp = p; bscore = sc; }
                if (runmin < blen || runmin < P.minMatch) end_dir = true;
              }
            }
          }
        }
        if (end_dir) {
          if (dir == 0) { dir = 1; k = 1; runmin = 0xffffffffu; }
          else { fin = true; res = bscore > 0 ? F::r0_pack(blen, bp) : (typename F::R0T)0; }
        }
        if (fin) { r0[s] = res; q = LZS_IDLE; }
      }
    }
  }
}

// ---- pass 2: look-ahead 1, both literal states ---------------------------------------------------------------
__device__ __forceinline__ int lzs_scale58(int sc) { return sc * 5 / 8; }   // C division: truncates toward zero like the reference

template <typename IdxT>
__device__ void lzs_scan1_tile(const IdxT* __restrict__ s_sa, const u16* __restrict__ s_lcp, const u8* __restrict__ s_bwt,
                               const LzsDesc& d, const LzsParams P, u32* next_row,
                               const typename LzsFmt<IdxT>::R0T* __restrict__ r0, typename LzsFmt<IdxT>::FT* __restrict__ f) {
  typedef LzsFmt<IdxT> F;
  typedef typename F::FT FT;
  const u32 lane = lane_id(), n = d.n, lo = d.lo, t1 = d.t1;
  u32 q = LZS_IDLE, s = 0, k = 0, runmin = 0, dir = 0, ci = 0;
  u32 blen0 = 0, bp0 = 0, blit0 = 0, blen1 = 0, bp1 = 0, blit1 = 0;
  int bs0 = 0, bs1 = 0;
  bool stop0 = false, stop1 = false;
  for (;;) {
    bool exhausted = false;
    for (int tries = 0; tries < 4; ++tries) {   // rows that need no look-ahead scan finish at once: fill the lanes again
      const u32 need = __ballot_sync(ZQ_FULL, q == LZS_IDLE);
      if (__popc(need) < (tries ? 8 : 1)) break;
      u32 base = 0;
      if (lane == 0) base = atomicAdd(next_row, (u32)__popc(need));
      base = __shfl_sync(ZQ_FULL, base, 0);
      if (base >= t1) { exhausted = true; break; }
      if (q == LZS_IDLE) {
        const u32 r = base + (u32)__popc(need & lanemask_lt());
        if (r < t1) {
          s = s_sa[r - lo];
          // this row's job: position i = s-1 (row of suffix i+1); the row of suffix 0 takes position n-1, which has no look-ahead
          const u32 i = s > 0 ? s - 1 : n - 1;
          const typename F::R0T a = r0[i];
          if (a == F::R0_DEFER) { f[2 * (u64)i] = F::F_DEFER; f[2 * (u64)i + 1] = F::F_DEFER; }
          else {
            const u32 bl = a ? F::r0_blen(a) : P.minMatch - 1, bpp = a ? F::r0_bp(a) : 0u;
            const int bsc = a ? (int)(bl * 8u) - zq_bitlen(i - bpp) - 11 : 0;
            const bool cont = s > 0 && P.lookahead >= 1 && bsc > 0 && bl >= P.minMatch && (s >> P.checkbits) == (i >> P.checkbits);
            if (!cont) { const FT dd = lzs_decide<IdxT>(P, i - bpp, bl, 0, bsc); f[2 * (u64)i] = dd; f[2 * (u64)i + 1] = dd; }
            else {
              q = r; dir = 0; k = 1; runmin = 0xffffffffu; ci = s_bwt[r - lo];
              blen0 = blen1 = bl; bp0 = bp1 = bpp; blit0 = blit1 = 0; bs0 = bs1 = bsc; stop0 = stop1 = false;
            }
          }
        }
      }
    }
    if (exhausted && __all_sync(ZQ_FULL, q == LZS_IDLE)) break;
#pragma unroll
    for (int r = 0; r < LZS_ROUND; ++r) {
      if (q != LZS_IDLE) {
        bool end_dir = false, defer = false;
        const bool inr = k <= P.bucket && (dir == 0 ? q >= k : q + k < n);
        if (!inr) end_dir = true;
        else {
          const u32 x = dir == 0 ? q - k : q + k;
          runmin = min(runmin, (u32)s_lcp[(dir == 0 ? x + 1 : x) - lo]);
          const int ub = lzs_scale58((int)((1u + runmin) * 8u) - 12);
          if (ub <= bs0) stop0 = true;
          if (ub <= bs1) stop1 = true;
          if (stop0 && stop1) end_dir = true;
          else {
            const u32 p1 = s_sa[x - lo];
            const u32 bw = s_bwt[x - lo];
            ++k;
            if (p1 != 0 && p1 < s) {            // p = p1 - 1 < i
              if (runmin >= ZQ_LCP_CAP) defer = true;
              else {
                const u32 l = 1u + runmin;
                const u32 l1 = bw == ci ? 0u : 1u;
                const int base = (int)((l - l1) * 8u) - zq_bitlen(s - p1) - 11;
                const bool brk = l < P.minMatch || l > 255;
                if (!stop0) {
                  const int sc = lzs_scale58(base - (l1 ? 4 : 0));
                  if (sc > bs0) { blen0 = l; bp0 = p1 - 1; blit0 = l1; bs0 = sc; }
                  if (l < blen0 || brk) stop0 = true;
                }
                if (!stop1) {
                  const int sc = lzs_scale58(base);
                  if (sc > bs1) { blen1 = l; bp1 = p1 - 1; blit1 = l1; bs1 = sc; }
                  if (l < blen1 || brk) stop1 = true;
                }
                if (stop0 && stop1) end_dir = true;
              }
            }
          }
        }
        const u32 i = s - 1;
        if (defer) { f[2 * (u64)i] = F::F_DEFER; f[2 * (u64)i + 1] = F::F_DEFER; q = LZS_IDLE; }
        else if (end_dir) {
          if (dir == 0) { dir = 1; k = 1; runmin = 0xffffffffu; stop0 = stop1 = false; }
          else {
            f[2 * (u64)i] = lzs_decide<IdxT>(P, i - bp0, blen0, blit0, bs0);
            f[2 * (u64)i + 1] = lzs_decide<IdxT>(P, i - bp1, blen1, blit1, bs1);
            q = LZS_IDLE;
          }
        }
      }
    }
  }
}

// PASS 0: k_lz_scan0, PASS 1: k_lz_scan1.  Persistent CTAs pull tiles (unit, first row) from a counter.
template <typename IdxT, int PASS>
__global__ void __launch_bounds__(LZS_NT)
k_lz_scan(const ZqUnit* __restrict__ units, const ZqPlan* __restrict__ plans, const int* __restrict__ todo,
          const u32* __restrict__ tile_first, int ntodo, u8* __restrict__ work_base, u32* tile_ctr) {
  ZQ_DYN_SMEM(smem_raw);
  LzsSmem<IdxT>& sm = *reinterpret_cast<LzsSmem<IdxT>*>(smem_raw);
  const u32 tid = threadIdx.x;
  if (tid == 0) {
    zq_mbar_init(&sm.bar[0], 1); zq_mbar_init(&sm.bar[1], 1);
    lzs_fetch<IdxT, PASS == 1>(sm, 0, units, todo, tile_first, ntodo, work_base, tile_ctr);
  }
  __syncthreads();
  u32 use0 = 0, use1 = 0;
  for (int cur = 0;; cur ^= 1) {
    if (tid == 0) lzs_fetch<IdxT, PASS == 1>(sm, cur ^ 1, units, todo, tile_first, ntodo, work_base, tile_ctr);
    const LzsDesc d = sm.d[cur];
    if (!d.valid) break;
    if (cur == 0) { zq_mbar_wait(&sm.bar[0], use0 & 1u); ++use0; } else { zq_mbar_wait(&sm.bar[1], use1 & 1u); ++use1; }
    const LzsParams P = lzs_params(plans[d.plan]);
    const LzsView<IdxT> v = lzs_view<IdxT>(work_base, d.work_off, d.n);
    if (PASS == 0) lzs_scan0_tile<IdxT>(sm.sa[cur], sm.lcp[cur], d, P, &sm.next_row[cur], v.r0);
    else lzs_scan1_tile<IdxT>(sm.sa[cur], sm.lcp[cur], sm.bwt[cur], d, P, &sm.next_row[cur], v.r0, v.f);
    __syncthreads();
  }
}

// ---- walk ---------------------------------------------------------------------------------------------------------
struct LzToken { u32 pos, lit, mlen, off; };   // lit literals starting at pos, then (mlen > 0) a match

// exact decision at position i (the reference's scan with byte compares where the LCP array is capped), warp
// cooperative.  Returns match?; lengths are not bounded by the packed formats here (up to 3 << 14).
struct LzsDecision { bool match; u32 off, blen, blit; };
template <typename IdxT>
__device__ LzsDecision lzs_decide_exact(const u8* __restrict__ in, u32 n, const LzsView<IdxT>& v, const LzsParams& Q, u32 i, u32 lit) {
  LzParams P;
  P.level = Q.level; P.minMatch = Q.minMatch; P.lookahead = Q.lookahead; P.bucket = Q.bucket; P.rb = 0; P.checkbits = Q.checkbits;
  LzBest b; b.blen = P.minMatch - 1; b.bp = 0; b.blit = 0; b.bscore = 0;
  const u32 lmax = min(3u << 14, n - i);
  for (u32 h = 0; h <= P.lookahead; ++h) {
    const u32 pos = i + h;
    if (pos >= n || (pos >> P.checkbits) != (i >> P.checkbits)) continue;
    const u32 q = v.isa[pos];
    const LzChunk ch = lz_chunk_issue(v.sa, v.lcp, v.bwt, n, q, P.bucket, true);
    lz_scan_pos(in, n, v.sa, v.lcp, v.bwt, P, i, h, lit, lmax, q, ch, b);
    if (b.bscore <= 0 || b.blen < P.minMatch) break;
  }
  LzsDecision r;
  r.off = i - b.bp; r.blen = b.blen; r.blit = b.blit;
  r.match = r.off > 0 && b.bscore > 0 &&
            b.blen - b.blit >= P.minMatch + (P.level == 2 ? (u32)(r.off >= (1u << 16)) + (u32)(r.off >= (1u << 24)) : 0u);
  return r;
}

// One warp per block.  tok_off[t] = first token slot of todo entry t; ntok[t] = tokens written.
template <typename IdxT>
__global__ void __launch_bounds__(128)
k_lz_walk(const u8* __restrict__ in_base, const ZqUnit* __restrict__ units, const ZqPlan* __restrict__ plans,
          const int* __restrict__ todo, int ntodo, u8* __restrict__ work_base, const u64* __restrict__ tok_off,
          LzToken* __restrict__ tok_base, u32* __restrict__ ntok, u32* next_unit) {
  typedef LzsFmt<IdxT> F;
  typedef typename F::FT FT;
  const u32 lane = lane_id();
  for (;;) {
    int t = 0;
    if (lane == 0) t = (int)atomicAdd(next_unit, 1u);
    t = __shfl_sync(ZQ_FULL, t, 0);
    if (t >= ntodo) break;
    const ZqUnit u = units[todo[t]];
    const LzsParams P = lzs_params(plans[u.plan]);
    const u32 n = u.n;
    const u8* in = in_base + u.in_off;
    const LzsView<IdxT> v = lzs_view<IdxT>(work_base, u.work_off, n);
    LzToken* tok = tok_base + tok_off[t];
    const u32 maxLiteral = 1u << 12;
    u32 i = 0, lit = 0, nt = 0;
    LzToken mine; mine.pos = mine.lit = mine.mlen = mine.off = 0;
    // tokens are handed to lane (nt & 31) and stored 32 at a time
#define LZS_PUSH(POS, LIT, MLEN, OFF)                                                                   \
  do {                                                                                                  \
    if (lane == (nt & 31u)) { mine.pos = (POS); mine.lit = (LIT); mine.mlen = (MLEN); mine.off = (OFF); } \
    ++nt;                                                                                               \
    if ((nt & 31u) == 0) *(uint4*)&tok[nt - 32 + lane] = make_uint4(mine.pos, mine.lit, mine.mlen, mine.off); \
  } while (0)
    u32 w0 = 0;
    FT c0 = 0, c1 = 0, n0 = 0, n1 = 0;   // decisions of the current / next window for lit == 0 / lit > 0
    if (lane < n) { c0 = v.f[2 * (u64)lane]; c1 = v.f[2 * (u64)lane + 1]; }
    if (32 + lane < n) { n0 = v.f[2 * (u64)(32 + lane)]; n1 = v.f[2 * (u64)(32 + lane) + 1]; }
    while (i < n) {
      if (i >= w0 + 32) {
        if (i < w0 + 64) { w0 += 32; c0 = n0; c1 = n1; }
        else { w0 = i & ~31u; c0 = c1 = 0; if (w0 + lane < n) { c0 = v.f[2 * (u64)(w0 + lane)]; c1 = v.f[2 * (u64)(w0 + lane) + 1]; } }
        n0 = n1 = 0;
        if (w0 + 32 + lane < n) { n0 = v.f[2 * (u64)(w0 + 32 + lane)]; n1 = v.f[2 * (u64)(w0 + 32 + lane) + 1]; }
      }
      // positions of this window that do not end as a plain literal, per state
      const u32 m0 = __ballot_sync(ZQ_FULL, c0 != 0), m1 = __ballot_sync(ZQ_FULL, c1 != 0);
      const u32 wend = min(32u, n - w0);
      while (i < w0 + wend) {
        const u32 j = i - w0;
        if (lit > 0) {   // hop over the literals up to the next candidate position of the lit > 0 state
          const u32 ahead = m1 >> j;
          u32 skip = ahead ? (u32)(__ffs(ahead) - 1) : wend - j;
          skip = min(skip, wend - j);
          if (skip) {
            const u32 room = maxLiteral - lit;
            if (skip >= room) { skip = room; lit += skip; i += skip; LZS_PUSH(i - lit, lit, 0u, 0u); lit = 0; }
            else { lit += skip; i += skip; }
            continue;
          }
        }
        const FT dd = __shfl_sync(ZQ_FULL, lit ? c1 : c0, j);
        LzsDecision dc;
        if (dd & F::F_DEFER) dc = lzs_decide_exact<IdxT>(in, n, v, P, i, lit);
        else { dc.match = (dd & F::F_MATCH) != 0; dc.blit = F::f_blit(dd); dc.blen = F::f_blen(dd); dc.off = F::f_off(dd); }
        if (dc.match) {
          const u32 blit = dc.blit, blen = dc.blen, off = dc.off;
          lit += blit;
          LZS_PUSH(i + blit - lit, lit, blen - blit, off);
          lit = 0;
          i += blen;
        } else {
          ++lit; ++i;
          if (lit >= maxLiteral) { LZS_PUSH(i - lit, lit, 0u, 0u); lit = 0; }
        }
      }
    }
    if (lit) LZS_PUSH(n - lit, lit, 0u, 0u);
    if ((nt & 31u) && lane < (nt & 31u)) *(uint4*)&tok[(nt & ~31u) + lane] = make_uint4(mine.pos, mine.lit, mine.mlen, mine.off);
#undef LZS_PUSH
    if (lane == 0) ntok[t] = nt;
  }
}

// ---- emit ---------------------------------------------------------------------------------------------------------
// bits a token occupies (level 1) / 8 x bytes (level 2)
__device__ __forceinline__ u64 lz_token_bits(const LzToken& k, u32 level, u32 minMatch, u32 rb) {
  u64 bits = 0;
  if (level == 1) {
    if (k.lit) bits += 2 + (2 * (zq_bitlen(k.lit) - 1) + 1) + 8ull * k.lit;
    if (k.mlen) {
      const u32 off = k.off + (1u << rb) - 1;
      const u32 lo = (u32)zq_bitlen(off) - 1 - rb;
      bits += 5 + (2 * (zq_bitlen(k.mlen >> 2) - 1) + 1) + 2 + rb + lo;
    }
  } else {
    bits += 8ull * (k.lit + (k.lit + 63) / 64);
    if (k.mlen) {
      const u32 ob = k.off - 1 < (1u << 16) ? 3 : k.off - 1 < (1u << 24) ? 4 : 5;
      u32 len = k.mlen; const u32 mm = minMatch;
      while (len > 0) { const u32 l1 = len > mm * 2 + 63 ? mm + 63 : len > mm + 63 ? len - mm : len; bits += 8ull * ob; len -= l1; }
    }
  }
  return bits;
}

// writes bit strings at absolute bit positions of a zeroed, 4-byte aligned buffer (shared or global memory)
struct BitOut {
  u32* w;
  __device__ __forceinline__ void put(u64 bitpos, u64 code, u32 nbits) {   // nbits <= 57
    if (!nbits) return;
    code &= nbits < 64 ? ((1ull << nbits) - 1) : ~0ull;
    const u64 wi = bitpos >> 5; const u32 sh = (u32)bitpos & 31u;
    const u64 lo = code << sh;                       // bits for words wi, wi+1
    if ((u32)lo) atomicOr(&w[wi], (u32)lo);
    if (sh + nbits > 32 && (u32)(lo >> 32)) atomicOr(&w[wi + 1], (u32)(lo >> 32));
    if (sh + nbits > 64) { const u32 hi = (u32)(code >> (64 - sh)); if (hi) atomicOr(&w[wi + 2], hi); }
  }
};

constexpr u32 LZE_SMEM_STREAM = 65536 + 65536 / 32 + 64 + 64;   // a 64 KiB block's stream fits shared memory
constexpr int LZE_NT = 256;
constexpr u32 LZE_LONG = 24;      // literal runs longer than this are copied by a whole warp
constexpr u32 LZE_QCAP = 1024;

struct LzeSmem {
  u64 wsum[LZE_NT / 32];
  u64 carry;
  u32 nlong, pad;
  u32 qtok[LZE_QCAP];
  alignas(128) u32 stream[LZE_SMEM_STREAM / 4];
};

// literal bytes src[0..cnt) at bit position pos; `lanes` threads (this one is number `me`) share the copy
__device__ __forceinline__ void lze_copy_literals(BitOut& bo, u64 pos, const u8* __restrict__ src, u32 cnt, u32 me, u32 lanes) {
  for (u32 j = me * 4; j < cnt; j += lanes * 4) {
    const u32 k = min(4u, cnt - j);
    u32 v4 = 0;
    for (u32 b = 0; b < k; ++b) v4 |= (u32)src[j + b] << (8 * b);
    bo.put(pos + 8ull * j, v4, 8 * k);
  }
}

// One CTA per block: tokens -> bit offsets (scan) -> bits.  Blocks whose stream fits LZE_SMEM_STREAM are assembled in
// shared memory and stored with one bulk copy; larger ones are OR-ed into the (zeroed) global stream.
__global__ void __launch_bounds__(LZE_NT)
k_lz_emit(const u8* __restrict__ in_base, const ZqUnit* __restrict__ units, const ZqPlan* __restrict__ plans,
          const int* __restrict__ todo, int ntodo, const u64* __restrict__ tok_off, const LzToken* __restrict__ tok_base,
          const u32* __restrict__ ntok, u64* __restrict__ bitpos_base, u8* __restrict__ lz_base, u32* __restrict__ lz_len,
          u32* __restrict__ err_flag) {
  ZQ_DYN_SMEM(smem_raw);
  LzeSmem& sm = *reinterpret_cast<LzeSmem*>(smem_raw);
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int t = blockIdx.x; t < ntodo; t += gridDim.x) {
    const int ui = todo[t];
    const ZqUnit u = units[ui];
    const ZqPlan pl = plans[u.plan];
    const u32 level = pl.lz_level, minMatch = pl.args[2], rb = pl.args[0] > 4 ? pl.args[0] - 4 : 0;
    const u8* __restrict__ in = in_base + u.in_off;
    const LzToken* tok = tok_base + tok_off[t];
    u64* bp = bitpos_base + tok_off[t];
    const u32 nt = ntok[t];
    const bool in_smem = u.lz_cap <= LZE_SMEM_STREAM;
    if (tid == 0) { sm.carry = 0; sm.nlong = 0; }
    if (in_smem) for (u32 k = tid; k < (u.lz_cap + 3) / 4; k += LZE_NT) sm.stream[k] = 0;
    __syncthreads();
    // exclusive scan of the token bit lengths
    for (u32 b0 = 0; b0 < nt; b0 += LZE_NT) {
      const u32 k = b0 + tid;
      const u64 mine = k < nt ? lz_token_bits(tok[k], level, minMatch, rb) : 0;
      u64 inc = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const u64 x = __shfl_up_sync(ZQ_FULL, inc, o); if (lane >= (u32)o) inc += x; }
      if (lane == 31) sm.wsum[warp] = inc;
      __syncthreads();
      u64 pre = sm.carry;
      for (u32 w = 0; w < warp; ++w) pre += sm.wsum[w];
      if (k < nt) bp[k] = pre + inc - mine;
      __syncthreads();
      if (tid == LZE_NT - 1) sm.carry = pre + inc;
      __syncthreads();
    }
    const u64 total_bits = sm.carry;
    const u64 total_bytes = (total_bits + 7) >> 3;
    if (tid == 0) { lz_len[ui] = (u32)total_bytes; if (total_bytes > u.lz_cap) atomicOr(err_flag, 1u); }
    if (total_bytes > u.lz_cap) { __syncthreads(); continue; }
    BitOut bo; bo.w = in_smem ? sm.stream : (u32*)(lz_base + u.lz_off);
    for (u32 k = tid; k < nt; k += LZE_NT) {
      const LzToken tk = tok[k];
      u64 pos = bp[k];
      if (level == 1) {
        if (tk.lit) {
          u32 nb; const u32 g = gamma_code(tk.lit, &nb);
          bo.put(pos, (u64)g << 2, nb + 2); pos += nb + 2;
          if (tk.lit > LZE_LONG) {
            const u32 slot = atomicAdd(&sm.nlong, 1u);
            if (slot < LZE_QCAP) sm.qtok[slot] = k; else lze_copy_literals(bo, pos, in + tk.pos, tk.lit, 0, 1);
          } else lze_copy_literals(bo, pos, in + tk.pos, tk.lit, 0, 1);
          pos += 8ull * tk.lit;
        }
        if (tk.mlen) {
          const u32 off = tk.off + (1u << rb) - 1;
          const u32 lo = (u32)zq_bitlen(off) - 1 - rb;
          u32 nb; const u32 g = gamma_code(tk.mlen >> 2, &nb);
          const u64 head = (u64)((lo + 8) >> 3) | ((u64)(lo & 7) << 2) | ((u64)g << 5) | ((u64)(tk.mlen & 3) << (5 + nb));
          bo.put(pos, head, 7 + nb); pos += 7 + nb;
          bo.put(pos, ((u64)off & ((1u << rb) - 1)) | ((u64)((off >> rb) & ((1u << lo) - 1)) << rb), rb + lo);
        }
      } else {
        u32 lit = tk.lit; const u8* src = in + tk.pos;
        while (lit > 0) {
          const u32 l1 = lit > 64 ? 64 : lit;
          bo.put(pos, l1 - 1, 8); pos += 8;
          lze_copy_literals(bo, pos, src, l1, 0, 1); pos += 8ull * l1;
          src += l1; lit -= l1;
        }
        u32 len = tk.mlen; const u32 mm = minMatch; const u32 off = tk.off - 1;
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
    // long literal runs (level 1): one warp per run
    const u32 nl = min(sm.nlong, LZE_QCAP);
    for (u32 q = warp; q < nl; q += LZE_NT / 32) {
      const u32 k = sm.qtok[q];
      const LzToken tk = tok[k];
      u32 nb; (void)gamma_code(tk.lit, &nb);
      lze_copy_literals(bo, bp[k] + nb + 2, in + tk.pos, tk.lit, lane, 32);
    }
    __syncthreads();
    if (in_smem) {   // one bulk store of the finished stream (16-byte granules; the slot is padded to 16)
      zq_fence_async_smem();
      __syncthreads();
      if (tid == 0 && total_bytes) { zq_bulk_s2g(lz_base + u.lz_off, sm.stream, (u32)((total_bytes + 15) & ~15ull)); zq_bulk_commit_wait(); }
      __syncthreads();
    }
  }
}

}  // namespace zqdev
