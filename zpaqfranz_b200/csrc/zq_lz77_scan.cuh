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
// tile when it is done, at round boundaries every LZS_ROUND_STEPS steps: ~65-70 % of the lanes do useful steps instead
// of 13 % with one row per lane per loop.
// Bit-exact by construction: the per-row state machine is the reference's loop with one exact pruning rule (no
// remaining neighbour can beat the best score once 8*(h+runmin)-12, scaled, is not above it).
#pragma once
#include "zq_lz77.cuh"

namespace zqdev {

constexpr u32 LZS_HALO = 128;    // rows of context on either side of a tile (bucket <= 127)
constexpr u32 LZS_TILE = 4096;   // rows per tile
constexpr u32 LZS_ROWS = LZS_TILE + 2 * LZS_HALO;
constexpr int LZS_NB = 4;        // tiles resident per CTA (ring of bulk-copy buffers)
constexpr int LZS_NT = 512;      // threads per CTA of the scan kernels
#ifndef LZS_ROUND_STEPS
#define LZS_ROUND_STEPS 6      // scan steps between two event rounds, look-ahead-0 pass
#endif
#ifndef LZS_ROUND_STEPS1
#define LZS_ROUND_STEPS1 10    // ... look-ahead-1 pass (its rows are fewer and longer)
#endif
#ifndef LZS_OCC1
#define LZS_OCC1 2      // CTAs per SM the look-ahead-1 pass is compiled for
#endif
constexpr u32 LZS_CAP = 255;     // saturated LCP of a packed row: the position goes to the exact evaluator

// per-unit arrays behind sa | isa | lcp | bwt | pk in the work region: r0[n+1] (indexed by CONSUMER row) then f[n][2]
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
  // both decisions of a position with one store
#ifdef LZS_STORE_SPLIT
  static __device__ __forceinline__ void f_store2(u32* fo, u32 a, u32 b) { fo[0] = a; fo[1] = b; }
#else
  static __device__ __forceinline__ void f_store2(u32* fo, u32 a, u32 b) { *reinterpret_cast<uint2*>(fo) = make_uint2(a, b); }
#endif
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
#ifdef LZS_STORE_SPLIT
  static __device__ __forceinline__ void f_store2(u64* fo, u64 a, u64 b) { fo[0] = a; fo[1] = b; }
#else
  static __device__ __forceinline__ void f_store2(u64* fo, u64 a, u64 b) { *reinterpret_cast<ulonglong2*>(fo) = make_ulonglong2(a, b); }
#endif
};

__host__ __device__ inline u32 lzs_tiles(u32 n) { return (n + LZS_TILE - 1) / LZS_TILE; }

template <typename IdxT>
struct LzsView {
  const IdxT* sa; const IdxT* isa; const u16* lcp; const u8* bwt; const typename LzsPack<IdxT>::T* pk;
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
  v.pk = (const typename LzsPack<IdxT>::T*)r;
  r += zq_pk_bytes(n, sizeof(IdxT));
  v.r0 = (typename LzsFmt<IdxT>::R0T*)r;
  v.f = (typename LzsFmt<IdxT>::FT*)(r + zq_align128(((u64)n + 1) * sizeof(typename LzsFmt<IdxT>::R0T)));
  return v;
}

struct LzsDesc {   // one resident tile; everything the event round needs, so that it costs a few shared-memory loads
  u32 valid, n, t1, adj;          // adj: shared-memory index of row r is r + adj
  u32 minMatch, bucket, lookahead, checkbits, level, pad;
  const void* isa; void* r0; void* f;
};

template <typename IdxT>
struct LzsSmem {
  ZqMbar full[LZS_NB];      // bulk copy of the tile in this buffer has landed (one phase per reuse)
  u32 next_row[LZS_NB];     // next unclaimed row of the tile
  u32 exited[LZS_NB];       // lanes that have left the tile; the last one out refills the buffer
  LzsDesc d[LZS_NB];
  alignas(128) typename LzsPack<IdxT>::T pk[LZS_NB][LZS_ROWS];
};

// One lane: claim the next tile of the launch for buffer `sl` and start its bulk copy (or mark the buffer invalid
// when the launch has no tiles left; the barrier phase completes either way).
template <typename IdxT>
__device__ void lzs_fetch(LzsSmem<IdxT>& sm, u32 sl, const ZqUnit* __restrict__ units, const ZqPlan* __restrict__ plans,
                          const int* __restrict__ todo, const u32* __restrict__ tile_first, int ntodo, u32 uniform_tpu,
                          u8* work_base, u32* tile_ctr) {
  LzsDesc& d = sm.d[sl];
  const u32 tile = atomicAdd(tile_ctr, 1u);
  sm.exited[sl] = 0;
  const u32 total = uniform_tpu ? uniform_tpu * (u32)ntodo : tile_first[ntodo];
  if (tile >= total) { d.valid = 0; d.t1 = 0; sm.next_row[sl] = 0; zq_mbar_expect_tx(&sm.full[sl], 0); return; }
  int lo = 0;
  u32 first_tile;
  if (uniform_tpu) { lo = (int)(tile / uniform_tpu); first_tile = (u32)lo * uniform_tpu; }   // every block has the same number of tiles
  else {
    int hi = ntodo - 1;      // last t with tile_first[t] <= tile
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (tile_first[mid] <= tile) lo = mid; else hi = mid - 1; }
    first_tile = tile_first[lo];
  }
  const ZqUnit u = units[todo[lo]];
  const ZqPlan& pl = plans[u.plan];
  const u32 n = u.n;
  const LzsView<IdxT> v = lzs_view<IdxT>(work_base, u.work_off, n);
  d.valid = 1; d.n = n;
  d.minMatch = pl.args[2]; d.bucket = (1u << pl.args[4]) - 1; d.lookahead = pl.args[6]; d.checkbits = 17 + pl.args[0]; d.level = pl.lz_level;
  d.isa = v.isa; d.r0 = v.r0; d.f = v.f;
  const u32 t0 = (tile - first_tile) * LZS_TILE;
  d.t1 = min(n, t0 + LZS_TILE);
  const u32 first = t0 >= LZS_HALO ? t0 - LZS_HALO : 0u;
  d.adj = sl * LZS_ROWS - first;
  const u32 rhi = min(d.t1 + LZS_HALO, (n + 63u) & ~63u);
  const u32 bytes = (rhi - first) * (u32)sizeof(typename LzsPack<IdxT>::T);
  sm.next_row[sl] = t0;
  zq_mbar_expect_tx(&sm.full[sl], bytes);
  zq_bulk_g2s(sm.pk[sl], v.pk + first, bytes, &sm.full[sl]);
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
__device__ __forceinline__ typename LzsFmt<IdxT>::FT lzs_decide(u32 minMatch, u32 level, u32 off, u32 blen, u32 blit, int bscore) {
  typedef LzsFmt<IdxT> F;
  const bool match = off > 0 && bscore > 0 &&
                     blen - blit >= minMatch + (level == 2 ? (u32)(off >= (1u << 16)) + (u32)(off >= (1u << 24)) : 0u);
  return match ? F::f_pack(off, blen, blit) : (typename F::FT)0;
}
__device__ __forceinline__ int lzs_scale58(int sc) { return sc * 5 / 8; }   // C division: truncates toward zero like the reference
// The look-ahead pass only ever compares scaled scores with a best score that is positive (a position without a match
// has no look-ahead, Z:19412), which makes the truncation irrelevant below zero:
//   lzs_ub58(rm)  = scale58(8*(1+rm) - 12) = trunc(5*rm - 2.5) = 5*rm - 3 for rm >= 1; for rm == 0 both are negative
//   lzs_pos58(x)  = floor(5*x/8): equal to scale58(x) for x >= 0, and like it not above zero for x < 0
#ifdef LZS_SCALE_EXACT
__device__ __forceinline__ int lzs_ub58(u32 rm) { return lzs_scale58((int)(rm * 8u) - 4); }
__device__ __forceinline__ int lzs_pos58(int x) { return lzs_scale58(x); }
#else
__device__ __forceinline__ int lzs_ub58(u32 rm) { return (int)(rm * 5u) - 3; }
__device__ __forceinline__ int lzs_pos58(int x) { return (x * 5) >> 3; }
#endif

// Both passes.  Per-lane state machine, the scan step written without branches so the warp stays converged: `left`
// counts the steps the lane may still take in its current direction (0 = row done / no row), x is the shared-memory
// index of the next neighbour; the switch from the backward to the forward direction is part of the step.
// Every LZS_ROUND_STEPS (look-ahead 1: LZS_ROUND_STEPS1) steps an event round runs, warp-converged except for the result stores:
//   * lanes whose row is done store its result;
//   * free lanes take rows from the warp's queue; when the queue is empty the warp SWEEPS: it claims 32 rows of its
//     current tile with one atomic, evaluates all of them at once (PASS 1: coalesced load of r0, rows the look-ahead
//     cannot improve are decided on the spot) and queues the ones that need a scan.
// Warps move through the tiles of the CTA's ring on their own: nobody waits at a tile boundary for a lane that is
// still on a 254-step row.  A warp leaves a tile when it has swept past it and all its rows from that tile are
// finished; the last warp out refills the buffer (next tile of the launch, bulk copy).  A warp that finds the next
// tile not landed yet keeps scanning what it has and tests the barrier again next round.
//   PASS 0: look-ahead 0 of position s (row q); the result is scattered to r0[isa[s+1]], the row that continues it.
//   PASS 1: row q continues position s-1 with look-ahead 1 for both literal states -> f[s-1][0..1].
template <typename IdxT>
struct LzsQueue { u32 q0[32]; typename LzsFmt<IdxT>::R0T a[32]; };

template <typename IdxT, int PASS>
__global__ void __launch_bounds__(LZS_NT, PASS == 0 ? 3 : LZS_OCC1)
k_lz_scan(const ZqUnit* __restrict__ units, const ZqPlan* __restrict__ plans, const int* __restrict__ todo,
          const u32* __restrict__ tile_first, int ntodo, u32 uniform_tpu, u8* __restrict__ work_base, u32* tile_ctr) {
  typedef LzsFmt<IdxT> F;
  typedef LzsPack<IdxT> K;
  typedef typename K::T PT;
  typedef typename F::R0T R0T;
  typedef typename F::FT FT;
  ZQ_DYN_SMEM(smem_raw);
  LzsSmem<IdxT>& sm = *reinterpret_cast<LzsSmem<IdxT>*>(smem_raw);
  LzsQueue<IdxT>* wqs = reinterpret_cast<LzsQueue<IdxT>*>(smem_raw + sizeof(LzsSmem<IdxT>));
  const u32 tid = threadIdx.x, lane = tid & 31;
  LzsQueue<IdxT>& wq = wqs[tid >> 5];
  constexpr u32 NWARPS = LZS_NT / 32;
  if (tid == 0) for (int b = 0; b < LZS_NB; ++b) zq_mbar_init(&sm.full[b], 1);
  __syncthreads();
  if (tid < (u32)LZS_NB) lzs_fetch<IdxT>(sm, tid, units, plans, todo, tile_first, ntodo, uniform_tpu, work_base, tile_ctr);
  __syncthreads();
  const PT* __restrict__ s_pk = &sm.pk[0][0];
  // warp-uniform: tile being swept, its barrier still to be tested, launch out of tiles, queue, rows out per buffer
  u32 wseq = 0, qhead = 0, qlen = 0, pend = 0, swept = 0, held = 0, dead = 0;
  bool wwait = true, wdone = false;
  // per lane
  // PASS 1 runs ONE state machine per row: first for "literals pending" (no penalty).  The "no literal pending" state
  // scores a candidate that needs a leading literal 4 lower (Z:19421); it can only end differently if the first run
  // ACCEPTED such a candidate (a rejected one is rejected with the penalty too, and nothing else differs) -- true for
  // one row in ten, which then runs a second time with the penalty on (`pen`).
  bool have = false, defer = false, stop0 = false, pen = false, div = false;
  u32 qq = 0, s = 0, x = 0, carry = 0, runmin = 0, aux = 0, lb = 0;   // aux: PASS 0 consumer row, PASS 1 own BWT byte
  u32 blen0 = 0, bp0 = 0, blit0 = 0;
  R0T keep = 0;                                                       // PASS 1: the row's r0 word, for the second run
  int left = 0, fwd_left = 0, bwd_left = 0, dstep = 1, bs0 = 0, mm = 0;
  for (;;) {
    // ---- event round ----
    bool fin = have && left == 0;
    if (fin) {
      const LzsDesc& d = sm.d[lb];
      if (PASS == 0) ((R0T*)d.r0)[aux] = defer ? F::R0_DEFER : (bs0 > 0 ? F::r0_pack(blen0, bp0) : (R0T)0);
      else {
        const u32 i = s - 1;
        FT* fo = (FT*)d.f + 2 * (u64)i;
        const FT dd = defer ? F::F_DEFER : lzs_decide<IdxT>(d.minMatch, d.level, i - bp0, blen0, blit0, bs0);
        if (pen) fo[0] = dd;                       // second run: the "no literal pending" decision
        else if (div && !defer) {                  // first run accepted a candidate with a leading literal: run again
          fo[1] = dd;
          pen = true; div = false; fin = false;
          const u32 bl = F::r0_blen(keep), bpp = F::r0_bp(keep);
          blen0 = bl; bp0 = bpp; blit0 = 0; bs0 = (int)(bl * 8u) - zq_bitlen(i - bpp) - 11; stop0 = false;
          x = qq - 1; dstep = -1; runmin = 0xffffffffu; carry = K::lcp(s_pk[qq]);
          left = bwd_left;
          if (left == 0) { x = qq + 1; dstep = 1; left = fwd_left; }
        } else F::f_store2(fo, dd, dd);
      }
      if (fin) have = false;
    }
    pend -= __reduce_add_sync(ZQ_FULL, fin ? 1u << (8 * lb) : 0u);
    // leave the tiles this warp is through with
    for (u32 m = swept & held; m; m &= m - 1) {
      const u32 b = (u32)__ffs(m) - 1;
      if ((pend >> (8 * b)) & 255u) continue;
      held &= ~(1u << b); swept &= ~(1u << b);
      if (lane == 0 && atomicAdd(&sm.exited[b], 1u) == NWARPS - 1)
        lzs_fetch<IdxT>(sm, b, units, plans, todo, tile_first, ntodo, uniform_tpu, work_base, tile_ctr);
      __syncwarp();
    }
    // hand rows to the free lanes
    u32 freemask = __ballot_sync(ZQ_FULL, !have);
    while (freemask) {
      if (qlen == 0) {
        if (wdone) break;
        const u32 b = wseq % LZS_NB;
        if (dead & (1u << b)) { ++wseq; continue; }                          // this buffer got no tile any more
        if (wwait) {
          // (one lane looks, everyone gets the same answer: the warp's bookkeeping must not diverge)
          const bool landed = __shfl_sync(ZQ_FULL, (lane == 0 && zq_mbar_test(&sm.full[b], (wseq / LZS_NB) & 1u)) ? 1 : 0, 0) != 0;
          if (!landed) break;                                                 // not landed yet: try again next round
          wwait = false;
          if (!sm.d[b].valid) {   // the launch ran out of tiles when this buffer was to be refilled; others may still hold some
            dead |= 1u << b; ++wseq; wwait = true;
            wdone = dead == (1u << LZS_NB) - 1;
            continue;
          }
          held |= 1u << b;
        }
        const LzsDesc& d = sm.d[b];
        u32 base = 0;
        if (lane == 0) base = atomicAdd(&sm.next_row[b], 32u);
        base = __shfl_sync(ZQ_FULL, base, 0);
        if (base >= d.t1) { swept |= 1u << b; ++wseq; wwait = true; continue; }   // (left at the next event round)
        const u32 r = base + lane;
        bool push = r < d.t1;
        const u32 q0 = r + d.adj;
        R0T a = 0;
        if (PASS == 1 && push) {
          const u32 n = d.n;
          const PT w = s_pk[q0];
          const u32 ss = K::sa(w);
          // this row's job: position i = s-1 (it is the row of suffix i+1); the row of suffix 0 takes position n-1,
          // which has no look-ahead
          const u32 i = ss > 0 ? ss - 1 : n - 1;
          a = ((const R0T*)d.r0)[ss > 0 ? r : n];
          FT* fo = (FT*)d.f + 2 * (u64)i;
          if (a == F::R0_DEFER) { F::f_store2(fo, F::F_DEFER, F::F_DEFER); push = false; }
          else {
            const u32 bl = a ? F::r0_blen(a) : d.minMatch - 1, bpp = a ? F::r0_bp(a) : 0u;
            const int bsc = a ? (int)(bl * 8u) - zq_bitlen(i - bpp) - 11 : 0;
            const bool cont = ss > 0 && d.lookahead >= 1 && bsc > 0 && bl >= d.minMatch && (ss >> d.checkbits) == (i >> d.checkbits);
            // nothing to gain from the look-ahead either when the rows next to q share too little with it: the
            // first step of each direction would already be pruned (ub <= best score)
            const u32 lcf = r + 1 < n ? K::lcp(s_pk[q0 + 1]) : 0u;
            const bool nogain = lzs_ub58(max(K::lcp(w), lcf)) <= bsc;      // (only looked at when cont, i.e. bsc > 0)
            if (!cont || nogain) { const FT dd = lzs_decide<IdxT>(d.minMatch, d.level, i - bpp, bl, 0, bsc); F::f_store2(fo, dd, dd); push = false; }
          }
        }
        const u32 pm = __ballot_sync(ZQ_FULL, push);
        if (push) { const u32 slot = (qhead + (u32)__popc(pm & lanemask_lt())) & 31u; wq.q0[slot] = q0; if (PASS == 1) wq.a[slot] = a; }
        qlen = (u32)__popc(pm);
        pend += qlen << (8 * b);
        __syncwarp();
        if (qlen == 0) continue;
      }
      const u32 take = min((u32)__popc(freemask), qlen);
      const u32 rank = (u32)__popc(freemask & lanemask_lt());
      if (!have && rank < take) {
        const u32 slot = (qhead + rank) & 31u;
        const u32 q0 = wq.q0[slot];
        lb = q0 / LZS_ROWS;
        const LzsDesc& d = sm.d[lb];
        const u32 n = d.n, r = q0 - d.adj;
        const PT w = s_pk[q0];
        const u32 ss = K::sa(w);
        if (PASS == 0) {
          aux = ss + 1 < n ? (u32)((const IdxT*)d.isa)[ss + 1] : n;   // in flight during the scan, needed when the row is done
          blen0 = d.minMatch - 1; bp0 = 0; bs0 = 0;
        } else {
          const R0T a = wq.a[slot];
          const u32 i = ss - 1;
          const u32 bl = F::r0_blen(a), bpp = F::r0_bp(a);
          const int bsc = (int)(bl * 8u) - zq_bitlen(i - bpp) - 11;
          aux = K::bwt(w); keep = a;
          blen0 = bl; bp0 = bpp; blit0 = 0; bs0 = bsc; stop0 = false; pen = false; div = false;
        }
        qq = q0; s = ss; carry = K::lcp(w);
        have = true; x = qq - 1; dstep = -1; runmin = 0xffffffffu; defer = false;
        left = bwd_left = (int)min(d.bucket, r);
        fwd_left = (int)min(d.bucket, n - 1 - r);
        mm = (int)d.minMatch;
        if (left == 0) { x = qq + 1; dstep = 1; left = fwd_left; }   // first row of the block: forward only
        // (a block of one row has nothing to scan: left stays 0 and the row is stored at the next event round)
      }
      __syncwarp();
      qhead = (qhead + take) & 31u; qlen -= take;
      freemask = __ballot_sync(ZQ_FULL, !have);
    }
    if (wdone && held == 0 && freemask == ZQ_FULL) break;
    if (__all_sync(ZQ_FULL, left == 0)) {   // nobody has a step to take: a tile is landing, or rows finished at once
      if (freemask == ZQ_FULL) __nanosleep(100);
      continue;
    }
    // ---- scan steps ----
#pragma unroll
    for (int r = 0; r < (PASS == 0 ? LZS_ROUND_STEPS : LZS_ROUND_STEPS1); ++r) {
      const bool inr = left > 0;
      const PT w = s_pk[inr ? x : qq];
      const u32 p = K::sa(w), lc = K::lcp(w);
      const u32 e = dstep < 0 ? carry : lc;
      carry = lc;
      const u32 rm = min(runmin, e);
      const bool capped = rm >= LZS_CAP;
      bool go;          // keep going in this direction
      bool dnow;        // this candidate needs the exact evaluator
      if (PASS == 0) {
        const int t8 = (int)(rm * 8u) - 12;
        const bool live = inr && t8 > bs0;                   // exact pruning: nothing left can beat the best score
        const bool valid = live && p < s;
        const int sc = t8 - (31 - __clz(s - p));             // 8*l - lg(s-p) - 11
        const bool take = valid && !capped && sc > bs0;
        blen0 = take ? rm : blen0; bp0 = take ? p : bp0; bs0 = take ? sc : bs0;
        dnow = valid && capped;
        go = live && !dnow && !(valid && (int)rm < max((int)blen0, mm));
      } else {
        const u32 bw = K::bwt(w);
        const int ub = lzs_ub58(rm);                         // ((1+rm)*8 - 12) * 5/8
        stop0 = stop0 || ub <= bs0;
        const bool live = inr && !stop0;
        const bool valid = live && p != 0 && p < s;          // candidate p-1 < i
        const u32 l = 1u + rm;
        const u32 l1 = bw == aux ? 0u : 1u;
        const int base = (int)((l - l1) * 8u) - (32 - __clz(s - p)) - 11;
        const int sc = lzs_pos58(base - ((pen && l1) ? 4 : 0));
        const bool ok = valid && !capped;
        const bool t0 = ok && sc > bs0;
        blen0 = t0 ? l : blen0; bp0 = t0 ? p - 1 : bp0; blit0 = t0 ? l1 : blit0; bs0 = t0 ? sc : bs0;
        div = div || (t0 && l1 != 0);
        stop0 = stop0 || (ok && ((int)l < max((int)blen0, mm)));   // (l > 255 cannot happen below the cap)
        dnow = valid && capped;
        go = live && !dnow && !stop0;
      }
      defer = defer || dnow;
      // next step: same direction, or the first forward neighbour once the backward direction has ended
      const bool sw = inr && !dnow && dstep < 0 && (!go || left == 1);   // ended by the rules, or out of backward neighbours
      x = sw ? qq + 1 : x + dstep;
      runmin = sw ? 0xffffffffu : rm;
      left = sw ? fwd_left : (go ? left - 1 : 0);
      dstep = sw ? 1 : dstep;
      if (PASS == 1) stop0 = stop0 && !sw;
    }
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
// EXACT = false: the lean form (no exact evaluator in the kernel: half the registers, twice the warps per SM); a block
// in which the walk meets a deferred position is abandoned with ntok[t] = LZW_REDO and walked again by the EXACT = true
// form, which is launched right after and looks at those blocks only.
constexpr u32 LZW_REDO = 0xffffffffu;
template <typename IdxT, bool EXACT>
__global__ void __launch_bounds__(128, EXACT ? 6 : 12)
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
    if (EXACT && ntok[t] != LZW_REDO) continue;
    const ZqUnit u = units[todo[t]];
    const LzsParams P = lzs_params(plans[u.plan]);
    const u32 n = u.n;
    const u8* in = in_base + u.in_off;
    bool redo = false;
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
    while (i < n && !redo) {
      if (i >= w0 + 32) {
        if (i < w0 + 64) { w0 += 32; c0 = n0; c1 = n1; }
        else { w0 = i & ~31u; c0 = c1 = 0; if (w0 + lane < n) { c0 = v.f[2 * (u64)(w0 + lane)]; c1 = v.f[2 * (u64)(w0 + lane) + 1]; } }
        n0 = n1 = 0;
        if (w0 + 32 + lane < n) { n0 = v.f[2 * (u64)(w0 + 32 + lane)]; n1 = v.f[2 * (u64)(w0 + 32 + lane) + 1]; }
      }
      // positions of this window that do not end as a plain literal, per state
      const u32 m0 = __ballot_sync(ZQ_FULL, c0 != 0), m1 = __ballot_sync(ZQ_FULL, c1 != 0);
      const u32 wend = min(32u, n - w0);
      while (i < w0 + wend && !redo) {
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
        if (dd & F::F_DEFER) {
          if (EXACT) dc = lzs_decide_exact<IdxT>(in, n, v, P, i, lit);
          else { redo = true; dc.match = false; dc.blit = dc.blen = dc.off = 0; }
        } else { dc.match = (dd & F::F_MATCH) != 0; dc.blit = F::f_blit(dd); dc.blen = F::f_blen(dd); dc.off = F::f_off(dd); }
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
    if (redo) { if (lane == 0) ntok[t] = LZW_REDO; continue; }
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
